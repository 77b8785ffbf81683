// vloam_b200 — the ceres::Solve of one outer pass as a sequence of WIDE launches: every Levenberg-Marquardt
// evaluation is one gn_accumulate launch over all residual blocks of all streams (grid = tiles x streams, every SM busy,
// 28-double partial normal equations per tile) followed by one gn_step launch (a warp per stream: fixed-order sum of the
// tile partials, 6 x 6 Cholesky and the Ceres trust-region bookkeeping of gn_solver.cuh).  Between the two sits the one
// real exchange step of the path: with the residuals of a stream split across GPUs (BASELINE configs[4]) the partials are
// all-reduced there (ncclAllReduce, see capi.cu).  The one-CTA-per-stream kernels (lo_solve, lm_solve) remain for small
// batches, where launch count matters more than width.
#pragma once
#include "common.cuh"
#include "gn_solver.cuh"

namespace vb {

struct Profiler;

// Residual record shared by laser odometry and laser mapping.
struct GNResidual {
  double v[7];   // edge: a(3), b(3); plane: n(3), d0
  float px, py, pz;
  int type;      // 0 none, 1 edge (LidarEdgeFactor), 2 plane (LidarPlaneFactor / LidarPlaneNormFactor: r = n . (q p + t) + d0);
                 // 3 / 4: edge / plane with the motion-distortion ratio s in v[6] / v[4] (laser_odometry.h:90 DISTORTION)
};

// Storage of the records: structure of arrays, so that consecutive threads read consecutive addresses — seven planes of
// doubles and one float4 (point + type) per record.  Block `k` (a stream, or a stream's corner / surf half) holds `n` slots:
// v[(k * 7 + plane) * n + i], p[k * n + i].  (Round 1 kept 72-byte records: every 8-byte load of a warp touched 32 sectors.)
struct GNRecArray {
  double* v;
  float4* p;
  int n;
};
__device__ __forceinline__ void gn_store(const GNRecArray& A, size_t block, int i, const GNResidual& R) {
  double* v = A.v + block * 7 * (size_t)A.n + i;
#pragma unroll
  for (int k = 0; k < 7; ++k) v[(size_t)k * A.n] = R.v[k];
  A.p[block * (size_t)A.n + i] = make_float4(R.px, R.py, R.pz, __int_as_float(R.type));
}
__device__ __forceinline__ int gn_type(const float4& p) { return __float_as_int(p.w); }

constexpr int kGnTiles = 8;          // CTAs per stream of gn_accumulate
constexpr int kGnThreads = 256;

struct GNState {                     // per stream, lives in global memory between the launches of one solve
  LMCore core;
  double evalX[7];                   // the point the next gn_accumulate evaluates at (x0, then the LM candidates)
  int active;                        // the stream has a problem to solve in this pass
  int pad;
};

// A per-stream field inside an array of per-stream structs: base + b * stride bytes.
struct Strided {
  const void* base; size_t stride;
  template <typename T> __device__ __forceinline__ const T* at(int b) const { return reinterpret_cast<const T*>(static_cast<const char*>(base) + (size_t)b * stride); }
  template <typename T> __device__ __forceinline__ T* at_mut(int b) const { return reinterpret_cast<T*>(const_cast<char*>(static_cast<const char*>(base)) + (size_t)b * stride); }
};

struct GNProblemView {
  GNRecArray rec;                    // the records; stream b owns the blocks b * blocksPerStream + (0 .. blocksPerStream - 1)
  int blocksPerStream;               // 1 (laser odometry) or 2 (laser mapping: corner / surf queries)
  Strided count[2];                  // int: live records of each array (count[i].base == nullptr: fixedCount[i])
  int fixedCount[2];
  Strided active;                    // int flag (nullptr: always active)
  Strided x;                         // double[7]: the parameters (read at the start, written back at the end)
  Strided trace;                     // SolveTrace
  int slerp;                         // the records may be of type 3 / 4 (motion-distortion ratio)
};

void launch_gn_solve(Profiler* prof, cudaStream_t st, int B, const GNProblemView& pv, GNState* gs, double* partial /*[B][kGnTiles][28]*/,
                     int max_iterations, int kidAccumulate, int kidStep, void* ncclComm /*or nullptr*/);

}  // namespace vb
