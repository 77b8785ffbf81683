// vloam_b200 — wide Gauss-Newton / Levenberg-Marquardt solve (see gn_split.cuh).
#include "gn_split.cuh"

#include "gn_jet.cuh"

#include "internal.h"

namespace vb {

// gn_begin: grid (ceil(B / 128)), block 128.  x0 -> evalX, activity flag.
__global__ void gn_begin(const GNProblemView pv, GNState* __restrict__ gs, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  GNState& G = gs[b];
  G.active = pv.active.base ? (*pv.active.at<int>(b) != 0) : 1;
  const double* x = pv.x.at<double>(b);
  for (int i = 0; i < 7; ++i) { G.evalX[i] = x[i]; G.core.x[i] = x[i]; }
  G.core.done = G.active ? 0 : 1;
}

// gn_accumulate: grid (kGnTiles, B), block 256.  One evaluation of (J'J, J'r, cost) at evalX over every residual block of
// the stream: analytic Jacobians per block, warp-shuffle + block reduction to 28 doubles per tile.  The summation order
// is fixed (thread-strided, warp tree, warp order, then tile order in gn_step): results are reproducible run to run.
template <bool SLERP>   // SLERP: the records may carry a motion-distortion ratio (types 3 / 4); a separate instance keeps the registers of the shipped one low
__global__ void __launch_bounds__(kGnThreads) gn_accumulate(const GNProblemView pv, const GNState* __restrict__ gs, double* __restrict__ partial) {
  __shared__ double s_red[28];
  __shared__ double s_scratch[32 * 28];
  const int b = blockIdx.y;
  const GNState& G = gs[b];
  if (!G.active || G.core.done) return;       // uniform over the CTA
  double acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.0;
  const double q[4] = {G.evalX[0], G.evalX[1], G.evalX[2], G.evalX[3]};
  const double t[3] = {G.evalX[4], G.evalX[5], G.evalX[6]};
  for (int a = 0; a < pv.blocksPerStream; ++a) {
    const int n = pv.count[a].base ? *pv.count[a].at<int>(b) : pv.fixedCount[a];
    const size_t block = (size_t)b * pv.blocksPerStream + a;
    const float4* P = pv.rec.p + block * (size_t)pv.rec.n;
    const double* V = pv.rec.v + block * 7 * (size_t)pv.rec.n;
    const size_t pl = (size_t)pv.rec.n;
    for (int i = blockIdx.x * kGnThreads + threadIdx.x; i < n; i += kGnTiles * kGnThreads) {
      const float4 p = P[i];
      const int type = gn_type(p);
      if (type == 0) continue;
      // type 1 / 2: the shipped configuration (interpolation ratio s == 1, closed-form Jacobians); 3 / 4: the same blocks with
      // a per-point ratio (DISTORTION == true) through Eigen's slerp, by dual numbers
      if (type == 1 || type == 3) {
        const double pa[3] = {V[i], V[pl + i], V[2 * pl + i]}, pb[3] = {V[3 * pl + i], V[4 * pl + i], V[5 * pl + i]};
        if (!SLERP || type == 1) edge_block(q, t, p, pa, pb, acc); else edge_block_slerp(q, t, p, pa, pb, V[6 * pl + i], acc);
      } else {
        const double nn[3] = {V[i], V[pl + i], V[2 * pl + i]};
        if (!SLERP || type == 2) plane_block(q, t, p, nn, V[3 * pl + i], acc); else plane_block_slerp(q, t, p, nn, V[3 * pl + i], V[4 * pl + i], acc);
      }
    }
  }
  block_reduce28(acc, s_red, s_scratch);
  if (threadIdx.x < 28) partial[((size_t)b * kGnTiles + blockIdx.x) * 28 + threadIdx.x] = s_red[threadIdx.x];
}

// gn_step: grid (ceil(B / 4)), block 128: one warp per stream.  phase 0 follows the evaluation at x0, phase k > 0 the
// evaluation of the k-th candidate.
__global__ void __launch_bounds__(128) gn_step(const GNProblemView pv, GNState* __restrict__ gs, const double* __restrict__ partial, int B,
                                               int phase, int max_iterations) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), l = lane_id();
  if (b >= B) return;
  GNState& G = gs[b];
  if (!G.active || G.core.done) return;       // uniform over the warp
  if (l < 28) {
    double s = 0.0;
    for (int tIdx = 0; tIdx < kGnTiles; ++tIdx) s += partial[((size_t)b * kGnTiles + tIdx) * 28 + l];
    G.core.red[l] = s;
  }
  __syncwarp();
  if (l == 0) {
    LMCore& S = G.core;
    SolveTrace* tr = pv.trace.at_mut<SolveTrace>(b);
    if (phase == 0) {
      S.euclid = 0;
      double x0[7];
      for (int i = 0; i < 7; ++i) x0[i] = G.evalX[i];
      lm_begin(S, tr, x0);
      lm_prepare_step(S, tr, max_iterations);
    } else {
      lm_finish_step(S, tr);
      if (!S.done) lm_prepare_step(S, tr, max_iterations);
    }
    if (S.done) {
      for (int i = 0; i < 7; ++i) tr->para[i] = S.x[i];
      tr->termination = S.termination;
      double* x = pv.x.at_mut<double>(b);
      for (int i = 0; i < 7; ++i) x[i] = S.x[i];
    } else {
      for (int i = 0; i < 7; ++i) G.evalX[i] = S.cand[i];
    }
  }
}

// Hook for the exchange step (point-sharded streams): defined in capi.cu, which owns the NCCL binding.
void gn_allreduce_partials(void* ncclComm, double* partial, size_t count, cudaStream_t st);

void launch_gn_solve(Profiler* prof, cudaStream_t st, int B, const GNProblemView& pv, GNState* gs, double* partial, int max_iterations,
                     int kidAccumulate, int kidStep, void* ncclComm) {
  VB_LAUNCH(prof, kidStep, st, gn_begin<<<(B + 127) / 128, 128, 0, st>>>(pv, gs, B));
  // evaluation 0 at x0, then one evaluation per LM iteration: a stream that is done (converged, failed or out of
  // iterations) is skipped by both kernels, so the late launches cost a few microseconds each
  for (int phase = 0; phase <= max_iterations; ++phase) {
    if (pv.slerp) VB_LAUNCH(prof, kidAccumulate, st, gn_accumulate<true><<<dim3(kGnTiles, B), kGnThreads, 0, st>>>(pv, gs, partial));
    else VB_LAUNCH(prof, kidAccumulate, st, gn_accumulate<false><<<dim3(kGnTiles, B), kGnThreads, 0, st>>>(pv, gs, partial));
    if (ncclComm) gn_allreduce_partials(ncclComm, partial, (size_t)B * kGnTiles * 28, st);
    VB_LAUNCH(prof, kidStep, st, gn_step<<<(B + 3) / 4, 128, 0, st>>>(pv, gs, partial, B, phase, max_iterations));
  }
}

}  // namespace vb
