// vloam_b200 — shared device/host definitions for the sm_100a kernels.
//
// Layout conventions (DESIGN.md "Data layout in HBM"):
//   * a handle owns B independent LiDAR streams ("batch"); every per-stream
//     array is a slab of a [B][capacity] device buffer, so one launch covers the
//     whole batch with blockIdx.y = stream.
//   * clouds are arrays of float4 (x, y, z, intensity) — pcl::PointXYZI's 16-byte
//     record (reference include/lidar_odometry_mapping/common.h:43) — so a warp
//     reads 512 contiguous bytes.
//   * poses, residuals and normal equations are double (the reference does the
//     same: laser_odometry.cpp:158-165, lidarFactor.hpp:24-42).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

constexpr int kMaxRings = 64;         // N_SCANS <= 64 (scan_registration.cpp:56)
constexpr int kSectors = 6;           // scan_registration.cpp:317
constexpr int kSharpPerSector = 2;    // scan_registration.cpp:335
constexpr int kLessSharpPerSector = 20;  // scan_registration.cpp:341
constexpr int kFlatPerSector = 4;     // scan_registration.cpp:391
constexpr int kRingCap = 4096;        // max points of one ring handled by the per-ring CTA
constexpr int kSectorCap = 1024;      // max points of one sector (power of two)
constexpr int kMaxSharp = kMaxRings * kSectors * kSharpPerSector;          // 768
constexpr int kMaxLessSharp = kMaxRings * kSectors * kLessSharpPerSector;  // 7680
constexpr int kMaxFlat = kMaxRings * kSectors * kFlatPerSector;            // 1536
constexpr int kClassifyBlock = 1024;  // points per block in classify / scatter

// status bits reported per stream (vloam_b200.h: VLOAM_STREAM_*)
constexpr int kStatusEmpty = 1;        // no point survived NaN / range filters
constexpr int kStatusRingOverflow = 2; // a ring exceeded kRingCap or a sector kSectorCap
constexpr int kStatusVoxelOverflow = 4;  // PCL's "leaf size too small" path was taken (input returned unfiltered)
constexpr int kStatusCapacity = 8;     // the scan had more points than the handle's capacity: the excess was ignored

// Per-stream scan-registration bookkeeping, resident on the device.
struct SRHeader {
  int n_in;
  int firstValid, lastValid;
  float startOri, endOri;
  int halfIdx;    // smallest kept index whose un-wrapped azimuth passed startOri + pi
  int cloudSize;  // points kept (the reference's `count`)
  int status;
  int ringCount[kMaxRings];
  int ringStart[kMaxRings + 1];
  int ringLessFlat[kMaxRings];  // size of each ring's down-sampled less-flat cloud
  int secCount[kMaxRings * kSectors * 3];  // sharp / lessSharp / flat picked per (ring, sector)
  // packed feature clouds of this scan (filled by sr_pack)
  int nSharp, nLessSharp, nFlat, nLessFlat;
  int ringStartLessSharp[kMaxRings + 1];
  int ringStartLessFlat[kMaxRings + 1];
};

// Uniform xy-column grid over one target cloud (the kd-tree replacement of laserOdometry, laser_odometry.cpp:525-526).
// Points are counting-sorted by column; column (ix, iy) owns sorted[cellStart[iy*nx+ix] .. cellStart[iy*nx+ix+1]).
constexpr int kGridCap = 40000;    // max columns per grid (the column table lives in shared memory while it is built)
constexpr float kGridCell = 1.0f;  // column size (m); grows by 1.25x until the cloud's xy extent fits kGridCap columns
struct GridHeader {
  float minx, miny, c, inv_c;
  int nx, ny, n;
  // The reference classifies target points by int(intensity), which equals the point's true ring R or, when its
  // relTime is negative, R - 1.  With that property (ringsOk) the two walks of laser_odometry.cpp:279-324 / 368-417
  // visit exactly an index interval that follows from the true ring offsets and these two per-ring tables.
  int ringsOk;
  int ringStart[kMaxRings + 2];  // true ring offsets of the ring-major cloud; [64] = [65] = n
  int firstFull[kMaxRings + 1];  // first index in ring R whose int(intensity) == R      (INT_MAX if none)
  int lastLow[kMaxRings + 1];    // last index in ring R whose int(intensity) == R - 1   (-1 if none)
};

// Levenberg-Marquardt trace record (mirrors oracle::LMIteration) for parity read-out.
struct LMRecord {
  double cost, candidate_cost, model_cost_change, relative_decrease, radius;
  double step_is_valid, step_is_successful;
};
constexpr int kMaxLMRecords = 8;  // iteration 0 + up to 7 further records are kept (later ones dropped)

struct SolveTrace {
  int n_records;
  int termination;
  int n_corner, n_plane;
  double para[7];
  LMRecord rec[kMaxLMRecords];
};

// Per-stream laser-odometry state.
struct LOState {
  double para_q[4];  // q_last_curr (x, y, z, w)      laser_odometry.cpp:84-87
  double para_t[3];  // t_last_curr                   laser_odometry.cpp:88-90
  double q_w[4];     // q_w_curr                      laser_odometry.cpp:80
  double t_w[3];     // t_w_curr                      laser_odometry.cpp:81
  int corner_correspondence, plane_correspondence;
  SolveTrace trace[2];
};


// ---------------------------------------------------------------------------------------------
// Point-sharded solve (SURVEY.md section 8e layout (ii), BASELINE configs[4]): every rank holds the same clouds, owns a
// slice of each stream's correspondences and the per-iteration normal equations are summed across ranks INSIDE the
// solve kernel through peer memory (NVLink P2P loads of the other ranks' exchange slots).  One slot per stream; a slot
// holds two generations of the vector (parity of the sequence number) and the sequence number of the newest one.
constexpr int kMaxShard = 8;
struct ShardSlot {
  double v[2][32];
  unsigned long long flag;     // sequence number of the newest published vector
  unsigned long long pad[7];
};
struct ShardView {
  int rank = 0, world = 1;
  ShardSlot* peer[kMaxShard] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [rank][B]
  unsigned long long* seq = nullptr;   // [B] local exchange counters (identical on every rank by construction)
  int* error = nullptr;                // set when a peer did not answer in time
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// squared distance exactly as flann::L2_Simple / the reference's window search
// accumulate it: ((dx*dx + dy*dy) + dz*dz) in float, no FMA contraction.
__device__ __forceinline__ float sqdist_f(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Eigen quaternion * vector: v + w*(2 u x v) + u x (2 u x v), q = (x, y, z, w).
__device__ __forceinline__ void quat_rotate(const double q[4], double vx, double vy, double vz, double out[3]) {
  double uvx = q[1] * vz - q[2] * vy, uvy = q[2] * vx - q[0] * vz, uvz = q[0] * vy - q[1] * vx;
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[0] = vx + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  out[1] = vy + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  out[2] = vz + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
}

}  // namespace vb
