// vloam_b200 — scanRegistration on sm_100a.
//
// Replaces vloam::ScanRegistration::input
// (reference src/lidar_odometry_mapping/src/scan_registration.cpp:131-449).
// One launch per stage covers the whole batch (blockIdx.y / blockIdx.x = stream).
//
//   sr_find_ends     :157-176  first/last valid point -> startOri / endOri
//   sr_classify      :157-158,186-262  NaN + range filter, elevation -> ring id, half-sweep index, ring histograms
//   sr_scan          :276-281  ring offsets (exclusive scan of the per-block histograms)
//   sr_scatter       :264-266,276-281  stable ring-major compaction + intensity = ring + 0.1*relTime
//   sr_curvature     :288-307  11-point curvature, strict left-to-right float sums (no FMA)
//   sr_pick_features :312-422  warp per ring: greedy sharp / flat picks with neighbour suppression (sort-free arg-max)
//   sr_less_flat_voxel :424-439  per ring: less-flat gather + pcl::VoxelGrid(0.2) restated in shared memory
//   sr_pack                    ring-major packing of the four feature clouds
//
// Azimuths go through vb_fdlibm::atan2f_fd (fdlibm_atan2f.h): the C library's atan2f bit for bit, so relTime / intensity and the
// 2*pi unwrapping decisions are the reference's.
// Bit-level decisions (ring id, curvature, thresholds, voxel keys) use explicit
// round-to-nearest intrinsics so nvcc cannot contract them into FMAs: the CPU
// reference is built without FMA (CMakeLists.txt:5-6) and the index sets depend on it.
#include <math_constants.h>

#include <cstdlib>

#include <cuda_pipeline.h>

#include "common.cuh"
#include "fdlibm_atan2f.h"
#include "internal.h"
#include "tma_bulk.cuh"

namespace vb {

__device__ __forceinline__ bool point_valid(float x, float y, float z, float thres2) {
  // pcl::removeNaNFromPointCloud (:157) then removeClosedPointCloud (:114-117)
  if (!isfinite(x) || !isfinite(y) || !isfinite(z)) return false;
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  return !(r2 < thres2);
}

// :192-226.  Returns ring id or -1.  atan/sqrt evaluated in double.
__device__ __noinline__ int ring_of_exact(float x, float y, float z, int n_scans) {
  const float xy2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const double a = atan((double)z / sqrt((double)xy2));
  const float angle = (float)(__ddiv_rn(__dmul_rn(a, 180.0), 3.14159265358979323846));
  int id;
  if (n_scans == 16) {
    id = (int)(__dadd_rn(__ddiv_rn(__dadd_rn((double)angle, 15.0), 2.0), 0.5));
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 32) {
    id = (int)(__ddiv_rn(__dmul_rn(__dadd_rn((double)angle, 92.0 / 3.0), 3.0), 4.0));
    if (id > n_scans - 1 || id < 0) return -1;
  } else {
    if ((double)angle >= -8.83)
      id = (int)(__dadd_rn(__dmul_rn(__dsub_rn(2.0, (double)angle), 3.0), 0.5));
    else
      id = n_scans / 2 + (int)(__dadd_rn(__dmul_rn(__dsub_rn(-8.83, (double)angle), 2.0), 0.5));
    if ((double)angle > 2.0 || (double)angle < -24.33 || id > 50 || id < 0) return -1;
  }
  return id;
}

// Same decisions as ring_of_exact.  For the HDL-64 table a float estimate of the elevation is used when it is
// farther than kAngleErr from every decision threshold (bin edges, -8.83, 2, -24.33): the estimate differs from the
// exactly-rounded `angle` by far less than that (atanf / rsqrt / mul: a few ulp, < 3e-5 deg), so both give the
// same ring; otherwise (about 0.1 % of the points) the double-precision path decides.
__device__ __forceinline__ int ring_of(float x, float y, float z, int n_scans) {
  if (n_scans != 64) return ring_of_exact(x, y, z, n_scans);
  constexpr float kAngleErr = 2e-4f;
  const float xy2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const float af = atanf(z * rsqrtf(xy2)) * 57.29577951308232f;
  bool safe = fabsf(af - (-8.83f)) > kAngleErr && fabsf(af - 2.0f) > kAngleErr && fabsf(af - (-24.33f)) > kAngleErr && isfinite(af);
  int id = 0;
  if (af >= -8.83f) {
    const float t = (2.0f - af) * 3.0f + 0.5f;
    const float fr = t - floorf(t);
    safe = safe && fr > 3.0f * kAngleErr + 1e-5f && fr < 1.0f - 3.0f * kAngleErr - 1e-5f;
    id = (int)t;
  } else {
    const float t = (-8.83f - af) * 2.0f + 0.5f;
    const float fr = t - floorf(t);
    safe = safe && fr > 2.0f * kAngleErr + 1e-5f && fr < 1.0f - 2.0f * kAngleErr - 1e-5f;
    id = 32 + (int)t;
  }
  if (!safe) return ring_of_exact(x, y, z, n_scans);
  if (af > 2.0f || af < -24.33f || id > 50 || id < 0) return -1;
  return id;
}

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------------
// sr_find_ends: grid (B), block kEndsThreads.
constexpr int kEndsThreads = 1024;
__global__ void __launch_bounds__(kEndsThreads) sr_find_ends(const float* __restrict__ xyz, int stride, size_t slab_floats,
                                                     const int* __restrict__ n_points, int cap, float min_range,
                                                     SRHeader* __restrict__ hdr) {
  const int b = blockIdx.x;
  const float* p = xyz + (size_t)b * slab_floats;
  // A count beyond the handle's capacity (or the caller's slab) is clamped and reported (kStatusCapacity): the device entry
  // point cannot validate counts that live in device memory, and every later kernel sizes its loops from h.n_in.
  const int nmax = (int)min((size_t)cap, slab_floats / (size_t)stride);
  const int nraw = n_points[b];
  const int n = min(max(nraw, 0), nmax);
  SRHeader& h = hdr[b];
  const float thres2 = __fmul_rn(min_range, min_range);
  __shared__ int s_first, s_last;
  if (threadIdx.x == 0) { s_first = 0x7fffffff; s_last = -1; }
  __syncthreads();
  // 8192 points per step (8 independent loads per thread): a scan can open / close with thousands of no-return points
  constexpr int kPer = 8;
  for (int base = 0; base < n; base += kEndsThreads * kPer) {
    int first = 0x7fffffff;
#pragma unroll
    for (int u = kPer - 1; u >= 0; --u) {
      const int i = base + u * kEndsThreads + threadIdx.x;
      if (i < n && point_valid(p[(size_t)i * stride], p[(size_t)i * stride + 1], p[(size_t)i * stride + 2], thres2)) first = i;
    }
    if (first != 0x7fffffff) atomicMin(&s_first, first);
    __syncthreads();
    const bool found = s_first != 0x7fffffff;
    __syncthreads();
    if (found) break;
  }
  for (int top = n; top > 0; top -= kEndsThreads * kPer) {
    int last = -1;
#pragma unroll
    for (int u = kPer - 1; u >= 0; --u) {
      const int i = top - 1 - u * kEndsThreads - (int)threadIdx.x;
      if (i >= 0 && point_valid(p[(size_t)i * stride], p[(size_t)i * stride + 1], p[(size_t)i * stride + 2], thres2)) last = i;
    }
    if (last >= 0) atomicMax(&s_last, last);
    __syncthreads();
    const bool found = s_last >= 0;
    __syncthreads();
    if (found) break;
  }
  // zero the per-scan counters
  for (int i = threadIdx.x; i < kMaxRings; i += kEndsThreads) { h.ringCount[i] = 0; h.ringLessFlat[i] = 0; }
  for (int i = threadIdx.x; i < kMaxRings * kSectors * 3; i += kEndsThreads) h.secCount[i] = 0;
  if (threadIdx.x == 0) {
    h.n_in = n;
    h.halfIdx = 0x7fffffff;
    h.cloudSize = 0;
    h.nSharp = h.nLessSharp = h.nFlat = h.nLessFlat = 0;
    const int capBit = nraw > nmax ? kStatusCapacity : 0;
    if (s_last < 0) {
      h.firstValid = -1; h.lastValid = -1; h.startOri = 0.f; h.endOri = 0.f;
      h.status = kStatusEmpty | capBit;
    } else {
      h.firstValid = s_first; h.lastValid = s_last;
      const float x0 = p[(size_t)s_first * stride], y0 = p[(size_t)s_first * stride + 1];
      const float x1 = p[(size_t)s_last * stride], y1 = p[(size_t)s_last * stride + 1];
      // :166-176
      float startOri = -vb_fdlibm::atan2f_fd(y0, x0);
      float endOri = (float)((double)(-vb_fdlibm::atan2f_fd(y1, x1)) + 2 * kPi);
      if ((double)(__fsub_rn(endOri, startOri)) > 3 * kPi) {
        endOri = (float)((double)endOri - 2 * kPi);
      } else if ((double)(__fsub_rn(endOri, startOri)) < kPi) {
        endOri = (float)((double)endOri + 2 * kPi);
      }
      h.startOri = startOri; h.endOri = endOri;
      h.status = capBit;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// sr_classify: grid (nblk, B), block 256, 4 points per thread (strided for coalescing).
__global__ void __launch_bounds__(256) sr_classify(const float* __restrict__ xyz, int stride, size_t slab_floats,
                                                    float min_range, int n_scans, SRHeader* __restrict__ hdr,
                                                    uint8_t* __restrict__ ring8, int cap, int* __restrict__ blockHist,
                                                    int nblk) {
  const int b = blockIdx.y, blk = blockIdx.x;
  SRHeader& h = hdr[b];
  const int n = h.n_in;
  const float* p = xyz + (size_t)b * slab_floats;
  const float thres2 = __fmul_rn(min_range, min_range);
  const float startOri = h.startOri;
  __shared__ int hist[kMaxRings];
  __shared__ int s_half;
  if (threadIdx.x < kMaxRings) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_half = 0x7fffffff;
  __syncthreads();
  // Stage the tile's coordinates in shared memory with 16-byte loads (packed xyz, the common case), so the global
  // loads are wide, coalesced and all in flight at once; the per-point reads below are stride-3 words: conflict free.
  __shared__ __align__(16) float tile[kClassifyBlock * 3];
  {
    const int base = blk * kClassifyBlock;
    const float* src = p + (size_t)base * stride;
    if (stride == 3 && base + kClassifyBlock <= n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(tile);
#pragma unroll
      for (int k = 0; k < 3; ++k) d4[k * 256 + threadIdx.x] = s4[k * 256 + threadIdx.x];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = j * 256 + threadIdx.x;
        if (base + k < n) {
          const float* q = src + (size_t)k * stride;
          tile[3 * k] = q[0]; tile[3 * k + 1] = q[1]; tile[3 * k + 2] = q[2];
        }
      }
    }
  }
  __syncthreads();
  int myHalf = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = j * 256 + threadIdx.x;
    const int i = blk * kClassifyBlock + k;
    int ring = -1;
    if (i < n) {
      const float x = tile[3 * k], y = tile[3 * k + 1], z = tile[3 * k + 2];
      if (point_valid(x, y, z, thres2)) ring = ring_of(x, y, z, n_scans);
      if (ring >= 0) {
        // :234-250, the not-yet-halfPassed branch: is this the point that flips halfPassed?
        float ori = -vb_fdlibm::atan2f_fd(y, x);
        if ((double)ori < (double)startOri - kPi / 2) ori = (float)((double)ori + 2 * kPi);
        else if ((double)ori > (double)startOri + kPi * 3 / 2) ori = (float)((double)ori - 2 * kPi);
        if ((double)__fsub_rn(ori, startOri) > kPi) myHalf = min(myHalf, i);
      }
      ring8[(size_t)b * cap + i] = (uint8_t)(ring < 0 ? 255 : ring);
    }
    // warp-aggregated histogram update
    const unsigned m = __match_any_sync(0xffffffffu, ring);
    if (ring >= 0 && (int)lane_id() == __ffs(m) - 1) atomicAdd(&hist[ring], __popc(m));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) myHalf = min(myHalf, __shfl_xor_sync(0xffffffffu, myHalf, o));
  if (lane_id() == 0 && myHalf != 0x7fffffff) atomicMin(&s_half, myHalf);
  __syncthreads();
  if (threadIdx.x < kMaxRings) blockHist[((size_t)b * nblk + blk) * kMaxRings + threadIdx.x] = hist[threadIdx.x];
  if (threadIdx.x == 0 && s_half != 0x7fffffff) atomicMin(&h.halfIdx, s_half);
}

// ---------------------------------------------------------------------------------------------
// sr_scan: grid (B), block 64.  blockHist[b][blk][r] -> exclusive offset of (ring r, block blk) in the output cloud.
__global__ void __launch_bounds__(64) sr_scan(SRHeader* __restrict__ hdr, int* __restrict__ blockHist, int nblk) {
  const int b = blockIdx.x, r = threadIdx.x;
  SRHeader& h = hdr[b];
  int* bh = blockHist + (size_t)b * nblk * kMaxRings;
  const int used = (h.n_in + kClassifyBlock - 1) / kClassifyBlock;
  int tot = 0;
  for (int k = 0; k < used; ++k) tot += bh[k * kMaxRings + r];
  __shared__ int s_cnt[kMaxRings], s_start[kMaxRings + 1];
  s_cnt[r] = tot;
  __syncthreads();
  if (r == 0) {
    int acc = 0;
    for (int i = 0; i < kMaxRings; ++i) { s_start[i] = acc; acc += s_cnt[i]; }
    s_start[kMaxRings] = acc;
    h.cloudSize = acc;
    h.ringStart[kMaxRings] = acc;
    if (acc == 0) atomicOr(&h.status, kStatusEmpty);
  }
  __syncthreads();
  h.ringCount[r] = tot;
  h.ringStart[r] = s_start[r];
  if (tot > kRingCap) atomicOr(&h.status, kStatusRingOverflow);
  int acc = s_start[r];
  for (int k = 0; k < used; ++k) {
    const int t = bh[k * kMaxRings + r];
    bh[k * kMaxRings + r] = acc;
    acc += t;
  }
}

// ---------------------------------------------------------------------------------------------
// sr_scatter: grid (nblk, B), block 1024 (one point per thread, index order == thread order).
__global__ void __launch_bounds__(1024) sr_scatter(const float* __restrict__ xyz, int stride, size_t slab_floats,
                                                    const SRHeader* __restrict__ hdr, const uint8_t* __restrict__ ring8,
                                                    int cap, const int* __restrict__ blockOff, int nblk,
                                                    float4* __restrict__ cloud) {
  const int b = blockIdx.y, blk = blockIdx.x;
  const SRHeader& h = hdr[b];
  const int n = h.n_in;
  const int i = blk * kClassifyBlock + threadIdx.x;
  __shared__ int wh[kMaxRings][32];  // [ring][warp]: lanes of the scan below read consecutive words
  for (int k = threadIdx.x; k < 32 * kMaxRings; k += 1024) (&wh[0][0])[k] = 0;
  __syncthreads();
  int ring = -1;
  if (i < n) { const int r8 = ring8[(size_t)b * cap + i]; ring = r8 == 255 ? -1 : r8; }
  const unsigned m = __match_any_sync(0xffffffffu, ring);
  const int warp = threadIdx.x >> 5;
  const int rank = __popc(m & ((1u << lane_id()) - 1u));
  if (ring >= 0 && rank == 0) wh[ring][warp] = __popc(m);
  __syncthreads();
  // exclusive scan over the 32 warps for each ring: warp w handles rings 2w, 2w+1; lane = source warp.
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = warp * 2 + rr;
    const int v = wh[r][lane_id()];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((int)lane_id() >= o) s += t; }
    wh[r][lane_id()] = s - v + blockOff[((size_t)b * nblk + blk) * kMaxRings + r];
  }
  __syncthreads();
  // (computing the point's own part before the ranking barriers was measured slower: 179 vs 164 us at 128 streams)
  if (ring >= 0) {
    const float* p = xyz + (size_t)b * slab_floats + (size_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    const float startOri = h.startOri, endOri = h.endOri;
    float ori = -vb_fdlibm::atan2f_fd(y, x);
    if (i <= h.halfIdx) {  // :235-250
      if ((double)ori < (double)startOri - kPi / 2) ori = (float)((double)ori + 2 * kPi);
      else if ((double)ori > (double)startOri + kPi * 3 / 2) ori = (float)((double)ori - 2 * kPi);
    } else {  // :251-262
      ori = (float)((double)ori + 2 * kPi);
      if ((double)ori < (double)endOri - kPi * 3 / 2) ori = (float)((double)ori + 2 * kPi);
      else if ((double)ori > (double)endOri + kPi / 2) ori = (float)((double)ori - 2 * kPi);
    }
    const float relTime = __fdiv_rn(__fsub_rn(ori, startOri), __fsub_rn(endOri, startOri));
    const float intensity = (float)__dadd_rn((double)ring, __dmul_rn(0.1, (double)relTime));  // :265
    cloud[(size_t)b * cap + wh[ring][warp] + rank] = make_float4(x, y, z, intensity);
  }
}

// ---------------------------------------------------------------------------------------------
// sr_curvature: grid (ceil(tiles / tilesPerCta), B), block 256; a CTA walks `tilesPerCta` consecutive 1024-point tiles
// through a two-stage pipeline: tile k+1 streams into shared memory while tile k is computed, so loads stay in flight for
// the whole life of the CTA.  Every thread produces 4 consecutive curvatures from 14 points held in registers; rows of 8
// points sit at a 144-byte stride in shared memory (one float4 of padding every 8) so that stride-4 register fill is
// bank-conflict free.  21 B of HBM traffic per point (16 read + 4 + 1 written).
//
// Two ways of moving a tile, same arithmetic (template parameter TMA):
//   false (default)  per-thread 16-byte cp.async (LDGSTS), 4 per thread and tile, no bounds tests on interior tiles;
//   true             the TMA unit: one cp.async.bulk per 128-byte row (the padding forbids longer runs), 130 per tile, issued
//                    by one warp in a uniform loop, completion counted in bytes on an mbarrier per stage.
// Measured on B200, 64 streams per launch: LDGSTS 36.1 us (3.7 TB/s, 57 % of the measured HBM peak), TMA rows 42.4 us —
// 130 small bulk copies per 16 KB tile cost more than they save, so LDGSTS stays the default (VLOAM_SR_CURV_TMA=1 selects
// the other).
__device__ __forceinline__ int curv_pad(int e) { return e + (e >> 3); }
constexpr int kCurvTile = 1024;
constexpr int kCurvElems = kCurvTile + 10, kCurvRows = (kCurvElems + 7) / 8;     // 1034 points with the halo = 129 rows + 2 points
constexpr int kCurvTileSmem = kCurvTile + 10 + (kCurvTile + 10) / 8 + 1;
template <bool TMA>
__device__ __forceinline__ void sr_curvature_body(const SRHeader* __restrict__ hdr, const float4* __restrict__ cloud,
                                                  int cap, float* __restrict__ curv, uint8_t* __restrict__ gapflag,
                                                  int tilesPerCta) {
  const int b = blockIdx.y;
  const int size = hdr[b].cloudSize;
  const int ntiles = (size + kCurvTile - 1) / kCurvTile;
  const int first = blockIdx.x * tilesPerCta, last = min(first + tilesPerCta, ntiles);
  if (first >= last) return;
  const float4* c = cloud + (size_t)b * cap;
  __shared__ __align__(16) float4 tile[2][kCurvTileSmem];
  __shared__ __align__(8) uint64_t mbar[2];
  if (TMA) {
    if (threadIdx.x == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    __syncthreads();
  }
  auto issue = [&](int tileIdx, int buf) {   // tile[buf][pad(e)] <- cloud[tileIdx * 1024 - 5 + e], zero outside the cloud
    const int base = tileIdx * kCurvTile - 5;
    const int lo = max(0, -base), hi = min(kCurvElems, size - base);      // tile elements [lo, hi) exist
    if (TMA) {
      if (threadIdx.x == 0) {
        // bytes the bulk copies will deliver: the rows that lie entirely inside [lo, hi)
        const int rlo = (lo + 7) >> 3, rfull = hi >> 3;                   // full rows: rlo <= r < min(rfull, 129)
        int pts = max(min(rfull, kCurvRows - 1) - rlo, 0) * 8;
        if (hi == kCurvElems && lo <= 8 * (kCurvRows - 1)) pts += kCurvElems - 8 * (kCurvRows - 1);   // the 2-point tail row
        mbar_arrive_expect_tx(&mbar[buf], (unsigned)pts * 16u);
      }
      if (threadIdx.x < 32) {       // one warp issues the copies: the instruction is warp-uniform (UBLKCP takes uniform registers)
        fence_proxy_async();        // this buffer was last read with ordinary loads (two iterations ago)
        for (int r = 0; r < kCurvRows; ++r) {
          const int e0 = 8 * r, e1 = min(e0 + 8, kCurvElems);
          if (e0 >= lo && e1 <= hi) {
            if (threadIdx.x == 0) bulk_g2s(&tile[buf][curv_pad(e0)], c + base + e0, (unsigned)(e1 - e0) * 16u, &mbar[buf]);
          } else if ((int)threadIdx.x < e1 - e0) {   // a row that sticks out of the cloud (first / last tile only)
            const int e = e0 + threadIdx.x;
            tile[buf][curv_pad(e)] = (e >= lo && e < hi) ? c[base + e] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    } else {
      if (lo == 0 && hi == kCurvElems) {      // interior tile: no bounds tests
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int e = k * 256 + threadIdx.x; __pipeline_memcpy_async(&tile[buf][curv_pad(e)], c + base + e, sizeof(float4)); }
        if (threadIdx.x < 10) { const int e = kCurvTile + threadIdx.x; __pipeline_memcpy_async(&tile[buf][curv_pad(e)], c + base + e, sizeof(float4)); }
      } else {
        for (int e = threadIdx.x; e < kCurvElems; e += 256) {
          if (e >= lo && e < hi) __pipeline_memcpy_async(&tile[buf][curv_pad(e)], c + base + e, sizeof(float4));
          else tile[buf][curv_pad(e)] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      __pipeline_commit();
    }
  };
  unsigned phase0 = 0u, phase1 = 0u;
  issue(first, 0);
  for (int tix = first; tix < last; ++tix) {
    const int buf = (tix - first) & 1;
    if (tix + 1 < last) issue(tix + 1, buf ^ 1);
    if (TMA) {
      if (buf == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&mbar[1], phase1); phase1 ^= 1u; }
    } else {
      if (tix + 1 < last) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int i0 = tix * kCurvTile + 4 * threadIdx.x;
    if (i0 < size) {
      float4 t[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) t[k] = tile[buf][curv_pad(4 * threadIdx.x + k)];  // t[k] == cloud[i0 - 5 + k]
      float out[4];
      unsigned gaps = 0;  // byte o = [ |p[i+1] - p[i]|^2 > 0.05 ], the neighbour-suppression break test (:355-358, :367-370)
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        {
          const float gx = __fsub_rn(t[o + 6].x, t[o + 5].x), gy = __fsub_rn(t[o + 6].y, t[o + 5].y), gz = __fsub_rn(t[o + 6].z, t[o + 5].z);
          const float g2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
          if ((double)g2 > 0.05 || i0 + o + 1 >= size) gaps |= 1u << (8 * o);
        }
        // :290-301: p[i-5] + p[i-4] + p[i-3] + p[i-2] + p[i-1] - 10*p[i] + p[i+1] + ... + p[i+5], left to right
        float dx = __fadd_rn(t[o].x, t[o + 1].x), dy = __fadd_rn(t[o].y, t[o + 1].y), dz = __fadd_rn(t[o].z, t[o + 1].z);
#pragma unroll
        for (int k = 2; k <= 4; ++k) { dx = __fadd_rn(dx, t[o + k].x); dy = __fadd_rn(dy, t[o + k].y); dz = __fadd_rn(dz, t[o + k].z); }
        dx = __fsub_rn(dx, __fmul_rn(10.f, t[o + 5].x)); dy = __fsub_rn(dy, __fmul_rn(10.f, t[o + 5].y)); dz = __fsub_rn(dz, __fmul_rn(10.f, t[o + 5].z));
#pragma unroll
        for (int k = 6; k <= 10; ++k) { dx = __fadd_rn(dx, t[o + k].x); dy = __fadd_rn(dy, t[o + k].y); dz = __fadd_rn(dz, t[o + k].z); }
        const float v = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // :303
        const int i = i0 + o;
        out[o] = (i >= 5 && i < size - 5) ? v : 0.f;
      }
      float* dst = curv + (size_t)b * cap + i0;
      *reinterpret_cast<unsigned*>(gapflag + (size_t)b * cap + i0) = gaps;  // cap is a multiple of 1024: always in bounds
      if (i0 + 3 < size) {
        *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
      } else {
        for (int o = 0; o < 4 && i0 + o < size; ++o) dst[o] = out[o];
      }
    }
    __syncthreads();   // everyone is done with tile[buf] before the next iteration refills it
  }
}
__global__ void __launch_bounds__(256) sr_curvature(const SRHeader* __restrict__ hdr, const float4* __restrict__ cloud, int cap,
                                                     float* __restrict__ curv, uint8_t* __restrict__ gapflag, int tilesPerCta) {
  sr_curvature_body<false>(hdr, cloud, cap, curv, gapflag, tilesPerCta);
}
__global__ void __launch_bounds__(256) sr_curvature_tma(const SRHeader* __restrict__ hdr, const float4* __restrict__ cloud, int cap,
                                                         float* __restrict__ curv, uint8_t* __restrict__ gapflag, int tilesPerCta) {
  sr_curvature_body<true>(hdr, cloud, cap, curv, gapflag, tilesPerCta);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory bitonic sort of n (power of two) 64-bit keys, ascending, by the whole block.
__device__ void bitonic_sort_u64(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a = keys[i], bq = keys[ixj];
        if ((a > bq) == up) { keys[i] = bq; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
}
// Same, for `nseg` independent segments of length n laid out back to back.
__device__ void bitonic_sort_u64_segments(unsigned long long* keys, int n, int nseg) {
  const int half = n >> 1;
  const int shift = 31 - __clz(half);  // half is a power of two
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int w = threadIdx.x; w < half * nseg; w += blockDim.x) {
        const int seg = w >> shift, t = w & (half - 1);
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool up = (i & k) == 0;
        unsigned long long* ks = keys + (size_t)seg * n;
        const unsigned long long a = ks[i], bq = ks[ixj];
        if ((a > bq) == up) { ks[i] = bq; ks[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// sr_pick_features: one WARP per ring, grid (kMaxRings / 4, B), block 128.
// Restates the sort + greedy walk of scan_registration.cpp:319-422 without sorting: the walk over the sorted sector
// picks, each time, the largest (sharp) or smallest (flat) still-eligible (curvature, index) key, and eligibility only
// ever shrinks, so repeated warp-wide arg-max / arg-min over the sector yields the same picks in the same order.
// A lane keeps its share of the sector's curvatures in registers and an "alive" bit mask; neighbour suppression uses
// the per-pair gap flags written by sr_curvature, so the cloud itself is never read here.
template <int CAP>
struct PickSmem {
  uint8_t picked[4][CAP];
  uint8_t gap[4][CAP];
};

template <int ITEMS>
__device__ void pick_ring(const float* __restrict__ cv, uint8_t* picked, const uint8_t* gap, int rs, int SI, int EI,
                          int8_t* __restrict__ lab, int* __restrict__ fidx, int* __restrict__ secCount) {
  const int l = lane_id();
  for (int j = 0; j < kSectors; ++j) {
    const int sp = SI + (EI - SI) * j / 6;
    const int ep = SI + (EI - SI) * (j + 1) / 6 - 1;
    float c[ITEMS];
    unsigned alive = 0;
#pragma unroll
    for (int m = 0; m < ITEMS; ++m) {
      const int pos = sp + l + 32 * m;
      c[m] = 0.f;
      if (pos <= ep) { c[m] = cv[pos]; if (!picked[pos - rs]) alive |= 1u << m; }
    }
    __syncwarp();   // the reads of `picked` above are ordered before the writes of suppress() below (racecheck: write-after-read)
    // mark `ind` and its suppressed neighbours picked (:351-376 / :396-420); every lane updates its alive mask
    auto suppress = [&](int ind) {
      // lanes 0..4: forward steps 1..5 = pairs (ind+l, ind+l+1); lanes 5..9: backward steps = pairs (ind-m-1, ind-m)
      bool stop = false;
      if (l < 5) stop = gap[ind + l - rs] != 0;
      else if (l < 10) stop = gap[ind - (l - 5) - 1 - rs] != 0;
      const unsigned sb = __ballot_sync(0xffffffffu, stop);
      const int nf = min(5, (int)__ffs((sb & 0x1fu) | 0x20u) - 1);
      const int nb = min(5, (int)__ffs(((sb >> 5) & 0x1fu) | 0x20u) - 1);
      const int r0 = ind - nb, r1 = ind + nf;
      if (l <= r1 - r0) picked[r0 + l - rs] = 1;
      // at most one of this lane's positions (32 apart) lies in [r0, r1]
      const int pos0 = r0 + ((l - (r0 - sp)) & 31);
      if (pos0 <= r1 && pos0 >= sp && pos0 <= ep) alive &= ~(1u << ((pos0 - sp - l) >> 5));
      __syncwarp();
    };
    int nSharp = 0, nLess = 0, nFlat = 0;
    // sharp / less sharp (:327-378)
    for (int pickNo = 1; pickNo <= 21; ++pickNo) {
      // arg-max of (curvature bits, index) over the eligible elements: two hardware warp reductions
      unsigned bc = 0, bi = 0;
#pragma unroll
      for (int m = 0; m < ITEMS; ++m)
        if (((alive >> m) & 1u) && (double)c[m] > 0.1) {
          const unsigned cb = __float_as_uint(c[m]);  // curvature >= 0: the bit pattern orders like the value
          if (cb >= bc) { bc = cb; bi = (unsigned)(sp + l + 32 * m); }  // m ascending -> larger index wins ties
        }
      const unsigned mc = __reduce_max_sync(0xffffffffu, bc);
      if (mc == 0u || pickNo > 20) break;  // nothing eligible (eligible curvatures are > 0.1), or the 21st candidate (:346-349)
      const int ind = (int)__reduce_max_sync(0xffffffffu, bc == mc ? bi : 0u);
      if (l == 0) {
        if (pickNo <= 2) { lab[ind] = 2; fidx[j * 26 + nSharp] = ind; } else lab[ind] = 1;
        fidx[j * 26 + 2 + nLess] = ind;
      }
      if (pickNo <= 2) ++nSharp;
      ++nLess;
      suppress(ind);
    }
    // flat (:380-422)
    for (int pickNo = 1; pickNo <= 4; ++pickNo) {
      unsigned bc = 0xffffffffu, bi = 0xffffffffu;
#pragma unroll
      for (int m = 0; m < ITEMS; ++m)
        if (((alive >> m) & 1u) && (double)c[m] < 0.1) {
          const unsigned cb = __float_as_uint(c[m]);
          if (cb < bc) { bc = cb; bi = (unsigned)(sp + l + 32 * m); }  // m ascending -> smaller index wins ties
        }
      const unsigned mc = __reduce_min_sync(0xffffffffu, bc);
      if (mc == 0xffffffffu) break;
      const int ind = (int)__reduce_min_sync(0xffffffffu, bc == mc ? bi : 0xffffffffu);
      if (l == 0) { lab[ind] = -1; fidx[j * 26 + 22 + nFlat] = ind; }
      ++nFlat;
      if (pickNo >= 4) break;  // :390-394: the 4th flat point is not suppressed
      suppress(ind);
    }
    if (l == 0) { secCount[j * 3 + 0] = nSharp; secCount[j * 3 + 1] = nLess; secCount[j * 3 + 2] = nFlat; }
  }
}

// CAP = 2048 handles rings of up to 2048 points (every HDL-64 ring at 10 Hz) at twice the occupancy; the CAP = 4096
// instance only picks up the longer rings.
template <int CAP>
__global__ void __launch_bounds__(128, 7) sr_pick_features(SRHeader* __restrict__ hdr, const float* __restrict__ curv,
                                                         const uint8_t* __restrict__ gapflag, int cap,
                                                         int8_t* __restrict__ label_out, int* __restrict__ featIdx) {
  __shared__ PickSmem<CAP> S;
  const int b = blockIdx.y, w = threadIdx.x >> 5, ring = blockIdx.x * 4 + w, l = lane_id();
  SRHeader& h = hdr[b];
  const int rs = h.ringStart[ring], re = h.ringStart[ring + 1];
  const int len = re - rs;
  const int SI = rs + 5, EI = re - 6;  // scanStartInd / scanEndInd (:278-280)
  int8_t* lab = label_out + (size_t)b * cap;
  if (CAP == 2048 ? len > 2048 : (len <= 2048 || len > kRingCap)) {
    if (CAP != 2048 && len > kRingCap) {
      for (int i = l; i < len; i += 32) lab[rs + i] = 0;
      if (l == 0) atomicOr(&h.status, kStatusRingOverflow);
    }
    return;
  }
  for (int i = l; i < len; i += 32) lab[rs + i] = 0;
  if (EI - SI < 6) return;  // :314
  const int maxn = (EI - SI + 5) / 6 + 1;
  if (maxn > kSectorCap) { if (l == 0) atomicOr(&h.status, kStatusRingOverflow); return; }
  const uint8_t* gf = gapflag + (size_t)b * cap + rs;
  for (int i = l; i < len; i += 32) { S.picked[w][i] = 0; S.gap[w][i] = gf[i]; }
  __syncwarp();
  const float* cv = curv + (size_t)b * cap;
  int* fidx = featIdx + ((size_t)b * kMaxRings + ring) * kSectors * 26;
  int* sc = h.secCount + ring * kSectors * 3;
  if constexpr (CAP == 2048) {  // len <= 2048 -> sectors of at most 341 points
    pick_ring<12>(cv, S.picked[w], S.gap[w], rs, SI, EI, lab, fidx, sc);
  } else {
    if (maxn <= 512) pick_ring<16>(cv, S.picked[w], S.gap[w], rs, SI, EI, lab, fidx, sc);
    else pick_ring<32>(cv, S.picked[w], S.gap[w], rs, SI, EI, lab, fidx, sc);
  }
}

// ---------------------------------------------------------------------------------------------
// sr_less_flat_voxel: grid (kMaxRings, B), block 256, dynamic shared memory = sizeof(VoxelSmem).
// Per ring: gather the less-flat candidates (label <= 0 inside [scanStartInd, scanEndInd), :424-430) and run
// pcl::VoxelGrid(0.2) on them (:433-437; semantics restated in oracle/voxel_grid.hpp).  The key sort is a stable
// LSD radix sort (8-bit digits, warp match_any ranking) of the 32-bit voxel keys: stability keeps points of one
// voxel in input order, which fixes the float summation order of the centroid.
template <int CAP>
struct VoxelSmem {
  unsigned key[2][CAP];
  unsigned short pos[2][CAP];
  int lf[CAP];           // cloud indices of the less-flat candidates (ring order)
  unsigned short runStart[CAP + 2];   // first candidate of every run of equal voxel keys, in ring order; [runs] = m
  int off[8][256];       // per-warp digit offsets
  int scan[256 + 1];
  float red[6 * 8];
};

__device__ int block_exclusive_scan(int v, int* scan /*[257]*/) {
  // 256 threads; returns the exclusive prefix of v, scan[256] = total
  const int w = threadIdx.x >> 5, l = lane_id();
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += t; }
  __shared__ int wsum[8];
  if (l == 31) wsum[w] = s;
  __syncthreads();
  int off = 0;
  for (int k = 0; k < w; ++k) off += wsum[k];
  if (threadIdx.x == 255) scan[256] = off + s;
  __syncthreads();
  return off + s - v;
}

// Stable radix sort of S.key[0][0..m) / S.pos[0][0..m) by the low `bits` bits; returns the buffer index holding the result.
template <int CAP>
__device__ int voxel_radix_sort(VoxelSmem<CAP>& S, int m, int bits) {
  const int w = threadIdx.x >> 5, l = lane_id();
  const int per = (m + 7) / 8;                 // contiguous elements owned by each warp
  const int w0 = min(w * per, m), w1 = min(w0 + per, m);
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    const unsigned* kin = S.key[cur];
    const unsigned short* pin = S.pos[cur];
    unsigned* kout = S.key[cur ^ 1];
    unsigned short* pout = S.pos[cur ^ 1];
    for (int d = l; d < 256; d += 32) S.off[w][d] = 0;
    __syncwarp();
    // (A) per-warp digit totals (order is irrelevant here: shared-memory atomics)
    for (int k = w0 + l; k < w1; k += 32) atomicAdd(&S.off[w][(kin[k] >> shift) & 255], 1);
    __syncthreads();
    // (B) exclusive scan in (digit, warp) order: thread t owns digit t
    {
      const int d = threadIdx.x;
      int c[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = S.off[q][d]; tot += c[q]; }
      int run = block_exclusive_scan(tot, S.scan);
#pragma unroll
      for (int q = 0; q < 8; ++q) { S.off[q][d] = run; run += c[q]; }
    }
    __syncthreads();
    // (C) stable scatter
    for (int base = w0; base < w1; base += 32) {
      const int k = base + l;
      const bool act = k < w1;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const unsigned key = kin[k];
        const int d = (key >> shift) & 255;
        const unsigned peers = __match_any_sync(amask, d);
        const int rank = __popc(peers & ((1u << l) - 1u));
        const int dst = S.off[w][d] + rank;
        kout[dst] = key;
        pout[dst] = pin[k];
        __syncwarp(amask);
        if (l == __ffs(peers) - 1) S.off[w][d] += __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

template <int CAP>
__device__ __forceinline__ void voxel_ring(VoxelSmem<CAP>& S, SRHeader* __restrict__ hdr, const float4* __restrict__ cloud, int cap,
                                           const int8_t* __restrict__ label, float4* __restrict__ lessFlatStage, int b, int ring) {
  SRHeader& h = hdr[b];
  const float4* c = cloud + (size_t)b * cap;
  const int8_t* lab = label + (size_t)b * cap;
  const int rs = h.ringStart[ring], re = h.ringStart[ring + 1];
  const int len = re - rs;
  const int SI = rs + 5, EI = re - 6;
  if (CAP == 2048 ? len > 2048 : (len <= 2048 || len > kRingCap)) return;  // the other instance's ring (or overflow)
  if (EI - SI < 6 || (EI - SI + 5) / 6 + 1 > kSectorCap) return;             // ringLessFlat stays 0
  // ---- less-flat candidates = positions [SI, EI) with label <= 0, in order (:424-430).
  // Thread t owns a contiguous run of positions, so one block scan yields the order-preserving compaction.
  const int span = EI - SI;
  int m;
  {
    const int per = (span + 255) / 256;
    const int k0 = min((int)threadIdx.x * per, span), k1 = min(k0 + per, span);
    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += lab[SI + k] <= 0 ? 1 : 0;
    int pos = block_exclusive_scan(cnt, S.scan);
    for (int k = k0; k < k1; ++k) if (lab[SI + k] <= 0) S.lf[pos++] = SI + k;
    m = S.scan[256];
    __syncthreads();
  }
  if (m == 0) return;
  // ---- pcl::VoxelGrid, leaf 0.2
  const float inv = __fdiv_rn(1.0f, 0.2f);
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  for (int k = threadIdx.x; k < m; k += 256) {
    const float4 p = c[S.lf[k]];
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  if (lane_id() == 0) {
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) { S.red[a * 8 + w] = mn[a]; S.red[(3 + a) * 8 + w] = mx[a]; }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = S.red[a * 8], hi = S.red[(3 + a) * 8];
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, S.red[a * 8 + w]); hi = fmaxf(hi, S.red[(3 + a) * 8 + w]); }
    mn[a] = lo; mx[a] = hi;
  }
  const long long dx = (long long)(__fmul_rn(__fsub_rn(mx[0], mn[0]), inv)) + 1;
  const long long dy = (long long)(__fmul_rn(__fsub_rn(mx[1], mn[1]), inv)) + 1;
  const long long dz = (long long)(__fmul_rn(__fsub_rn(mx[2], mn[2]), inv)) + 1;
  float4* stage = lessFlatStage + (size_t)b * cap + rs;
  if (dx * dy * dz > 2147483647LL) {  // PCL: "Leaf size is too small": output = input
    for (int k = threadIdx.x; k < m; k += 256) stage[k] = c[S.lf[k]];
    if (threadIdx.x == 0) { h.ringLessFlat[ring] = m; atomicOr(&h.status, kStatusVoxelOverflow); }
    return;
  }
  int minb[3], divb[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    minb[a] = (int)floorf(__fmul_rn(mn[a], inv));
    divb[a] = (int)floorf(__fmul_rn(mx[a], inv)) - minb[a] + 1;
  }
  const int mul1 = divb[0], mul2 = divb[0] * divb[1];
  const unsigned maxKey = (unsigned)((long long)divb[0] * divb[1] * divb[2] - 1);  // divb product <= dx*dy*dz + slack; see below
  for (int k = threadIdx.x; k < m; k += 256) {
    const float4 p = c[S.lf[k]];
    const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)minb[0]);
    const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)minb[1]);
    const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)minb[2]);
    S.key[1][k] = (unsigned)(i0 + i1 * mul1 + i2 * mul2);     // per-point keys (scratch: the sort's second buffer)
  }
  __syncthreads();
  // ---- runs of equal keys along the ring.  Consecutive points of a ring mostly stay in one 0.2 m voxel for a few
  // steps and a voxel is rarely entered twice, so sorting RUNS (about a third as many as points) is enough: a voxel's runs
  // end up side by side in ring order (stable sort) and the ordered float sum is chained through them point by point —
  // the same additions in the same order as over the individually sorted points.
  int nr;
  {
    const int per = (m + 255) / 256;
    const int k0 = min((int)threadIdx.x * per, m), k1 = min(k0 + per, m);
    int cnt = 0;
    for (int k = k0; k < k1; ++k) cnt += (k == 0 || S.key[1][k] != S.key[1][k - 1]) ? 1 : 0;
    int r = block_exclusive_scan(cnt, S.scan);
    for (int k = k0; k < k1; ++k)
      if (k == 0 || S.key[1][k] != S.key[1][k - 1]) { S.runStart[r] = (unsigned short)k; S.key[0][r] = S.key[1][k]; S.pos[0][r] = (unsigned short)r; ++r; }
    nr = S.scan[256];
    if (threadIdx.x == 0) S.runStart[nr] = (unsigned short)m;
    __syncthreads();
  }
  // keys are < divb[0]*divb[1]*divb[2]; if that product does not fit 32 bits fall back to all 32 key bits
  int bits = 32;
  if ((long long)divb[0] * divb[1] * divb[2] <= 0xffffffffLL) bits = maxKey ? 32 - __clz(maxKey) : 1;
  const int cur = voxel_radix_sort<CAP>(S, nr, bits);
  const unsigned* keys = S.key[cur];
  const unsigned short* rid = S.pos[cur];
  // ---- one output per voxel = per head among the sorted runs; thread t owns a contiguous range of sorted runs
  {
    const int per = (nr + 255) / 256;
    const int q0 = min((int)threadIdx.x * per, nr), q1 = min(q0 + per, nr);
    int nh = 0;
    for (int q = q0; q < q1; ++q) nh += (q == 0 || keys[q] != keys[q - 1]) ? 1 : 0;
    int opos = block_exclusive_scan(nh, S.scan);
    for (int q = q0; q < q1; ++q) {
      if (!(q == 0 || keys[q] != keys[q - 1])) continue;
      const unsigned vox = keys[q];
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      for (int qq = q; qq < nr && keys[qq] == vox; ++qq) {        // the voxel's runs, in ring order
        const int r = rid[qq];
        for (int t = S.runStart[r]; t < S.runStart[r + 1]; ++t) {  // the run's points, in ring order
          const float4 pp = c[S.lf[t]];
          sx = __fadd_rn(sx, pp.x); sy = __fadd_rn(sy, pp.y); sz = __fadd_rn(sz, pp.z); si = __fadd_rn(si, pp.w);
          ++cnt;
        }
      }
      const float nf = (float)cnt;
      stage[opos++] = make_float4(__fdiv_rn(sx, nf), __fdiv_rn(sy, nf), __fdiv_rn(sz, nf), __fdiv_rn(si, nf));
    }
  }
  const int outBase = S.scan[256];
  if (threadIdx.x == 0) h.ringLessFlat[ring] = outBase;
}

// grid (R, B): block x takes rings x, x + R, ...  The CAP = 2048 instance runs with R = kMaxRings (one ring per CTA);
// the CAP = 4096 instance, which normally finds nothing to do, with a small R.
template <int CAP>
__global__ void __launch_bounds__(256) sr_less_flat_voxel(SRHeader* __restrict__ hdr, const float4* __restrict__ cloud, int cap,
                                                           const int8_t* __restrict__ label, float4* __restrict__ lessFlatStage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VoxelSmem<CAP>& S = *reinterpret_cast<VoxelSmem<CAP>*>(smem_raw);
  for (int ring = blockIdx.x; ring < kMaxRings; ring += gridDim.x) {
    voxel_ring<CAP>(S, hdr, cloud, cap, label, lessFlatStage, blockIdx.y, ring);
    if (gridDim.x < kMaxRings) __syncthreads();   // shared memory is reused by the next ring
  }
}

// ---------------------------------------------------------------------------------------------
// sr_pack: grid (kMaxRings + kPackParts, B), block 256.
//   blockIdx.x < 64 : copy ring x's down-sampled less-flat points to their packed position
//   blockIdx.x >= 64: pack sharp / less-sharp / flat clouds (ring-major, sector-major, pick order), one slice of the
//                     (ring, sector, pick) slots per block; the first of them publishes the counts and ring offsets.
constexpr int kPackParts = 8;
__global__ void __launch_bounds__(256) sr_pack(SRHeader* __restrict__ hdr, const float4* __restrict__ cloud, int cap,
                                                const int* __restrict__ featIdx, const float4* __restrict__ lessFlatStage,
                                                float4* __restrict__ sharp, int* __restrict__ sharpIdx,
                                                float4* __restrict__ lessSharp, int* __restrict__ lessSharpIdx,
                                                float4* __restrict__ flat, int* __restrict__ flatIdx,
                                                float4* __restrict__ lessFlat) {
  const int b = blockIdx.y;
  SRHeader& h = hdr[b];
  if (blockIdx.x < kMaxRings) {
    const int ring = blockIdx.x;
    __shared__ int s_off;
    if (threadIdx.x < 32) {   // offset of this ring = sum of the earlier rings' counts (one warp, two loads per lane)
      const int l = threadIdx.x;
      int v = (l < ring ? h.ringLessFlat[l] : 0) + (l + 32 < ring ? h.ringLessFlat[l + 32] : 0);
      v = __reduce_add_sync(0xffffffffu, v);
      if (l == 0) s_off = v;
    }
    __syncthreads();
    const int off = s_off;
    const int n = h.ringLessFlat[ring];
    const float4* src = lessFlatStage + (size_t)b * cap + h.ringStart[ring];
    float4* dst = lessFlat + (size_t)b * cap + off;
    for (int k = threadIdx.x; k < n; k += 256) dst[k] = src[k];
    return;
  }
  // blocks 64 .. 64 + kPackParts - 1: the picked features, kPackParts slices of the (ring, sector, pick) slots.  Every block
  // forms the exclusive prefix of the per-sector counts itself (coalesced load + one warp scan per feature kind).
  constexpr int kSec = kMaxRings * kSectors;        // 384
  __shared__ int pre[3][kSec + 1];
  __shared__ int cnt[kSec * 3];
  const int part = blockIdx.x - kMaxRings;
  const float4* c = cloud + (size_t)b * cap;
  for (int i = threadIdx.x; i < kSec * 3; i += 256) cnt[i] = h.secCount[i];
  __syncthreads();
  if (threadIdx.x < 96) {
    const int kind = threadIdx.x >> 5, l = threadIdx.x & 31;
    constexpr int per = kSec / 32;                  // 12 sectors per lane
    int sum = 0;
#pragma unroll
    for (int q = 0; q < per; ++q) sum += cnt[(l * per + q) * 3 + kind];
    int sc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (l >= o) sc += t; }
    int run = sc - sum;
#pragma unroll
    for (int q = 0; q < per; ++q) { pre[kind][l * per + q] = run; run += cnt[(l * per + q) * 3 + kind]; }
    if (l == 31) pre[kind][kSec] = run;
  }
  __syncthreads();
  const int* fi = featIdx + (size_t)b * kMaxRings * kSectors * 26;
  constexpr int kSlots = kSec * 26, kSlice = (kSlots + kPackParts - 1) / kPackParts;
  for (int w = part * kSlice + threadIdx.x; w < min((part + 1) * kSlice, kSlots); w += 256) {
    const int s = w / 26, k = w % 26;
    int kind, kk;
    if (k < 2) { kind = 0; kk = k; } else if (k < 22) { kind = 1; kk = k - 2; } else { kind = 2; kk = k - 22; }
    if (kk < cnt[s * 3 + kind]) {
      const int ind = fi[w];
      const int dst = pre[kind][s] + kk;
      const float4 p = c[ind];
      if (kind == 0) { sharp[(size_t)b * kMaxSharp + dst] = p; sharpIdx[(size_t)b * kMaxSharp + dst] = ind; }
      else if (kind == 1) { lessSharp[(size_t)b * kMaxLessSharp + dst] = p; lessSharpIdx[(size_t)b * kMaxLessSharp + dst] = ind; }
      else { flat[(size_t)b * kMaxFlat + dst] = p; flatIdx[(size_t)b * kMaxFlat + dst] = ind; }
    }
  }
  if (part == 0 && threadIdx.x <= kMaxRings) {
    const int r = threadIdx.x;
    h.ringStartLessSharp[r] = pre[1][r * kSectors];  // r == 64 -> total
    int off = 0;
    for (int q = 0; q < r; ++q) off += h.ringLessFlat[q];
    h.ringStartLessFlat[r] = off;
    if (r == kMaxRings) {
      h.nSharp = pre[0][kMaxRings * kSectors];
      h.nLessSharp = pre[1][kMaxRings * kSectors];
      h.nFlat = pre[2][kMaxRings * kSectors];
      h.nLessFlat = off;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Opt-in shared-memory sizes are per-device function attributes: set (and checked) once per context, on its device.
cudaError_t sr_prepare_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sr_less_flat_voxel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoxelSmem<2048>));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(sr_less_flat_voxel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VoxelSmem<4096>));
  return e;
}

// host-side launcher (called from capi.cu)
void launch_scan_registration(Profiler* prof, cudaStream_t st, int B, int cap, const float* xyz, int stride, size_t slab_floats,
                              const int* n_points_dev, float min_range, int n_scans, SRHeader* hdr, uint8_t* ring8,
                              int* blockHist, float4* cloud, float* curv, uint8_t* gapflag, int8_t* label, int* featIdx,
                              float4* lessFlatStage, float4* sharp, int* sharpIdx, float4* lessSharp, int* lessSharpIdx,
                              float4* flat, int* flatIdx, float4* lessFlat) {
  const int nblk = (cap + kClassifyBlock - 1) / kClassifyBlock;
  VB_LAUNCH(prof, K_SR_FIND_ENDS, st, sr_find_ends<<<B, kEndsThreads, 0, st>>>(xyz, stride, slab_floats, n_points_dev, cap, min_range, hdr));
  VB_LAUNCH(prof, K_SR_CLASSIFY, st, sr_classify<<<dim3(nblk, B), 256, 0, st>>>(xyz, stride, slab_floats, min_range, n_scans, hdr, ring8, cap, blockHist, nblk));
  VB_LAUNCH(prof, K_SR_SCAN, st, sr_scan<<<B, 64, 0, st>>>(hdr, blockHist, nblk));
  VB_LAUNCH(prof, K_SR_SCATTER, st, sr_scatter<<<dim3(nblk, B), 1024, 0, st>>>(xyz, stride, slab_floats, hdr, ring8, cap, blockHist, nblk, cloud));
  {
    // tiles per CTA: long pipelines when the batch alone fills the machine, one tile per CTA for small batches
    const int tiles = (cap + kCurvTile - 1) / kCurvTile;
    static const int perEnv = [] { const char* e = getenv("VLOAM_SR_CURV_TILES"); return e ? atoi(e) : 0; }();
    const int per = perEnv > 0 ? perEnv : ((size_t)B * tiles >= 4096 ? 4 : 1);
    static const bool useTma = [] { const char* e = getenv("VLOAM_SR_CURV_TMA"); return e && e[0] == '1'; }();
    if (useTma) VB_LAUNCH(prof, K_SR_CURVATURE, st, sr_curvature_tma<<<dim3((tiles + per - 1) / per, B), 256, 0, st>>>(hdr, cloud, cap, curv, gapflag, per));
    else VB_LAUNCH(prof, K_SR_CURVATURE, st, sr_curvature<<<dim3((tiles + per - 1) / per, B), 256, 0, st>>>(hdr, cloud, cap, curv, gapflag, per));
  }
  VB_LAUNCH(prof, K_SR_PICK, st, sr_pick_features<2048><<<dim3(kMaxRings / 4, B), 128, 0, st>>>(hdr, curv, gapflag, cap, label, featIdx));
  VB_LAUNCH(prof, K_SR_VOXEL, st, sr_less_flat_voxel<2048><<<dim3(kMaxRings, B), 256, sizeof(VoxelSmem<2048>), st>>>(hdr, cloud, cap, label, lessFlatStage));
  if (cap > 2048 + 0) {  // rings longer than 2048 points (only possible when a scan has more than 2048 points at all)
    VB_LAUNCH(prof, K_SR_PICK, st, sr_pick_features<4096><<<dim3(kMaxRings / 4, B), 128, 0, st>>>(hdr, curv, gapflag, cap, label, featIdx));
    VB_LAUNCH(prof, K_SR_VOXEL, st, sr_less_flat_voxel<4096><<<dim3(4, B), 256, sizeof(VoxelSmem<4096>), st>>>(hdr, cloud, cap, label, lessFlatStage));
  }
  VB_LAUNCH(prof, K_SR_PACK, st, sr_pack<<<dim3(kMaxRings + kPackParts, B), 256, 0, st>>>(hdr, cloud, cap, featIdx, lessFlatStage, sharp, sharpIdx,
                                                                              lessSharp, lessSharpIdx, flat, flatIdx, lessFlat));
}

}  // namespace vb
