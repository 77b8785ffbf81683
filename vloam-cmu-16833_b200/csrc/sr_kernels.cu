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
//   sr_ring_features :312-439  per ring: 6 sector sorts, greedy sharp/flat picks with neighbour suppression,
//                               less-flat gather + pcl::VoxelGrid(0.2) restated in shared memory
//   sr_pack                    ring-major packing of the four feature clouds
//
// Bit-level decisions (ring id, curvature, thresholds, voxel keys) use explicit
// round-to-nearest intrinsics so nvcc cannot contract them into FMAs: the CPU
// reference is built without FMA (CMakeLists.txt:5-6) and the index sets depend on it.
#include <math_constants.h>

#include "common.cuh"
#include "internal.h"

namespace vb {

__device__ __forceinline__ bool point_valid(float x, float y, float z, float thres2) {
  // pcl::removeNaNFromPointCloud (:157) then removeClosedPointCloud (:114-117)
  if (!isfinite(x) || !isfinite(y) || !isfinite(z)) return false;
  const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  return !(r2 < thres2);
}

// :192-226.  Returns ring id or -1.  atan/sqrt evaluated in double.
__device__ __forceinline__ int ring_of(float x, float y, float z, int n_scans) {
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

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------------
// sr_find_ends: grid (B), block 256.
__global__ void __launch_bounds__(256) sr_find_ends(const float* __restrict__ xyz, int stride, size_t slab_floats,
                                                     const int* __restrict__ n_points, float min_range,
                                                     SRHeader* __restrict__ hdr) {
  const int b = blockIdx.x;
  const float* p = xyz + (size_t)b * slab_floats;
  const int n = n_points[b];
  SRHeader& h = hdr[b];
  const float thres2 = __fmul_rn(min_range, min_range);
  __shared__ int s_first, s_last;
  if (threadIdx.x == 0) { s_first = 0x7fffffff; s_last = -1; }
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    bool v = false;
    if (i < n) v = point_valid(p[(size_t)i * stride], p[(size_t)i * stride + 1], p[(size_t)i * stride + 2], thres2);
    if (v) atomicMin(&s_first, i);
    __syncthreads();
    const bool found = s_first != 0x7fffffff;
    __syncthreads();
    if (found) break;
  }
  for (int top = n; top > 0; top -= 256) {
    const int i = top - 1 - (int)threadIdx.x;
    bool v = false;
    if (i >= 0) v = point_valid(p[(size_t)i * stride], p[(size_t)i * stride + 1], p[(size_t)i * stride + 2], thres2);
    if (v) atomicMax(&s_last, i);
    __syncthreads();
    const bool found = s_last >= 0;
    __syncthreads();
    if (found) break;
  }
  // zero the per-scan counters
  for (int i = threadIdx.x; i < kMaxRings; i += 256) { h.ringCount[i] = 0; h.ringLessFlat[i] = 0; }
  for (int i = threadIdx.x; i < kMaxRings * kSectors * 3; i += 256) h.secCount[i] = 0;
  if (threadIdx.x == 0) {
    h.n_in = n;
    h.halfIdx = 0x7fffffff;
    h.cloudSize = 0;
    h.nSharp = h.nLessSharp = h.nFlat = h.nLessFlat = 0;
    if (s_last < 0) {
      h.firstValid = -1; h.lastValid = -1; h.startOri = 0.f; h.endOri = 0.f;
      h.status = kStatusEmpty;
    } else {
      h.firstValid = s_first; h.lastValid = s_last;
      const float x0 = p[(size_t)s_first * stride], y0 = p[(size_t)s_first * stride + 1];
      const float x1 = p[(size_t)s_last * stride], y1 = p[(size_t)s_last * stride + 1];
      // :166-176
      float startOri = -atan2f(y0, x0);
      float endOri = (float)((double)(-atan2f(y1, x1)) + 2 * kPi);
      if ((double)(__fsub_rn(endOri, startOri)) > 3 * kPi) {
        endOri = (float)((double)endOri - 2 * kPi);
      } else if ((double)(__fsub_rn(endOri, startOri)) < kPi) {
        endOri = (float)((double)endOri + 2 * kPi);
      }
      h.startOri = startOri; h.endOri = endOri;
      h.status = 0;
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
  int myHalf = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = blk * kClassifyBlock + j * 256 + threadIdx.x;
    int ring = -1;
    if (i < n) {
      const float x = p[(size_t)i * stride], y = p[(size_t)i * stride + 1], z = p[(size_t)i * stride + 2];
      if (point_valid(x, y, z, thres2)) ring = ring_of(x, y, z, n_scans);
      if (ring >= 0) {
        // :234-250, the not-yet-halfPassed branch: is this the point that flips halfPassed?
        float ori = -atan2f(y, x);
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
  __shared__ int wh[32][kMaxRings];
  for (int k = threadIdx.x; k < 32 * kMaxRings; k += 1024) (&wh[0][0])[k] = 0;
  __syncthreads();
  int ring = -1;
  if (i < n) { const int r8 = ring8[(size_t)b * cap + i]; ring = r8 == 255 ? -1 : r8; }
  const unsigned m = __match_any_sync(0xffffffffu, ring);
  const int warp = threadIdx.x >> 5;
  const int rank = __popc(m & ((1u << lane_id()) - 1u));
  if (ring >= 0 && rank == 0) wh[warp][ring] = __popc(m);
  __syncthreads();
  // exclusive scan over the 32 warps for each ring: warp w handles rings 2w, 2w+1; lane = source warp.
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = warp * 2 + rr;
    const int v = wh[lane_id()][r];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if ((int)lane_id() >= o) s += t; }
    wh[lane_id()][r] = s - v + blockOff[((size_t)b * nblk + blk) * kMaxRings + r];
  }
  __syncthreads();
  if (ring >= 0) {
    const float* p = xyz + (size_t)b * slab_floats + (size_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    const float startOri = h.startOri, endOri = h.endOri;
    float ori = -atan2f(y, x);
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
    cloud[(size_t)b * cap + wh[warp][ring] + rank] = make_float4(x, y, z, intensity);
  }
}

// ---------------------------------------------------------------------------------------------
// sr_curvature: grid (ceil(cap/256), B), block 256.  20 B of HBM traffic per point (16 read + 4 written).
__global__ void __launch_bounds__(256) sr_curvature(const SRHeader* __restrict__ hdr, const float4* __restrict__ cloud,
                                                     int cap, float* __restrict__ curv) {
  const int b = blockIdx.y;
  const int size = hdr[b].cloudSize;
  const int base = blockIdx.x * 256;
  if (base >= size) return;
  const float4* c = cloud + (size_t)b * cap;
  __shared__ float4 tile[256 + 10];
  const int i = base + threadIdx.x;
  {
    const int g = i - 5;
    tile[threadIdx.x] = (g >= 0 && g < size) ? c[g] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < 10) {
      const int g2 = base + 256 - 5 + threadIdx.x;
      tile[256 + threadIdx.x] = (g2 >= 0 && g2 < size) ? c[g2] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  if (i >= size) return;
  float out = 0.f;
  if (i >= 5 && i < size - 5) {
    const float4* t = &tile[threadIdx.x];  // t[k] == cloud[i - 5 + k]
    // :290-301: p[i-5] + p[i-4] + p[i-3] + p[i-2] + p[i-1] - 10*p[i] + p[i+1] + ... + p[i+5], left to right
    float dx = __fadd_rn(t[0].x, t[1].x), dy = __fadd_rn(t[0].y, t[1].y), dz = __fadd_rn(t[0].z, t[1].z);
#pragma unroll
    for (int k = 2; k <= 4; ++k) { dx = __fadd_rn(dx, t[k].x); dy = __fadd_rn(dy, t[k].y); dz = __fadd_rn(dz, t[k].z); }
    dx = __fsub_rn(dx, __fmul_rn(10.f, t[5].x)); dy = __fsub_rn(dy, __fmul_rn(10.f, t[5].y)); dz = __fsub_rn(dz, __fmul_rn(10.f, t[5].z));
#pragma unroll
    for (int k = 6; k <= 10; ++k) { dx = __fadd_rn(dx, t[k].x); dy = __fadd_rn(dy, t[k].y); dz = __fadd_rn(dz, t[k].z); }
    out = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // :303
  }
  curv[(size_t)b * cap + i] = out;
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

__device__ __forceinline__ float gap2(const float4 a, const float4 b) {
  // :355-358: diff = p[a] - p[b]; dx*dx + dy*dy + dz*dz
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// warp-cooperative neighbour suppression (:353-376 / :397-420).  `ind` is the picked index (cloud
// coordinates), `rs` the ring start; picked[] is indexed ring-relative.
__device__ __forceinline__ void mark_neighbours(const float4* __restrict__ c, int ind, int rs, uint8_t* picked) {
  const int l = lane_id();
  bool stop = false;
  if (l < 5) stop = (double)gap2(c[ind + l + 1], c[ind + l]) > 0.05;          // forward step l+1
  else if (l < 10) stop = (double)gap2(c[ind - (l - 5) - 1], c[ind - (l - 5)]) > 0.05;  // backward step l-4
  const unsigned sb = __ballot_sync(0xffffffffu, stop);
  const int nf = min(5, (int)__ffs((sb & 0x1fu) | 0x20u) - 1);          // forward steps before the first break
  const int nb = min(5, (int)__ffs(((sb >> 5) & 0x1fu) | 0x20u) - 1);   // backward steps before the first break
  if (l < nf) picked[ind + l + 1 - rs] = 1;
  else if (l >= 5 && l - 5 < nb) picked[ind - (l - 5) - 1 - rs] = 1;
  __syncwarp();
}

// sr_ring_features: grid (kMaxRings, B), block 256, dynamic shared memory (see sr_ring_smem_bytes()).
struct RingSmem {
  unsigned long long keys[kSectors * kSectorCap];  // sector sort keys; re-used for the voxel sort (kRingCap keys)
  int lf[kRingCap];                                // cloud indices of the less-flat candidates (ring order)
  uint8_t picked[kRingCap];
  int8_t label[kRingCap];
  int scan[256 + 1];
  float red[6 * 8];
  int misc[16];
};
static_assert(kSectors * kSectorCap >= kRingCap, "voxel keys alias the sector keys");
size_t sr_ring_smem_bytes() { return sizeof(RingSmem); }

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

__global__ void __launch_bounds__(256) sr_ring_features(SRHeader* __restrict__ hdr, const float4* __restrict__ cloud,
                                                         const float* __restrict__ curv, int cap,
                                                         int8_t* __restrict__ label_out, int* __restrict__ featIdx,
                                                         float4* __restrict__ lessFlatStage) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RingSmem& S = *reinterpret_cast<RingSmem*>(smem_raw);
  const int b = blockIdx.y, ring = blockIdx.x;
  SRHeader& h = hdr[b];
  const float4* c = cloud + (size_t)b * cap;
  const float* cv = curv + (size_t)b * cap;
  const int rs = h.ringStart[ring], re = h.ringStart[ring + 1];
  const int len = re - rs;
  const int SI = rs + 5, EI = re - 6;  // scanStartInd / scanEndInd (:278-280)
  int8_t* lab = label_out + (size_t)b * cap;
  int* fidx = featIdx + ((size_t)b * kMaxRings + ring) * kSectors * 26;
  if (len > kRingCap || EI - SI < 6) {  // :314
    for (int i = threadIdx.x; i < len; i += 256) lab[rs + i] = 0;
    return;
  }
  int sp[kSectors], ep[kSectors];
  int maxn = 0;
#pragma unroll
  for (int j = 0; j < kSectors; ++j) {
    sp[j] = SI + (EI - SI) * j / 6;
    ep[j] = SI + (EI - SI) * (j + 1) / 6 - 1;
    maxn = max(maxn, ep[j] - sp[j] + 1);
  }
  if (maxn > kSectorCap) {
    if (threadIdx.x == 0) atomicOr(&h.status, kStatusRingOverflow);
    for (int i = threadIdx.x; i < len; i += 256) lab[rs + i] = 0;
    return;
  }
  const int P = next_pow2(maxn);
  // ---- phase A: keys (curvature bits, index); curvature >= 0 so the float bit pattern orders as an unsigned int
#pragma unroll
  for (int j = 0; j < kSectors; ++j) {
    const int n = ep[j] - sp[j] + 1;
    for (int k = threadIdx.x; k < P; k += 256)
      S.keys[j * P + k] = k < n ? (((unsigned long long)__float_as_uint(cv[sp[j] + k]) << 32) | (unsigned)(sp[j] + k))
                                : 0xffffffffffffffffull;
  }
  for (int i = threadIdx.x; i < len; i += 256) { S.picked[i] = 0; S.label[i] = 0; }
  __syncthreads();
  // ---- phase B: six sector sorts, ascending by (curvature, index)  (:323-324)
  bitonic_sort_u64_segments(S.keys, P, kSectors);
  // ---- phase C: greedy picks, sectors in order (suppression marks leak into the next sector, :353-376)
  if (threadIdx.x < 32) {
    const int l = lane_id();
    for (int j = 0; j < kSectors; ++j) {
      const int n = ep[j] - sp[j] + 1;
      const unsigned long long* ks = S.keys + j * P;
      int nSharp = 0, nLess = 0, nFlat = 0;
      // sharp / less sharp: walk from the largest curvature (:327-378)
      int largestPickedNum = 0;
      int k = n - 1;
      while (k >= 0) {
        const int kk = k - l;
        bool elig = false, below = false;
        int ind = 0;
        if (kk >= 0) {
          const unsigned long long key = ks[kk];
          ind = (int)(unsigned)key;
          const float cvv = __uint_as_float((unsigned)(key >> 32));
          below = !((double)cvv > 0.1);
          elig = !below && S.picked[ind - rs] == 0;
        }
        const unsigned eb = __ballot_sync(0xffffffffu, elig);
        const unsigned bb = __ballot_sync(0xffffffffu, below);
        // lanes are in descending-curvature order; ignore eligibles that come after the first `below` lane
        const int firstBelow = bb ? __ffs(bb) - 1 : 32;
        const unsigned ebv = eb & (firstBelow == 32 ? 0xffffffffu : ((1u << firstBelow) - 1u));
        if (!ebv) {
          if (bb) break;
          k -= 32;
          continue;
        }
        const int f = __ffs(ebv) - 1;
        const int pind = __shfl_sync(0xffffffffu, ind, f);
        largestPickedNum++;
        if (largestPickedNum > 20) break;  // :346-349
        if (l == 0) {
          if (largestPickedNum <= 2) { S.label[pind - rs] = 2; fidx[j * 26 + nSharp] = pind; }
          else S.label[pind - rs] = 1;
          fidx[j * 26 + 2 + nLess] = pind;
          S.picked[pind - rs] = 1;
        }
        if (largestPickedNum <= 2) nSharp++;
        nLess++;
        __syncwarp();
        mark_neighbours(c, pind, rs, S.picked);
        k = k - f - 1;
      }
      // flat: walk from the smallest curvature (:380-422)
      int smallestPickedNum = 0;
      k = 0;
      while (k < n) {
        const int kk = k + l;
        bool elig = false, above = false;
        int ind = 0;
        if (kk < n) {
          const unsigned long long key = ks[kk];
          ind = (int)(unsigned)key;
          const float cvv = __uint_as_float((unsigned)(key >> 32));
          above = !((double)cvv < 0.1);
          elig = !above && S.picked[ind - rs] == 0;
        }
        const unsigned eb = __ballot_sync(0xffffffffu, elig);
        const unsigned ab = __ballot_sync(0xffffffffu, above);
        const int firstAbove = ab ? __ffs(ab) - 1 : 32;
        const unsigned ebv = eb & (firstAbove == 32 ? 0xffffffffu : ((1u << firstAbove) - 1u));
        if (!ebv) {
          if (ab) break;
          k += 32;
          continue;
        }
        const int f = __ffs(ebv) - 1;
        const int pind = __shfl_sync(0xffffffffu, ind, f);
        if (l == 0) { S.label[pind - rs] = -1; fidx[j * 26 + 22 + nFlat] = pind; }
        nFlat++;
        smallestPickedNum++;
        if (smallestPickedNum >= 4) break;  // :390-394 (before marking: SURVEY Q2)
        if (l == 0) S.picked[pind - rs] = 1;
        __syncwarp();
        mark_neighbours(c, pind, rs, S.picked);
        k = k + f + 1;
      }
      if (l == 0) {
        h.secCount[(ring * kSectors + j) * 3 + 0] = nSharp;
        h.secCount[(ring * kSectors + j) * 3 + 1] = nLess;
        h.secCount[(ring * kSectors + j) * 3 + 2] = nFlat;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < len; i += 256) lab[rs + i] = S.label[i];
  // ---- phase D: less-flat candidates = positions [SI, EI) with label <= 0, in order (:424-430)
  const int span = EI - SI;  // positions SI .. EI-1
  int m = 0;
  for (int base = 0; base < span; base += 256) {
    const int k = base + threadIdx.x;
    const int flag = (k < span && S.label[SI + k - rs] <= 0) ? 1 : 0;
    const int pos = block_exclusive_scan(flag, S.scan);
    if (flag) S.lf[m + pos] = SI + k;
    m += S.scan[256];
    __syncthreads();
  }
  // ---- phase E: pcl::VoxelGrid, leaf 0.2 (:433-437); PCL semantics restated in oracle/voxel_grid.hpp
  if (m == 0) { if (threadIdx.x == 0) h.ringLessFlat[ring] = 0; return; }
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
  const int PV = next_pow2(m);
  for (int k = threadIdx.x; k < PV; k += 256) {
    unsigned long long key = 0xffffffffffffffffull;
    if (k < m) {
      const float4 p = c[S.lf[k]];
      const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)minb[0]);
      const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)minb[1]);
      const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)minb[2]);
      const unsigned idx = (unsigned)(i0 + i1 * mul1 + i2 * mul2);
      key = ((unsigned long long)idx << 32) | (unsigned)k;
    }
    S.keys[k] = key;
  }
  __syncthreads();
  bitonic_sort_u64(S.keys, PV);
  // segment heads -> centroid of (x, y, z, intensity), summed in ascending input order, divided by float(n)
  int outBase = 0;
  for (int base = 0; base < m; base += 256) {
    const int k = base + threadIdx.x;
    int head = 0;
    if (k < m) head = (k == 0) || ((unsigned)(S.keys[k] >> 32) != (unsigned)(S.keys[k - 1] >> 32));
    const int pos = block_exclusive_scan(head, S.scan);
    if (head) {
      const unsigned vox = (unsigned)(S.keys[k] >> 32);
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      for (int q = k; q < m && (unsigned)(S.keys[q] >> 32) == vox; ++q) {
        const float4 p = c[S.lf[(unsigned)S.keys[q]]];
        sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
        ++cnt;
      }
      const float nf = (float)cnt;
      stage[outBase + pos] = make_float4(__fdiv_rn(sx, nf), __fdiv_rn(sy, nf), __fdiv_rn(sz, nf), __fdiv_rn(si, nf));
    }
    outBase += S.scan[256];
    __syncthreads();
  }
  if (threadIdx.x == 0) h.ringLessFlat[ring] = outBase;
}

// ---------------------------------------------------------------------------------------------
// sr_pack: grid (kMaxRings + 1, B), block 256.
//   blockIdx.x < 64 : copy ring x's down-sampled less-flat points to their packed position
//   blockIdx.x == 64: pack sharp / less-sharp / flat clouds (ring-major, sector-major, pick order) and
//                     publish the counts and ring offsets of the packed clouds.
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
    int off = 0;
    for (int r = 0; r < ring; ++r) off += h.ringLessFlat[r];
    const int n = h.ringLessFlat[ring];
    const float4* src = lessFlatStage + (size_t)b * cap + h.ringStart[ring];
    float4* dst = lessFlat + (size_t)b * cap + off;
    for (int k = threadIdx.x; k < n; k += 256) dst[k] = src[k];
    return;
  }
  __shared__ int pre[3][kMaxRings * kSectors + 1];
  const float4* c = cloud + (size_t)b * cap;
  if (threadIdx.x < 3) {
    int acc = 0;
    for (int s = 0; s < kMaxRings * kSectors; ++s) { pre[threadIdx.x][s] = acc; acc += h.secCount[s * 3 + threadIdx.x]; }
    pre[threadIdx.x][kMaxRings * kSectors] = acc;
  }
  __syncthreads();
  const int* fi = featIdx + (size_t)b * kMaxRings * kSectors * 26;
  for (int w = threadIdx.x; w < kMaxRings * kSectors * 26; w += 256) {
    const int s = w / 26, k = w % 26;
    int kind, kk;
    if (k < 2) { kind = 0; kk = k; } else if (k < 22) { kind = 1; kk = k - 2; } else { kind = 2; kk = k - 22; }
    if (kk < h.secCount[s * 3 + kind]) {
      const int ind = fi[w];
      const int dst = pre[kind][s] + kk;
      const float4 p = c[ind];
      if (kind == 0) { sharp[(size_t)b * kMaxSharp + dst] = p; sharpIdx[(size_t)b * kMaxSharp + dst] = ind; }
      else if (kind == 1) { lessSharp[(size_t)b * kMaxLessSharp + dst] = p; lessSharpIdx[(size_t)b * kMaxLessSharp + dst] = ind; }
      else { flat[(size_t)b * kMaxFlat + dst] = p; flatIdx[(size_t)b * kMaxFlat + dst] = ind; }
    }
  }
  if (threadIdx.x <= kMaxRings) {
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
// host-side launcher (called from capi.cu)
void launch_scan_registration(Profiler* prof, cudaStream_t st, int B, int cap, const float* xyz, int stride, size_t slab_floats,
                              const int* n_points_dev, float min_range, int n_scans, SRHeader* hdr, uint8_t* ring8,
                              int* blockHist, float4* cloud, float* curv, int8_t* label, int* featIdx,
                              float4* lessFlatStage, float4* sharp, int* sharpIdx, float4* lessSharp, int* lessSharpIdx,
                              float4* flat, int* flatIdx, float4* lessFlat) {
  const int nblk = (cap + kClassifyBlock - 1) / kClassifyBlock;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(sr_ring_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RingSmem));
    attr_set = true;
  }
  VB_LAUNCH(prof, K_SR_FIND_ENDS, st, sr_find_ends<<<B, 256, 0, st>>>(xyz, stride, slab_floats, n_points_dev, min_range, hdr));
  VB_LAUNCH(prof, K_SR_CLASSIFY, st, sr_classify<<<dim3(nblk, B), 256, 0, st>>>(xyz, stride, slab_floats, min_range, n_scans, hdr, ring8, cap, blockHist, nblk));
  VB_LAUNCH(prof, K_SR_SCAN, st, sr_scan<<<B, 64, 0, st>>>(hdr, blockHist, nblk));
  VB_LAUNCH(prof, K_SR_SCATTER, st, sr_scatter<<<dim3(nblk, B), 1024, 0, st>>>(xyz, stride, slab_floats, hdr, ring8, cap, blockHist, nblk, cloud));
  VB_LAUNCH(prof, K_SR_CURVATURE, st, sr_curvature<<<dim3((cap + 255) / 256, B), 256, 0, st>>>(hdr, cloud, cap, curv));
  VB_LAUNCH(prof, K_SR_RING_FEATURES, st, sr_ring_features<<<dim3(kMaxRings, B), 256, sizeof(RingSmem), st>>>(hdr, cloud, curv, cap, label, featIdx, lessFlatStage));
  VB_LAUNCH(prof, K_SR_PACK, st, sr_pack<<<dim3(kMaxRings + 1, B), 256, 0, st>>>(hdr, cloud, cap, featIdx, lessFlatStage, sharp, sharpIdx,
                                                                              lessSharp, lessSharpIdx, flat, flatIdx, lessFlat));
}

}  // namespace vb
