// vloam_b200 — block-wide helpers on global-memory arrays: exclusive scan and a stable LSD radix sort by one CTA of
// 1024 threads (used by the map voxel filters, the map insertion and the visual-odometry depth buckets).
#pragma once
#include "common.cuh"

namespace vb {

// ---------------------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, val) pairs living in global memory, by one CTA of 1024 threads.
// Warp w owns a contiguous range; digit offsets are kept per warp in shared memory (see sr_less_flat_voxel).
struct SortSmem {
  int off[32][256];
  int wsum[32];
  int total;
};
static __device__ int block_exclusive_scan1024(int v, SortSmem& S) {  // 1024 threads; S.total = block total
  const int w = threadIdx.x >> 5, l = lane_id();
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += t; }
  if (l == 31) S.wsum[w] = s;
  __syncthreads();
  if (w == 0) {
    const int x = S.wsum[l];
    int sx = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sx, o); if (l >= o) sx += t; }
    S.wsum[l] = sx - x;
    if (l == 31) S.total = sx;
  }
  __syncthreads();
  const int r = S.wsum[w] + s - v;
  __syncthreads();
  return r;
}
// Returns 0 if the result is in (kA, vA), 1 if in (kB, vB).
static __device__ int cta_radix_sort(unsigned* kA, unsigned* vA, unsigned* kB, unsigned* vB, int n, int bits, SortSmem& S) {
  const int w = threadIdx.x >> 5, l = lane_id();
  const int per = (n + 31) / 32;
  const int w0 = min(w * per, n), w1 = min(w0 + per, n);
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    const unsigned* kin = cur ? kB : kA;
    const unsigned* vin = cur ? vB : vA;
    unsigned* kout = cur ? kA : kB;
    unsigned* vout = cur ? vA : vB;
    for (int d = l; d < 256; d += 32) S.off[w][d] = 0;
    __syncwarp();
    // per-warp digit histogram: four independent loads in flight per lane, shared-memory atomics (order is irrelevant here)
    for (int base = w0; base < w1; base += 128) {
      unsigned kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int k = base + u * 32 + l; kk[u] = k < w1 ? kin[k] : 0u; }
#pragma unroll
      for (int u = 0; u < 4; ++u) if (base + u * 32 + l < w1) atomicAdd(&S.off[w][(kk[u] >> shift) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
      const int d = threadIdx.x;
      int tot = 0;
      for (int q = 0; q < 32; ++q) tot += S.off[q][d];
      // exclusive scan over the 256 digit totals by warps 0..7
      int s = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += t; }
      if (l == 31) S.wsum[w] = s;
      __syncwarp();
      // (threads >= 256 idle; a named barrier over the first 256 threads)
      asm volatile("bar.sync 1, 256;");
      int basew = 0;
      for (int q = 0; q < w; ++q) basew += S.wsum[q];
      int run = basew + s - tot;
      for (int q = 0; q < 32; ++q) { const int c = S.off[q][d]; S.off[q][d] = run; run += c; }
    }
    __syncthreads();
    // stable scatter: ranks inside a 32-key step by match_any; four steps' keys and values are loaded up front
    for (int base = w0; base < w1; base += 128) {
      unsigned kk[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = base + u * 32 + l;
        kk[u] = k < w1 ? kin[k] : 0u;
        vv[u] = k < w1 ? vin[k] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool act = base + u * 32 + l < w1;
        const unsigned amask = __ballot_sync(0xffffffffu, act);
        if (act) {
          const int d = (kk[u] >> shift) & 255;
          const unsigned peers = __match_any_sync(amask, d);
          const int rank = __popc(peers & ((1u << l) - 1u));
          const int dst = S.off[w][d] + rank;
          kout[dst] = kk[u];
          vout[dst] = vv[u];
          __syncwarp(amask);
          if (l == __ffs(peers) - 1) S.off[w][d] += __popc(peers);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}


}  // namespace vb
