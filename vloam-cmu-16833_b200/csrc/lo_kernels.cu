// vloam_b200 — laserOdometry on sm_100a.
//
// Replaces vloam::LaserOdometry::solveLO
// (reference src/lidar_odometry_mapping/src/laser_odometry.cpp:187-536) and the Ceres
// problem it builds from lidarFactor.hpp:14-106:
//
//   lo_associate  :266-444  TransformToStart (:149-167), exact 1-NN in the last scan's
//                           less-sharp / less-flat cloud, ring-window 2nd / 3rd neighbour
//   lo_solve      :457-463 + :477-478  the whole ceres::Solve (Huber 0.1, quaternion manifold,
//                           Levenberg-Marquardt, <= 4 iterations) for one outer pass in ONE
//                           launch: per-residual analytic Jacobians, warp-shuffle reduction of
//                           the 6x6 J'J / 6x1 J'r / cost (28 doubles), 6x6 Cholesky and the
//                           trust-region bookkeeping on-chip — no host round trip per iteration.
//
// The Ceres semantics restated here are documented in oracle/ceres_lm.hpp.
#include <cstdlib>

#include "common.cuh"
#include "gn_solver.cuh"
#include "gn_split.cuh"
#include "internal.h"

namespace vb {


__device__ __forceinline__ int decode_order(unsigned long long k, int nT) {
  const unsigned o = (unsigned)k;
  return (o & 0x80000000u) ? nT - (int)(o & 0x7fffffffu) : (int)o;
}

// Eigen::Quaterniond::Identity().slerp(s, q) (laser_odometry.cpp:159), q and the result as (x, y, z, w).
__device__ __forceinline__ void slerp_from_identity(double s, const double q[4], double out[4]) {
  const double one = 1.0 - 2.220446049250313e-16;
  const double d = q[3], absD = fabs(d);
  double scale0, scale1;
  if (absD >= one) { scale0 = 1.0 - s; scale1 = s; }
  else { const double theta = acos(absD), sinTheta = sin(theta); scale0 = sin((1.0 - s) * theta) / sinTheta; scale1 = sin(s * theta) / sinTheta; }
  if (d < 0.0) scale1 = -scale1;
  out[0] = scale1 * q[0]; out[1] = scale1 * q[1]; out[2] = scale1 * q[2]; out[3] = scale0 + scale1 * q[3];
}

// The reference walks the ring-major target cloud away from `closest` in both directions, classifying every point
// by int(intensity) and stopping at the first point more than NEARBY_SCAN = 2.5 rings away (laser_odometry.cpp:
// 279-324 / 368-417).  Literal restatement, 32 points per step: a later candidate replaces the incumbent only if
// strictly closer, so the winner is the minimum over (distance, visiting order) with the forward walk (ascending j)
// visited before the backward walk (descending j).  Results are warp-uniform.
__device__ void window_walk_literal(const float4* __restrict__ T, int nT, int closest, int id, bool isCorner, float sx,
                                    float sy, float sz, unsigned long long& k2, unsigned long long& k3) {
  const int l = lane_id();
  k2 = 0xffffffffffffffffull; k3 = 0xffffffffffffffffull;
  for (int base = closest + 1; base < nT; base += 32) {  // forward
    const int j = base + l;
    bool brk = false, live = j < nT;
    int rid = 0;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) { t = T[j]; rid = (int)t.w; brk = (double)rid > (double)id + 2.5; }
    const unsigned bm = __ballot_sync(0xffffffffu, brk);
    if (bm) live = live && l < __ffs(bm) - 1;
    if (live) {
      const float d = sqdist_f(t.x, t.y, t.z, sx, sy, sz);
      if ((double)d < 25.0) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
        if (isCorner) { if (rid > id) k2 = key < k2 ? key : k2; }
        else if (rid <= id) k2 = key < k2 ? key : k2;
        else k3 = key < k3 ? key : k3;
      }
    }
    if (bm) break;
  }
  for (int base = closest - 1; base >= 0; base -= 32) {  // backward
    const int j = base - l;
    bool brk = false, live = j >= 0;
    int rid = 0;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) { t = T[j]; rid = (int)t.w; brk = (double)rid < (double)id - 2.5; }
    const unsigned bm = __ballot_sync(0xffffffffu, brk);
    if (bm) live = live && l < __ffs(bm) - 1;
    if (live) {
      const float d = sqdist_f(t.x, t.y, t.z, sx, sy, sz);
      if ((double)d < 25.0) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (0x80000000u + (unsigned)(nT - j));
        if (isCorner) { if (rid < id) k2 = key < k2 ? key : k2; }
        else if (rid >= id) k2 = key < k2 ? key : k2;
        else k3 = key < k3 ? key : k3;
      }
    }
    if (bm) break;
  }
  k2 = warp_min_u64(k2);
  k3 = warp_min_u64(k3);
}

// ---------------------------------------------------------------------------------------------
// lo_associate_brute: one warp per query, exhaustive 1-NN (validation / fallback path, VLOAM_LO_BRUTE=1).  grid (ceil((kMaxSharp + kMaxFlat) / 8), B), block 256.
//   query slots [0, kMaxSharp)            : corner features vs cornerLast
//   query slots [kMaxSharp, +kMaxFlat)    : plane  features vs surfLast
// corr[b][slot] = (closest, ind2, ind3, valid)
__global__ void __launch_bounds__(256) lo_associate_brute(const SRHeader* __restrict__ hdrCur, const SRHeader* __restrict__ hdrLast,
                                                     const LOState* __restrict__ lo,
                                                     const float4* __restrict__ sharp, const float4* __restrict__ flat,
                                                     const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                     int cap, int4* __restrict__ corr) {
  const int b = blockIdx.y;
  const int slot = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int l = lane_id();
  const SRHeader& hc = hdrCur[b];
  const SRHeader& hl = hdrLast[b];
  const bool isCorner = slot < kMaxSharp;
  const int qi = isCorner ? slot : slot - kMaxSharp;
  const int nq = isCorner ? hc.nSharp : hc.nFlat;
  if (slot >= kMaxSharp + kMaxFlat) return;
  int4 out = make_int4(-1, -1, -1, 0);
  if (qi < nq) {
    const float4 p = isCorner ? sharp[(size_t)b * kMaxSharp + qi] : flat[(size_t)b * kMaxFlat + qi];
    // TransformToStart, DISTORTION == false: un = q_last_curr * p + t_last_curr, rounded to float (:158-165)
    double un[3];
    quat_rotate(lo[b].para_q, (double)p.x, (double)p.y, (double)p.z, un);
    const float sx = (float)(un[0] + lo[b].para_t[0]);
    const float sy = (float)(un[1] + lo[b].para_t[1]);
    const float sz = (float)(un[2] + lo[b].para_t[2]);
    const float4* T = isCorner ? cornerLast + (size_t)b * kMaxLessSharp : surfLast + (size_t)b * cap;
    const int nT = isCorner ? hl.nLessSharp : hl.nLessFlat;
    // exact 1-NN (kd-tree replacement): ties -> lower index
    unsigned long long best = 0xffffffffffffffffull;
    for (int j = l; j < nT; j += 32) {
      const float4 t = T[j];
      const float d = sqdist_f(sx, sy, sz, t.x, t.y, t.z);
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
      best = key < best ? key : best;
    }
    best = warp_min_u64(best);
    if (nT > 0 && (double)__uint_as_float((unsigned)(best >> 32)) < 25.0) {  // DISTANCE_SQ_THRESHOLD (:272)
      const int closest = (int)((unsigned)best >> 12);
      const int id = (int)((unsigned)best & 63u);  // closestPointScanID = int(intensity) of the closest point (:275)
      unsigned long long k2, k3;
      window_walk_literal(T, nT, closest, id, isCorner, sx, sy, sz, k2, k3);
      auto decode = [&](unsigned long long k) { return decode_order(k, nT); };
      if (isCorner) {
        if (k2 != 0xffffffffffffffffull) out = make_int4(closest, decode(k2), -1, 1);
      } else if (k2 != 0xffffffffffffffffull && k3 != 0xffffffffffffffffull) {
        out = make_int4(closest, decode(k2), decode(k3), 1);
      }
    }
  }
  if (l == 0) corr[(size_t)b * (kMaxSharp + kMaxFlat) + slot] = out;
}


// ---------------------------------------------------------------------------------------------
// Uniform xy-column grid: the exact replacement of pcl::KdTreeFLANN::setInputCloud / nearestKSearch for laser
// odometry (laser_odometry.cpp:525-526, 269, 356).
//
// Exactness argument.  Columns partition the xy plane into c x c squares.  After the columns
// [qx-k, qx+k] x [qy-k, qy+k] around the query's column have been visited, every unvisited point p satisfies
// |p - q| >= |p - q|_xy >= R_k, where R_k is the distance from q to the border of the visited block.  The search
// stops at the first k with best <= R_k (shrunk by a 1e-5 relative + 1e-4 m margin that dominates all float rounding
// in the column assignment), or R_k >= 5 m: beyond 5 m the reference rejects the match (DISTANCE_SQ_THRESHOLD = 25).
// Distances are the same float expression the brute-force kernel uses and ties break on the original index, so
// both kernels return identical results.
__device__ __forceinline__ int cell_coord(float v, float mn, float inv_c) { return (int)floorf((v - mn) * inv_c); }
// Column-sorted copies keep (original index << 12 | true ring << 6 | int(intensity)) in the w lane; both ring fields are in
// [0, 63].  int(intensity) is the "scan id" the reference's window tests read; the true ring orders a column (below).
__device__ __forceinline__ int pack_index_ring(int j, int trueRing, int rid) { return (j << 12) | ((trueRing & 63) << 6) | (rid & 63); }
__device__ __forceinline__ int packed_index(float w) { return (int)((unsigned)__float_as_int(w) >> 12); }
__device__ __forceinline__ int packed_ring(float w) { return __float_as_int(w) & 63; }
__device__ __forceinline__ int packed_pair(float w) { return (__float_as_int(w) >> 7) & 31; }   // true ring / 2

// lo_build_grid: grid (2, B), block 1024, dynamic smem = (kGridCap + 1) ints.  blockIdx.x: 0 = corner cloud, 1 = surf.
// Counting sort by column with the column table in shared memory (a global-memory table was measured 14x slower).
__global__ void __launch_bounds__(1024) lo_build_grid(const SRHeader* __restrict__ hdrCur, const float4* __restrict__ lessSharp,
                                                       const float4* __restrict__ lessFlat, int cap, GridHeader* __restrict__ ghdr,
                                                       int* __restrict__ cellStartAll, int* __restrict__ /*cursorAll*/,
                                                       float4* __restrict__ sortedC, float4* __restrict__ sortedS) {
  extern __shared__ int cells[];
  __shared__ float s_red[4][32];
  __shared__ int s_firstFull[kMaxRings + 1], s_lastLow[kMaxRings + 1], s_ringStart[kMaxRings + 2];
  __shared__ int s_mono, s_nx, s_ny;
  __shared__ float s_c, s_minx, s_miny;
  __shared__ int s_wsum[32];
  const int which = blockIdx.x, b = blockIdx.y;
  const float4* T = which == 0 ? lessSharp + (size_t)b * kMaxLessSharp : lessFlat + (size_t)b * cap;
  float4* S = which == 0 ? sortedC + (size_t)b * kMaxLessSharp : sortedS + (size_t)b * cap;
  const int n = which == 0 ? hdrCur[b].nLessSharp : hdrCur[b].nLessFlat;
  const int* trueStart = which == 0 ? hdrCur[b].ringStartLessSharp : hdrCur[b].ringStartLessFlat;
  GridHeader& G = ghdr[b * 2 + which];
  int* cs = cellStartAll + (size_t)(b * 2 + which) * (kGridCap + 1);
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  if (n == 0) {
    if (tid == 0) { G.n = 0; G.nx = 0; G.ny = 0; G.ringsOk = 1; G.c = 1.f; G.inv_c = 1.f; G.minx = 0.f; G.miny = 0.f; }
    return;
  }
  // ---- bounding box in xy
  float mnx = 3.0e38f, mny = 3.0e38f, mxx = -3.0e38f, mxy = -3.0e38f;
  for (int j = tid; j < n; j += 1024) {
    const float4 p = T[j];
    mnx = fminf(mnx, p.x); mxx = fmaxf(mxx, p.x); mny = fminf(mny, p.y); mxy = fmaxf(mxy, p.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (l == 0) { s_red[0][w] = mnx; s_red[1][w] = mxx; s_red[2][w] = mny; s_red[3][w] = mxy; }
  if (tid <= kMaxRings) { s_firstFull[tid] = 0x7fffffff; s_lastLow[tid] = -1; }
  if (tid < kMaxRings + 2) s_ringStart[tid] = tid <= kMaxRings ? trueStart[tid] : n;
  if (tid == 0) s_mono = 1;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 32; ++i) {
      mnx = fminf(mnx, s_red[0][i]); mxx = fmaxf(mxx, s_red[1][i]); mny = fminf(mny, s_red[2][i]); mxy = fmaxf(mxy, s_red[3][i]);
    }
    float c = kGridCell;
    int nx, ny;
    while (true) {
      nx = (int)floorf((mxx - mnx) / c) + 1;
      ny = (int)floorf((mxy - mny) / c) + 1;
      if ((long long)nx * ny <= kGridCap) break;
      c *= 1.25f;
    }
    s_c = c; s_nx = nx; s_ny = ny; s_minx = mnx; s_miny = mny;
  }
  __syncthreads();
  const float c = s_c, inv_c = 1.0f / c, minx = s_minx, miny = s_miny;
  const int nx = s_nx, ny = s_ny, ncells = nx * ny;
  for (int i = tid; i <= ncells; i += 1024) cells[i] = 0;
  __syncthreads();
  // ---- histogram + ring bookkeeping
  for (int j = tid; j < n; j += 1024) {
    const float4 p = T[j];
    const int ix = min(max(cell_coord(p.x, minx, inv_c), 0), nx - 1);
    const int iy = min(max(cell_coord(p.y, miny, inv_c), 0), ny - 1);
    atomicAdd(&cells[iy * nx + ix], 1);
    // true ring of point j (ring-major cloud): largest R with ringStart[R] <= j
    int R = 0;
#pragma unroll
    for (int step = 32; step > 0; step >>= 1) if (R + step <= kMaxRings - 1 && s_ringStart[R + step] <= j) R += step;
    const int rid = (int)p.w;
    if (rid == R) atomicMin(&s_firstFull[R], j);
    else if (rid == R - 1) atomicMax(&s_lastLow[R], j);
    else s_mono = 0;
  }
  __syncthreads();
  // ---- exclusive scan of the column counts: thread t owns a contiguous chunk
  const int chunk = (ncells + 1023) / 1024;
  const int c0 = min(tid * chunk, ncells), c1 = min(c0 + chunk, ncells);
  int sum = 0;
  for (int i = c0; i < c1; ++i) sum += cells[i];
  int sc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (l >= o) sc += t; }
  if (l == 31) s_wsum[w] = sc;
  __syncthreads();
  if (w == 0) {
    int v = s_wsum[l], sv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sv, o); if (l >= o) sv += t; }
    s_wsum[l] = sv - v;
  }
  __syncthreads();
  int run = s_wsum[w] + sc - sum;
  for (int i = c0; i < c1; ++i) { const int t = cells[i]; cells[i] = run; cs[i] = run; run += t; }
  if (tid == 0) { cells[ncells] = n; cs[ncells] = n; }
  __syncthreads();
  // ---- scatter, one pair of rings at a time: inside a column the points end up grouped by ring pair in ascending
  // order (arbitrary inside a pair), so the ring-window pass of the search can jump to its rings instead of reading the
  // whole column.  The cloud is ring-major, so a ring pair is a contiguous index range; w carries the original index
  // (tie-breaks, result), the true ring and int(intensity), the "scan id" of the reference's window tests (:275, :285 ...).
  // A ring pair rarely holds more than 1024 points, so a pair is one step of the CTA followed by a barrier: the load of the
  // NEXT pair's point is issued before the barrier of the current one, so that a pair does not pay a full memory latency
  // (measured: 78 -> 76 us per 64-stream launch; the 32 barriers and the shared-memory atomics are what remains).
  auto place = [&](const float4 p, int j, int pr, int jm) {
    const int ix = min(max(cell_coord(p.x, minx, inv_c), 0), nx - 1);
    const int iy = min(max(cell_coord(p.y, miny, inv_c), 0), ny - 1);
    const int pos = atomicAdd(&cells[iy * nx + ix], 1);
    S[pos] = make_float4(p.x, p.y, p.z, __int_as_float(pack_index_ring(j, j >= jm ? 2 * pr + 1 : 2 * pr, (int)p.w)));
  };
  float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
  { const int j = s_ringStart[0] + tid; if (j < s_ringStart[2]) pn = T[j]; }
  for (int pr = 0; pr < kMaxRings / 2; ++pr) {
    const int j0 = s_ringStart[2 * pr], jm = s_ringStart[2 * pr + 1], j1 = s_ringStart[2 * pr + 2];
    const float4 pc = pn;
    if (pr + 1 < kMaxRings / 2) { const int j = j1 + tid; if (j < s_ringStart[2 * pr + 4]) pn = T[j]; }
    if (j1 <= j0) continue;                     // uniform: empty pair
    if (j0 + tid < j1) place(pc, j0 + tid, pr, jm);
    for (int j = j0 + tid + 1024; j < j1; j += 1024) place(T[j], j, pr, jm);
    __syncthreads();
  }
  for (int j = s_ringStart[kMaxRings] + tid; j < n; j += 1024) {   // points past the last ring offset (foreign clouds only)
    const float4 p = T[j];
    const int ix = min(max(cell_coord(p.x, minx, inv_c), 0), nx - 1);
    const int iy = min(max(cell_coord(p.y, miny, inv_c), 0), ny - 1);
    const int pos = atomicAdd(&cells[iy * nx + ix], 1);
    S[pos] = make_float4(p.x, p.y, p.z, __int_as_float(pack_index_ring(j, kMaxRings - 1, (int)p.w)));
  }
  if (tid == 0) {
    G.minx = minx; G.miny = miny; G.c = c; G.inv_c = inv_c; G.nx = nx; G.ny = ny; G.n = n; G.ringsOk = s_mono;
    for (int r = 0; r < kMaxRings + 2; ++r) G.ringStart[r] = s_ringStart[r];
    for (int r = 0; r <= kMaxRings; ++r) { G.firstFull[r] = s_firstFull[r]; G.lastLow[r] = s_lastLow[r]; }
  }
}

// The search runs in groups of 8 lanes (4 queries per warp): the per-query control flow (cell ordering, stopping
// tests, window tables) costs about a thousand warp instructions, far more than the point visits themselves, so
// sharing it between four queries is what matters.  All cross-lane operations use the group's lane mask.
constexpr int kGroup = 8;
__device__ __forceinline__ unsigned long long group_min_u64(unsigned gmask, unsigned long long v) {
#pragma unroll
  for (int o = kGroup / 2; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(gmask, v, o);
    v = t < v ? t : v;
  }
  return v;
}
struct GridView {  // the GridHeader fields the search needs, in registers
  float minx, miny, c, inv_c;
  int nx, ny;
};
// Visit shell k >= 2 (the outer ring of the (2k+1)^2 column block): rows qy-k and qy+k in full, the two end columns
// of the rows in between.  range(aa, bb) is called (by the whole group) for every run [aa, bb) of sorted-array positions
// the shell is made of.
template <typename F>
__device__ __forceinline__ void group_visit_shell(const GridView& G, const int* __restrict__ cs, int qx, int qy, int k, unsigned gmask,
                                                  int gl, F range) {
  const int nseg = 4 * k;
  for (int s0 = 0; s0 < nseg; s0 += kGroup) {
    int a = 0, bnd = 0;
    const int sgi = s0 + gl;
    if (sgi < nseg) {
      int iy, x0, x1;
      if (sgi == 0) { iy = qy - k; x0 = qx - k; x1 = qx + k; }
      else if (sgi == 1) { iy = qy + k; x0 = qx - k; x1 = qx + k; }
      else { const int m = sgi - 2; iy = qy - k + 1 + (m >> 1); x0 = x1 = (m & 1) ? qx + k : qx - k; }
      if (iy >= 0 && iy < G.ny) {
        x0 = max(x0, 0); x1 = min(x1, G.nx - 1);
        if (x0 <= x1) { a = cs[iy * G.nx + x0]; bnd = cs[iy * G.nx + x1 + 1]; }
      }
    }
    const int cnt = min(kGroup, nseg - s0);
    for (int sidx = 0; sidx < cnt; ++sidx) {
      const int aa = __shfl_sync(gmask, a, sidx, kGroup), bb = __shfl_sync(gmask, bnd, sidx, kGroup);
      range(aa, bb);
    }
  }
}
// Best-first visit of the 3 x 3 column block around the query: the query's own column first, then the 8 neighbours
// (one per lane) in order of their lower bound on the distance to any point inside them (distance from the query to
// the column's rectangle, shrunk by the rounding margin); the walk stops as soon as that bound exceeds `limit()` —
// the largest squared distance that could still improve a result.  `limit()` must be uniform across the group.
template <typename F, typename L>
__device__ __forceinline__ void group_visit_block_best_first(const GridView& G, const int* __restrict__ cs, float sx, float sy, int qx, int qy,
                                                             unsigned gmask, int gshift, int gl, F range, L limit) {
  if (qx >= 0 && qx < G.nx && qy >= 0 && qy < G.ny) {
    const int a0 = cs[qy * G.nx + qx], b0 = cs[qy * G.nx + qx + 1];
    range(a0, b0);
  }
  int a = 0, bnd = 0;
  unsigned lbBits = 0xffffffffu;
  {
    const int n = gl < 4 ? gl : gl + 1;  // neighbour cells 0..8 without the centre (4)
    const int cx = qx + (n % 3) - 1, cy = qy + (n / 3) - 1;
    if (cx >= 0 && cx < G.nx && cy >= 0 && cy < G.ny) {
      a = cs[cy * G.nx + cx]; bnd = cs[cy * G.nx + cx + 1];
      if (bnd > a) {
        const float x0 = G.minx + (float)cx * G.c, y0 = G.miny + (float)cy * G.c;
        const float ddx = fmaxf(fmaxf(x0 - sx, sx - (x0 + G.c)), 0.f), ddy = fmaxf(fmaxf(y0 - sy, sy - (y0 + G.c)), 0.f);
        const float lb = fmaxf(sqrtf(ddx * ddx + ddy * ddy) * (1.0f - 1e-5f) - 1e-4f, 0.f);
        lbBits = __float_as_uint(lb * lb);
      }
    }
  }
  while (true) {
    const unsigned m = __reduce_min_sync(gmask, lbBits);
    if (m == 0xffffffffu || __uint_as_float(m) > limit()) break;
    const int src = __ffs(__ballot_sync(gmask, lbBits == m) >> gshift) - 1;
    const int aa = __shfl_sync(gmask, a, src, kGroup), bb = __shfl_sync(gmask, bnd, src, kGroup);
    if (gl == src) lbBits = 0xffffffffu;
    range(aa, bb);
  }
}
// Radius within which the visited block [qx-k, qx+k] x [qy-k, qy+k] is guaranteed complete (<= 0 if none).
__device__ __forceinline__ float grid_safe_radius(const GridView& G, float sx, float sy, int qx, int qy, int k) {
  const float xl = sx - (G.minx + (float)(qx - k) * G.c), xr = (G.minx + (float)(qx + k + 1) * G.c) - sx;
  const float yl = sy - (G.miny + (float)(qy - k) * G.c), yr = (G.miny + (float)(qy + k + 1) * G.c) - sy;
  return fminf(fminf(xl, xr), fminf(yl, yr)) * (1.0f - 1e-5f) - 1e-4f;
}

// lo_associate: one 8-lane group per query, grid search.  grid (ceil((kMaxSharp + kMaxFlat) / 32), B), block 256.
// Same contract as lo_associate_brute.  U = candidate points a lane keeps in flight in the plain column walks (the walks
// are bound by the latency of those loads; more in flight costs registers, i.e. occupancy).
template <int U>
__device__ __forceinline__ void lo_associate_body(const SRHeader* __restrict__ hdrCur, const SRHeader* __restrict__ hdrLast,
                                                     const LOState* __restrict__ lo, const float4* __restrict__ sharp,
                                                     const float4* __restrict__ flat, const float4* __restrict__ cornerLast,
                                                     const float4* __restrict__ surfLast, int cap,
                                                     const GridHeader* __restrict__ ghdr, const int* __restrict__ cellStartAll,
                                                     const float4* __restrict__ sortedC, const float4* __restrict__ sortedS,
                                                     int4* __restrict__ corr, int shardRank, int shardWorld, int distortion) {
  const int b = blockIdx.y;
  const int lane = lane_id(), g = lane / kGroup, gl = lane % kGroup, gshift = g * kGroup;
  const unsigned gmask = 0xffu << gshift;
  const int q = blockIdx.x * 32 + (threadIdx.x >> 5) * 4 + g;
  if (q >= kMaxSharp + kMaxFlat) return;
  const SRHeader& hc = hdrCur[b];
  // Groups are dealt to the live queries first (sharp 0..nSharp-1, then flat 0..nFlat-1), so warps are full; the
  // remaining groups only clear the unused correspondence slots.
  const int nS = min(hc.nSharp, kMaxSharp), nF = min(hc.nFlat, kMaxFlat);
  int slot;
  if (q < nS) slot = q;
  else if (q < nS + nF) slot = kMaxSharp + (q - nS);
  else { const int r = q - (nS + nF); slot = r < kMaxSharp - nS ? nS + r : kMaxSharp + nF + (r - (kMaxSharp - nS)); }
  const bool isCorner = slot < kMaxSharp;
  const int which = isCorner ? 0 : 1;
  const int qi = isCorner ? slot : slot - kMaxSharp;
  const int nq = isCorner ? nS : nF;
  int4 out = make_int4(-1, -1, -1, 0);
  const GridHeader& GH = ghdr[b * 2 + which];
  const int nT = GH.n;
  // point-sharded: warps are dealt round-robin to the ranks; the other ranks' queries stay "no correspondence" here
  const bool mine = shardWorld <= 1 || ((q >> 2) % shardWorld) == shardRank;
  if (mine && qi < nq && nT > 0) {
    GridView G;
    G.minx = GH.minx; G.miny = GH.miny; G.c = GH.c; G.inv_c = GH.inv_c; G.nx = GH.nx; G.ny = GH.ny;
    const float4 p = isCorner ? sharp[(size_t)b * kMaxSharp + qi] : flat[(size_t)b * kMaxFlat + qi];
    // TransformToStart (:149-167): un = Identity.slerp(s, q_last_curr) * p + s * t_last_curr, rounded to float; s == 1 (the
    // shipped DISTORTION == false) keeps its exact short path
    double un[3];
    float sx, sy, sz;
    if (distortion) {
      const double s = (double)__fsub_rn(p.w, (float)(int)p.w) / 0.1;
      double qs[4];
      slerp_from_identity(s, lo[b].para_q, qs);
      quat_rotate(qs, (double)p.x, (double)p.y, (double)p.z, un);
      sx = (float)(un[0] + s * lo[b].para_t[0]); sy = (float)(un[1] + s * lo[b].para_t[1]); sz = (float)(un[2] + s * lo[b].para_t[2]);
    } else {
      quat_rotate(lo[b].para_q, (double)p.x, (double)p.y, (double)p.z, un);
      sx = (float)(un[0] + lo[b].para_t[0]); sy = (float)(un[1] + lo[b].para_t[1]); sz = (float)(un[2] + lo[b].para_t[2]);
    }
    const float4* T = isCorner ? cornerLast + (size_t)b * kMaxLessSharp : surfLast + (size_t)b * cap;
    const float4* S = isCorner ? sortedC + (size_t)b * kMaxLessSharp : sortedS + (size_t)b * cap;
    const int* cs = cellStartAll + (size_t)(b * 2 + which) * (kGridCap + 1);
    const int qx = cell_coord(sx, G.minx, G.inv_c), qy = cell_coord(sy, G.miny, G.inv_c);
    // ---- phase 1: exact nearest neighbour (laser_odometry.cpp:269 / :356)
    unsigned long long best = 0xffffffffffffffffull;
    auto visit1 = [&](int aa, int bb) {
      for (int t = aa + gl; t < bb; t += U * kGroup) {
        float4 tp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t + u * kGroup < bb) tp[u] = S[t + u * kGroup];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u * kGroup >= bb) break;
          const float d = sqdist_f(sx, sy, sz, tp[u].x, tp[u].y, tp[u].z);
          const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(tp[u].w);
          best = key < best ? key : best;   // order: distance, then original index (the ring bits sit below the index)
        }
      }
    };
    // a point at exactly the same distance could still win the index tie-break: the walk only stops on `>`.
    // (no candidate yet: the reduced bits are 0xffffffff = NaN, every comparison is false, the walk continues)
    group_visit_block_best_first(G, cs, sx, sy, qx, qy, gmask, gshift, gl, visit1,
                                 [&]() { return __uint_as_float(__reduce_min_sync(gmask, (unsigned)(best >> 32))); });
    best = group_min_u64(gmask, best);
    for (int k = 1;; ++k) {
      if (k > 1) { group_visit_shell(G, cs, qx, qy, k, gmask, gl, visit1); best = group_min_u64(gmask, best); }
      const float R = grid_safe_radius(G, sx, sy, qx, qy, k);
      if (R >= 5.0f) break;
      if (best != 0xffffffffffffffffull && R > 0.f && __uint_as_float((unsigned)(best >> 32)) <= R * R) break;
    }
    if (best != 0xffffffffffffffffull && (double)__uint_as_float((unsigned)(best >> 32)) < 25.0) {  // :272
      const int closest = (int)((unsigned)best >> 12);
      const int id = (int)((unsigned)best & 63u);  // closestPointScanID = int(intensity) of the closest point (:275)
      unsigned long long k2 = 0xffffffffffffffffull, k3 = 0xffffffffffffffffull;
      if (!GH.ringsOk) {
        // literal walk by the whole group's warp is not possible inside a group: do it lane-serially per group
        // (never taken for ring-major clouds produced by scan registration; kept for exactness on foreign clouds)
        if (gl == 0) {
          double m2 = 25.0, m3 = 25.0;
          int i2 = -1, i3 = -1;
          for (int j = closest + 1; j < nT; ++j) {
            const float4 t = T[j];
            const int rid = (int)t.w;
            if ((double)rid > (double)id + 2.5) break;
            const double d = (double)sqdist_f(t.x, t.y, t.z, sx, sy, sz);
            if (isCorner) { if (rid > id && d < m2) { m2 = d; i2 = j; } }
            else if (rid <= id && d < m2) { m2 = d; i2 = j; }
            else if (rid > id && d < m3) { m3 = d; i3 = j; }
          }
          for (int j = closest - 1; j >= 0; --j) {
            const float4 t = T[j];
            const int rid = (int)t.w;
            if ((double)rid < (double)id - 2.5) break;
            const double d = (double)sqdist_f(t.x, t.y, t.z, sx, sy, sz);
            if (isCorner) { if (rid < id && d < m2) { m2 = d; i2 = j; } }
            else if (rid >= id && d < m2) { m2 = d; i2 = j; }
            else if (rid < id && d < m3) { m3 = d; i3 = j; }
          }
          if (isCorner) { if (i2 >= 0) out = make_int4(closest, i2, -1, 1); }
          else if (i2 >= 0 && i3 >= 0) out = make_int4(closest, i2, i3, 1);
        }
      } else {
        // ---- phase 2: nearest point per class inside the +-2.5-ring window (:279-324 / :368-417).
        // Forward walk stops at the first j > closest with int(intensity) >= id + 3: that is the first such point of
        // true ring id + 3, else the start of ring id + 4 (every point there qualifies).  Backward walk stops at the
        // last j < closest with int(intensity) <= id - 3: the last such point of true ring id - 2, else the point just
        // before that ring.  So the walks visit exactly the indices [lo_j, hi_j) \ {closest}.
        int hi_j = nT, lo_j = 0;
        if (id + 3 <= kMaxRings - 1) hi_j = min(GH.firstFull[id + 3], GH.ringStart[min(id + 4, kMaxRings)]);
        if (id - 2 >= 0) lo_j = max(GH.lastLow[id - 2], GH.ringStart[id - 2] - 1) + 1;
        // Columns are ordered by ring pair (true ring / 2).  [lo_j, hi_j) only holds rings id-2 .. id+3, so a column
        // contributes the sub-range of pairs [pairLo, pairHi]: found by an 8-way probing lower bound (one probe per
        // lane), then walked until the first pair beyond pairHi.
        const int pairLo = max(id - 2, 0) >> 1, pairHi = min(id + 3, kMaxRings - 1) >> 1;
        auto consider = [&](const float4 tp) {
          const int j = packed_index(tp.w);
          if (j < lo_j || j >= hi_j || j == closest) return;
          const float d = sqdist_f(tp.x, tp.y, tp.z, sx, sy, sz);
          if (!((double)d < 25.0)) return;
          const int rid = packed_ring(tp.w);
          const bool fwd = j > closest;
          const unsigned order = fwd ? (unsigned)j : 0x80000000u + (unsigned)(nT - j);
          const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | order;
          const bool classA = fwd ? (rid <= id) : (rid >= id);  // same-ring side of the walk
          if (isCorner) { if (!classA) k2 = key < k2 ? key : k2; }
          else if (classA) k2 = key < k2 ? key : k2;
          else k3 = key < k3 ? key : k3;
        };
        // one column [aa, bb): jump to the window's ring pairs, stop after them
        auto visit2 = [&](int aa, int bb) {
          int lo = aa, hi = bb;
          while (hi - lo > kGroup) {
            const int step = (hi - lo + kGroup) / (kGroup + 1);
            const int pidx = lo + (gl + 1) * step - 1;
            const bool ge = pidx >= hi || packed_pair(S[pidx].w) >= pairLo;
            const unsigned m = (__ballot_sync(gmask, ge) >> gshift) & 0xffu;
            if (m == 0u) { lo = min(lo + kGroup * step, hi); }
            else { const int f = __ffs(m) - 1; hi = min(lo + (f + 1) * step, hi); lo += f * step; }
          }
          for (int base = lo; base < bb; base += kGroup) {
            const int t = base + gl;
            bool beyond = false;
            if (t < bb) {
              const float4 tp = S[t];
              beyond = packed_pair(tp.w) > pairHi;
              consider(tp);
            }
            if (__ballot_sync(gmask, beyond) & gmask) break;   // sorted by pair: nothing of interest further on
          }
        };
        // a run of several columns (shell rows): not ordered as a whole, every point is tested
        auto visit2_run = [&](int aa, int bb) {
          for (int t = aa + gl; t < bb; t += U * kGroup) {
            float4 tp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (t + u * kGroup < bb) tp[u] = S[t + u * kGroup];
#pragma unroll
            for (int u = 0; u < U; ++u) if (t + u * kGroup < bb) consider(tp[u]);
          }
        };
        // a column can still matter while its bound is below the worst of the needed classes (25 = none found yet)
        group_visit_block_best_first(G, cs, sx, sy, qx, qy, gmask, gshift, gl, visit2, [&]() {
          const unsigned d2 = __reduce_min_sync(gmask, (unsigned)(k2 >> 32));
          float lim = d2 == 0xffffffffu ? 25.0f : __uint_as_float(d2);
          if (!isCorner) {
            const unsigned d3 = __reduce_min_sync(gmask, (unsigned)(k3 >> 32));
            lim = fmaxf(lim, d3 == 0xffffffffu ? 25.0f : __uint_as_float(d3));
          }
          return lim;
        });
        k2 = group_min_u64(gmask, k2);
        k3 = group_min_u64(gmask, k3);
        for (int k = 1;; ++k) {
          if (k > 1) { group_visit_shell(G, cs, qx, qy, k, gmask, gl, visit2_run); k2 = group_min_u64(gmask, k2); k3 = group_min_u64(gmask, k3); }
          const float R = grid_safe_radius(G, sx, sy, qx, qy, k);
          if (R >= 5.0f) break;
          if (R > 0.f) {
            const float R2 = R * R;
            const bool ok2 = k2 != 0xffffffffffffffffull && __uint_as_float((unsigned)(k2 >> 32)) <= R2;
            const bool ok3 = isCorner || (k3 != 0xffffffffffffffffull && __uint_as_float((unsigned)(k3 >> 32)) <= R2);
            if (ok2 && ok3) break;
          }
        }
        if (isCorner) {
          if (k2 != 0xffffffffffffffffull) out = make_int4(closest, decode_order(k2, nT), -1, 1);
        } else if (k2 != 0xffffffffffffffffull && k3 != 0xffffffffffffffffull) {
          out = make_int4(closest, decode_order(k2, nT), decode_order(k3, nT), 1);
        }
      }
    }
  }
  if (gl == 0) corr[(size_t)b * (kMaxSharp + kMaxFlat) + slot] = out;
}

#define VB_LO_ASSOC_ARGS                                                                                                         \
  const SRHeader *__restrict__ hdrCur, const SRHeader *__restrict__ hdrLast, const LOState *__restrict__ lo,                     \
      const float4 *__restrict__ sharp, const float4 *__restrict__ flat, const float4 *__restrict__ cornerLast,                  \
      const float4 *__restrict__ surfLast, int cap, const GridHeader *__restrict__ ghdr, const int *__restrict__ cellStartAll,   \
      const float4 *__restrict__ sortedC, const float4 *__restrict__ sortedS, int4 *__restrict__ corr, int shardRank,          \
      int shardWorld, int distortion
#define VB_LO_ASSOC_PASS hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, ghdr, cellStartAll, sortedC, sortedS, corr, shardRank, shardWorld, distortion
// Two register budgets of the same body: 64 registers (4 CTAs / SM) and <= 40 (6 CTAs / SM); the kernel is latency
// bound, so which one wins is an occupancy question settled by measurement (VLOAM_LO_ASSOC_OCC=4|5|6|8; measured on B200 at 128 streams: 334 / 314 / 293 us for 4 / 5 / 6, so 6 is the default).
__global__ void __launch_bounds__(256, 4) lo_associate(VB_LO_ASSOC_ARGS) { lo_associate_body<4>(VB_LO_ASSOC_PASS); }
__global__ void __launch_bounds__(256, 5) lo_associate_occ5(VB_LO_ASSOC_ARGS) { lo_associate_body<2>(VB_LO_ASSOC_PASS); }
__global__ void __launch_bounds__(256, 6) lo_associate_occ6(VB_LO_ASSOC_ARGS) { lo_associate_body<1>(VB_LO_ASSOC_PASS); }
__global__ void __launch_bounds__(256, 8) lo_associate_occ8(VB_LO_ASSOC_ARGS) { lo_associate_body<1>(VB_LO_ASSOC_PASS); }


// ---------------------------------------------------------------------------------------------
// lo_solve: grid (B), block 256.  One CTA solves one stream's pass.
__device__ __forceinline__ void lo_solve_body(const SRHeader* __restrict__ hdrCur, LOState* __restrict__ lo,
                                                 const float4* __restrict__ sharp, const float4* __restrict__ flat,
                                                 const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                 int cap, const int4* __restrict__ corr, int pass, int max_iterations,
                                                 int integrate, const ShardView sv) {
  __shared__ LMShared S;
  const int b = blockIdx.x;
  LOState& st = lo[b];
  SolveTrace* tr = &st.trace[pass];
  const int4* cr = corr + (size_t)b * (kMaxSharp + kMaxFlat);
  const float4* sh = sharp + (size_t)b * kMaxSharp;
  const float4* fl = flat + (size_t)b * kMaxFlat;
  const float4* CL = cornerLast + (size_t)b * kMaxLessSharp;
  const float4* SL = surfLast + (size_t)b * cap;
  const int nSharp = hdrCur[b].nSharp, nFlat = hdrCur[b].nFlat;

  // Stage the per-correspondence geometry in shared memory once: it does not change between LM iterations, and the
  // gathers (correspondence record -> 2-3 target points) are the longest dependent chain of an evaluation.
  //   edge : a, b as six floats packed into G[0..2];   plane: unit normal in G[0..2], d0 in G[3].
  constexpr int kSlots = kMaxSharp + kMaxFlat;
  extern __shared__ double lo_dyn[];
  double* G = lo_dyn;                                            // [4][kSlots]
  unsigned char* kind = reinterpret_cast<unsigned char*>(G + 4 * kSlots);   // 0 none, 1 edge, 2 plane
  __shared__ int s_nc, s_np;
  if (threadIdx.x == 0) {
    s_nc = 0; s_np = 0;
    for (int i = 0; i < 4; ++i) S.x[i] = st.para_q[i];
    for (int i = 0; i < 3; ++i) S.x[4 + i] = st.para_t[i];
  }
  __syncthreads();
  {
    int nc = 0, np = 0;
    for (int s = threadIdx.x; s < kSlots; s += blockDim.x) {
      const bool isCorner = s < kMaxSharp;
      const int qi = isCorner ? s : s - kMaxSharp;
      unsigned char kd = 0;
      if (qi < (isCorner ? nSharp : nFlat)) {
        const int4 c = cr[s];
        if (c.w) {
          if (isCorner) {
            const float4 A = CL[c.x], Bp = CL[c.y];
            G[0 * kSlots + s] = __hiloint2double(__float_as_int(A.x), __float_as_int(A.y));
            G[1 * kSlots + s] = __hiloint2double(__float_as_int(A.z), __float_as_int(Bp.x));
            G[2 * kSlots + s] = __hiloint2double(__float_as_int(Bp.y), __float_as_int(Bp.z));
            kd = 1; nc++;
          } else {
            const float4 Jp = SL[c.x], Lp = SL[c.y], Mp = SL[c.z];
            // ljm_norm = normalize((j - l) x (j - m))  (lidarFactor.hpp:68-69)
            const double ax = (double)Jp.x - (double)Lp.x, ay = (double)Jp.y - (double)Lp.y, az = (double)Jp.z - (double)Lp.z;
            const double bx = (double)Jp.x - (double)Mp.x, by = (double)Jp.y - (double)Mp.y, bz = (double)Jp.z - (double)Mp.z;
            double n[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
            const double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
            if (z2 > 0.0) { const double nn = sqrt(z2); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
            G[0 * kSlots + s] = n[0]; G[1 * kSlots + s] = n[1]; G[2 * kSlots + s] = n[2];
            G[3 * kSlots + s] = -(n[0] * (double)Jp.x + n[1] * (double)Jp.y + n[2] * (double)Jp.z);
            kd = 2; np++;
          }
        }
      }
      kind[s] = kd;
    }
    // correspondence counts (:348, :441)
    nc = __reduce_add_sync(0xffffffffu, nc);
    np = __reduce_add_sync(0xffffffffu, np);
    if (lane_id() == 0) { atomicAdd(&s_nc, nc); atomicAdd(&s_np, np); }
    __syncthreads();
    if (sv.world > 1) {   // point-sharded: this rank only holds its slice of the correspondences
      if (threadIdx.x == 0) { S.red[0] = (double)s_nc; S.red[1] = (double)s_np; }
      __syncthreads();
      shard_allreduce(sv, b, S.red, 2);
      if (threadIdx.x == 0) { s_nc = (int)S.red[0]; s_np = (int)S.red[1]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) { st.corner_correspondence = s_nc; st.plane_correspondence = s_np; tr->n_corner = s_nc; tr->n_plane = s_np; }
  }

  auto evaluate = [&](const double* x) {
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double q[4] = {x[0], x[1], x[2], x[3]};
    const double t[3] = {x[4], x[5], x[6]};
    for (int s = threadIdx.x; s < kSlots; s += blockDim.x) {
      const int kd = kind[s];
      if (!kd) continue;
      if (kd == 1) {
        const float4 p = sh[s];
        const double g0 = G[0 * kSlots + s], g1 = G[1 * kSlots + s], g2 = G[2 * kSlots + s];
        const double a[3] = {(double)__int_as_float(__double2hiint(g0)), (double)__int_as_float(__double2loint(g0)),
                             (double)__int_as_float(__double2hiint(g1))};
        const double bb[3] = {(double)__int_as_float(__double2loint(g1)), (double)__int_as_float(__double2hiint(g2)),
                              (double)__int_as_float(__double2loint(g2))};
        edge_block(q, t, p, a, bb, acc);
      } else {
        const float4 p = fl[s - kMaxSharp];
        const double n[3] = {G[0 * kSlots + s], G[1 * kSlots + s], G[2 * kSlots + s]};
        plane_block(q, t, p, n, G[3 * kSlots + s], acc);
      }
    }
    block_reduce28(acc, S.red, S.scratch);
    if (sv.world > 1) shard_allreduce(sv, b, S.red, 28);   // J'J, J'r and the cost summed over the ranks, in the kernel
  };

  lm_solve_block(S, tr, max_iterations, false, evaluate);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) st.para_q[i] = S.x[i];
    for (int i = 0; i < 3; ++i) st.para_t[i] = S.x[4 + i];
    if (integrate) {  // :477-478
      double rt[3];
      quat_rotate(st.q_w, st.para_t[0], st.para_t[1], st.para_t[2], rt);
      st.t_w[0] += rt[0]; st.t_w[1] += rt[1]; st.t_w[2] += rt[2];
      double qn[4];
      quat_mul(st.q_w, st.para_q, qn);
      for (int i = 0; i < 4; ++i) st.q_w[i] = qn[i];
    }
  }
}

#define VB_LO_SOLVE_ARGS                                                                                                       \
  const SRHeader *__restrict__ hdrCur, LOState *__restrict__ lo, const float4 *__restrict__ sharp, const float4 *__restrict__ flat, \
      const float4 *__restrict__ cornerLast, const float4 *__restrict__ surfLast, int cap, const int4 *__restrict__ corr, int pass, \
      int max_iterations, int integrate, const ShardView sv
// Two register budgets (as lm_solve): 212 registers, one CTA per SM, or <= 128 with two per SM (VLOAM_LO_SOLVE_REGS=128):
// while a solve runs, the registers it leaves free decide how much of the other handles' kernels its SM can take.
__global__ void __launch_bounds__(256) lo_solve(VB_LO_SOLVE_ARGS) {
  lo_solve_body(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass, max_iterations, integrate, sv);
}
__global__ void __launch_bounds__(256, 2) lo_solve_r128(VB_LO_SOLVE_ARGS) {
  lo_solve_body(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass, max_iterations, integrate, sv);
}

// ---------------------------------------------------------------------------------------------
// Wide solve (gn_split.cuh).  lo_stage_residuals: grid (B), block 256 — the per-correspondence geometry of lidarFactor.hpp:14-106
// as residual records (what lo_solve stages in shared memory), plus the correspondence counts (:348, :441).
__global__ void __launch_bounds__(256) lo_stage_residuals(const SRHeader* __restrict__ hdrCur, LOState* __restrict__ lo,
                                                           const float4* __restrict__ sharp, const float4* __restrict__ flat,
                                                           const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                           int cap, const int4* __restrict__ corr, int pass, const GNRecArray recAll,
                                                           double* __restrict__ counts /*[B][2] or nullptr*/, int distortion) {
  constexpr int kSlots = kMaxSharp + kMaxFlat;
  const int b = blockIdx.x;
  LOState& st = lo[b];
  const int4* cr = corr + (size_t)b * kSlots;
  const float4* sh = sharp + (size_t)b * kMaxSharp;
  const float4* fl = flat + (size_t)b * kMaxFlat;
  const float4* CL = cornerLast + (size_t)b * kMaxLessSharp;
  const float4* SL = surfLast + (size_t)b * cap;
  const int nSharp = hdrCur[b].nSharp, nFlat = hdrCur[b].nFlat;
  __shared__ int s_nc, s_np;
  if (threadIdx.x == 0) { s_nc = 0; s_np = 0; }
  __syncthreads();
  int nc = 0, np = 0;
  for (int s = threadIdx.x; s < kSlots; s += blockDim.x) {
    const bool isCorner = s < kMaxSharp;
    const int qi = isCorner ? s : s - kMaxSharp;
    GNResidual R;
    R.type = 0; R.px = R.py = R.pz = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) R.v[i] = 0.0;
    if (qi < (isCorner ? nSharp : nFlat)) {
      const int4 c = cr[s];
      if (c.w) {
        const float4 p = isCorner ? sh[qi] : fl[qi];
        R.px = p.x; R.py = p.y; R.pz = p.z;
        const double s = distortion ? (double)__fsub_rn(p.w, (float)(int)p.w) / 0.1 : 1.0;    // :329-335 / :425-431
        if (isCorner) {
          const float4 A = CL[c.x], Bp = CL[c.y];
          R.v[0] = A.x; R.v[1] = A.y; R.v[2] = A.z; R.v[3] = Bp.x; R.v[4] = Bp.y; R.v[5] = Bp.z;
          R.v[6] = s; R.type = distortion ? 3 : 1; nc++;
        } else {
          const float4 Jp = SL[c.x], Lp = SL[c.y], Mp = SL[c.z];
          // ljm_norm = normalize((j - l) x (j - m))  (lidarFactor.hpp:68-69)
          const double ax = (double)Jp.x - (double)Lp.x, ay = (double)Jp.y - (double)Lp.y, az = (double)Jp.z - (double)Lp.z;
          const double bx = (double)Jp.x - (double)Mp.x, by = (double)Jp.y - (double)Mp.y, bz = (double)Jp.z - (double)Mp.z;
          double n[3] = {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx};
          const double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
          if (z2 > 0.0) { const double nn = sqrt(z2); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
          R.v[0] = n[0]; R.v[1] = n[1]; R.v[2] = n[2];
          R.v[3] = -(n[0] * (double)Jp.x + n[1] * (double)Jp.y + n[2] * (double)Jp.z);
          R.v[4] = s; R.type = distortion ? 4 : 2; np++;
        }
      }
    }
    gn_store(recAll, (size_t)b, s, R);
  }
  nc = __reduce_add_sync(0xffffffffu, nc);
  np = __reduce_add_sync(0xffffffffu, np);
  if (lane_id() == 0) { atomicAdd(&s_nc, nc); atomicAdd(&s_np, np); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (counts) { counts[2 * b] = (double)s_nc; counts[2 * b + 1] = (double)s_np; }    // summed across ranks before lo_gn_finish reads them
    else { st.corner_correspondence = s_nc; st.plane_correspondence = s_np; st.trace[pass].n_corner = s_nc; st.trace[pass].n_plane = s_np; }
  }
}
// after the last gn_step of a pass: the (rank-summed) counts and the pose integration of :477-478
__global__ void lo_gn_finish(LOState* __restrict__ lo, int B, int pass, int integrate, const double* __restrict__ counts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LOState& st = lo[b];
  if (counts) {
    st.corner_correspondence = (int)counts[2 * b]; st.plane_correspondence = (int)counts[2 * b + 1];
    st.trace[pass].n_corner = st.corner_correspondence; st.trace[pass].n_plane = st.plane_correspondence;
  }
  if (integrate) {
    double rt[3];
    quat_rotate(st.q_w, st.para_t[0], st.para_t[1], st.para_t[2], rt);
    st.t_w[0] += rt[0]; st.t_w[1] += rt[1]; st.t_w[2] += rt[2];
    double qn[4];
    quat_mul(st.q_w, st.para_q, qn);
    for (int i = 0; i < 4; ++i) st.q_w[i] = qn[i];
  }
}

__global__ void lo_init_state(LOState* lo, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LOState& s = lo[b];
  s.para_q[0] = s.para_q[1] = s.para_q[2] = 0.0; s.para_q[3] = 1.0;
  s.para_t[0] = s.para_t[1] = s.para_t[2] = 0.0;
  s.q_w[0] = s.q_w[1] = s.q_w[2] = 0.0; s.q_w[3] = 1.0;
  s.t_w[0] = s.t_w[1] = s.t_w[2] = 0.0;
  s.corner_correspondence = s.plane_correspondence = 0;
  s.trace[0].n_records = s.trace[1].n_records = 0;
}

// !detach_VO_LO: para_q / para_t are overwritten with the VO prior at the top of BOTH outer passes, before the
// association (laser_odometry.cpp:223-236).  prior: [B][7] = q(xyzw), t.
__global__ void lo_set_motion(LOState* lo, const double* __restrict__ prior, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < 4; ++i) lo[b].para_q[i] = prior[b * 7 + i];
  for (int i = 0; i < 3; ++i) lo[b].para_t[i] = prior[b * 7 + 4 + i];
}

// pose[b][16] = q_last_curr(4) t_last_curr(3) q_w_curr(4) t_w_curr(3) corner_correspondence plane_correspondence
__global__ void lo_export_pose(const LOState* __restrict__ lo, double* __restrict__ pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const LOState& s = lo[b];
  double* o = pose + (size_t)b * 16;
  for (int i = 0; i < 4; ++i) { o[i] = s.para_q[i]; o[7 + i] = s.q_w[i]; }
  for (int i = 0; i < 3; ++i) { o[4 + i] = s.para_t[i]; o[11 + i] = s.t_w[i]; }
  o[14] = s.corner_correspondence; o[15] = s.plane_correspondence;
}
void launch_lo_export(Profiler* prof, cudaStream_t st, const LOState* lo, double* pose, int B) {
  VB_LAUNCH(prof, K_LO_EXPORT, st, lo_export_pose<<<(B + 127) / 128, 128, 0, st>>>(lo, pose, B));
}
void launch_lo_set_motion(Profiler* prof, cudaStream_t st, LOState* lo, const double* motion, int B) {
  VB_LAUNCH(prof, K_LO_SET_MOTION, st, lo_set_motion<<<(B + 127) / 128, 128, 0, st>>>(lo, motion, B));
}

// Checkpoint / resume: overwrite the accumulated odometry pose q_w_curr / t_w_curr (laser_odometry.cpp:80-81) — a stream
// that starts in the middle of a trajectory.  pose: [B][7] = q(xyzw), t.
__global__ void lo_set_pose(LOState* lo, const double* __restrict__ pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int i = 0; i < 4; ++i) lo[b].q_w[i] = pose[b * 7 + i];
  for (int i = 0; i < 3; ++i) lo[b].t_w[i] = pose[b * 7 + 4 + i];
}
void launch_lo_set_pose(Profiler* prof, cudaStream_t st, LOState* lo, const double* pose, int B) {
  VB_LAUNCH(prof, K_LO_SET_MOTION, st, lo_set_pose<<<(B + 127) / 128, 128, 0, st>>>(lo, pose, B));
}

void launch_lo_init(Profiler* prof, cudaStream_t st, LOState* lo, int B) {
  VB_LAUNCH(prof, K_LO_INIT, st, lo_init_state<<<(B + 127) / 128, 128, 0, st>>>(lo, B));
}

static bool lo_use_brute() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VLOAM_LO_BRUTE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

constexpr int kLoSolveDynSmem = (kMaxSharp + kMaxFlat) * (4 * (int)sizeof(double) + 1);
constexpr int kLoSplitMinBatch = 1 << 30;   // batch size from which the wide solve is the default: set by measurement
void gn_allreduce_partials(void* ncclComm, double* partial, size_t count, cudaStream_t st);   // capi.cu
// Opt-in shared-memory sizes are per-device function attributes: set (and checked) once per context, on its device.
cudaError_t lo_prepare_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lo_build_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (kGridCap + 1) * (int)sizeof(int));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lo_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, kLoSolveDynSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lo_solve_r128, cudaFuncAttributeMaxDynamicSharedMemorySize, kLoSolveDynSmem);
  return e;
}

void launch_lo_build_grid(Profiler* prof, cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const float4* lessSharp,
                          const float4* lessFlat, const LOGrid* g) {
  const int smem = (kGridCap + 1) * (int)sizeof(int);
  VB_LAUNCH(prof, K_LO_BUILD_GRID, st, lo_build_grid<<<dim3(2, B), 1024, smem, st>>>(hdrCur, lessSharp, lessFlat, cap, g->hdr, g->cellStart, g->cursor,
                                                                                  g->sorted[0], g->sorted[1]));
}

void launch_lo_pass(Profiler* prof, cudaStream_t st, int B, int cap, const SRHeader* hdrCur, const SRHeader* hdrLast, LOState* lo,
                    const float4* sharp, const float4* flat, const float4* cornerLast, const float4* surfLast,
                    const LOGrid* g, int4* corr, int pass, int max_iterations, int integrate, const double* prior,
                    const ShardView* shard, int solverMode, bool distortion) {
  const ShardView sv = shard ? *shard : ShardView();
  if (prior) VB_LAUNCH(prof, K_LO_SET_MOTION, st, lo_set_motion<<<(B + 127) / 128, 128, 0, st>>>(lo, prior, B));
  if (lo_use_brute() && sv.world <= 1 && !distortion)
    VB_LAUNCH(prof, K_LO_ASSOCIATE_BRUTE, st, lo_associate_brute<<<dim3((kMaxSharp + kMaxFlat + 7) / 8, B), 256, 0, st>>>(
                                                  hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, corr));
  else
  {
    static const int occ = [] { const char* e = getenv("VLOAM_LO_ASSOC_OCC"); return e ? atoi(e) : 6; }();
    const dim3 grid((kMaxSharp + kMaxFlat + 31) / 32, B);
    if (occ >= 8)
      VB_LAUNCH(prof, K_LO_ASSOCIATE, st, lo_associate_occ8<<<grid, 256, 0, st>>>(
                                              hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, g->hdr, g->cellStart,
                                              g->sorted[0], g->sorted[1], corr, sv.rank, sv.world, distortion ? 1 : 0));
    else if (occ >= 6)
      VB_LAUNCH(prof, K_LO_ASSOCIATE, st, lo_associate_occ6<<<grid, 256, 0, st>>>(
                                              hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, g->hdr, g->cellStart,
                                              g->sorted[0], g->sorted[1], corr, sv.rank, sv.world, distortion ? 1 : 0));
    else if (occ == 5)
      VB_LAUNCH(prof, K_LO_ASSOCIATE, st, lo_associate_occ5<<<grid, 256, 0, st>>>(
                                              hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, g->hdr, g->cellStart,
                                              g->sorted[0], g->sorted[1], corr, sv.rank, sv.world, distortion ? 1 : 0));
    else
      VB_LAUNCH(prof, K_LO_ASSOCIATE, st, lo_associate<<<grid, 256, 0, st>>>(
                                              hdrCur, hdrLast, lo, sharp, flat, cornerLast, surfLast, cap, g->hdr, g->cellStart,
                                              g->sorted[0], g->sorted[1], corr, sv.rank, sv.world, distortion ? 1 : 0));
  }
  // vloam_lidar_params::solver_mode (0 = by batch size); the NCCL exchange needs the wide layout, the in-kernel peer exchange the narrow one
  // (the motion-distortion blocks exist in the wide layout only)
  const bool split = g->gnState != nullptr && (g->ncclComm != nullptr || distortion || (sv.world <= 1 && (solverMode == 2 || (solverMode == 0 && B >= kLoSplitMinBatch))));
  if (split) {
    // wide solve (gn_split.cuh): residual records once per pass, then one launch over all streams per LM evaluation
    double* counts = g->ncclComm ? g->gnCounts : nullptr;
    VB_LAUNCH(prof, K_LO_STEP, st, lo_stage_residuals<<<B, 256, 0, st>>>(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass, GNRecArray{g->gnRecV, g->gnRecP, kMaxSharp + kMaxFlat}, counts, distortion ? 1 : 0));
    if (counts) gn_allreduce_partials(g->ncclComm, counts, (size_t)B * 2, st);
    GNProblemView pv{};
    pv.rec = GNRecArray{g->gnRecV, g->gnRecP, kMaxSharp + kMaxFlat}; pv.blocksPerStream = 1;
    pv.fixedCount[0] = kMaxSharp + kMaxFlat;
    pv.x = Strided{&lo[0].para_q[0], sizeof(LOState)};
    pv.trace = Strided{&lo[0].trace[pass], sizeof(LOState)};
    pv.slerp = distortion ? 1 : 0;
    launch_gn_solve(prof, st, B, pv, g->gnState, g->gnPartial, max_iterations, K_LO_ACCUMULATE, K_LO_STEP, g->ncclComm);
    VB_LAUNCH(prof, K_LO_STEP, st, lo_gn_finish<<<(B + 127) / 128, 128, 0, st>>>(lo, B, pass, integrate, counts));
    return;
  }
  static const int regsEnv = [] { const char* e = getenv("VLOAM_LO_SOLVE_REGS"); return e ? atoi(e) : 0; }();
  const bool regs128 = regsEnv == 128;     // measured: no gain for the step at any batch size (41.65 k vs 41.68 k scans/s), so only on request
  if (regs128)
    VB_LAUNCH(prof, K_LO_SOLVE, st, lo_solve_r128<<<B, 256, kLoSolveDynSmem, st>>>(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass,
                                                                      max_iterations, integrate, sv));
  else
    VB_LAUNCH(prof, K_LO_SOLVE, st, lo_solve<<<B, 256, kLoSolveDynSmem, st>>>(hdrCur, lo, sharp, flat, cornerLast, surfLast, cap, corr, pass,
                                                                 max_iterations, integrate, sv));
}

}  // namespace vb
