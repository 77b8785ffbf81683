// vloam_b200 — laserMapping on sm_100a.
//
// Replaces vloam::LaserMapping::input / solveMapping
// (reference src/lidar_odometry_mapping/src/laser_mapping.cpp:167-196, 198-708):
//
//   lm_prepare        :186-216 initial guess, :218-402 rolling 21x21x11 cube grid shift (a permutation of the cube
//                     table, no point moves), :404-430 the 5x5x3 valid-cube list and sub-map offsets
//   lm_voxel_segments :432-440 / :689-702 pcl::VoxelGrid on global-memory segments (scan features, map cubes):
//                     bbox -> voxel keys -> stable LSD radix sort -> ordered centroid sums
//   lm_grid_*         :452-453 the kd-tree replacement: the sub-map counting-sorted into 1.001 m xy columns.  A match
//                     needs its 5th neighbour within 1 m (:479, :547), so the 3x3 column block around a query is
//                     always sufficient — no expansion, exact by construction.
//   lm_associate      :472-581 exact 5-NN, PCA line test (corner) / least-squares plane fit + 0.2 m check (surf)
//   lm_solve          :609-617 the whole ceres::Solve of one outer pass in one launch (gn_solver.cuh)
//   lm_insert_*       :636-683 transformUpdate, scan points into cubes (stable by cube)
//   lm_rebuild_*      :689-702 re-filter every valid cube, compact the map into the other buffer
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/vloam_b200.h"
#include "common.cuh"
#include "cta_sort.cuh"
#include "gn_solver.cuh"
#include "gn_split.cuh"
#include "internal.h"

namespace vb {

constexpr int kCubeW = 21, kCubeH = 21, kCubeD = 11, kCubes = kCubeW * kCubeH * kCubeD;  // laser_mapping.h:110-114
constexpr int kMaxValid = 125;                                                           // laser_mapping.h:116
// LMState::counters
enum { kCntScans = 0, kCntSolved, kCntIters0, kCntIters1, kCntIndexedCorner, kCntIndexedSurf, kCntWorkCorner, kCntWorkSurf, kCntMerged,
       kCntMergedInPlace, kCntFiltered, kCntAppended, kCntVoxelsInserted, kCntRepacks, kCntQueries, kCntFactors };
// LMState::error bits (vloam_get_lm_status)
constexpr int kLmErrWorkList = 1;   // more than kMaxWork cubes would be rewritten by one scan: the scan's points were not inserted
constexpr int kLmErrScratch = 2;    // the re-filter scratch would overflow: the scan's points were not inserted
constexpr int kLmErrCapacity = 4;   // << kind: map_capacity_points exceeded for corner (4) / surf (8): that map keeps its previous content
constexpr int kMaxWork = 200;   // cubes rewritten per scan: the valid ones plus cubes that received points
// Column index of a cube (the kd-tree replacement, see lm_associate): 50 x 50 xy-columns of 1.001 m per 50 m cube.
constexpr float kCubeCell = 1.001f;
constexpr int kCubeCellsX = 50, kCubeCells = kCubeCellsX * kCubeCellsX;
// Inside the cube the points are ordered by 4 m z-layer, then by column (row-major): a query only reads, per z-layer its
// +-1.001 m interval touches (one or two) and per column row (three), ONE contiguous run covering its three columns.
constexpr int kZBins = 13;       // 13 x 4 m cover the 50 m cube
constexpr float kZBin = 4.0f;
constexpr int kZCells = kZBins * kCubeCells;
// One index table ("slot") per indexed cube:  int hdr[16]: hdr[z] = first sorted position of z-layer z (z < 13), hdr[13] = n,
// hdr[14] = mode;  then, mode 0: unsigned short rel[13][2500] = start of every column inside its z-layer (the end of a layer's
// last column is the next layer's start);  mode 1 (a cube of more than 65 535 points: 16-bit offsets do not reach):
// int flat[2501] = column starts of a column-only order, no z-layers.
constexpr int kTabHdr = 16;
constexpr int kTabInts = kTabHdr + (kZCells + 1) / 2 + 2;   // 16 268 ints = 65 072 bytes, a multiple of 16
static_assert(kTabInts % 4 == 0 && kTabInts >= kTabHdr + kCubeCells + 1, "slot layout");
constexpr int kTabSlots = 384;   // index tables per stream and kind (>= the 125 valid cubes + the cubes one scan can touch)

struct LMState {
  double parameters[7];                 // q_w_curr (x,y,z,w), t_w_curr      laser_mapping.cpp:74-83
  double q_wmap_wodom[4], t_wmap_wodom[3];
  double q_wodom[4], t_wodom[3];
  int cenW, cenH, cenD;                 // laserCloudCenWidth / Height / Depth
  int validNum;
  int validInd[kMaxValid];
  int validPrefix[2][kMaxValid + 1];    // offsets of each valid cube inside the concatenated sub-map
  int fromMapNum[2];
  int stackNum[2];
  int solved;                           // the map was large enough (:448)
  int poolEnd[2];                       // first free slot of the current map pool (bump allocator of cube slabs)
  int cur[2];                           // which of the two pools holds this stream's corner / surf map
  int compact[2];                       // this scan's write-back re-packs the whole map into the other pool
  int repacks[2];                       // re-packs so far
  int tabEnd[2];                        // column-table slots handed out so far
  int tabLimit;                         // slots that may be handed out (kTabSlots; smaller only to exercise the recycling in tests)
  int buildNum[2], buildList[2][kMaxValid];   // valid cubes whose column index has to be (re)built before the association
  short entryNext[kMaxValid];           // next valid-list entry naming the same cube (-1: none); heads in LMDevice::entryHead
  int workNum[2];
  int workCube[2][kMaxWork], workFilter[2][kMaxWork], workNew0[2][kMaxWork], workNewN[2][kMaxWork];
  int workIn0[2][kMaxWork + 1];         // offsets of each work cube's (old ++ new) input inside the concat buffer
  int workOutN[2][kMaxWork], workFixed[2][kMaxWork];
  int workDirect[2][kMaxWork];          // the cube's new content already sits in its slab (patched in place by lm_refilter)
  int error;                            // this scan's error bits (kLmErr*), cleared by lm_prepare
  int errorEver;                        // OR of every scan's error bits since the handle was created
  int applied[2];                       // lm_place committed this scan's rewritten cubes of the kind
  // cumulative work counters since the handle was created (vloam_get_lm_counters; bench.py reports their per-scan means)
  int counters[16];
  unsigned long long knnQueries, knnCandidates;   // filled only while debug statistics are on (vloam_lidar_set_debug_stats)
  SolveTrace trace[2];
};

// Residual record produced by lm_fit for one down-sampled scan point: v = edge a(3), b(3) / plane n(3), d; p = the point;
// type 0 none, 1 edge (LidarEdgeFactor), 2 plane (LidarPlaneNormFactor).  The record of gn_split.cuh.
using LMResidual = GNResidual;

struct LMDevice {
  int B = 0, cap = 0, mapCap = 0;
  vloam_lidar_params p{};
  Profiler* prof = nullptr;
  bool allocated = false, reset_valid = true, ran = false, debugStats = false;
  int curTab = 0;                  // the map = cube slabs in mapPts[LMState::cur] addressed by the tables [curTab]
  LMState* st = nullptr;
  // Cube tables [B][2][kCubes]: a cube is the slab [off, off + cap) of its pool holding cnt points; fix = the cube is a
  // fixed point of its voxel filter (cta_voxel_filter) and has not received a point since.
  int* cubeOff[2] = {nullptr, nullptr};
  int* cubeCnt[2] = {nullptr, nullptr};
  int* cubeCap[2] = {nullptr, nullptr};
  int* cubeFix[2] = {nullptr, nullptr};
  int* cubeTab[2] = {nullptr, nullptr};   // slot of the cube's column table in tabPool, -1 = no valid index
  int* tabPool = nullptr;                 // [B][2][kTabSlots][kTabInts] index tables (layout: kTabHdr)
  short* entryHead = nullptr;             // [B][kCubes] first entry of the valid list naming this cube, -1 = not in the sub-map
  float4* mapPts[2] = {nullptr, nullptr}; // two pools [B][2][mapCap]; a stream changes pool only when its map is re-packed
  float4* snap = nullptr;                 // [B][2][mapCap] laserCloud{Corner,Surf}FromMap of the last scan (debug_keep_submap)
  float4* pubBuf = nullptr;               // staging for the clouds LaserMapping::publish sends (allocated on first use)
  int* pubCount = nullptr;
  size_t pubCap = 0;
  float4* stack = nullptr;                // [B][2][cap]  down-sampled scan (laserCloudCornerStack / SurfStack)
  float4* stackW = nullptr;               // [B][2][cap]  the same points in the map frame (pointAssociateToMap)
  unsigned* keyA = nullptr; unsigned* valA = nullptr; unsigned* keyB = nullptr; unsigned* valB = nullptr;  // [B][2][workCap]
  float4* concat = nullptr;               // [B][2][workCap] refilter inputs (old ++ new per work cube)
  float4* staged = nullptr;               // [B][2][workCap] refilter outputs
  float4* sorted = nullptr;               // [B][2][mapCap] column-sorted copy of every indexed cube at its slab's offset, w = index in the cube
  GNRecArray res{nullptr, nullptr, 0};    // [B][2] blocks of cap records (gn_split.cuh): residual blocks of the current pass
  GNState* gnState = nullptr;             // [B] wide solve (gn_split.cuh): per-stream trust-region state between launches
  double* gnPartial = nullptr;            // [B][kGnTiles][28] partial normal equations of one evaluation
  void* ncclComm = nullptr;               // point-sharded streams: the partials are all-reduced across ranks (capi.cu)
  int shardRank = 0, shardWorld = 1;      // ... and the queries are dealt round-robin to the ranks
  uint8_t* fitType = nullptr;             // [2 passes][B][2][cap] factor type per query and outer pass (vloam_get_lm_queries)
  int* cubeOf = nullptr;                  // [B][2][cap] cube id of every down-sampled scan point (map frame)
  int* nnPos = nullptr;                   // [B][2][5][cap] positions (in `sorted`) of the five nearest map points per query
  double* pose = nullptr;                 // [B][16]
  short* workOf = nullptr;                // [B][2][kCubes]
  short* liveList = nullptr;              // [B][2][kCubes] cubes a re-pack has to move
  int* liveNum = nullptr;                 // [B][2]
  size_t workCap = 0;
};

// pcl::VoxelGrid<PointXYZI> on one global-memory segment by one CTA (1024 threads).  Semantics: oracle/voxel_grid.hpp.
// Returns the number of output points (valid in all threads).
//
// *fixedPoint (uniform) is set when the OUTPUT is provably a fixed point of the filter, i.e. filtering it again would
// return it bit for bit: every centroid lies in the voxel it was averaged over.  Then the next filter sees one point
// per voxel, the same voxel lattice (floor(x / leaf) does not depend on the bounding box) and hence the same
// lexicographic key order, and the "mean" of a single point is the point itself ((0 + x) / 1 == x; sums never produce
// -0).  The map re-filter (laser_mapping.cpp:689-702) uses this to skip cubes that received no point since they were
// last filtered — the reference filters them again and gets the same cloud back.
__device__ int cta_voxel_filter(const float4* __restrict__ in, int n, float leaf, float4* __restrict__ out, unsigned* kA,
                                unsigned* vA, unsigned* kB, unsigned* vB, SortSmem& S, float* red /*[6*32]*/, int* fixedPoint) {
  *fixedPoint = 1;
  if (n == 0) return 0;
  const float inv = __fdiv_rn(1.0f, leaf);
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int k = threadIdx.x; k < n; k += 1024) {
    const float4 p = in[k];
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  const int w = threadIdx.x >> 5, l = lane_id();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if (l == 0) { red[a * 32 + w] = mn[a]; red[(3 + a) * 32 + w] = mx[a]; }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = red[a * 32], hi = red[(3 + a) * 32];
    for (int q = 1; q < 32; ++q) { lo = fminf(lo, red[a * 32 + q]); hi = fmaxf(hi, red[(3 + a) * 32 + q]); }
    mn[a] = lo; mx[a] = hi;
  }
  __syncthreads();
  const long long dx = (long long)(__fmul_rn(__fsub_rn(mx[0], mn[0]), inv)) + 1;
  const long long dy = (long long)(__fmul_rn(__fsub_rn(mx[1], mn[1]), inv)) + 1;
  const long long dz = (long long)(__fmul_rn(__fsub_rn(mx[2], mn[2]), inv)) + 1;
  if (dx * dy * dz > 2147483647LL) {  // PCL: "Leaf size is too small": output = input
    for (int k = threadIdx.x; k < n; k += 1024) out[k] = in[k];
    *fixedPoint = 0;
    return n;
  }
  int minb[3], divb[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    minb[a] = (int)floorf(__fmul_rn(mn[a], inv));
    divb[a] = (int)floorf(__fmul_rn(mx[a], inv)) - minb[a] + 1;
  }
  const int mul1 = divb[0], mul2 = divb[0] * divb[1];
  for (int k = threadIdx.x; k < n; k += 1024) {
    const float4 p = in[k];
    const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)minb[0]);
    const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)minb[1]);
    const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)minb[2]);
    kA[k] = (unsigned)(i0 + i1 * mul1 + i2 * mul2);
    vA[k] = (unsigned)k;
  }
  __syncthreads();
  const long long prod = (long long)divb[0] * divb[1] * divb[2];
  int bits = 32;
  if (prod <= 0xffffffffLL) { const unsigned mk = (unsigned)(prod - 1); bits = mk ? 32 - __clz(mk) : 1; }
  const int cur = cta_radix_sort(kA, vA, kB, vB, n, bits, S);
  const unsigned* keys = cur ? kB : kA;
  const unsigned* vals = cur ? vB : vA;
  // Segment heads -> centroid of (x, y, z, intensity), summed in ascending input order, divided by float(n).
  // Pass 1: heads per 32-entry window -> exclusive prefix; pass 2: the sorted position of every voxel's first entry;
  // pass 3: one THREAD per voxel adds its run strictly left to right (the float sum is order-dependent, so a voxel is a
  // serial chain, but the chains of 32 neighbouring voxels run side by side in a warp).  The first version walked the
  // windows with a head lane pulling its run out of the other lanes by shuffle: every window then costs as many rounds as
  // its longest run with a quarter of the lanes working — 53 % of the kernel's instructions (ncu, profiles/r02h).
  unsigned* winHeads = cur ? kA : kB;                 // the sort's spare key buffer: heads per window -> exclusive prefix
  unsigned* headPos = cur ? vA : vB;                  // the sort's spare value buffer: first sorted position of every voxel
  const int nwin = (n + 31) >> 5;
  for (int v = w; v < nwin; v += 32) {
    const int q = v * 32 + l;
    const bool head = q < n && (q == 0 || keys[q] != keys[q - 1]);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (l == 0) winHeads[v] = __popc(hm);
  }
  __syncthreads();
  {
    const int per = (nwin + 1023) / 1024;
    const int v0 = min((int)threadIdx.x * per, nwin), v1 = min(v0 + per, nwin);
    int sum = 0;
    for (int v = v0; v < v1; ++v) sum += (int)winHeads[v];
    int run = block_exclusive_scan1024(sum, S);
    for (int v = v0; v < v1; ++v) { const int t = (int)winHeads[v]; winHeads[v] = (unsigned)run; run += t; }
  }
  const int total = S.total;
  __syncthreads();
  for (int v = w; v < nwin; v += 32) {
    const int q = v * 32 + l;
    const bool head = q < n && (q == 0 || keys[q] != keys[q - 1]);
    const unsigned hm = __ballot_sync(0xffffffffu, head);
    if (head) headPos[(int)winHeads[v] + __popc(hm & ((1u << l) - 1u))] = (unsigned)q;
  }
  __syncthreads();
  int inside = 1;
  for (int v = threadIdx.x; v < total; v += 1024) {
    const int q0 = (int)headPos[v], q1 = v + 1 < total ? (int)headPos[v + 1] : n;
    const float4 first = in[vals[q0]];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    sx = __fadd_rn(sx, first.x); sy = __fadd_rn(sy, first.y); sz = __fadd_rn(sz, first.z); si = __fadd_rn(si, first.w);
    for (int q = q0 + 1; q < q1; q += 4) {             // four gathers in flight, added in order
      float4 pp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (q + u < q1) pp[u] = in[vals[q + u]];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q + u < q1) { sx = __fadd_rn(sx, pp[u].x); sy = __fadd_rn(sy, pp[u].y); sz = __fadd_rn(sz, pp[u].z); si = __fadd_rn(si, pp[u].w); }
    }
    const float nf = (float)(q1 - q0);
    const float4 c = make_float4(__fdiv_rn(sx, nf), __fdiv_rn(sy, nf), __fdiv_rn(sz, nf), __fdiv_rn(si, nf));
    // the voxel's lattice coordinates are those of its first member
    if (floorf(__fmul_rn(c.x, inv)) != floorf(__fmul_rn(first.x, inv)) || floorf(__fmul_rn(c.y, inv)) != floorf(__fmul_rn(first.y, inv)) ||
        floorf(__fmul_rn(c.z, inv)) != floorf(__fmul_rn(first.z, inv))) inside = 0;
    out[v] = c;
  }
  *fixedPoint = __syncthreads_and(inside);
  return total;
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cube_coord(double v, int cen) {  // :207-216 / :643-652
  int c = (int)((v + 25.0) / 50.0) + cen;
  if (v + 25.0 < 0) c--;
  return c;
}

struct CubeTables { int* off; int* cnt; int* cap; int* fix; int* tab; };   // [B][2][kCubes] each
struct MapPools { float4* p[2]; };                                // [B][2][mapCap] each
__device__ __forceinline__ float4* stream_map(const MapPools& pools, const LMState& st, int b, int kind, int mapCap) {
  return (st.cur[kind] ? pools.p[1] : pools.p[0]) + ((size_t)b * 2 + kind) * mapCap;
}

// lm_prepare: grid (B), block kPrepThreads.  Table ping-pong: reads tables `src`, writes the shifted tables to `dst`.
// Also: the valid-cube list, the cube -> valid-entry lists the association walks, and column-table slots for the valid
// cubes that have no usable column index yet (built next by lm_index_build).
// The kernel is a chain of short dependent steps on one stream's tables, i.e. latency: the table copy runs 1024 wide, the
// per-valid-cube look-ups are fetched by one thread per (kind, entry) into shared memory, and only the two order-dependent
// loops (prefix of the sub-map, slot hand-out) stay serial — on shared memory.  (The first version did the look-ups in the
// serial loops: ~300 dependent global loads, two thirds of the kernel's 67 us.)
constexpr int kPrepThreads = 1024;
__global__ void __launch_bounds__(kPrepThreads) lm_prepare(LMState* __restrict__ stAll, const LOState* __restrict__ lo,
                                                            const CubeTables src, const CubeTables dst, short* __restrict__ entryHeadAll,
                                                            int resetValid) {
  const int b = blockIdx.x;
  LMState& st = stAll[b];
  short* entryHead = entryHeadAll + (size_t)b * kCubes;
  __shared__ int sh[3];
  __shared__ int s_gc[2], s_vn, s_dup, s_total[2];
  __shared__ int s_vind[kMaxValid], s_cnt[2][kMaxValid], s_tab[2][kMaxValid];
  __shared__ unsigned char s_head[kMaxValid];
  if (threadIdx.x == 0) {
    if (resetValid) st.validNum = 0;  // LaserMapping::reset (:127-131)
    st.error = 0;                     // error bits describe one scan (the sticky copy is errorEver)
    // input(), :182-195: q_w_curr = q_wmap_wodom * q_wodom_curr; t_w_curr = q_wmap_wodom * t_wodom_curr + t_wmap_wodom
    for (int i = 0; i < 4; ++i) st.q_wodom[i] = lo[b].q_w[i];
    for (int i = 0; i < 3; ++i) st.t_wodom[i] = lo[b].t_w[i];
    double q[4], t[3];
    quat_mul(st.q_wmap_wodom, st.q_wodom, q);
    quat_rotate(st.q_wmap_wodom, st.t_wodom[0], st.t_wodom[1], st.t_wodom[2], t);
    for (int i = 0; i < 4; ++i) st.parameters[i] = q[i];
    for (int i = 0; i < 3; ++i) st.parameters[4 + i] = t[i] + st.t_wmap_wodom[i];
    // :207-216
    int cI = cube_coord(st.parameters[4], st.cenW), cJ = cube_coord(st.parameters[5], st.cenH), cK = cube_coord(st.parameters[6], st.cenD);
    // :218-402: each while-iteration moves every cube one step and clears the plane that wrapped around
    int sI = 0, sJ = 0, sK = 0;
    while (cI < 3) { cI++; st.cenW++; sI++; }
    while (cI >= kCubeW - 3) { cI--; st.cenW--; sI--; }
    while (cJ < 3) { cJ++; st.cenH++; sJ++; }
    while (cJ >= kCubeH - 3) { cJ--; st.cenH--; sJ--; }
    while (cK < 3) { cK++; st.cenD++; sK++; }
    while (cK >= kCubeD - 3) { cK--; st.cenD--; sK--; }
    sh[0] = sI; sh[1] = sJ; sh[2] = sK;
    // :404-420 valid cubes in the reference's loop order
    int vn = st.validNum;
    const int vn0 = vn;
    for (int i = cI - 2; i <= cI + 2; i++)
      for (int j = cJ - 2; j <= cJ + 2; j++)
        for (int k = cK - 1; k <= cK + 1; k++)
          if (i >= 0 && i < kCubeW && j >= 0 && j < kCubeH && k >= 0 && k < kCubeD && vn < kMaxValid)
            st.validInd[vn++] = i + kCubeW * j + kCubeW * kCubeH * k;
    st.validNum = vn;
    s_vn = vn;
    for (int v = 0; v < vn; ++v) s_vind[v] = st.validInd[v];
    // cube -> entries of the valid list naming it, ascending.  One entry per cube unless the caller skipped reset()
    // (the list then keeps growing and names cubes twice); only that case needs the chained insertion below.
    for (int v = 0; v < vn; ++v) { st.entryNext[v] = -1; s_head[v] = 1; }
    s_dup = vn0 > 0 ? 1 : 0;
  }
  __syncthreads();
  const int sI = sh[0], sJ = sh[1], sK = sh[2];
  for (int kind = 0; kind < 2; ++kind) {
    const size_t tb = ((size_t)b * 2 + kind) * kCubes;
    for (int c = threadIdx.x; c < kCubes; c += kPrepThreads) {
      const int i = c % kCubeW, j = (c / kCubeW) % kCubeH, k = c / (kCubeW * kCubeH);
      const int si = i - sI, sj = j - sJ, sk = k - sK;  // new[i] = old[i - shift]; wrapped planes are cleared
      int o = 0, n = 0, cp = 0, fx = 0, tbs = -1;        // (a cleared cube's slab and table slot are reclaimed later)
      if (si >= 0 && si < kCubeW && sj >= 0 && sj < kCubeH && sk >= 0 && sk < kCubeD) {
        const int s = si + kCubeW * sj + kCubeW * kCubeH * sk;
        o = src.off[tb + s]; n = src.cnt[tb + s]; cp = src.cap[tb + s]; fx = src.fix[tb + s]; tbs = src.tab[tb + s];
      }
      dst.off[tb + c] = o; dst.cnt[tb + c] = n; dst.cap[tb + c] = cp; dst.fix[tb + c] = fx; dst.tab[tb + c] = n > 0 ? tbs : -1;
    }
  }
  for (int c = threadIdx.x; c < kCubes; c += kPrepThreads) entryHead[c] = -1;
  __syncthreads();
  const int vn = s_vn;
  if (threadIdx.x == 0) {
    if (s_dup) {   // duplicates possible: literal chained insertion (head = the smallest entry naming the cube)
      for (int v = vn - 1; v >= 0; --v) { const int c = s_vind[v]; st.entryNext[v] = entryHead[c]; entryHead[c] = (short)v; }
      for (int v = 0; v < vn; ++v) s_head[v] = entryHead[s_vind[v]] == v ? 1 : 0;
    } else {
      for (int v = 0; v < vn; ++v) entryHead[s_vind[v]] = (short)v;
    }
  }
  // the shifted tables' rows of the valid cubes, one thread per (kind, entry)
  if (threadIdx.x < 2 * kMaxValid) {
    const int kind = threadIdx.x / kMaxValid, v = threadIdx.x % kMaxValid;
    if (v < vn) {
      const size_t tb = ((size_t)b * 2 + kind) * kCubes;
      s_cnt[kind][v] = dst.cnt[tb + s_vind[v]];
      s_tab[kind][v] = dst.tab[tb + s_vind[v]];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int kind = threadIdx.x;
    int acc = 0, need = 0;
    for (int v = 0; v < vn; ++v) {
      st.validPrefix[kind][v] = acc; acc += s_cnt[kind][v];
      if (s_head[v] && s_cnt[kind][v] > 0 && s_tab[kind][v] < 0) ++need;
    }
    st.validPrefix[kind][vn] = acc;
    st.fromMapNum[kind] = acc;
    s_gc[kind] = st.tabEnd[kind] + need > st.tabLimit ? 1 : 0;   // out of table slots: drop every index of this kind, start over
    s_total[kind] = acc;
  }
  __syncthreads();
  for (int kind = 0; kind < 2; ++kind)
    if (s_gc[kind]) for (int c = threadIdx.x; c < kCubes; c += kPrepThreads) dst.tab[((size_t)b * 2 + kind) * kCubes + c] = -1;
  __syncthreads();
  if (threadIdx.x < 2) {
    const int kind = threadIdx.x;
    const size_t tb = ((size_t)b * 2 + kind) * kCubes;
    const int solved = (s_total[0] > 10 && s_total[1] > 50) ? 1 : 0;  // :448
    int te = s_gc[kind] ? 0 : st.tabEnd[kind], nb = 0;
    if (solved)
      for (int v = 0; v < vn; ++v) {
        if (!s_head[v] || s_cnt[kind][v] == 0 || (!s_gc[kind] && s_tab[kind][v] >= 0)) continue;
        const int c = s_vind[v];
        dst.tab[tb + c] = te++;               // <= kMaxValid new slots after a reset of the slot counter: always fits
        st.buildList[kind][nb++] = c;
      }
    st.tabEnd[kind] = te;
    st.buildNum[kind] = nb;
    st.counters[kCntIndexedCorner + kind] += nb;
    if (kind == 0) {
      st.solved = solved;
      st.trace[0].n_records = st.trace[1].n_records = 0;
      st.trace[0].n_corner = st.trace[0].n_plane = st.trace[1].n_corner = st.trace[1].n_plane = 0;
      st.counters[kCntScans]++; st.counters[kCntSolved] += solved;
    }
  }
}

// lm_voxel_stack: grid (2, B), block 1024.  VoxelGrid of the scan's corner (lineRes) / surf (planeRes) features (:432-440).
__global__ void __launch_bounds__(1024) lm_voxel_stack(LMState* __restrict__ stAll, const SRHeader* __restrict__ hdr,
                                                        const float4* __restrict__ cornerLast, const float4* __restrict__ surfLast,
                                                        int cap, float lineRes, float planeRes, float4* __restrict__ stack,
                                                        unsigned* kA, unsigned* vA, unsigned* kB, unsigned* vB, size_t workCap) {
  __shared__ SortSmem S;
  __shared__ float red[6 * 32];
  const int kind = blockIdx.x, b = blockIdx.y;
  const float4* in = kind == 0 ? cornerLast + (size_t)b * kMaxLessSharp : surfLast + (size_t)b * cap;
  const int n = kind == 0 ? hdr[b].nLessSharp : hdr[b].nLessFlat;
  const size_t so = ((size_t)b * 2 + kind) * workCap;
  int fixedPoint;
  const int m = cta_voxel_filter(in, n, kind == 0 ? lineRes : planeRes, stack + ((size_t)b * 2 + kind) * cap, kA + so, vA + so,
                                 kB + so, vB + so, S, red, &fixedPoint);
  if (threadIdx.x == 0) stAll[b].stackNum[kind] = m;
}

// ---------------------------------------------------------------------------------------------------------------
// Sub-map column index.  Sub-map index g (the position in laserCloudCornerFromMap / SurfFromMap, :422-428) maps to
// valid cube v = upper_bound(validPrefix, g) - 1 and offset g - validPrefix[v] inside that cube.
__device__ __forceinline__ float4 submap_point(const LMState& st, int kind, const int* __restrict__ off,
                                               const float4* __restrict__ map, int g) {
  int lo = 0, hi = st.validNum;  // find v with prefix[v] <= g < prefix[v + 1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (st.validPrefix[kind][mid] <= g) lo = mid; else hi = mid; }
  return map[off[st.validInd[lo]] + (g - st.validPrefix[kind][lo])];
}
// ---------------------------------------------------------------------------------------------------------------
// Column index of one cube: the cube's points counting-sorted into 50 x 50 xy-columns of 1.001 m (origin = the cube's
// min corner), kept beside the cube until the cube's content changes.  Replaces the kd-trees the reference rebuilds
// over the whole sub-map on every scan (:452-453): here only the cubes a scan rewrites are re-indexed.
// A match needs its 5th neighbour within 1 m (:479, :547) and |dx| < 1 m moves a point by at most one 1.001 m column
// (the 0.1 % margin dominates the float rounding of the column coordinate), so the 3 x 3 column block around the query
// in every cube the 1.001 m box around the query touches holds every point that can matter: one visit, exact.
__device__ __forceinline__ float cube_min_coord(int idx, int cen) { return (float)((idx - cen) * 50.0 - 25.0); }
__device__ __forceinline__ int cube_cell(float v, float mn) { return (int)floorf((v - mn) * (1.0f / kCubeCell)); }
__device__ __forceinline__ int cube_cell_clamped(float v, float mn) { return min(max(cube_cell(v, mn), 0), kCubeCellsX - 1); }
__device__ __forceinline__ int cube_zbin(float z, float mnz) { return min(max((int)floorf((z - mnz) * (1.0f / kZBin)), 0), kZBins - 1); }
// One CTA of NT threads (a multiple of 32, <= 1024) builds the index of one cube.  pts[0..n): the cube's slab; tab[kTabInts]
// (global): the cube's table slot; sortedOut[0..n): the copy ordered by (z-layer, column) with w = index inside the cube.
// Shared memory (dynamic, kIndexSmemBytes): one 16-bit counter per (z-layer, column) cell, two to a word — first the
// histogram, then the running position inside the cell, advanced by shared-memory atomics (a half cannot overflow into its
// neighbour: the whole cube holds at most 65 535 points in this mode) — plus the layer starts.  A larger cube is indexed by
// column only, with 32-bit counters in the same shared memory (mode 1).
constexpr int kIndexSmemBytes = ((kZCells + 1) / 2 + kZBins + 3 + 32) * (int)sizeof(int);
template <int NT>
__device__ void cta_build_cube_index(const float4* __restrict__ pts, int n, float minX, float minY, float minZ, int* __restrict__ tab,
                                     float4* __restrict__ sortedOut, int* smem) {
  const int l = lane_id(), w = threadIdx.x >> 5;
  auto col_of = [&](const float4& p) { return cube_cell_clamped(p.y, minY) * kCubeCellsX + cube_cell_clamped(p.x, minX); };
  if (n > 65535) {
    // ---- mode 1: column-only index, 32-bit counters
    int* s_col = smem;                      // [kCubeCells + 1]
    int* s_w = smem + kCubeCells + 1;       // [32]
    int* flat = tab + kTabHdr;
    for (int i = threadIdx.x; i <= kCubeCells; i += NT) s_col[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) atomicAdd(&s_col[col_of(pts[i])], 1);
    __syncthreads();
    constexpr int per = (kCubeCells + NT - 1) / NT;
    const int c0 = min((int)threadIdx.x * per, kCubeCells), c1 = min(c0 + per, kCubeCells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += s_col[c];
    int sc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (l >= o) sc += t; }
    if (l == 31) s_w[w] = sc;
    __syncthreads();
    int run = sc - sum;
    for (int q = 0; q < w; ++q) run += s_w[q];
    for (int c = c0; c < c1; ++c) { const int t = s_col[c]; s_col[c] = run; flat[c] = run; run += t; }
    if (threadIdx.x == NT - 1) flat[kCubeCells] = n;
    if (threadIdx.x < kTabHdr) tab[threadIdx.x] = threadIdx.x == 0 ? 0 : threadIdx.x == 14 ? 1 : n;   // one "layer"; mode 1
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += NT) {
      const float4 p = pts[i];
      sortedOut[atomicAdd(&s_col[col_of(p)], 1)] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
    __syncthreads();
    return;
  }
  // ---- mode 0: (z-layer, column) cells, packed 16-bit counters
  unsigned* s_pack = reinterpret_cast<unsigned*>(smem);            // [(kZCells + 1) / 2]
  int* s_layer = smem + (kZCells + 1) / 2;                         // [kZBins + 1] layer starts
  int* s_w = s_layer + kZBins + 3;                                 // [32]
  auto cell_of = [&](const float4& p) { return cube_zbin(p.z, minZ) * kCubeCells + col_of(p); };
  for (int i = threadIdx.x; i < (kZCells + 1) / 2; i += NT) s_pack[i] = 0u;
  __syncthreads();
  // (four independent loads in flight per thread: the passes over the cube are latency bound otherwise)
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * NT) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * NT; if (i < n) p[u] = pts[i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * NT < n) { const int c = cell_of(p[u]); atomicAdd(&s_pack[c >> 1], (c & 1) ? 0x10000u : 1u); }
  }
  __syncthreads();
  // Column starts relative to the layer start = an exclusive scan inside every z-layer: one warp per layer (13 of the
  // warps), a lane per word of a 32-word row (consecutive lanes -> consecutive banks), one pass.  The layer totals fall out
  // of the same pass; their prefix gives the layer starts the scatter and the table header need.
  constexpr int kLayerWords = kCubeCells / 2;                      // 1250 words of two 16-bit counters per layer
  constexpr int kRows = (kLayerWords + 31) / 32;
  static_assert(NT / 32 >= kZBins, "one warp per z-layer");
  if (w < kZBins) {
    const int w0 = w * kLayerWords;
    int carry = 0;
    for (int r = 0; r < kRows; ++r) {
      const int jj = r * 32 + l;
      const unsigned v = jj < kLayerWords ? s_pack[w0 + jj] : 0u;
      const int lo = (int)(v & 0xffffu), cnt = lo + (int)(v >> 16);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += t; }
      if (jj < kLayerWords) {
        const int start = carry + incl - cnt;                      // both cells of a word lie in one layer (2500 is even)
        s_pack[w0 + jj] = (unsigned)start | ((unsigned)(start + lo) << 16);
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (l == 0) s_w[w] = carry;                                    // points in layer w
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int z = 0; z < kZBins; ++z) { s_layer[z] = run; run += s_w[z]; }
    s_layer[kZBins] = run;                                         // == n
  }
  __syncthreads();
  if (threadIdx.x < kTabHdr) tab[threadIdx.x] = threadIdx.x <= kZBins ? s_layer[threadIdx.x] : 0;   // hdr[13] = n, hdr[14] = mode 0
  // the table goes to global memory as it sits in shared memory: coalesced 32-bit words
  {
    unsigned* relw = reinterpret_cast<unsigned*>(tab + kTabHdr);
    for (int i = threadIdx.x; i < (kZCells + 1) / 2; i += NT) relw[i] = s_pack[i];
  }
  __syncthreads();
  // scatter: the same counters now run from the column's start; the atomic returns a point's position inside its layer
  // (start + points of the column placed so far <= size of the layer <= 65 535: a half never carries into its neighbour)
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * NT) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * NT; if (i < n) p[u] = pts[i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * NT;
      if (i < n) {
        const int c = cell_of(p[u]);
        const unsigned old = atomicAdd(&s_pack[c >> 1], (c & 1) ? 0x10000u : 1u);
        const int within = (int)((c & 1) ? (old >> 16) : (old & 0xffffu));
        sortedOut[s_layer[c / kCubeCells] + within] = make_float4(p[u].x, p[u].y, p[u].z, __int_as_float(i));
      }
    }
  }
  __syncthreads();
}
// lm_index_build: grid (32, 2, B), block kIndexThreads: CTAs walk lm_prepare's build list.
constexpr int kIndexThreads = 512;
__global__ void __launch_bounds__(kIndexThreads, 2) lm_index_build(const LMState* __restrict__ stAll, const CubeTables T, const MapPools pools,
                                                       int mapCap, int* __restrict__ tabPool, float4* __restrict__ sorted) {
  extern __shared__ int idx_smem[];
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  for (int u = blockIdx.x; u < st.buildNum[kind]; u += gridDim.x) {
    const int c = st.buildList[kind][u];
    const size_t tb = ((size_t)b * 2 + kind) * kCubes;
    const int off = T.off[tb + c], n = T.cnt[tb + c], slot = T.tab[tb + c];
    cta_build_cube_index<kIndexThreads>(stream_map(pools, st, b, kind, mapCap) + off, n, cube_min_coord(c % kCubeW, st.cenW),
                                        cube_min_coord((c / kCubeW) % kCubeH, st.cenH), cube_min_coord(c / (kCubeW * kCubeH), st.cenD),
                                        tabPool + (((size_t)b * 2 + kind) * kTabSlots + slot) * kTabInts,
                                        sorted + ((size_t)b * 2 + kind) * mapCap + off, idx_smem);
  }
}

// laserCloudCornerFromMap / SurfFromMap (:422-428) kept for inspection (debug_keep_submap): grid (nblk, 2, B), block 256
__global__ void __launch_bounds__(256) lm_snapshot_submap(const LMState* __restrict__ stAll, const int* __restrict__ cubeOff,
                                                           const MapPools pools, int mapCap, float4* __restrict__ snap) {
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  const float4* map = stream_map(pools, st, b, kind, mapCap);
  for (int g = blockIdx.x * 256 + threadIdx.x; g < st.fromMapNum[kind]; g += gridDim.x * 256)
    snap[((size_t)b * 2 + kind) * mapCap + g] = submap_point(st, kind, cubeOff + ((size_t)b * 2 + kind) * kCubes, map, g);
}

// ---------------------------------------------------------------------------------------------------------------
// Small dense kernels of the association (same algorithms as oracle/small_linalg.hpp).
// Both are written with compile-time indices only (unrolled loops, conditional swaps instead of index arrays) so that
// the small matrices live in registers: lm_fit runs them once per thread.
__device__ __forceinline__ void sym_eig3_dev(const double A[9], double evals[3], double evecs[3][3]) {
  double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-300 || off <= 1e-32 * diag) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
      }
  }
  // ascending eigenvalues: the bubble sort of oracle/small_linalg.hpp on (value, column) pairs instead of an index array
  double e[3] = {a[0][0], a[1][1], a[2][2]};
  double col[3][3];   // col[k][r] = v[r][k]
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int r = 0; r < 3; ++r) col[k][r] = v[r][k];
  auto cswap = [&](int i, int j) {   // compile-time i, j after inlining
    if (e[j] < e[i]) {
      const double te = e[i]; e[i] = e[j]; e[j] = te;
#pragma unroll
      for (int r = 0; r < 3; ++r) { const double tc = col[i][r]; col[i][r] = col[j][r]; col[j][r] = tc; }
    }
  };
  cswap(0, 1); cswap(1, 2); cswap(0, 1);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    evals[k] = e[k];
    double n = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) n += col[k][r] * col[k][r];
    n = sqrt(n);
#pragma unroll
    for (int r = 0; r < 3; ++r) evecs[k][r] = col[k][r] / n;
  }
}
__device__ __forceinline__ void colpiv_qr_solve_5x3_dev(const double Ain[15], const double bin[5], double x[3]) {
  constexpr int M = 5;
  double A[M][3], b[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    b[i] = bin[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) A[i][c] = Ain[i * 3 + c];
  }
  int perm[3] = {0, 1, 2};
  double maxnorm2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) s += A[i][c] * A[i][c];
    maxnorm2 = fmax(maxnorm2, s);
  }
  const double eps = 2.220446049250313e-16;
  const double thresh = maxnorm2 * (eps / M) * (eps / M);
  int rank = 0;
  bool active = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!active) continue;
    int best = k; double bestn = -1.0;
#pragma unroll
    for (int c = k; c < 3; ++c) {
      double s = 0;
#pragma unroll
      for (int i = k; i < M; ++i) s += A[i][c] * A[i][c];
      if (s > bestn) { bestn = s; best = c; }
    }
    if (bestn <= thresh) { active = false; continue; }
#pragma unroll
    for (int c = k + 1; c < 3; ++c)
      if (best == c) {   // swap columns k and c (compile-time indices)
#pragma unroll
        for (int i = 0; i < M; ++i) { const double t = A[i][k]; A[i][k] = A[i][c]; A[i][c] = t; }
        const int t = perm[k]; perm[k] = perm[c]; perm[c] = t;
      }
    const double nrm = sqrt(bestn);
    const double alpha = A[k][k] > 0 ? -nrm : nrm;
    double v[M];
#pragma unroll
    for (int i = 0; i < M; ++i) v[i] = i >= k ? A[i][k] : 0.0;
    v[k] -= alpha;
    double vn = 0;
#pragma unroll
    for (int i = k; i < M; ++i) vn += v[i] * v[i];
    if (vn > 0) {
#pragma unroll
      for (int c = k; c < 3; ++c) {
        double s = 0;
#pragma unroll
        for (int i = k; i < M; ++i) s += v[i] * A[i][c];
        s = 2.0 * s / vn;
#pragma unroll
        for (int i = k; i < M; ++i) A[i][c] -= s * v[i];
      }
      double s = 0;
#pragma unroll
      for (int i = k; i < M; ++i) s += v[i] * b[i];
      s = 2.0 * s / vn;
#pragma unroll
      for (int i = k; i < M; ++i) b[i] -= s * v[i];
    }
    ++rank;
  }
  double y[3] = {0, 0, 0};
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    if (k >= rank) continue;
    double s = b[k];
#pragma unroll
    for (int c = k + 1; c < 3; ++c) if (c < rank) s -= A[k][c] * y[c];
    y[k] = s / A[k][k];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) if (perm[k] == j) x[j] = y[k];
}

// Association (:472-581) in two kernels, so that each runs at its own register budget / occupancy:
//   lm_knn  exact 5-NN, one 8-lane group per down-sampled scan point (4 points per warp): the candidates of the 3 x 3
//           column block (in every cube the query's 1.001 m box touches) are dealt to the lanes, every lane keeps its own
//           sorted top-5, the group merges them; writes the five positions (or -1: fifth neighbour not within 1 m).
//   lm_fit  one THREAD per point: PCA line test (corner) or least-squares plane fit + 0.2 m check (surf) on the five
//           neighbours, in double like the reference.  (Done by whole warps this part ran 32 times redundantly.)
// lm_knn: grid (kKnnGrid, 2, B), block 128: one THREAD per down-sampled scan point, grid-stride.
// Candidates of a query = per cube its 1.001 m box touches (one, unless the query sits at a cube border), per z-layer its
// +-1.001 m interval touches (<= 2), per column row (<= 3): the run of sorted positions covering its three columns.  On
// the benchmark map that is ~25 candidates per query instead of the 142 of a z-blind 3 x 3 column block, so the search is a
// short serial scan: a thread per query needs no cross-lane merge at all.  (Round 1 and the first z-layered version used
// an 8-lane group per query: ncu counted ~390 warp instructions per query, two thirds of them the per-query set-up,
// shuffles and the five-round group merge, each executed by all 8 lanes.)
constexpr int kKnnGrid = 48, kKnnThreads = 128;
template <bool STATS>
__device__ __forceinline__ void lm_knn_body(const LMState* __restrict__ stAll, const float4* __restrict__ stack, int cap,
                                            const CubeTables& T, const short* __restrict__ entryHeadAll,
                                            const int* __restrict__ tabPool, const float4* __restrict__ sorted,
                                            int mapCap, int* __restrict__ nnPos /*[B][2][5][cap]*/, LMState* statsOut, int shardRank,
                                            int shardWorld) {
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  if (!st.solved) return;
  const int nq = st.stackNum[kind];
  const size_t tb = ((size_t)b * 2 + kind) * kCubes;
  const short* entryHead = entryHeadAll + (size_t)b * kCubes;
  const int* tabs = tabPool + ((size_t)b * 2 + kind) * kTabSlots * kTabInts;
  const float4* S = sorted + ((size_t)b * 2 + kind) * mapCap;
  const int cenW = st.cenW, cenH = st.cenH, cenD = st.cenD;
  const float4* stk = stack + ((size_t)b * 2 + kind) * cap;
  int* outPos = nnPos + ((size_t)b * 2 + kind) * cap * 5;
  const double qq[4] = {st.parameters[0], st.parameters[1], st.parameters[2], st.parameters[3]};
  const double t0 = st.parameters[4], t1 = st.parameters[5], t2 = st.parameters[6];
  unsigned nCand = 0, nQuer = 0;      // debug statistics (only accumulated into global memory when STATS)
  for (int qi = blockIdx.x * kKnnThreads + threadIdx.x; qi < nq; qi += gridDim.x * kKnnThreads) {
    unsigned long long bk[5];
    int bp[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { bk[i] = 0xffffffffffffffffull; bp[i] = -1; }
    unsigned wbits = 0x3f7fffffu;   // the largest float below 1.0f (non-negative floats order like their bit patterns)
    if (shardWorld > 1 && qi % shardWorld != shardRank) {     // point-sharded: another rank's query (no factor on this rank)
#pragma unroll
      for (int i = 0; i < 5; ++i) outPos[(size_t)i * cap + qi] = -1;
      continue;
    }
    const float4 po = stk[qi];
    // pointAssociateToMap (:146-155): double transform, rounded to float
    double w[3];
    quat_rotate(qq, (double)po.x, (double)po.y, (double)po.z, w);
    const float sx = (float)(w[0] + t0), sy = (float)(w[1] + t1), sz = (float)(w[2] + t2);
    // Only a candidate closer than 1 m can matter (a query whose 5th neighbour is not within 1 m is dropped, :479 / :547,
    // and then every true neighbour is), and none farther than the current 5th best.  The distance is
    // (dx^2 + dy^2) + dz^2 in float: never below dz^2, so the z term alone prunes first.
    // (Tried and rejected: noting passing candidates in a per-thread shared-memory list and forming the top-5 at the end of
    // the query, to keep the ~45-instruction insertion out of the candidate loop: 185 us instead of 122 — without the
    // shrinking threshold far more candidates reach the distance computation and the list.)
    auto consider = [&](const float4 tp, unsigned gBase, int pos) {
      const float dz = __fsub_rn(sz, tp.z);
      if (__float_as_uint(__fmul_rn(dz, dz)) > wbits) return;
      const float d = sqdist_f(sx, sy, sz, tp.x, tp.y, tp.z);
      if (__float_as_uint(d) > wbits) return;
      unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (gBase + (unsigned)__float_as_int(tp.w));
      if (key < bk[4]) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (key < bk[i]) { const unsigned long long tk = bk[i]; bk[i] = key; key = tk; const int tq = bp[i]; bp[i] = pos; pos = tq; }
        wbits = min(wbits, (unsigned)(bk[4] >> 32));
      }
    };
    // Every cube the 1.001 m box around the query touches.  A superset is harmless (a cube that holds nothing near the query
    // contributes empty or far-away runs), so the cube range is computed in float with a margin that dominates the rounding:
    // floor((v + 25) / 50) + cen, the cube coordinate of laser_mapping.cpp:643-652 for in-range values.
    const int ci0 = max((int)floorf((sx + 23.995f) * 0.02f) + cenW, 0), ci1 = min((int)floorf((sx + 26.005f) * 0.02f) + cenW, kCubeW - 1);
    const int cj0 = max((int)floorf((sy + 23.995f) * 0.02f) + cenH, 0), cj1 = min((int)floorf((sy + 26.005f) * 0.02f) + cenH, kCubeH - 1);
    const int ck0 = max((int)floorf((sz + 23.995f) * 0.02f) + cenD, 0), ck1 = min((int)floorf((sz + 26.005f) * 0.02f) + cenD, kCubeD - 1);
    for (int ck = ck0; ck <= ck1; ++ck)
      for (int cj = cj0; cj <= cj1; ++cj)
        for (int ci = ci0; ci <= ci1; ++ci) {
          const int c = ci + kCubeW * cj + kCubeW * kCubeH * ck;
          int e = entryHead[c];
          if (e < 0) continue;                       // not part of the sub-map (:404-420)
          const int slot = T.tab[tb + c];
          if (slot < 0) continue;                    // empty cube
          const int* tab = tabs + (size_t)slot * kTabInts;
          const int off = T.off[tb + c];
          const float mnx = cube_min_coord(ci, cenW), mny = cube_min_coord(cj, cenH), mnz = cube_min_coord(ck, cenD);
          const int qx = cube_cell(sx, mnx), qy = cube_cell(sy, mny);
          const int x0 = max(qx - 1, 0), x1 = min(qx + 1, kCubeCellsX - 1);
          const int r0 = max(qy - 1, 0), r1 = min(qy + 1, kCubeCellsX - 1);
          if (x0 > x1 || r0 > r1) continue;
          // the z-layers the +-1.001 m interval touches (the clamped bin function of the build is monotone, so every point
          // within 1 m in z lies in one of them); a cube in mode 1 has a single layer
          const bool flat = tab[14] != 0;
          const int zb0 = flat ? 0 : cube_zbin(sz - 1.001f, mnz), zb1 = flat ? 0 : cube_zbin(sz + 1.001f, mnz);
          // ... once per entry of the valid list naming it: the sub-map index (the tie-break of the k-NN order) of a
          // point = offset of that entry's copy of the cube in laserCloud*FromMap + index inside the cube
          for (; e >= 0; e = st.entryNext[e]) {
            const unsigned gBase = (unsigned)st.validPrefix[kind][e];
            for (int z = zb0; z <= zb1; ++z) {
              const int L = flat ? 0 : tab[z];
              const unsigned short* rel = reinterpret_cast<const unsigned short*>(tab + kTabHdr) + z * kCubeCells;
              for (int row = r0; row <= r1; ++row) {
                int ra, re;
                if (flat) {
                  ra = tab[kTabHdr + row * kCubeCellsX + x0]; re = tab[kTabHdr + row * kCubeCellsX + x1 + 1];
                } else {
                  ra = L + (int)rel[row * kCubeCellsX + x0];
                  re = (row == kCubeCellsX - 1 && x1 == kCubeCellsX - 1) ? tab[z + 1] : L + (int)rel[row * kCubeCellsX + x1 + 1];   // a layer ends where the next begins
                }
                if (STATS) nCand += (unsigned)(re - ra);
                // two candidates in flight
                int t = ra;
                for (; t + 1 < re; t += 2) {
                  const float4 p0 = S[off + t], p1 = S[off + t + 1];
                  consider(p0, gBase, off + t);
                  consider(p1, gBase, off + t + 1);
                }
                if (t < re) consider(S[off + t], gBase, off + t);
              }
            }
          }
        }
    if (STATS) ++nQuer;
    const bool ok = bk[4] != 0xffffffffffffffffull && (double)__uint_as_float((unsigned)(bk[4] >> 32)) < 1.0;  // :479 / :547
#pragma unroll
    for (int i = 0; i < 5; ++i) outPos[(size_t)i * cap + qi] = ok ? bp[i] : -1;
  }
  if (STATS && statsOut != nullptr) {
    nCand = __reduce_add_sync(__activemask(), nCand); nQuer = __reduce_add_sync(__activemask(), nQuer);
    if (lane_id() == 0 && nQuer) { atomicAdd(&statsOut[b].knnCandidates, (unsigned long long)nCand); atomicAdd(&statsOut[b].knnQueries, (unsigned long long)nQuer); }
  }
}
#define VB_LM_KNN_ARGS                                                                                                            \
  const LMState *__restrict__ stAll, const float4 *__restrict__ stack, int cap, const CubeTables T,                              \
      const short *__restrict__ entryHeadAll, const int *__restrict__ tabPool, const float4 *__restrict__ sorted, int mapCap,    \
      int *__restrict__ nnPos, LMState *statsOut, int shardRank, int shardWorld
// Measured on B200 (64 streams per launch): this body, runs walked one after the other with two candidates in flight, 80
// registers: 127 us; all run bounds loaded up front + four candidates in flight: 161 us at 80 registers (spills), 147 us at 126.
// The occ4 variant stays selectable (VLOAM_LM_KNN_OCC=4).
__global__ void __launch_bounds__(kKnnThreads, 6) lm_knn(VB_LM_KNN_ARGS) { lm_knn_body<false>(stAll, stack, cap, T, entryHeadAll, tabPool, sorted, mapCap, nnPos, statsOut, shardRank, shardWorld); }
__global__ void __launch_bounds__(kKnnThreads, 4) lm_knn_occ4(VB_LM_KNN_ARGS) { lm_knn_body<false>(stAll, stack, cap, T, entryHeadAll, tabPool, sorted, mapCap, nnPos, statsOut, shardRank, shardWorld); }
__global__ void __launch_bounds__(kKnnThreads) lm_knn_stats(VB_LM_KNN_ARGS) { lm_knn_body<true>(stAll, stack, cap, T, entryHeadAll, tabPool, sorted, mapCap, nnPos, statsOut, shardRank, shardWorld); }
// grid (nblk, 2, B), block 128: one thread per point
__global__ void __launch_bounds__(128) lm_fit(LMState* __restrict__ stAll, const float4* __restrict__ stack, int cap,
                                               const float4* __restrict__ sorted, int mapCap, const int* __restrict__ nnPos,
                                               const GNRecArray res, uint8_t* __restrict__ fitType /*[B][2][cap] of this pass*/, int pass) {
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  if (!st.solved) return;
  int nFactors = 0;
  const int nq = st.stackNum[kind];
  const float4* S = sorted + ((size_t)b * 2 + kind) * mapCap;
  const float4* stk = stack + ((size_t)b * 2 + kind) * cap;
  const int* inPos = nnPos + ((size_t)b * 2 + kind) * cap * 5;
  for (int qi = blockIdx.x * 128 + threadIdx.x; qi < nq; qi += gridDim.x * 128) {
    const float4 po = stk[qi];
    LMResidual R;
    R.type = 0; R.px = po.x; R.py = po.y; R.pz = po.z;
#pragma unroll
    for (int i = 0; i < 7; ++i) R.v[i] = 0.0;
    int pos[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) pos[j] = inPos[(size_t)j * cap + qi];
    if (pos[4] >= 0) {
      double P[5][3];
#pragma unroll
      for (int j = 0; j < 5; ++j) { const float4 tp = S[pos[j]]; P[j][0] = tp.x; P[j][1] = tp.y; P[j][2] = tp.z; }
      if (kind == 0) {  // :481-516
        double c[3] = {0, 0, 0};
        for (int j = 0; j < 5; ++j) { c[0] = c[0] + P[j][0]; c[1] = c[1] + P[j][1]; c[2] = c[2] + P[j][2]; }
        c[0] = c[0] / 5.0; c[1] = c[1] / 5.0; c[2] = c[2] / 5.0;
        double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < 5; ++j) {
          const double d[3] = {P[j][0] - c[0], P[j][1] - c[1], P[j][2] - c[2]};
          for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) cov[r * 3 + q] += d[r] * d[q];
        }
        double ev[3], evec[3][3];
        sym_eig3_dev(cov, ev, evec);
        if (ev[2] > 3 * ev[1]) {
          R.type = 1;
          for (int i = 0; i < 3; ++i) { R.v[i] = 0.1 * evec[2][i] + c[i]; R.v[3 + i] = -0.1 * evec[2][i] + c[i]; }
        }
      } else {  // :545-580
        double A[15], bb[5] = {-1, -1, -1, -1, -1}, n[3];
        for (int j = 0; j < 5; ++j) { A[j * 3] = P[j][0]; A[j * 3 + 1] = P[j][1]; A[j * 3 + 2] = P[j][2]; }
        colpiv_qr_solve_5x3_dev(A, bb, n);
        const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        const double d = 1 / nn;
        n[0] /= nn; n[1] /= nn; n[2] /= nn;
        bool ok = true;
        for (int j = 0; j < 5; ++j)
          if (fabs(n[0] * P[j][0] + n[1] * P[j][1] + n[2] * P[j][2] + d) > 0.2) { ok = false; break; }
        if (ok) { R.type = 2; R.v[0] = n[0]; R.v[1] = n[1]; R.v[2] = n[2]; R.v[3] = d; }
      }
    }
    gn_store(res, (size_t)b * 2 + kind, qi, R);
    fitType[((size_t)b * 2 + kind) * cap + qi] = (uint8_t)R.type;   // parity read-out: which queries produced a factor in this pass
    nFactors += R.type != 0;
  }
  // residual blocks of the pass (:517 / :581), for the trace and the wide solve: one atomic per warp
  nFactors = __reduce_add_sync(0xffffffffu, nFactors);
  if (lane_id() == 0 && nFactors) atomicAdd(kind == 0 ? &stAll[b].trace[pass].n_corner : &stAll[b].trace[pass].n_plane, nFactors);
}

// lm_solve: one outer pass of :458-626 (the association was just done by lm_associate).  grid (kLmCluster, B), block 256,
// one thread-block CLUSTER per stream: ~13 k residual blocks of double-precision Jacobians are too much for one SM per
// evaluation (the kernel was FP64-bound at one CTA per stream), so the residuals are dealt to the cluster's CTAs, every
// CTA reduces its share to 28 doubles, and the partial sums are exchanged through distributed shared memory.  Every CTA
// adds them in rank order and runs the (cheap, deterministic) trust-region bookkeeping itself, so all of them hold
// bit-identical state and no broadcast is needed; only rank 0 writes results.
constexpr int kLmClusterMax = 8;
constexpr int kGnSplitMinBatch = 1 << 30;   // batch size from which the wide solve (gn_split.cuh) is the default: set by measurement     // cluster size is a launch attribute (1, 2, 4 or 8): see lm_run
__device__ __forceinline__ void lm_solve_body(LMState* __restrict__ stAll, const GNRecArray res,
                                                                                    int cap, int pass, int max_iterations, int lastPass) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ LMShared S;
  __shared__ int s_cnt[2];
  __shared__ double s_part[2][28];     // this CTA's partial sums, two generations (one cluster barrier per evaluation)
  __shared__ SolveTrace s_trace;       // ranks > 0 keep their (identical) trace here
  const int b = blockIdx.y;
  const int rank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  LMState& st = stAll[b];
  if (!st.solved) return;              // uniform over the cluster
  SolveTrace* tr = rank == 0 ? &st.trace[pass] : &s_trace;
  const size_t pl = (size_t)res.n;
  const float4* Pc = res.p + ((size_t)b * 2 + 0) * pl;
  const float4* Ps = res.p + ((size_t)b * 2 + 1) * pl;
  const double* Vc = res.v + ((size_t)b * 2 + 0) * 7 * pl;
  const double* Vs = res.v + ((size_t)b * 2 + 1) * 7 * pl;
  const int nc = st.stackNum[0], ns = st.stackNum[1];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) for (int i = 0; i < 7; ++i) S.x[i] = st.parameters[i];
  __syncthreads();
  if (rank == 0) {
    int a = 0, c = 0;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) a += gn_type(Pc[i]) == 1;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) c += gn_type(Ps[i]) == 2;
    a = __reduce_add_sync(0xffffffffu, a); c = __reduce_add_sync(0xffffffffu, c);
    if (lane_id() == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], c); }
    __syncthreads();
    if (threadIdx.x == 0) { tr->n_corner = s_cnt[0]; tr->n_plane = s_cnt[1]; }
  }
  int gen = 0;
  const int first = rank * blockDim.x + threadIdx.x, stride = csize * blockDim.x;
  auto evaluate = [&](const double* x) {
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double q[4] = {x[0], x[1], x[2], x[3]};
    const double t[3] = {x[4], x[5], x[6]};
    for (int i = first; i < nc; i += stride) {
      const float4 p = Pc[i];
      if (gn_type(p) != 1) continue;
      const double a[3] = {Vc[i], Vc[pl + i], Vc[2 * pl + i]}, bb[3] = {Vc[3 * pl + i], Vc[4 * pl + i], Vc[5 * pl + i]};
      edge_block(q, t, p, a, bb, acc);
    }
    for (int i = first; i < ns; i += stride) {
      const float4 p = Ps[i];
      if (gn_type(p) != 2) continue;
      const double n[3] = {Vs[i], Vs[pl + i], Vs[2 * pl + i]};
      plane_block(q, t, p, n, Vs[3 * pl + i], acc);
    }
    block_reduce28(acc, S.red, S.scratch);
    if (threadIdx.x < 28) s_part[gen][threadIdx.x] = S.red[threadIdx.x];
    cluster.sync();
    if (threadIdx.x < 28) {
      double sum = 0.0;
      for (int r = 0; r < csize; ++r) sum += cluster.map_shared_rank(&s_part[gen][0], r)[threadIdx.x];
      S.red[threadIdx.x] = sum;
    }
    gen ^= 1;
    __syncthreads();
  };
  lm_solve_block(S, tr, max_iterations, false, evaluate);
  if (rank == 0 && threadIdx.x == 0) {
    for (int i = 0; i < 7; ++i) st.parameters[i] = S.x[i];
    st.counters[pass == 0 ? kCntIters0 : kCntIters1] += tr->n_records - 1;       // evaluations after the initial one = LM iterations run
    if (lastPass) st.counters[kCntFactors] += tr->n_corner + tr->n_plane;
  }
  cluster.sync();                      // nobody leaves while a peer may still read its partial sums
}

// Two register budgets: 226 registers (one CTA per SM) and <= 128 (two per SM, some spilling); chosen by measurement
// (VLOAM_LM_SOLVE_REGS=128), DESIGN.md section 5.
__global__ void __launch_bounds__(256) lm_solve(LMState* __restrict__ stAll, const GNRecArray res, int cap, int pass, int max_iterations, int lastPass) {
  lm_solve_body(stAll, res, cap, pass, max_iterations, lastPass);
}
__global__ void __launch_bounds__(256, 2) lm_solve_r128(LMState* __restrict__ stAll, const GNRecArray res, int cap, int pass, int max_iterations, int lastPass) {
  lm_solve_body(stAll, res, cap, pass, max_iterations, lastPass);
}

// Wide solve (gn_split.cuh): book-keeping after the last gn_step of a pass.
__global__ void lm_gn_finish(LMState* __restrict__ stAll, int B, int pass, int lastPass) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LMState& st = stAll[b];
  if (!st.solved) return;
  const SolveTrace& tr = st.trace[pass];
  st.counters[pass == 0 ? kCntIters0 : kCntIters1] += tr.n_records - 1;
  if (lastPass) st.counters[kCntFactors] += tr.n_corner + tr.n_plane;
}

// ---------------------------------------------------------------------------------------------------------------
// Position of a point on the voxel lattice of its cube's filter (pcl::VoxelGrid: voxel = floor(p / leaf) per axis), relative
// to the cube's min corner, packed z | y | x with `s` bits per axis.  pcl::VoxelGrid emits voxels in ascending
// i + j * dx + k * dx * dy of bounding-box-relative coordinates, i.e. in lexicographic (z, y, x) lattice order whatever the
// box: ascending order of this key.  Valid for points inside the cube (all points of a cube are) and 3 s <= 30 bits.
struct VoxLattice { float inv; int bx, by, bz, s; };
__device__ __forceinline__ VoxLattice vox_lattice(int cube, int cenW, int cenH, int cenD, float leaf, int s) {
  VoxLattice L;
  L.inv = __fdiv_rn(1.0f, leaf);
  L.bx = (int)floorf(__fmul_rn(cube_min_coord(cube % kCubeW, cenW), L.inv)) - 1;
  L.by = (int)floorf(__fmul_rn(cube_min_coord((cube / kCubeW) % kCubeH, cenH), L.inv)) - 1;
  L.bz = (int)floorf(__fmul_rn(cube_min_coord(cube / (kCubeW * kCubeH), cenD), L.inv)) - 1;
  L.s = s;
  return L;
}
__device__ __forceinline__ unsigned vox_key(const VoxLattice& L, float x, float y, float z) {
  const unsigned mask = (1u << L.s) - 1u;
  const unsigned rx = (unsigned)((int)floorf(__fmul_rn(x, L.inv)) - L.bx) & mask;
  const unsigned ry = (unsigned)((int)floorf(__fmul_rn(y, L.inv)) - L.by) & mask;
  const unsigned rz = (unsigned)((int)floorf(__fmul_rn(z, L.inv)) - L.bz) & mask;
  return (rz << (2 * L.s)) | (ry << L.s) | rx;
}
// bits per axis for a leaf size (host): the lattice spans 50 / leaf + 3 voxels of a 50 m cube; 0 = too fine for a 30-bit key
static int vox_axis_bits(double leaf) {
  if (!(leaf > 0.0)) return 0;
  const double span = 50.0 / leaf + 4.0;
  int s = 1;
  while ((double)(1 << s) < span && s < 31) ++s;
  return 3 * s <= 30 ? s : 0;
}

// lm_insert_keys: grid (2, B), block 1024.  transformUpdate (:636, :140-144), map-frame coordinates of the stack
// points and their cube ids, stable sort by cube id (:639-683 push the points in stack order), work list.
//
// Work list = the cubes this scan rewrites: every cube that receives points (a valid cube is re-filtered, :689-702; any
// other cube only grows) plus every non-empty valid cube that is not known to be a fixed point of its filter.  A valid
// cube that received nothing since a filter left it in fixed-point form would come out of pcl::VoxelGrid unchanged
// (cta_voxel_filter), so it is not touched at all.
__global__ void __launch_bounds__(1024) lm_insert_keys(LMState* __restrict__ stAll, const float4* __restrict__ stack, int cap,
                                                        float4* __restrict__ stackW, const int* __restrict__ cubeCnt,
                                                        const int* __restrict__ cubeFix, int* __restrict__ cubeOfAll, float lineRes,
                                                        float planeRes, int bitsLine, int bitsPlane, unsigned* kA, unsigned* vA,
                                                        unsigned* kB, unsigned* vB, size_t workCap) {
  __shared__ SortSmem S;
  __shared__ int s_res;
  const int kind = blockIdx.x, b = blockIdx.y;
  LMState& st = stAll[b];
  const int n = st.stackNum[kind];
  const float4* in = stack + ((size_t)b * 2 + kind) * cap;
  float4* outW = stackW + ((size_t)b * 2 + kind) * cap;
  const size_t so = ((size_t)b * 2 + kind) * workCap;
  unsigned* ka = kA + so; unsigned* va = vA + so; unsigned* kb = kB + so; unsigned* vb = vB + so;
  int* cubeOf = cubeOfAll + ((size_t)b * 2 + kind) * cap;
  const float leaf = kind == 0 ? lineRes : planeRes;
  const int axisBits = kind == 0 ? bitsLine : bitsPlane;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float4 p = in[i];
    double w[3];
    quat_rotate(st.parameters, (double)p.x, (double)p.y, (double)p.z, w);
    const float x = (float)(w[0] + st.parameters[4]), y = (float)(w[1] + st.parameters[5]), z = (float)(w[2] + st.parameters[6]);
    outW[i] = make_float4(x, y, z, p.w);
    int cI = (int)(((double)x + 25.0) / 50.0) + st.cenW, cJ = (int)(((double)y + 25.0) / 50.0) + st.cenH, cK = (int)(((double)z + 25.0) / 50.0) + st.cenD;
    if ((double)x + 25.0 < 0) cI--;
    if ((double)y + 25.0 < 0) cJ--;
    if ((double)z + 25.0 < 0) cK--;
    unsigned key = 0xffffu, vkey = 0u;
    if (cI >= 0 && cI < kCubeW && cJ >= 0 && cJ < kCubeH && cK >= 0 && cK < kCubeD) {
      key = (unsigned)(cI + kCubeW * cJ + kCubeW * kCubeH * cK);
      if (axisBits) vkey = vox_key(vox_lattice((int)key, st.cenW, st.cenH, st.cenD, leaf, axisBits), x, y, z);
    }
    cubeOf[i] = (int)key; ka[i] = vkey; va[i] = (unsigned)i;
  }
  __syncthreads();
  // Order: by cube, inside a cube by voxel of the cube's filter lattice, inside a voxel by stack index (the reference
  // pushes the points in stack order, :639-683; only the order inside a voxel matters to the filter's sums).  Two stable
  // sorts: voxel key first, cube id second.
  int cur = 0;
  if (axisBits) {
    cur = cta_radix_sort(ka, va, kb, vb, n, 3 * axisBits, S);
    unsigned* kr = cur ? kb : ka;
    const unsigned* vr = cur ? vb : va;
    for (int i = threadIdx.x; i < n; i += 1024) kr[i] = (unsigned)cubeOf[vr[i]];
  } else {
    for (int i = threadIdx.x; i < n; i += 1024) ka[i] = (unsigned)cubeOf[i];
  }
  __syncthreads();
  cur ^= cur ? cta_radix_sort(kb, vb, ka, va, n, 16, S) : cta_radix_sort(ka, va, kb, vb, n, 16, S);
  if (threadIdx.x == 0) s_res = cur;
  __syncthreads();
  const unsigned* keys = s_res ? kb : ka;
  const unsigned* vals = s_res ? vb : va;
  // result always left in (kA, vA) so the next kernel needs no flag
  if (s_res) { for (int i = threadIdx.x; i < n; i += 1024) { ka[i] = keys[i]; va[i] = vals[i]; } }
  __syncthreads();
  // runs of equal cube id in the sorted key array: heads found in parallel, then a short serial work-list build
  constexpr int kBitWords = (kCubes + 31) / 32;
  __shared__ int s_nh, s_headCube[kMaxWork], s_headStart[kMaxWork], s_headEnd[kMaxWork];
  __shared__ unsigned s_validBits[kBitWords], s_listed[kBitWords];
  if (threadIdx.x == 0) s_nh = 0;
  for (int i = threadIdx.x; i < kBitWords; i += 1024) { s_validBits[i] = 0u; s_listed[i] = 0u; }
  __syncthreads();
  for (int v = threadIdx.x; v < st.validNum; v += 1024) { const int c = st.validInd[v]; atomicOr(&s_validBits[c >> 5], 1u << (c & 31)); }
  for (int i = threadIdx.x; i < n; i += 1024) {
    const unsigned c = ka[i];
    if (c != 0xffffu && (i == 0 || ka[i - 1] != c)) {
      const int h = atomicAdd(&s_nh, 1);
      if (h < kMaxWork) { s_headCube[h] = (int)c; s_headStart[h] = i; }
    }
  }
  __syncthreads();
  const int nh = min(s_nh, kMaxWork);
  for (int h = threadIdx.x; h < nh; h += 1024) {  // end of run h = start of the next different key
    int lo = s_headStart[h], hi = n;               // keys are sorted: binary search the first index with key > cube
    const unsigned c = (unsigned)s_headCube[h];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (ka[mid] <= c) lo = mid + 1; else hi = mid; }
    s_headEnd[h] = lo;
  }
  // the table rows the serial list build below needs, fetched side by side (it used to read them one dependent global load
  // at a time: a quarter of the kernel)
  __shared__ int s_headCnt[kMaxWork], s_validCube[kMaxValid], s_validCnt[kMaxValid];
  __shared__ unsigned char s_validOpen[kMaxValid];
  {
    const int* cnt = cubeCnt + ((size_t)b * 2 + kind) * kCubes;
    const int* fix = cubeFix + ((size_t)b * 2 + kind) * kCubes;
    for (int h = threadIdx.x; h < nh; h += 1024) s_headCnt[h] = cnt[s_headCube[h]];
    for (int v = threadIdx.x; v < st.validNum; v += 1024) {
      const int c = st.validInd[v];
      const int n0 = cnt[c];
      s_validCube[v] = c; s_validCnt[v] = n0; s_validOpen[v] = (n0 != 0 && !fix[c]) ? 1 : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_nh > kMaxWork) atomicOr(&st.error, kLmErrWorkList);
    const int vnum = st.validNum;
    int wn = 0, in0 = 0;
    for (int h = 0; h < nh; ++h) {                 // cubes that receive points (distinct by construction)
      const int c = s_headCube[h];
      s_listed[c >> 5] |= 1u << (c & 31);
      const int nNew = s_headEnd[h] - s_headStart[h];
      st.workCube[kind][wn] = c; st.workFilter[kind][wn] = (s_validBits[c >> 5] >> (c & 31)) & 1u;
      st.workNew0[kind][wn] = s_headStart[h]; st.workNewN[kind][wn] = nNew;
      st.workIn0[kind][wn] = in0; in0 += s_headCnt[h] + nNew; ++wn;
    }
    for (int v = 0; v < vnum; ++v) {               // valid cubes whose filter result is not known to be a fixed point
      if (!s_validOpen[v]) continue;
      const int c = s_validCube[v];
      if ((s_listed[c >> 5] >> (c & 31)) & 1u) continue;
      if (wn >= kMaxWork) { atomicOr(&st.error, kLmErrWorkList); break; }
      s_listed[c >> 5] |= 1u << (c & 31);
      st.workCube[kind][wn] = c; st.workFilter[kind][wn] = 1; st.workNew0[kind][wn] = 0; st.workNewN[kind][wn] = 0;
      st.workIn0[kind][wn] = in0; in0 += s_validCnt[v]; ++wn;
    }
    st.workIn0[kind][wn] = in0;
    st.workNum[kind] = wn;
    st.counters[kCntWorkCorner + kind] += wn;
    atomicAdd(&st.counters[kCntQueries], n);
    if ((size_t)in0 + (size_t)cap > workCap) atomicOr(&st.error, kLmErrScratch);
  }
}
__global__ void lm_transform_update(LMState* __restrict__ stAll, int B) {  // :140-144
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LMState& st = stAll[b];
  const double* q = st.q_wodom;
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double qi[4] = {-q[0] / n2, -q[1] / n2, -q[2] / n2, q[3] / n2};  // Eigen Quaternion::inverse()
  double qm[4], r[3];
  quat_mul(st.parameters, qi, qm);
  quat_rotate(qm, st.t_wodom[0], st.t_wodom[1], st.t_wodom[2], r);
  for (int i = 0; i < 4; ++i) st.q_wmap_wodom[i] = qm[i];
  for (int i = 0; i < 3; ++i) st.t_wmap_wodom[i] = st.parameters[4 + i] - r[i];
}

// Re-filter of the work list (LM:689-702 for the cubes that can change), one work cube per CTA; input = old cube ++ new points.
//
// Three paths:
//   merge   (lm_refilter_merge, 256 threads) the cube is a fixed point of its filter (one point per voxel, in lattice
//           order) and its new points arrive sorted by voxel (lm_insert_keys): nothing is sorted.  The new points' voxel
//           runs are located in the old cube by binary search; a run that hits an occupied voxel replaces that point by
//           the voxel's new centroid ((0 + old) + new_1 + ... in stack order, exactly the sum pcl::VoxelGrid forms over
//           old ++ new), a run in an empty voxel is inserted.  No insertion -> the few changed points are patched in the
//           slab itself (workDirect), else the merged cube is written to `staged`.
//   filter  (lm_refilter, 1024 threads) any other valid cube: the full voxel filter (sort) over old ++ new.
//   append  (lm_refilter) a cube outside the valid list only grows (:639-683), in stack order.
__device__ __forceinline__ bool refilter_merges(const LMState& st, int kind, int u, int nOld, int fixedPoint, int axisBits) {
  return st.workFilter[kind][u] && fixedPoint && nOld > 0 && st.workNewN[kind][u] > 0 && axisBits > 0;
}
template <int NT>
__device__ __forceinline__ int block_exclusive_scan_nt(int v, int* s_w /*[NT / 32 + 1]*/) {   // s_w[NT / 32] = block total
  const int w = threadIdx.x >> 5, l = lane_id();
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (l >= o) s += t; }
  __syncthreads();                 // s_w may still be read from a previous call
  if (l == 31) s_w[w] = s;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < NT / 32; ++q) { const int x = s_w[q]; if (q < w) base += x; tot += x; }
  if (threadIdx.x == 0) s_w[NT / 32] = tot;
  __syncthreads();
  return base + s - v;
}
constexpr int kMergeThreads = 256;
constexpr int kSortGrid = 16;     // CTAs per stream and kind of the sort-path re-filter (grid-stride over the work list)
__global__ void __launch_bounds__(kMergeThreads) lm_refilter_merge(LMState* __restrict__ stAll, const int* __restrict__ cubeOff,
                                                                    const int* __restrict__ cubeCnt, const int* __restrict__ cubeFix,
                                                                    const MapPools pools, int mapCap, const float4* __restrict__ stackW,
                                                                    int cap, const unsigned* __restrict__ vAins, float lineRes, float planeRes,
                                                                    int bitsLine, int bitsPlane, float4* __restrict__ concat,
                                                                    float4* __restrict__ staged, unsigned* kA, unsigned* vA, unsigned* kB,
                                                                    unsigned* vB, size_t workCap) {
  constexpr int NT = kMergeThreads;
  __shared__ int s_w[NT / 32 + 1];
  const int kind = blockIdx.y, b = blockIdx.z;
  LMState& st = stAll[b];
  if (st.error & (kLmErrWorkList | kLmErrScratch)) return;
  const int nWork = st.workNum[kind];
  for (int u = blockIdx.x; u < nWork; u += gridDim.x) {   // (a grid of one CTA per possible work cube would be mostly empty CTAs)
  const int c = st.workCube[kind][u];
  const size_t so = ((size_t)b * 2 + kind) * workCap;
  const size_t tb = ((size_t)b * 2 + kind) * kCubes;
  const int nOld = cubeCnt[tb + c], nNew = st.workNewN[kind][u];
  const int axisBits = kind == 0 ? bitsLine : bitsPlane;
  if (!refilter_merges(st, kind, u, nOld, cubeFix[tb + c], axisBits)) continue;   // lm_refilter's cube
  const int in0 = st.workIn0[kind][u];
  float4* old = stream_map(pools, st, b, kind, mapCap) + cubeOff[tb + c];
  const float4* sw = stackW + ((size_t)b * 2 + kind) * cap;
  const unsigned* ord = vAins + so + st.workNew0[kind][u];   // stack indices of this cube's new points: by voxel, then stack order
  float4* out = staged + so + in0;
  const int tid = threadIdx.x;
  const VoxLattice L = vox_lattice(c, st.cenW, st.cenH, st.cenD, kind == 0 ? lineRes : planeRes, axisBits);
  unsigned* keyNew = kA + so + in0;                       // [nNew] voxel key of every new point
  int* runStart = reinterpret_cast<int*>(vA + so + in0);  // [runs + 1] first new point of every voxel run
  int* runPos = reinterpret_cast<int*>(kB + so + in0);    // [runs] position in the old cube (bit 31: that voxel is occupied)
  int* insBefore = reinterpret_cast<int*>(vB + so + in0); // [runs] inserted runs before this one
  int* insPos = reinterpret_cast<int*>(concat + so + in0);  // [inserted] old-cube position of every inserted run (concat is free here)
  for (int t = tid; t < nNew; t += NT) { const float4 p = sw[ord[t]]; keyNew[t] = vox_key(L, p.x, p.y, p.z); }
  __syncthreads();
  const int per = (nNew + NT - 1) / NT;
  const int t0 = min(tid * per, nNew), t1 = min(t0 + per, nNew);
  int nh = 0;
  for (int t = t0; t < t1; ++t) nh += (t == 0 || keyNew[t] != keyNew[t - 1]) ? 1 : 0;
  int r0 = block_exclusive_scan_nt<NT>(nh, s_w);
  const int runs = s_w[NT / 32];
  for (int t = t0; t < t1; ++t) if (t == 0 || keyNew[t] != keyNew[t - 1]) runStart[r0++] = t;
  if (tid == 0) runStart[runs] = nNew;
  __syncthreads();
  for (int r = tid; r < runs; r += NT) {
    const unsigned vk = keyNew[runStart[r]];
    int lo = 0, hi = nOld;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float4 q = old[mid];
      if (vox_key(L, q.x, q.y, q.z) < vk) lo = mid + 1; else hi = mid;
    }
    int hit = 0;
    if (lo < nOld) { const float4 q = old[lo]; hit = vox_key(L, q.x, q.y, q.z) == vk; }
    runPos[r] = lo | (hit ? (int)0x80000000 : 0);
  }
  __syncthreads();
  const int perR = (runs + NT - 1) / NT;
  const int q0 = min(tid * perR, runs), q1 = min(q0 + perR, runs);
  int ni = 0;
  for (int r = q0; r < q1; ++r) ni += runPos[r] < 0 ? 0 : 1;
  int i0 = block_exclusive_scan_nt<NT>(ni, s_w);
  const int inserted = s_w[NT / 32];
  for (int r = q0; r < q1; ++r) {
    insBefore[r] = i0;
    if (runPos[r] >= 0) insPos[i0++] = runPos[r];
  }
  __syncthreads();
  const int m = nOld + inserted;
  const int direct = inserted == 0 ? 1 : 0;
  float4* dst = direct ? old : out;
  if (!direct) {
    // old point i moves up by the number of inserted runs placed at or before it (insPos is ascending)
    // (a thread's i only grows, so its count is carried along instead of searched for every point: the searches were 39 %
    // of the kernel's instructions)
    int lo = 0;
    for (int i = tid; i < nOld; i += NT) {
      while (lo < inserted && insPos[lo] <= i) ++lo;
      dst[i + lo] = old[i];
    }
    __syncthreads();
  }
  int inside = 1;
  for (int r = tid; r < runs; r += NT) {
    const int pos = runPos[r] & 0x7fffffff;
    const bool hit = runPos[r] < 0;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    if (hit) { const float4 o = old[pos]; sx = __fadd_rn(sx, o.x); sy = __fadd_rn(sy, o.y); sz = __fadd_rn(sz, o.z); si = __fadd_rn(si, o.w); cnt = 1; }
    for (int t = runStart[r]; t < runStart[r + 1]; ++t) {
      const float4 p = sw[ord[t]];
      sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
      ++cnt;
    }
    const float nf = (float)cnt;
    const float4 cen = make_float4(__fdiv_rn(sx, nf), __fdiv_rn(sy, nf), __fdiv_rn(sz, nf), __fdiv_rn(si, nf));
    if (vox_key(L, cen.x, cen.y, cen.z) != keyNew[runStart[r]]) inside = 0;
    dst[pos + insBefore[r]] = cen;
  }
  const int fixed = __syncthreads_and(inside);
  if (tid == 0) {
    st.workOutN[kind][u] = m; st.workFixed[kind][u] = fixed; st.workDirect[kind][u] = direct;
    atomicAdd(&st.counters[kCntMerged], 1); atomicAdd(&st.counters[kCntMergedInPlace], direct); atomicAdd(&st.counters[kCntVoxelsInserted], inserted);
  }
  }
}

// lm_refilter: grid (kSortGrid, 2, B), block 1024: the filter and append paths (rare once the map is in fixed-point form).
__global__ void __launch_bounds__(1024) lm_refilter(LMState* __restrict__ stAll, const int* __restrict__ cubeOff, const int* __restrict__ cubeCnt,
                                                     const int* __restrict__ cubeFix, const MapPools pools, int mapCap,
                                                     const float4* __restrict__ stackW,
                                                     int cap, const unsigned* __restrict__ vAins, float lineRes, float planeRes, int bitsLine, int bitsPlane,
                                                     float4* __restrict__ concat, float4* __restrict__ staged, unsigned* kA, unsigned* vA,
                                                     unsigned* kB, unsigned* vB, size_t workCap) {
  __shared__ SortSmem S;
  __shared__ float red[6 * 32];
  const int kind = blockIdx.y, b = blockIdx.z;
  LMState& st = stAll[b];
  if (st.error & (kLmErrWorkList | kLmErrScratch)) return;
  const int nWork = st.workNum[kind];
  for (int u = blockIdx.x; u < nWork; u += gridDim.x) {
  const int c = st.workCube[kind][u];
  const size_t so = ((size_t)b * 2 + kind) * workCap;
  const size_t tb = ((size_t)b * 2 + kind) * kCubes;
  const int in0 = st.workIn0[kind][u];
  const int nOld = cubeCnt[tb + c], nNew = st.workNewN[kind][u], n = nOld + nNew;
  if (refilter_merges(st, kind, u, nOld, cubeFix[tb + c], kind == 0 ? bitsLine : bitsPlane)) continue;   // lm_refilter_merge's cube
  const float4* old = stream_map(pools, st, b, kind, mapCap) + cubeOff[tb + c];
  const float4* sw = stackW + ((size_t)b * 2 + kind) * cap;
  const unsigned* ord = vAins + so + st.workNew0[kind][u];   // stack indices of this cube's new points: by voxel, then stack order
  float4* in = concat + so + in0;
  float4* out = staged + so + in0;
  const float leaf = kind == 0 ? lineRes : planeRes;
  const int tid = threadIdx.x;
  int m, fixed = 0;
  if (st.workFilter[kind][u]) {
    // ---------------- filter
    for (int i = tid; i < n; i += 1024) in[i] = i < nOld ? old[i] : sw[ord[i - nOld]];
    __syncthreads();
    // scratch keys for this cube live at the cube's input offset in the second half of the key arrays
    m = cta_voxel_filter(in, n, leaf, out, kA + so + in0, vA + so + in0, kB + so + in0, vB + so + in0, S, red, &fixed);
  } else {
    // ---------------- append: back to stack order (a 16-bit key covers every stack index of a scan up to 65 536 points;
    // larger scans sort on all 32 bits)
    unsigned* ka = kA + so + in0; unsigned* va = vA + so + in0; unsigned* kb = kB + so + in0; unsigned* vb = vB + so + in0;
    for (int t = tid; t < nNew; t += 1024) { ka[t] = ord[t]; va[t] = ord[t]; }
    __syncthreads();
    const int cur = cta_radix_sort(ka, va, kb, vb, nNew, cap <= 65536 ? 16 : 32, S);
    const unsigned* vs = cur ? vb : va;
    for (int i = tid; i < n; i += 1024) out[i] = i < nOld ? old[i] : sw[vs[i - nOld]];
    m = n;
  }
  if (tid == 0) {
    st.workOutN[kind][u] = m; st.workFixed[kind][u] = fixed; st.workDirect[kind][u] = 0;
    atomicAdd(&st.counters[st.workFilter[kind][u] ? kCntFiltered : kCntAppended], 1);
  }
  __syncthreads();   // shared memory is reused by the next cube
  }
}

// lm_place: grid (2, B), block 1024.  New cube tables `dst` from the post-shift tables `src`: a rewritten cube keeps its
// slab when the result fits, else gets a new slab (with head-room) from the pool's bump allocator; when the pool is
// exhausted the whole map is re-packed into the stream's other pool (st.compact, done by lm_compact_copy) with fresh
// head-room for every cube.  liveList: the non-empty cubes a re-pack has to move (the rewritten ones come from `staged`).
__device__ __forceinline__ int slab_headroom(int n) { return (n >> 2) + 256; }
__global__ void __launch_bounds__(1024) lm_place(LMState* __restrict__ stAll, const CubeTables src, const CubeTables dst,
                                                  short* __restrict__ workOfAll, short* __restrict__ liveListAll, int* __restrict__ liveNumAll,
                                                  int mapCap) {
  __shared__ SortSmem S;
  __shared__ int s_compact, s_nlive, s_err;
  const int kind = blockIdx.x, b = blockIdx.y;
  LMState& st = stAll[b];
  const size_t tb = ((size_t)b * 2 + kind) * kCubes;
  short* workOf = workOfAll + tb;
  short* liveList = liveListAll + tb;
  for (int c = threadIdx.x; c < kCubes; c += 1024) {
    workOf[c] = -1;
    dst.off[tb + c] = src.off[tb + c]; dst.cnt[tb + c] = src.cnt[tb + c]; dst.cap[tb + c] = src.cap[tb + c]; dst.fix[tb + c] = src.fix[tb + c];
    dst.tab[tb + c] = src.tab[tb + c];
  }
  // Only the bits lm_insert_keys raised (a whole kernel ago) are read here; the capacity bit is per kind and is raised and
  // acted upon by this CTA alone, so the two kinds' CTAs never see each other's decision half-way.
  if (threadIdx.x == 0) {
    s_compact = 0; s_nlive = 0; st.compact[kind] = 0; st.applied[kind] = 0;
    s_err = st.error & (kLmErrWorkList | kLmErrScratch);
  }
  __syncthreads();
  if (s_err) { if (threadIdx.x == 0) liveNumAll[b * 2 + kind] = 0; return; }   // nothing was re-filtered: the map keeps its pre-insertion state
  const int wn = st.workNum[kind];
  for (int u = threadIdx.x; u < wn; u += 1024) workOf[st.workCube[kind][u]] = (short)u;
  __syncthreads();
  if (threadIdx.x == 0) {
    int poolEnd = st.poolEnd[kind];
    for (int u = 0; u < wn; ++u) {
      const int c = st.workCube[kind][u], m = st.workOutN[kind][u];
      if (m > src.cap[tb + c]) {
        const int need = m + slab_headroom(m);
        if ((long long)poolEnd + need > (long long)mapCap) { s_compact = 1; break; }
        dst.off[tb + c] = poolEnd; dst.cap[tb + c] = need; poolEnd += need;
      }
      dst.cnt[tb + c] = m;
      dst.fix[tb + c] = st.workFilter[kind][u] ? st.workFixed[kind][u] : 0;
    }
    if (!s_compact) {
      st.poolEnd[kind] = poolEnd; st.applied[kind] = 1;
      // a rewritten cube is re-indexed by lm_write_back: in its old table slot, a new one, or (no slot left) lazily by
      // the next lm_prepare
      int te = st.tabEnd[kind];
      for (int u = 0; u < wn; ++u) {
        const int c = st.workCube[kind][u];
        if (st.workOutN[kind][u] == 0) dst.tab[tb + c] = -1;
        else if (dst.tab[tb + c] < 0 && te < st.tabLimit) dst.tab[tb + c] = te++;
      }
      st.tabEnd[kind] = te;
    }
  }
  __syncthreads();
  if (!s_compact) { if (threadIdx.x == 0) liveNumAll[b * 2 + kind] = 0; return; }
  // ---- re-pack: every cube gets a fresh slab in the other pool
  const int per = (kCubes + 1023) / 1024;
  const int c0 = min((int)threadIdx.x * per, kCubes), c1 = min(c0 + per, kCubes);
  int sum = 0;
  for (int c = c0; c < c1; ++c) sum += workOf[c] >= 0 ? st.workOutN[kind][workOf[c]] : src.cnt[tb + c];
  block_exclusive_scan1024(sum, S);
  const long long total = S.total;
  __syncthreads();
  if (total > (long long)mapCap) {
    // capacity exceeded: this kind's map keeps its pre-insertion tables (vloam_get_lm_status reports it).  The cubes
    // lm_refilter_merge patched inside their slab (no insertion: same size, same slab) did change; their filter flag
    // follows, lm_write_back re-indexes them.
    for (int c = threadIdx.x; c < kCubes; c += 1024) {
      const int u = workOf[c];
      const bool direct = u >= 0 && st.workDirect[kind][u];
      dst.off[tb + c] = src.off[tb + c]; dst.cnt[tb + c] = src.cnt[tb + c]; dst.cap[tb + c] = src.cap[tb + c];
      dst.fix[tb + c] = direct ? st.workFixed[kind][u] : src.fix[tb + c];
      dst.tab[tb + c] = src.tab[tb + c];
    }
    if (threadIdx.x == 0) { atomicOr(&st.error, kLmErrCapacity << kind); liveNumAll[b * 2 + kind] = 0; }
    return;
  }
  const long long spare = ((long long)mapCap - total) / 2;    // half of the free space becomes head-room, half stays for the bump allocator
  int capsum = 0;
  for (int c = c0; c < c1; ++c) {
    const int n = workOf[c] >= 0 ? st.workOutN[kind][workOf[c]] : src.cnt[tb + c];
    const long long share = total > 0 ? spare * n / total : 0;
    capsum += n > 0 ? n + (int)min((long long)slab_headroom(n), share) : 0;
  }
  int run = block_exclusive_scan1024(capsum, S);
  for (int c = c0; c < c1; ++c) {
    const int u = workOf[c];
    const int n = u >= 0 ? st.workOutN[kind][u] : src.cnt[tb + c];
    const long long share = total > 0 ? spare * n / total : 0;
    const int cp = n > 0 ? n + (int)min((long long)slab_headroom(n), share) : 0;
    dst.off[tb + c] = run; dst.cnt[tb + c] = n; dst.cap[tb + c] = cp;
    dst.fix[tb + c] = u >= 0 ? (st.workFilter[kind][u] ? st.workFixed[kind][u] : 0) : src.fix[tb + c];
    dst.tab[tb + c] = -1;              // every slab moves: the sorted copies (kept at the slab offsets) are void
    run += cp;
    if (u < 0 && n > 0) liveList[atomicAdd(&s_nlive, 1)] = (short)c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st.poolEnd[kind] = S.total; st.compact[kind] = 1; st.applied[kind] = 1; liveNumAll[b * 2 + kind] = s_nlive;
    int te = 0;                        // table slots start over; the rewritten cubes are re-indexed right away
    for (int u = 0; u < wn && te < st.tabLimit; ++u) if (st.workOutN[kind][u] > 0) dst.tab[tb + st.workCube[kind][u]] = te++;
    st.tabEnd[kind] = te;
  }
}
// lm_write_back: grid (kMaxWork, 2, B), block kIndexThreads: rewritten cube -> its slab (in the other pool when the map is re-packed;
// nothing to move when lm_refilter patched the slab in place), then the cube's column index is rebuilt from the new content.
__global__ void __launch_bounds__(kIndexThreads, 2) lm_write_back(const LMState* __restrict__ stAll, const CubeTables Tsrc, const CubeTables T,
                                                      const MapPools pools, int mapCap, const float4* __restrict__ staged, size_t workCap,
                                                      int* __restrict__ tabPool, float4* __restrict__ sorted) {
  extern __shared__ int idx_smem[];
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  if (st.error & (kLmErrWorkList | kLmErrScratch)) return;   // the scan's insertion was abandoned before the re-filter (lm_insert_keys)
  const bool applied = st.applied[kind] != 0;   // lm_place committed this kind's new tables
  const int nWork = st.workNum[kind];
  for (int u = blockIdx.x; u < nWork; u += gridDim.x) {
    const int c = st.workCube[kind][u], n = st.workOutN[kind][u];
    const size_t tb = ((size_t)b * 2 + kind) * kCubes;
    const bool direct = st.workDirect[kind][u] != 0;
    // placement failed (the map would not fit its pool even after a re-pack): only the cubes lm_refilter_merge already
    // patched inside their slab change — same size, same slab — and their index has to follow; everything else keeps
    // its pre-insertion state
    if (!applied && !direct) continue;
    const int off = T.off[tb + c], slot = T.tab[tb + c];
    const float4* src = direct ? stream_map(pools, st, b, kind, mapCap) + Tsrc.off[tb + c]
                               : staged + ((size_t)b * 2 + kind) * workCap + st.workIn0[kind][u];
    float4* dst = ((st.cur[kind] ^ st.compact[kind]) ? pools.p[1] : pools.p[0]) + ((size_t)b * 2 + kind) * mapCap + off;
    if (src != dst)
      for (int i0 = threadIdx.x; i0 < n; i0 += 4 * kIndexThreads) {
        float4 p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = i0 + k * kIndexThreads; if (i < n) p[k] = src[i]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = i0 + k * kIndexThreads; if (i < n) dst[i] = p[k]; }
      }
    if (slot < 0 || n == 0) continue;
    cta_build_cube_index<kIndexThreads>(src, n, cube_min_coord(c % kCubeW, st.cenW), cube_min_coord((c / kCubeW) % kCubeH, st.cenH),
                                        cube_min_coord(c / (kCubeW * kCubeH), st.cenD),
                                        tabPool + (((size_t)b * 2 + kind) * kTabSlots + slot) * kTabInts,
                                        sorted + ((size_t)b * 2 + kind) * mapCap + off, idx_smem);
  }
}
// lm_compact_copy: grid (16, 2, B), block 256: re-pack only — the cubes this scan did not rewrite move to the other pool.
__global__ void __launch_bounds__(256) lm_compact_copy(const LMState* __restrict__ stAll, const int* __restrict__ offSrc,
                                                        const int* __restrict__ offDst, const int* __restrict__ cntDst,
                                                        const short* __restrict__ liveListAll, const int* __restrict__ liveNumAll,
                                                        const MapPools pools, int mapCap) {
  const int kind = blockIdx.y, b = blockIdx.z;
  const LMState& st = stAll[b];
  if (!st.compact[kind]) return;
  const size_t tb = ((size_t)b * 2 + kind) * kCubes;
  const float4* from = (st.cur[kind] ? pools.p[1] : pools.p[0]) + ((size_t)b * 2 + kind) * mapCap;
  float4* to = (st.cur[kind] ? pools.p[0] : pools.p[1]) + ((size_t)b * 2 + kind) * mapCap;
  const int nl = liveNumAll[b * 2 + kind];
  for (int i = blockIdx.x; i < nl; i += gridDim.x) {
    const int c = liveListAll[tb + i];
    const float4* src = from + offSrc[tb + c];
    float4* dst = to + offDst[tb + c];
    const int n = cntDst[tb + c];
    for (int k = threadIdx.x; k < n; k += 256) dst[k] = src[k];
  }
}

// Also completes a re-pack: the stream's map now lives in its other pool.
__global__ void lm_export_pose(LMState* __restrict__ stAll, double* __restrict__ pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LMState& s = stAll[b];
  for (int kind = 0; kind < 2; ++kind) if (s.compact[kind]) { s.cur[kind] ^= 1; s.compact[kind] = 0; s.repacks[kind]++; s.counters[kCntRepacks]++; }
  double* o = pose + (size_t)b * 16;
  for (int i = 0; i < 7; ++i) o[i] = s.parameters[i];
  for (int i = 0; i < 4; ++i) o[7 + i] = s.q_wmap_wodom[i];
  for (int i = 0; i < 3; ++i) o[11 + i] = s.t_wmap_wodom[i];
  s.errorEver |= s.error;
  o[14] = s.error; o[15] = s.solved;
}
// A frame laserOdometry marks as skipped (laser_odometry.cpp:618-628) only refreshes the high-frequency pose
// (laser_mapping.cpp:186-190, published at :742-756): q_w_curr_highfreq = q_wmap_wodom * q_wodom_curr,
// t_w_curr_highfreq = q_wmap_wodom * t_wodom_curr + t_wmap_wodom.  q_w_curr / t_w_curr (st.parameters) and the map stay.
__global__ void lm_highfreq_pose(LMState* __restrict__ stAll, const LOState* __restrict__ lo, double* __restrict__ pose, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LMState& s = stAll[b];
  for (int i = 0; i < 4; ++i) s.q_wodom[i] = lo[b].q_w[i];     // :182-183 run for skipped frames too
  for (int i = 0; i < 3; ++i) s.t_wodom[i] = lo[b].t_w[i];
  double q[4], t[3];
  quat_mul(s.q_wmap_wodom, s.q_wodom, q);
  quat_rotate(s.q_wmap_wodom, s.t_wodom[0], s.t_wodom[1], s.t_wodom[2], t);
  double* o = pose + (size_t)b * 16;
  for (int i = 0; i < 4; ++i) o[i] = q[i];
  for (int i = 0; i < 3; ++i) o[4 + i] = t[i] + s.t_wmap_wodom[i];
  for (int i = 0; i < 4; ++i) o[7 + i] = s.q_wmap_wodom[i];
  for (int i = 0; i < 3; ++i) o[11 + i] = s.t_wmap_wodom[i];
  o[14] = 0.0; o[15] = 0.0;
}
__global__ void lm_init_state(LMState* stAll, int B, int tabLimit) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  LMState& s = stAll[b];
  for (int i = 0; i < 7; ++i) s.parameters[i] = i == 3 ? 1.0 : 0.0;           // :74-83
  for (int i = 0; i < 4; ++i) { s.q_wmap_wodom[i] = i == 3 ? 1.0 : 0.0; s.q_wodom[i] = i == 3 ? 1.0 : 0.0; }  // :87-91
  for (int i = 0; i < 3; ++i) { s.t_wmap_wodom[i] = 0.0; s.t_wodom[i] = 0.0; }
  s.cenW = 10; s.cenH = 10; s.cenD = 5;                                       // laser_mapping.h:76-78
  s.validNum = 0; s.fromMapNum[0] = s.fromMapNum[1] = 0; s.stackNum[0] = s.stackNum[1] = 0; s.solved = 0;
  s.poolEnd[0] = s.poolEnd[1] = 0; s.workNum[0] = s.workNum[1] = 0; s.error = 0; s.errorEver = 0; s.applied[0] = s.applied[1] = 0;
  for (int i = 0; i < 16; ++i) s.counters[i] = 0;
  s.knnQueries = 0ull; s.knnCandidates = 0ull;
  s.cur[0] = s.cur[1] = 0; s.compact[0] = s.compact[1] = 0; s.repacks[0] = s.repacks[1] = 0;
  s.tabEnd[0] = s.tabEnd[1] = 0; s.buildNum[0] = s.buildNum[1] = 0; s.tabLimit = tabLimit;
  s.trace[0].n_records = s.trace[1].n_records = 0;
}


// ===============================================================================================================
// host side
static cudaError_t lm_alloc(LMDevice* lm, cudaStream_t st) {
  if (lm->allocated) return cudaSuccess;
  const size_t B = lm->B, cap = lm->cap, mapCap = lm->mapCap;
  lm->workCap = mapCap + 2 * cap;  // refilter inputs (<= map + scan) after a cap-sized area for the scan's own sorts
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, st); } };
  A((void**)&lm->st, B * sizeof(LMState));
  for (int i = 0; i < 2; ++i) {
    A((void**)&lm->cubeOff[i], B * 2 * kCubes * sizeof(int)); A((void**)&lm->cubeCnt[i], B * 2 * kCubes * sizeof(int));
    A((void**)&lm->cubeCap[i], B * 2 * kCubes * sizeof(int)); A((void**)&lm->cubeFix[i], B * 2 * kCubes * sizeof(int));
    A((void**)&lm->cubeTab[i], B * 2 * kCubes * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(lm->cubeTab[i], 0xff, B * 2 * kCubes * sizeof(int), st);   // -1: no index
    A((void**)&lm->mapPts[i], B * 2 * mapCap * sizeof(float4));
  }
  if (lm->p.debug_keep_submap) A((void**)&lm->snap, B * 2 * mapCap * sizeof(float4));
  A((void**)&lm->liveList, B * 2 * kCubes * sizeof(short)); A((void**)&lm->liveNum, B * 2 * sizeof(int));
  A((void**)&lm->stack, B * 2 * cap * sizeof(float4)); A((void**)&lm->stackW, B * 2 * cap * sizeof(float4));
  A((void**)&lm->keyA, B * 2 * lm->workCap * 4); A((void**)&lm->valA, B * 2 * lm->workCap * 4);
  A((void**)&lm->keyB, B * 2 * lm->workCap * 4); A((void**)&lm->valB, B * 2 * lm->workCap * 4);
  A((void**)&lm->concat, B * 2 * lm->workCap * sizeof(float4)); A((void**)&lm->staged, B * 2 * lm->workCap * sizeof(float4));
  A((void**)&lm->tabPool, B * 2 * (size_t)kTabSlots * kTabInts * sizeof(int));
  A((void**)&lm->entryHead, B * kCubes * sizeof(short));
  A((void**)&lm->sorted, B * 2 * mapCap * sizeof(float4));
  A((void**)&lm->res.v, B * 2 * 7 * cap * sizeof(double)); A((void**)&lm->res.p, B * 2 * cap * sizeof(float4));
  lm->res.n = (int)cap;
  A((void**)&lm->fitType, 2 * B * 2 * cap);
  A((void**)&lm->gnState, B * sizeof(GNState)); A((void**)&lm->gnPartial, B * kGnTiles * 28 * sizeof(double));
  A((void**)&lm->nnPos, B * 2 * cap * 5 * sizeof(int));
  A((void**)&lm->cubeOf, B * 2 * cap * sizeof(int));
  A((void**)&lm->pose, B * 16 * sizeof(double));
  A((void**)&lm->workOf, B * 2 * kCubes * sizeof(short));
  if (e != cudaSuccess) return e;
  // the index builders keep one 16-bit counter per (z-layer, column) cell in shared memory: 65 KB, opt-in (per device)
  e = cudaFuncSetAttribute(lm_index_build, cudaFuncAttributeMaxDynamicSharedMemorySize, kIndexSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lm_write_back, cudaFuncAttributeMaxDynamicSharedMemorySize, kIndexSmemBytes);
  if (e != cudaSuccess) return e;
  // VLOAM_LM_TAB_SLOTS (tests only): hand out fewer column-table slots so that their recycling runs in a short sequence; the
  // value must exceed the number of occupied valid cubes
  int tabLimit = kTabSlots;
  if (const char* env = getenv("VLOAM_LM_TAB_SLOTS")) { const int v = atoi(env); if (v > 0 && v < kTabSlots) tabLimit = v; }
  VB_LAUNCH(lm->prof, K_LM_MISC, st, lm_init_state<<<(lm->B + 127) / 128, 128, 0, st>>>(lm->st, lm->B, tabLimit));
  lm->allocated = true;
  return cudaGetLastError();
}

cudaError_t lm_create(Profiler* prof, cudaStream_t, int B, int cap, const vloam_lidar_params* p, LMDevice** out) {
  LMDevice* lm = new LMDevice();
  lm->B = B; lm->cap = cap; lm->p = *p; lm->prof = prof;
  lm->mapCap = p->map_capacity_points > 0 ? p->map_capacity_points : (1 << 20);
  *out = lm;  // device memory is allocated on first use (vloam_laser_mapping / vloam_map_set_cube)
  return cudaSuccess;
}
void lm_destroy(LMDevice* lm) {
  if (!lm) return;
  if (lm->allocated) {
    cudaFree(lm->st);
    for (int i = 0; i < 2; ++i) { cudaFree(lm->cubeOff[i]); cudaFree(lm->cubeCnt[i]); cudaFree(lm->cubeCap[i]); cudaFree(lm->cubeFix[i]); cudaFree(lm->cubeTab[i]); cudaFree(lm->mapPts[i]); }
    cudaFree(lm->snap); cudaFree(lm->pubBuf); cudaFree(lm->pubCount); cudaFree(lm->liveList); cudaFree(lm->liveNum);
    cudaFree(lm->stack); cudaFree(lm->stackW); cudaFree(lm->keyA); cudaFree(lm->valA); cudaFree(lm->keyB); cudaFree(lm->valB);
    cudaFree(lm->concat); cudaFree(lm->staged); cudaFree(lm->tabPool); cudaFree(lm->entryHead); cudaFree(lm->sorted);
    cudaFree(lm->res.v); cudaFree(lm->res.p); cudaFree(lm->fitType); cudaFree(lm->gnState); cudaFree(lm->gnPartial); cudaFree(lm->nnPos); cudaFree(lm->cubeOf); cudaFree(lm->pose); cudaFree(lm->workOf);
  }
  delete lm;
}
void lm_reset(LMDevice* lm) { if (lm) lm->reset_valid = true; }
// CUDA-graph replay (capi.cu): the device buffers must exist before a capture starts, and a replayed lm_run leaves the same
// host-side flags behind as a launched one.
cudaError_t lm_ensure_alloc(LMDevice* lm, cudaStream_t st) { return lm_alloc(lm, st); }
void lm_note_run(LMDevice* lm, bool skip_frame) { if (lm && !skip_frame) { lm->reset_valid = false; lm->ran = true; } }
bool lm_graph_safe(const LMDevice* lm) { return lm && !lm->debugStats && lm->snap == nullptr && lm->ncclComm == nullptr; }

cudaError_t lm_run(LMDevice* lm, cudaStream_t st, const SRHeader* hdrCur, const float4* cornerLast, const float4* surfLast,
                   const LOState* lo, bool skip_frame) {
  (void)cudaGetLastError();
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  Profiler* prof = lm->prof;
  const int B = lm->B, cap = lm->cap, mapCap = lm->mapCap;
  if (skip_frame) {
    // :186-190: a skipped frame only refreshes the high-frequency pose; the map state is untouched
    VB_LAUNCH(prof, K_LM_MISC, st, lm_highfreq_pose<<<(B + 127) / 128, 128, 0, st>>>(lm->st, lo, lm->pose, B));
    return cudaGetLastError();
  }
  // the map = cube slabs in mapPts[LMState::cur] addressed by tables[ts].  lm_prepare writes the shifted tables to
  // tables[td]; everything up to the re-filter reads tables[td]; lm_place writes the post-insertion tables back to [ts].
  const int ts = lm->curTab, td = ts ^ 1;
  const CubeTables T_s{lm->cubeOff[ts], lm->cubeCnt[ts], lm->cubeCap[ts], lm->cubeFix[ts], lm->cubeTab[ts]};
  const CubeTables T_d{lm->cubeOff[td], lm->cubeCnt[td], lm->cubeCap[td], lm->cubeFix[td], lm->cubeTab[td]};
  const MapPools pools{{lm->mapPts[0], lm->mapPts[1]}};
  const float lineRes = (float)lm->p.mapping_line_resolution, planeRes = (float)lm->p.mapping_plane_resolution;
  static const bool noMerge = [] { const char* e = getenv("VLOAM_LM_NO_MERGE"); return e && e[0] == '1'; }();   // validation: always re-sort
  const int bitsLine = noMerge ? 0 : vox_axis_bits(lineRes), bitsPlane = noMerge ? 0 : vox_axis_bits(planeRes);
  VB_LAUNCH(prof, K_LM_PREPARE, st, lm_prepare<<<B, kPrepThreads, 0, st>>>(lm->st, lo, T_s, T_d, lm->entryHead, lm->reset_valid ? 1 : 0));
  lm->reset_valid = false;
  if (lm->snap)
    VB_LAUNCH(prof, K_LM_MISC, st, lm_snapshot_submap<<<dim3(256, 2, B), 256, 0, st>>>(lm->st, lm->cubeOff[td], pools, mapCap, lm->snap));
  // C4: VoxelGrid of the scan features (scratch: the first `cap` entries of each key/val slab)
  VB_LAUNCH(prof, K_LM_VOXEL, st, lm_voxel_stack<<<dim3(2, B), 1024, 0, st>>>(lm->st, hdrCur, cornerLast, surfLast, cap, lineRes, planeRes,
                                                                              lm->stack, lm->keyA, lm->valA, lm->keyB, lm->valB, lm->workCap));
  // C5: column index of the valid cubes that do not have one yet (new in the sub-map, seeded, or after a re-pack)
  VB_LAUNCH(prof, K_LM_GRID, st, lm_index_build<<<dim3(32, 2, B), kIndexThreads, kIndexSmemBytes, st>>>(lm->st, T_d, pools, mapCap, lm->tabPool, lm->sorted));
  // C6-C9: outer passes of association + LM
  for (int pass = 0; pass < lm->p.lm_outer_passes; ++pass) {
    const int tp = pass < 2 ? pass : 1;
    static const int knnOcc = [] { const char* e = getenv("VLOAM_LM_KNN_OCC"); return e ? atoi(e) : 6; }();
    if (lm->debugStats)
      VB_LAUNCH(prof, K_LM_ASSOCIATE, st, lm_knn_stats<<<dim3(kKnnGrid, 2, B), kKnnThreads, 0, st>>>(lm->st, lm->stack, cap, T_d, lm->entryHead, lm->tabPool, lm->sorted, mapCap, lm->nnPos, lm->st, lm->shardRank, lm->shardWorld));
    else if (knnOcc == 4)
      VB_LAUNCH(prof, K_LM_ASSOCIATE, st, lm_knn_occ4<<<dim3(kKnnGrid, 2, B), kKnnThreads, 0, st>>>(lm->st, lm->stack, cap, T_d, lm->entryHead, lm->tabPool, lm->sorted, mapCap, lm->nnPos, nullptr, lm->shardRank, lm->shardWorld));
    else
      VB_LAUNCH(prof, K_LM_ASSOCIATE, st, lm_knn<<<dim3(kKnnGrid, 2, B), kKnnThreads, 0, st>>>(lm->st, lm->stack, cap, T_d, lm->entryHead, lm->tabPool, lm->sorted, mapCap, lm->nnPos, nullptr, lm->shardRank, lm->shardWorld));
    VB_LAUNCH(prof, K_LM_FIT, st, lm_fit<<<dim3(64, 2, B), 128, 0, st>>>(lm->st, lm->stack, cap, lm->sorted, mapCap, lm->nnPos, lm->res,
                                                                         lm->fitType + (size_t)tp * B * 2 * cap, tp));
    const bool split = lm->ncclComm != nullptr || lm->p.solver_mode == 2 || (lm->p.solver_mode == 0 && B >= kGnSplitMinBatch);   // vloam_lidar_params::solver_mode
    if (split) {
      // wide solve: one launch over all residual blocks of all streams per evaluation + a warp-per-stream step
      GNProblemView pv{};
      pv.rec = lm->res; pv.blocksPerStream = 2;                  // corner / surf records of stream b: blocks 2 b, 2 b + 1
      pv.count[0] = Strided{&lm->st[0].stackNum[0], sizeof(LMState)}; pv.count[1] = Strided{&lm->st[0].stackNum[1], sizeof(LMState)};
      pv.active = Strided{&lm->st[0].solved, sizeof(LMState)};
      pv.x = Strided{&lm->st[0].parameters[0], sizeof(LMState)};
      pv.trace = Strided{&lm->st[0].trace[tp], sizeof(LMState)};
      launch_gn_solve(prof, st, B, pv, lm->gnState, lm->gnPartial, lm->p.lm_max_iterations, K_LM_ACCUMULATE, K_LM_STEP, lm->ncclComm);
      VB_LAUNCH(prof, K_LM_STEP, st, lm_gn_finish<<<(B + 127) / 128, 128, 0, st>>>(lm->st, B, tp, pass == lm->p.lm_outer_passes - 1 ? 1 : 0));
    } else {
      // one cluster per stream.  Measured on B200 at 16 streams per launch (two launches in flight): 132 / 92 / 73 / 98 us
      // for 1 / 2 / 4 / 8 CTAs per cluster — 8 costs more in barriers and remote reads than it gains.
      static const int forced = [] { const char* e = getenv("VLOAM_LM_CLUSTER"); return e ? atoi(e) : 0; }();
      int cs = forced > 0 ? forced : (B <= 20 ? 4 : B <= 40 ? 2 : 1);
      cs = cs >= 8 ? kLmClusterMax : cs >= 4 ? 4 : cs >= 2 ? 2 : 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      // register budget: from 32 streams per launch on, the SMs a solve occupies are wanted by the other handles' kernels, and the
      // 128-register variant leaves room for them (measured: kernel 3 % slower, step 2 % faster); VLOAM_LM_SOLVE_REGS=128|256 forces one
      static const int regsEnv = [] { const char* e = getenv("VLOAM_LM_SOLVE_REGS"); return e ? atoi(e) : 0; }();
      const bool regs128 = regsEnv == 128 || (regsEnv == 0 && B >= 32);
      if (regs128)
        VB_LAUNCH(prof, K_LM_SOLVE, st, cudaLaunchKernelEx(&cfg, lm_solve_r128, lm->st, lm->res, cap, tp, lm->p.lm_max_iterations,
                                                                   pass == lm->p.lm_outer_passes - 1 ? 1 : 0));
      else
        VB_LAUNCH(prof, K_LM_SOLVE, st, cudaLaunchKernelEx(&cfg, lm_solve, lm->st, lm->res, cap, tp, lm->p.lm_max_iterations,
                                                                   pass == lm->p.lm_outer_passes - 1 ? 1 : 0));
    }
  }
  // C10-C12: transformUpdate, insertion, re-filter of the cubes that can change, write-back
  VB_LAUNCH(prof, K_LM_MISC, st, lm_transform_update<<<(B + 127) / 128, 128, 0, st>>>(lm->st, B));
  VB_LAUNCH(prof, K_LM_INSERT, st, lm_insert_keys<<<dim3(2, B), 1024, 0, st>>>(lm->st, lm->stack, cap, lm->stackW, lm->cubeCnt[td], lm->cubeFix[td],
                                                                               lm->cubeOf, lineRes, planeRes, bitsLine, bitsPlane, lm->keyA, lm->valA, lm->keyB, lm->valB, lm->workCap));
  // the cube-sorted stack indices stay in valA[0 .. n); the per-cube filters use the key/val slabs from offset `cap` on
  VB_LAUNCH(prof, K_LM_REFILTER, st, lm_refilter_merge<<<dim3(32, 2, B), kMergeThreads, 0, st>>>(lm->st, lm->cubeOff[td], lm->cubeCnt[td], lm->cubeFix[td], pools, mapCap,
                                                                                         lm->stackW, cap, lm->valA, lineRes, planeRes, bitsLine, bitsPlane, lm->concat, lm->staged,
                                                                                         lm->keyA + cap, lm->valA + cap, lm->keyB + cap, lm->valB + cap, lm->workCap));
  VB_LAUNCH(prof, K_LM_REFILTER, st, lm_refilter<<<dim3(kSortGrid, 2, B), 1024, 0, st>>>(lm->st, lm->cubeOff[td], lm->cubeCnt[td], lm->cubeFix[td], pools, mapCap,
                                                                                         lm->stackW, cap, lm->valA, lineRes, planeRes, bitsLine, bitsPlane, lm->concat, lm->staged,
                                                                                         lm->keyA + cap, lm->valA + cap, lm->keyB + cap, lm->valB + cap, lm->workCap));
  VB_LAUNCH(prof, K_LM_PLACE, st, lm_place<<<dim3(2, B), 1024, 0, st>>>(lm->st, T_d, T_s, lm->workOf, lm->liveList, lm->liveNum, mapCap));
  VB_LAUNCH(prof, K_LM_PLACE, st, lm_compact_copy<<<dim3(16, 2, B), 256, 0, st>>>(lm->st, lm->cubeOff[td], lm->cubeOff[ts], lm->cubeCnt[ts], lm->liveList,
                                                                                   lm->liveNum, pools, mapCap));
  VB_LAUNCH(prof, K_LM_PLACE, st, lm_write_back<<<dim3(32, 2, B), kIndexThreads, kIndexSmemBytes, st>>>(lm->st, T_d, T_s, pools, mapCap, lm->staged, lm->workCap, lm->tabPool, lm->sorted));
  VB_LAUNCH(prof, K_LM_MISC, st, lm_export_pose<<<(B + 127) / 128, 128, 0, st>>>(lm->st, lm->pose, B));
  lm->ran = true;
  return cudaGetLastError();
}

static cudaError_t lm_fetch_state(LMDevice* lm, cudaStream_t st, int stream, LMState* out) {
  cudaError_t e = cudaMemcpyAsync(out, lm->st + stream, sizeof(LMState), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

cudaError_t lm_get_pose(LMDevice* lm, cudaStream_t st, double* pose_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  std::vector<double> h((size_t)lm->B * 16);
  e = cudaMemcpyAsync(h.data(), lm->pose, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < lm->B; ++b) for (int i = 0; i < 14; ++i) pose_out[(size_t)b * 14 + i] = h[(size_t)b * 16 + i];
  return cudaSuccess;     // per-stream map errors: lm_get_status
}

// status[B][2] = error bits of the last mapped scan, OR of the error bits of every scan so far
cudaError_t lm_get_status(LMDevice* lm, cudaStream_t st, int* status) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  std::vector<LMState> h(lm->B);
  e = cudaMemcpyAsync(h.data(), lm->st, h.size() * sizeof(LMState), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < lm->B; ++b) { status[2 * b] = h[b].error; status[2 * b + 1] = h[b].errorEver | h[b].error; }
  return cudaSuccess;
}

// counters[B][18]: LMState::counters (16, cumulative) + the k-NN debug statistics (queries, candidates; low 31 bits of the per-stream
// totals are enough for the benchmark's short statistics pass)
cudaError_t lm_get_counters(LMDevice* lm, cudaStream_t st, long long* counters) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  std::vector<LMState> h(lm->B);
  e = cudaMemcpyAsync(h.data(), lm->st, h.size() * sizeof(LMState), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < lm->B; ++b) {
    for (int i = 0; i < 16; ++i) counters[(size_t)b * 18 + i] = h[b].counters[i];
    counters[(size_t)b * 18 + 16] = (long long)h[b].knnQueries; counters[(size_t)b * 18 + 17] = (long long)h[b].knnCandidates;
  }
  return cudaSuccess;
}
void lm_set_debug_stats(LMDevice* lm, bool on) { if (lm) lm->debugStats = on; }
void lm_set_nccl(LMDevice* lm, void* comm, int rank, int world) { if (lm) { lm->ncclComm = comm; lm->shardRank = rank; lm->shardWorld = world > 0 ? world : 1; } }

// Queries (indices into the down-sampled corner / surf stack) that produced a residual block in outer pass `pass` of the last scan.
cudaError_t lm_get_queries(LMDevice* lm, cudaStream_t st, int stream, int pass, int kind, int* out, int capacity, int* n_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  LMState S;
  e = lm_fetch_state(lm, st, stream, &S);
  if (e != cudaSuccess) return e;
  int n = 0;
  if (lm->ran && S.solved && S.trace[pass].n_records > 0) {
    std::vector<uint8_t> t((size_t)S.stackNum[kind]);
    if (!t.empty()) {
      e = cudaMemcpyAsync(t.data(), lm->fitType + (((size_t)pass * lm->B + stream) * 2 + kind) * lm->cap, t.size(), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
    }
    for (size_t i = 0; i < t.size(); ++i)
      if (t[i]) { if (out && n < capacity) out[n] = (int)i; ++n; }
  }
  if (n_out) *n_out = n;
  return cudaSuccess;
}

// ---------------------------------------------------------------------------- the clouds LaserMapping::publish sends
// lm_register_cloud: /velodyne_cloud_registered (laser_mapping.cpp:797-805) — the full-resolution scan moved into the map
// frame by pointAssociateToMap with the mapping pose.
__global__ void __launch_bounds__(256) lm_register_cloud(const LMState* __restrict__ stAll, int b, const float4* __restrict__ cloud,
                                                          int n, float4* __restrict__ out) {
  const LMState& st = stAll[b];
  const double q[4] = {st.parameters[0], st.parameters[1], st.parameters[2], st.parameters[3]};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = cloud[i];
    double w[3];
    quat_rotate(q, (double)p.x, (double)p.y, (double)p.z, w);
    out[i] = make_float4((float)(w[0] + st.parameters[4]), (float)(w[1] + st.parameters[5]), (float)(w[2] + st.parameters[6]), p.w);
  }
}
// lm_gather_map: /laser_cloud_map (laser_mapping.cpp:778-790) — for every cube in index order its corner points, then its
// surf points.  Every CTA scans the 4851 cube sizes (cheap) and copies its share of the cubes.
constexpr int kGatherThreads = 256;
__global__ void __launch_bounds__(kGatherThreads) lm_gather_map(const LMState* __restrict__ stAll, int b, const int* __restrict__ cubeOff,
                                                                 const int* __restrict__ cubeCnt, const MapPools pools, int mapCap,
                                                                 float4* __restrict__ out, int* __restrict__ total) {
  __shared__ int s_start[kCubes + 1];
  __shared__ int s_w[kGatherThreads / 32 + 1];
  const LMState& st = stAll[b];
  const int* cnt0 = cubeCnt + ((size_t)b * 2 + 0) * kCubes;
  const int* cnt1 = cubeCnt + ((size_t)b * 2 + 1) * kCubes;
  constexpr int kPer = (kCubes + kGatherThreads - 1) / kGatherThreads;
  int sum = 0;
  for (int k = 0; k < kPer; ++k) { const int c = threadIdx.x * kPer + k; if (c < kCubes) sum += cnt0[c] + cnt1[c]; }
  int run = block_exclusive_scan_nt<kGatherThreads>(sum, s_w);
  for (int k = 0; k < kPer; ++k) { const int c = threadIdx.x * kPer + k; if (c < kCubes) { s_start[c] = run; run += cnt0[c] + cnt1[c]; } }
  if (threadIdx.x == 0) { s_start[kCubes] = s_w[kGatherThreads / 32]; if (blockIdx.x == 0) *total = s_w[kGatherThreads / 32]; }
  __syncthreads();
  for (int c = blockIdx.x; c < kCubes; c += gridDim.x) {
    const int n0 = cnt0[c], n = s_start[c + 1] - s_start[c];
    if (n == 0) continue;
    const float4* src0 = stream_map(pools, st, b, 0, mapCap) + cubeOff[((size_t)b * 2 + 0) * kCubes + c];
    const float4* src1 = stream_map(pools, st, b, 1, mapCap) + cubeOff[((size_t)b * 2 + 1) * kCubes + c];
    float4* dst = out + s_start[c];
    for (int i = threadIdx.x; i < n; i += kGatherThreads) dst[i] = i < n0 ? src0[i] : src1[i - n0];
  }
}
static cudaError_t lm_publish_scratch(LMDevice* lm, size_t points) {
  if (lm->pubCap >= points) return cudaSuccess;
  cudaFree(lm->pubBuf);
  lm->pubBuf = nullptr; lm->pubCap = 0;
  cudaError_t e = cudaMalloc((void**)&lm->pubBuf, points * sizeof(float4));
  if (e == cudaSuccess && !lm->pubCount) e = cudaMalloc((void**)&lm->pubCount, sizeof(int));
  if (e == cudaSuccess) lm->pubCap = points;
  return e;
}
cudaError_t lm_get_registered(LMDevice* lm, cudaStream_t st, int stream, const float4* cloud, int n, float* out, int capacity, int* n_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  if (n_out) *n_out = n;
  const int m = n < capacity ? n : capacity;
  if (m <= 0 || !out) return cudaSuccess;
  e = lm_publish_scratch(lm, (size_t)lm->cap);
  if (e != cudaSuccess) return e;
  lm_register_cloud<<<(m + 255) / 256, 256, 0, st>>>(lm->st, stream, cloud, m, lm->pubBuf);
  e = cudaMemcpyAsync(out, lm->pubBuf, (size_t)m * 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}
cudaError_t lm_get_map_cloud(LMDevice* lm, cudaStream_t st, int stream, float* out, int capacity, int* n_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  e = lm_publish_scratch(lm, (size_t)2 * lm->mapCap);
  if (e != cudaSuccess) return e;
  MapPools pools{{lm->mapPts[0], lm->mapPts[1]}};
  lm_gather_map<<<148, kGatherThreads, 0, st>>>(lm->st, stream, lm->cubeOff[lm->curTab], lm->cubeCnt[lm->curTab], pools, lm->mapCap, lm->pubBuf, lm->pubCount);
  int n = 0;
  e = cudaMemcpyAsync(&n, lm->pubCount, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  if (n_out) *n_out = n;
  const int m = n < capacity ? n : capacity;
  if (m > 0 && out) {
    e = cudaMemcpyAsync(out, lm->pubBuf, (size_t)m * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  return e;
}

cudaError_t lm_get_cloud(LMDevice* lm, cudaStream_t st, int stream, int which, float* out, int capacity, int* n_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  LMState S;
  e = lm_fetch_state(lm, st, stream, &S);
  if (e != cudaSuccess) return e;
  if (which == VLOAM_CLOUD_CORNER_STACK || which == VLOAM_CLOUD_SURF_STACK) {
    const int kind = which == VLOAM_CLOUD_CORNER_STACK ? 0 : 1;
    const int n = S.stackNum[kind], m = n < capacity ? n : capacity;
    if (n_out) *n_out = n;
    if (m > 0 && out) {
      e = cudaMemcpyAsync(out, lm->stack + ((size_t)stream * 2 + kind) * lm->cap, (size_t)m * 16, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    return e;
  }
  // laserCloudCornerFromMap / SurfFromMap of the last solveMapping (the pre-insertion sub-map): kept only on request
  const int kind = which == VLOAM_CLOUD_CORNER_MAP ? 0 : 1;
  if (!lm->snap) return cudaErrorNotSupported;   // vloam_lidar_params::debug_keep_submap was not set
  const int n = lm->ran ? S.fromMapNum[kind] : 0;
  if (n_out) *n_out = n;
  if (!out || capacity <= 0 || n == 0) return cudaSuccess;
  {
    const int m = n < capacity ? n : capacity;
    e = cudaMemcpyAsync(out, lm->snap + ((size_t)stream * 2 + kind) * lm->mapCap, (size_t)m * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e;
  }
}
cudaError_t lm_set_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, const float* xyzi, int n) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  LMState S;
  e = lm_fetch_state(lm, st, stream, &S);
  if (e != cudaSuccess) return e;
  // the association looks a point up through the cube its coordinates fall in (:643-652), like every point the
  // mapping itself inserts: refuse content that lies outside the cube
  for (int i = 0; i < n; ++i) {
    const double v[3] = {(double)xyzi[4 * i], (double)xyzi[4 * i + 1], (double)xyzi[4 * i + 2]};
    const int cen[3] = {S.cenW, S.cenH, S.cenD};
    int cc[3];
    for (int a = 0; a < 3; ++a) { cc[a] = (int)((v[a] + 25.0) / 50.0) + cen[a]; if (v[a] + 25.0 < 0) cc[a]--; }
    if (cc[0] + kCubeW * cc[1] + kCubeW * kCubeH * cc[2] != cube || cc[0] < 0 || cc[0] >= kCubeW || cc[1] < 0 || cc[1] >= kCubeH) return cudaErrorInvalidValue;
  }
  // a fresh slab from the pool's bump allocator (a slab the cube may have had is reclaimed by the next re-pack); the
  // content is arbitrary, so the cube is not marked as a fixed point of its voxel filter
  if ((long long)S.poolEnd[kind] + n > lm->mapCap) return cudaErrorMemoryAllocation;   // map_capacity_points exceeded (capi: VLOAM_E_CAPACITY)
  const size_t tb = ((size_t)stream * 2 + kind) * kCubes + cube;
  const int off = S.poolEnd[kind], zero = 0, none = -1;
  if (n) e = cudaMemcpyAsync(lm->mapPts[S.cur[kind]] + ((size_t)stream * 2 + kind) * lm->mapCap + off, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lm->cubeOff[lm->curTab] + tb, &off, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lm->cubeCnt[lm->curTab] + tb, &n, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lm->cubeCap[lm->curTab] + tb, &n, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lm->cubeFix[lm->curTab] + tb, &zero, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lm->cubeTab[lm->curTab] + tb, &none, sizeof(int), cudaMemcpyHostToDevice, st);
  const int newEnd = off + n;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&lm->st[stream].poolEnd[kind], &newEnd, sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

cudaError_t lm_get_cube(LMDevice* lm, cudaStream_t st, int stream, int kind, int cube, float* out, int capacity, int* n_out) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  const size_t tb = ((size_t)stream * 2 + kind) * kCubes + cube;
  int off = 0, n = 0, cur = 0;
  e = cudaMemcpyAsync(&off, lm->cubeOff[lm->curTab] + tb, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, lm->cubeCnt[lm->curTab] + tb, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&cur, &lm->st[stream].cur[kind], sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  if (n_out) *n_out = n;
  const int m = n < capacity ? n : capacity;
  if (m > 0 && out) {
    e = cudaMemcpyAsync(out, lm->mapPts[cur & 1] + ((size_t)stream * 2 + kind) * lm->mapCap + off, (size_t)m * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  return e;
}

cudaError_t lm_get_info(LMDevice* lm, cudaStream_t st, int* info) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  std::vector<LMState> h(lm->B);
  e = cudaMemcpyAsync(h.data(), lm->st, h.size() * sizeof(LMState), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < lm->B; ++b) {
    int* o = info + (size_t)b * 8;
    o[0] = h[b].cenW; o[1] = h[b].cenH; o[2] = h[b].cenD; o[3] = h[b].validNum;
    o[4] = h[b].fromMapNum[0]; o[5] = h[b].fromMapNum[1]; o[6] = h[b].stackNum[0]; o[7] = h[b].stackNum[1];
  }
  return cudaSuccess;
}

// stats[B][2][10] per stream and kind: points in the map, pool high-water mark, pool index, non-empty cubes, cubes in
// fixed-point form, cubes rewritten by the last scan, re-packs so far, slab capacity in use, points in the rewritten
// cubes, column-table slots in use
cudaError_t lm_get_map_stats(LMDevice* lm, cudaStream_t st, int* stats) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  std::vector<LMState> h(lm->B);
  std::vector<int> cnt((size_t)lm->B * 2 * kCubes), fix(cnt.size()), cp(cnt.size());
  e = cudaMemcpyAsync(h.data(), lm->st, h.size() * sizeof(LMState), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), lm->cubeCnt[lm->curTab], cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(fix.data(), lm->cubeFix[lm->curTab], fix.size() * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cp.data(), lm->cubeCap[lm->curTab], cp.size() * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  for (int b = 0; b < lm->B; ++b)
    for (int kind = 0; kind < 2; ++kind) {
      int* o = stats + ((size_t)b * 2 + kind) * 10;
      long long pts = 0, caps = 0; int occ = 0, fx = 0;
      for (int c = 0; c < kCubes; ++c) {
        const size_t i = ((size_t)b * 2 + kind) * kCubes + c;
        if (cnt[i] > 0) { pts += cnt[i]; caps += cp[i]; ++occ; fx += fix[i] ? 1 : 0; }
      }
      o[0] = (int)pts; o[1] = h[b].poolEnd[kind]; o[2] = h[b].cur[kind]; o[3] = occ; o[4] = fx; o[5] = h[b].workNum[kind];
      o[6] = h[b].repacks[kind]; o[7] = (int)caps;
      long long wpts = 0;
      for (int u = 0; u < h[b].workNum[kind] && u < kMaxWork; ++u) wpts += h[b].workOutN[kind][u];
      o[8] = (int)wpts; o[9] = h[b].tabEnd[kind];
    }
  return cudaSuccess;
}

cudaError_t lm_get_trace(LMDevice* lm, cudaStream_t st, int stream, int pass, double* records, int* info, double* para) {
  cudaError_t e = lm_alloc(lm, st);
  if (e != cudaSuccess) return e;
  LMState S;
  e = lm_fetch_state(lm, st, stream, &S);
  if (e != cudaSuccess) return e;
  const SolveTrace& t = S.trace[pass];
  if (info) { info[0] = t.n_records; info[1] = t.termination; info[2] = t.n_corner; info[3] = t.n_plane; }
  if (para) for (int i = 0; i < 7; ++i) para[i] = t.para[i];
  if (records) for (int i = 0; i < kMaxLMRecords; ++i) {
    const LMRecord& r = t.rec[i];
    double* o = records + (size_t)i * 7;
    o[0] = r.cost; o[1] = r.candidate_cost; o[2] = r.model_cost_change; o[3] = r.relative_decrease; o[4] = r.radius;
    o[5] = r.step_is_valid; o[6] = r.step_is_successful;
  }
  return cudaSuccess;
}

}  // namespace vb
