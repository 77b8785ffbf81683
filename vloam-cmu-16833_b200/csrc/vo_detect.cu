// vloam_b200 — key-point detection of the visual-odometry front end on the device:
//   ImageUtil::detKeypoints with DetectorType::ShiTomasi (/root/reference/src/visual_odometry/src/image_util.cpp:11-37)
//   = cv::goodFeaturesToTrack(img, 1024, 0.03, 7.5, Mat(), 5, false, 0.04).
// OpenCV's algorithm (imgproc corner.cpp / featureselect.cpp), restated in oracle/vo_frontend.py and pinned there against cv2:
//   vo_min_eigen         cornerMinEigenVal: 3 x 3 Sobel derivatives (scale folded into the smoothing taps), the products
//                        (Dx^2, Dx Dy, Dy^2), a 5 x 5 box sum in double precision (rows, then columns), the smaller eigenvalue;
//                        per-image maximum on the way out
//   vo_corner_candidates threshold at quality * max, 3 x 3 local maxima away from the border -> unordered candidate list (one
//                        atomic per CTA on the stream's counter)
//   vo_select_corners    sort by (value descending, address descending), then the greedy spacing pass.  The reference walks
//                        the sorted list serially ("keep a corner if no kept corner is closer than min_distance"); the same set
//                        comes out of rounds in which every undecided candidate looks at the earlier candidates within the
//                        radius: any of them kept -> dropped; all of them dropped -> kept; otherwise wait.  The earliest
//                        undecided candidate is always decided, a round decides thousands at once.  The rounds walk a growing
//                        prefix of the ranking and stop once max_corners are kept; with many candidates only the strongest ~4096
//                        (a histogram cut on the response) are ranked at all, the rest only if that prefix runs out.
// Every float operation that OpenCV's vectorised path fuses or does not fuse is written with the explicit intrinsic, so the
// response map carries the oracle's bits.
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/vloam_b200.h"
#include "common.cuh"
#include "cta_sort.cuh"
#include "internal.h"

namespace vb {

struct VODetect {
  int B = 0, H = 0, W = 0, capCand = 0, cells = 0, maxCorners = 0;
  uint8_t* img = nullptr;        // [B][H][W]: the frame the last run detected on (one of imgBuf)
  // Frames are double-buffered and uploaded on a copy stream of their own, so the upload of frame k + 1 (from pinned host
  // memory) overlaps the kernels of frame k: uploaded[j] orders the compute stream after the copy; freed[j], recorded on the
  // compute stream at the start of the NEXT run, tells the copy stream that everything enqueued for the frame in buffer j is done.
  uint8_t* imgBuf[2] = {nullptr, nullptr};
  cudaStream_t copyStream = nullptr;
  cudaEvent_t uploaded[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  bool freedValid[2] = {false, false};
  int cur = -1;                  // buffer of the last run (-1: none yet)
  float* eig = nullptr;          // [B][H][W]
  unsigned* maxBits = nullptr;   // [B] bits of the largest (positive) response
  int* nCand = nullptr;          // [B]
  unsigned *kA = nullptr, *vA = nullptr, *kB = nullptr, *vB = nullptr, *kC = nullptr, *vC = nullptr;   // [B][capCand] sort buffers
  uint8_t* state = nullptr;      // [B][capCand] 0 undecided, 1 kept, 2 dropped
  int* cellStart = nullptr;      // [B][cells + 1]
  int* cellFill = nullptr;       // [B][cells]
  int* cellItems = nullptr;      // [B][capCand] ranks, grouped by cell
  float* corners = nullptr;      // [B][maxCorners][2]
  int* nCorners = nullptr;       // [B]
  int* status = nullptr;         // [B] 1 = candidate list overflowed
};

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

constexpr int kTileW = 64, kTileH = 16, kHalo = 2;                 // 5 x 5 box
constexpr int kCovW = kTileW + 2 * kHalo, kCovH = kTileH + 2 * kHalo;
constexpr int kRawW = kCovW + 2, kRawH = kCovH + 2;                // + the 3 x 3 Sobel reach

// grid (ceil(W / 64), ceil(H / 16), B), block 256
__global__ void __launch_bounds__(256) vo_min_eigen(const uint8_t* __restrict__ imgAll, int H, int W, float* __restrict__ eigAll,
                                                    unsigned* __restrict__ maxBits) {
  __shared__ __align__(16) float cov[3][kCovH][kCovW];
  __shared__ __align__(16) double rs[3][kCovH][kTileW];
  __shared__ float s_max[8];
  float (*raw)[kRawW] = reinterpret_cast<float (*)[kRawW]>(&rs[0][0][0]);      // the staged pixels live in rs until the row sums overwrite them
  const int b = blockIdx.z;
  const uint8_t* img = imgAll + (size_t)b * H * W;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const float s1 = (float)(1.0 / (4.0 * 5.0 * 255.0));            // 1 / (2^(ksize-1) * block * 255), corner.cpp
  const float s2 = __fmul_rn(2.0f, s1);
  const int tail = W - W % 32;                                     // OpenCV's scalar tail columns: no fused operations
  // the tile + halo of the three products.  A CTA whose 22 x 70 pixel footprint lies inside the image stages it in shared memory
  // once (as float) and differentiates from there; a border CTA mirrors every access (BORDER_REFLECT_101 of the box filter applied to
  // the product at the mirrored pixel, and of the Sobel filter around it)
  auto products = [&](int j, int i, int x, float a00, float a01, float a02, float a10, float a12, float a20, float a21, float a22) {
    // Dx: derivative (-1, 0, 1) along x (exact), smoothing (s, 2s, s) along y as fma(s, r[y-1] + r[y+1], 2s * r[y])
    const float r0 = __fsub_rn(a02, a00), r1 = __fsub_rn(a12, a10), r2 = __fsub_rn(a22, a20);
    const float dx = __fmaf_rn(s1, __fadd_rn(r0, r2), __fmul_rn(s2, r1));
    // Dy: smoothing along x per row, then the difference of the rows below and above
    float rowm, rowp;
    if (x < tail) {
      rowm = __fmaf_rn(s1, a02, __fmaf_rn(s2, a01, __fmul_rn(s1, a00)));
      rowp = __fmaf_rn(s1, a22, __fmaf_rn(s2, a21, __fmul_rn(s1, a20)));
    } else {
      rowm = __fadd_rn(__fadd_rn(__fmul_rn(s1, a00), __fmul_rn(s2, a01)), __fmul_rn(s1, a02));
      rowp = __fadd_rn(__fadd_rn(__fmul_rn(s1, a20), __fmul_rn(s2, a21)), __fmul_rn(s1, a22));
    }
    const float dy = __fsub_rn(rowp, rowm);
    cov[0][j][i] = __fmul_rn(dx, dx); cov[1][j][i] = __fmul_rn(dx, dy); cov[2][j][i] = __fmul_rn(dy, dy);
  };
  const bool interior = x0 - kHalo - 1 >= 0 && x0 - kHalo - 1 + kRawW <= W && y0 - kHalo - 1 >= 0 && y0 - kHalo - 1 + kRawH <= H;   // (CTA-uniform)
  if (interior) {
    const uint8_t* src = img + (size_t)(y0 - kHalo - 1) * W + (x0 - kHalo - 1);
    for (int e = threadIdx.x; e < kRawH * kRawW; e += 256) {
      const int j = e / kRawW, i = e - j * kRawW;
      raw[j][i] = (float)src[(size_t)j * W + i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kCovH * kCovW; e += 256) {
      const int j = e / kCovW, i = e - j * kCovW;
      products(j, i, x0 - kHalo + i, raw[j][i], raw[j][i + 1], raw[j][i + 2], raw[j + 1][i], raw[j + 1][i + 2], raw[j + 2][i], raw[j + 2][i + 1], raw[j + 2][i + 2]);
    }
  } else {
    for (int e = threadIdx.x; e < kCovH * kCovW; e += 256) {
      const int j = e / kCovW, i = e % kCovW;
      const int y = reflect101(y0 - kHalo + j, H), x = reflect101(x0 - kHalo + i, W);   // the box filter's border: the product at the mirrored pixel
      const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
      products(j, i, x, img[(size_t)ym * W + xm], img[(size_t)ym * W + x], img[(size_t)ym * W + xp], img[(size_t)y * W + xm], img[(size_t)y * W + xp],
               img[(size_t)yp * W + xm], img[(size_t)yp * W + x], img[(size_t)yp * W + xp]);
    }
  }
  __syncthreads();
  // 5-tap row sums in double, left to right; a thread produces four neighbouring sums from eight products (two 16-byte loads)
  for (int e = threadIdx.x; e < 3 * kCovH * (kTileW / 4); e += 256) {
    const int ch = e / (kCovH * (kTileW / 4)), rem = e - ch * (kCovH * (kTileW / 4)), j = rem / (kTileW / 4), i = (rem - j * (kTileW / 4)) * 4;
    const float4 lo = *reinterpret_cast<const float4*>(&cov[ch][j][i]), hi = *reinterpret_cast<const float4*>(&cov[ch][j][i + 4]);
    const double c[8] = {(double)lo.x, (double)lo.y, (double)lo.z, (double)lo.w, (double)hi.x, (double)hi.y, (double)hi.z, (double)hi.w};
    double out[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      double t = c[o];
#pragma unroll
      for (int d = 1; d < 5; ++d) t = __dadd_rn(t, c[o + d]);
      out[o] = t;
    }
    *reinterpret_cast<double2*>(&rs[ch][j][i]) = make_double2(out[0], out[1]);
    *reinterpret_cast<double2*>(&rs[ch][j][i + 2]) = make_double2(out[2], out[3]);
  }
  __syncthreads();
  float vmax = 0.f;
  for (int e = threadIdx.x; e < kTileH * kTileW; e += 256) {
    const int j = e / kTileW, i = e % kTileW;
    const int y = y0 + j, x = x0 + i;
    if (y >= H || x >= W) continue;
    float box[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      double s = rs[ch][j][i];
#pragma unroll
      for (int d = 1; d < 5; ++d) s = __dadd_rn(s, rs[ch][j + d][i]);
      box[ch] = __double2float_rn(s);
    }
    // calcMinEigenVal: a = Sxx / 2, b = Sxy, c = Syy / 2; (a + c) - sqrt((a - c)^2 + b^2)
    const float a = __fmul_rn(box[0], 0.5f), bb = box[1], c = __fmul_rn(box[2], 0.5f);
    const float t = __fsub_rn(a, c);
    const float v = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(bb, bb))));
    eigAll[(size_t)b * H * W + (size_t)y * W + x] = v;
    vmax = fmaxf(vmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane_id() == 0) s_max[threadIdx.x >> 5] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {                                                             // one atomic per CTA on the stream's maximum
#pragma unroll
    for (int i = 1; i < 8; ++i) vmax = fmaxf(vmax, s_max[i]);
    if (vmax > 0.f) atomicMax(&maxBits[b], __float_as_uint(vmax));                    // positive floats order like their bits
  }
}

// grid (ceil(W / 32), ceil(H / 64), B), block (32, 8): threshold (to zero) + "equal to its 3 x 3 dilation", border excluded.
// A thread tests eight pixels of its column (rows y0 + ty + 8 j); the CTA reserves room for all its candidates with ONE atomic on
// the stream's counter (a counter per stream takes ~10^4 appends per frame: one atomic per warp serialised on that address and
// was two thirds of this kernel's time) and writes them in a CTA-local order — any order will do, vo_select_corners sorts by address.
constexpr int kCandRows = 8;
__global__ void __launch_bounds__(256) vo_corner_candidates(const float* __restrict__ eigAll, int H, int W, const unsigned* __restrict__ maxBits,
                                                            double quality, int capCand, unsigned* __restrict__ keyOfs, unsigned* __restrict__ valBits,
                                                            int* __restrict__ nCand, int* __restrict__ status) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.z;
  const float* eig = eigAll + (size_t)b * H * W;
  const int x = blockIdx.x * 32 + threadIdx.x, w = threadIdx.y, l = threadIdx.x;
  const float thr = (float)((double)__uint_as_float(maxBits[b]) * quality);      // threshold(eig, eig, maxVal * qualityLevel, 0, THRESH_TOZERO)
  float v[kCandRows];
  unsigned mask = 0;
#pragma unroll
  for (int j = 0; j < kCandRows; ++j) {
    const int y = blockIdx.y * (8 * kCandRows) + j * 8 + w;
    v[j] = 0.f;
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
      const float raw = eig[(size_t)y * W + x];
      v[j] = raw > thr ? raw : 0.f;
      if (v[j] != 0.f) {
        bool cand = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const float n = eig[(size_t)(y + dy) * W + (x + dx)];
            if ((n > thr ? n : 0.f) > v[j]) cand = false;
          }
        if (cand) mask |= 1u << j;
      }
    }
  }
  // exclusive position of this thread's candidates inside the CTA: lanes, then warps
  const int mine = __popc(mask);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (l >= o) incl += t; }
  if (l == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0 && l == 0) {
    int tot = 0;
    for (int i = 0; i < 8; ++i) { const int c = s_warp[i]; s_warp[i] = tot; tot += c; }
    s_base = tot ? atomicAdd(&nCand[b], tot) : 0;
  }
  __syncthreads();
  int pos = s_base + s_warp[w] + incl - mine;
#pragma unroll
  for (int j = 0; j < kCandRows; ++j) {
    if (!((mask >> j) & 1u)) continue;
    if (pos < capCand) {
      const int y = blockIdx.y * (8 * kCandRows) + j * 8 + w;
      // ascending sort of the complements = descending value, then descending address (greaterThanPtr, featureselect.cpp)
      keyOfs[(size_t)b * capCand + pos] = (unsigned)(H * W - 1 - (y * W + x));
      valBits[(size_t)b * capCand + pos] = ~__float_as_uint(v[j]);
    } else {
      status[b] = 1;
    }
    ++pos;
  }
}

// grid (B), block 1024, dynamic shared memory = sizeof(SortSmem)
__global__ void __launch_bounds__(1024, 1) vo_select_corners(int H, int W, int capCand, int cells, int gw, int gh, int cell, float minDist2,
                                                          int maxCorners, const int* __restrict__ nCand, const unsigned* __restrict__ maxBits, double quality,
                                                          unsigned* kAall, unsigned* vAall, unsigned* kBall, unsigned* vBall, unsigned* kCall, unsigned* vCall,
                                                          uint8_t* __restrict__ stateAll, int* __restrict__ cellStartAll,
                                                          int* __restrict__ cellFillAll, int* __restrict__ cellItemsAll, float* __restrict__ cornersAll,
                                                          int* __restrict__ nCorners) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem& S = *reinterpret_cast<SortSmem*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nAll = min(nCand[b], capCand);
  unsigned* kA = kAall + (size_t)b * capCand; unsigned* vA = vAall + (size_t)b * capCand;
  unsigned* kB = kBall + (size_t)b * capCand; unsigned* vB = vBall + (size_t)b * capCand;
  unsigned* kC = kCall + (size_t)b * capCand; unsigned* vC = vCall + (size_t)b * capCand;
  __shared__ int s_cut, s_kept;
  volatile uint8_t* state = stateAll + (size_t)b * capCand;
  int* cellStart = cellStartAll + (size_t)b * (cells + 1);
  int* cellFill = cellFillAll + (size_t)b * cells;
  int* cellItems = cellItemsAll + (size_t)b * capCand;
  float* corners = cornersAll + (size_t)b * maxCorners * 2;
  if (nAll == 0) { if (tid == 0) nCorners[b] = 0; return; }
  int bitsOfs = 1;
  while ((1ll << bitsOfs) < (long long)H * W) ++bitsOfs;
  // The output is the first maxCorners kept candidates of the ranking and the spacing pass below walks a prefix of it, so when
  // there are many candidates only the strongest kCutTarget (plus the rest of the response bin the cut falls into) are ranked:
  // a 4096-bin histogram of the (complemented) response bits between the frame's maximum and the threshold gives the bin, the
  // candidates up to it are copied aside and sorted.  Should that prefix run out before maxCorners are kept (dense clusters), the
  // whole list is ranked in a second attempt — the source list (kA, vA) is still intact then.
  constexpr int kCutTarget = 4096, kCutBins = 4096;
  for (int attempt = nAll > 2 * kCutTarget ? 0 : 1; attempt < 2; ++attempt) {
  int n = nAll;
  unsigned* kS = kA; unsigned* vS = vA;
  if (attempt == 0) {
    const float vmax = __uint_as_float(maxBits[b]);
    const unsigned lo = ~maxBits[b], hi = ~__float_as_uint((float)((double)vmax * quality));     // response' of every candidate lies in [lo, hi)
    int shift = 0;
    while (((hi - lo) >> shift) >= (unsigned)kCutBins) ++shift;
    int* hist = &S.off[0][0];
    for (int i = tid; i < kCutBins; i += 1024) hist[i] = 0;
    if (tid == 0) { s_cut = kCutBins - 1; s_kept = 0; }
    __syncthreads();
    for (int r = tid; r < nAll; r += 1024) atomicAdd(&hist[min((vA[r] - lo) >> shift, (unsigned)(kCutBins - 1))], 1);
    __syncthreads();
    int c4[kCutBins / 1024], sum = 0;
#pragma unroll
    for (int j = 0; j < kCutBins / 1024; ++j) { c4[j] = hist[tid * (kCutBins / 1024) + j]; sum += c4[j]; }
    int run = block_exclusive_scan1024(sum, S);
#pragma unroll
    for (int j = 0; j < kCutBins / 1024; ++j) {
      if (run < kCutTarget && run + c4[j] >= kCutTarget) s_cut = tid * (kCutBins / 1024) + j;      // exactly one bin crosses the target
      run += c4[j];
    }
    __syncthreads();
    const unsigned cutBin = (unsigned)s_cut;
    for (int r = tid; r < nAll; r += 1024) {
      const unsigned v = vA[r];
      if (min((v - lo) >> shift, (unsigned)(kCutBins - 1)) <= cutBin) { const int dst = atomicAdd(&s_kept, 1); kC[dst] = kA[r]; vC[dst] = v; }
    }
    __syncthreads();
    n = s_kept; kS = kC; vS = vC;
    __syncthreads();
  }
  // ---- order: address descending (unique, makes the unordered candidate list canonical), then value descending (stable)
  int cur = cta_radix_sort(kS, vS, kB, vB, n, bitsOfs, S);                       // (key = address', value = response')
  unsigned* k1 = cur ? kB : kS; unsigned* v1 = cur ? vB : vS;
  unsigned* k2 = cur ? kS : kB; unsigned* v2 = cur ? vS : vB;
  __syncthreads();
  cur = cta_radix_sort(v1, k1, v2, k2, n, 32, S);                                // (key = response', value = address')
  const unsigned* ofsSorted = cur ? k2 : k1;                                      // rank -> complemented address
  unsigned* pos = cur ? k1 : k2;                                                   // the other buffer pair is free now:
  unsigned* cellxy = cur ? v1 : v2;                                                // rank -> x | y << 16 and its cell, same packing
  __syncthreads();
  // ---- candidates grouped by 'cell' x 'cell' pixel cells (featureselect.cpp's grid; cell >= min_distance)
  for (int c = tid; c <= cells; c += 1024) { cellStart[c] = 0; if (c < cells) cellFill[c] = 0; }
  __syncthreads();
  for (int r = tid; r < n; r += 1024) {
    const int ofs = H * W - 1 - (int)ofsSorted[r];
    const int y = ofs / W, x = ofs - y * W, xc = x / cell, yc = y / cell;
    pos[r] = (unsigned)x | ((unsigned)y << 16);
    cellxy[r] = (unsigned)xc | ((unsigned)yc << 16);
    atomicAdd(&cellStart[yc * gw + xc], 1);
    state[r] = 0;
  }
  __syncthreads();
  {  // exclusive scan of the cell counts, a contiguous chunk per thread
    const int per = (cells + 1023) / 1024;
    const int c0 = min(tid * per, cells), c1 = min(c0 + per, cells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cellStart[c];
    int run = block_exclusive_scan1024(sum, S);
    for (int c = c0; c < c1; ++c) { const int t = cellStart[c]; cellStart[c] = run; run += t; }
    if (tid == 0) cellStart[cells] = n;
  }
  __syncthreads();
  for (int r = tid; r < n; r += 1024) {
    const unsigned cc = cellxy[r];
    const int c = (int)(cc >> 16) * gw + (int)(cc & 0xffffu);
    cellItems[cellStart[c] + atomicAdd(&cellFill[c], 1)] = r;
  }
  __syncthreads();
  // ---- the greedy spacing pass as rounds (see the file header), over a growing prefix of the ranking: a candidate depends on
  // earlier ranks only and the output is the first maxCorners kept ones, so the pass stops at the first chunk boundary behind
  // which maxCorners candidates are kept (featureselect.cpp breaks out of its loop at the same count)
  constexpr int kChunk = 2048;
  int limit = 0, keptTotal = 0;
  while (limit < n && keptTotal < maxCorners) {
    const int lo = limit, hi = min(n, lo + kChunk);
    for (int round = 0; round <= hi - lo; ++round) {
      int undecided = 0;
      for (int r = lo + tid; r < hi; r += 1024) {
        if (state[r] != 0) continue;
        const unsigned p = pos[r], cc = cellxy[r];
        const int x = (int)(p & 0xffffu), y = (int)(p >> 16), xc = (int)(cc & 0xffffu), yc = (int)(cc >> 16);
        bool dropped = false, wait = false;
        for (int yy = max(yc - 1, 0); yy <= min(yc + 1, gh - 1) && !dropped; ++yy)
          for (int xx = max(xc - 1, 0); xx <= min(xc + 1, gw - 1) && !dropped; ++xx) {
            const int c = yy * gw + xx;
            for (int t = cellStart[c]; t < cellStart[c + 1]; ++t) {
              const int q = cellItems[t];
              if (q >= r) continue;                                  // only earlier candidates can have been kept before this one
              const unsigned pq = pos[q];
              const float dx = (float)(x - (int)(pq & 0xffffu)), dy = (float)(y - (int)(pq >> 16));
              if (!(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < minDist2)) continue;
              const uint8_t sq = state[q];
              if (sq == 1) { dropped = true; break; }
              if (sq == 0) wait = true;
            }
          }
        if (dropped) state[r] = 2;
        else if (!wait) state[r] = 1;
        else undecided = 1;
      }
      if (!__syncthreads_or(undecided)) break;
    }
    for (int r0 = lo; r0 < hi; r0 += 1024) keptTotal += __syncthreads_count(r0 + tid < hi && state[r0 + tid] == 1);
    limit = hi;
  }
  if (attempt == 0 && keptTotal < maxCorners && n < nAll) { __syncthreads(); continue; }      // the prefix ran out: rank everything
  // ---- the first maxCorners kept candidates, in rank order
  {
    const int per = (limit + 1023) / 1024;
    const int r0 = min(tid * per, limit), r1 = min(r0 + per, limit);
    int cnt = 0;
    for (int r = r0; r < r1; ++r) cnt += state[r] == 1 ? 1 : 0;
    int p = block_exclusive_scan1024(cnt, S);
    const int total = S.total;
    for (int r = r0; r < r1 && p < maxCorners; ++r)
      if (state[r] == 1) { const unsigned q = pos[r]; corners[2 * p] = (float)(q & 0xffffu); corners[2 * p + 1] = (float)(q >> 16); ++p; }
    if (tid == 0) nCorners[b] = min(total, maxCorners);
  }
  break;
  }
}

}  // namespace

void vo_detect_destroy(VODetect* d) {
  if (!d) return;
  for (int j = 0; j < 2; ++j) {
    cudaFree(d->imgBuf[j]);
    if (d->uploaded[j]) cudaEventDestroy(d->uploaded[j]);
    if (d->freed[j]) cudaEventDestroy(d->freed[j]);
  }
  if (d->copyStream) { cudaStreamSynchronize(d->copyStream); cudaStreamDestroy(d->copyStream); }
  cudaFree(d->eig); cudaFree(d->maxBits); cudaFree(d->nCand); cudaFree(d->kA); cudaFree(d->vA); cudaFree(d->kB); cudaFree(d->vB); cudaFree(d->kC); cudaFree(d->vC);
  cudaFree(d->state); cudaFree(d->cellStart); cudaFree(d->cellFill); cudaFree(d->cellItems); cudaFree(d->corners); cudaFree(d->nCorners); cudaFree(d->status);
  delete d;
}

static cudaError_t vo_detect_alloc(VODetect** pd, int B, int H, int W, int cell, int maxCorners) {
  VODetect* d = *pd;
  const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
  if (d && d->B == B && d->H == H && d->W == W && d->cells == gw * gh && d->maxCorners == maxCorners) return cudaSuccess;
  vo_detect_destroy(d);
  *pd = d = new VODetect();
  d->B = B; d->H = H; d->W = W; d->cells = gw * gh; d->maxCorners = maxCorners;
  d->capCand = (H * W + 3) / 4 + 32;      // 3 x 3 local maxima of distinct values cannot be denser; plateaus of equal values can (reported)
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 4); };
  const size_t px = (size_t)B * H * W, nc = (size_t)B * d->capCand;
  A((void**)&d->imgBuf[0], px); A((void**)&d->imgBuf[1], px); A((void**)&d->eig, px * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->copyStream, cudaStreamNonBlocking);
  for (int j = 0; j < 2; ++j) {
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->uploaded[j], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->freed[j], cudaEventDisableTiming);
  } A((void**)&d->maxBits, B * sizeof(unsigned)); A((void**)&d->nCand, B * sizeof(int));
  A((void**)&d->kA, nc * 4); A((void**)&d->vA, nc * 4); A((void**)&d->kB, nc * 4); A((void**)&d->vB, nc * 4); A((void**)&d->kC, nc * 4); A((void**)&d->vC, nc * 4); A((void**)&d->state, nc);
  A((void**)&d->cellStart, (size_t)B * (d->cells + 1) * sizeof(int)); A((void**)&d->cellFill, (size_t)B * d->cells * sizeof(int));
  A((void**)&d->cellItems, nc * sizeof(int)); A((void**)&d->corners, (size_t)B * maxCorners * 2 * sizeof(float));
  A((void**)&d->nCorners, B * sizeof(int)); A((void**)&d->status, B * sizeof(int));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(vo_select_corners, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
  if (e != cudaSuccess) { vo_detect_destroy(d); *pd = nullptr; }      // (a half-built state must not pass the shape test of the next call)
  return e;
}

// images: [B][H][W] (8-bit), host or device memory (the copy direction is inferred).  Results stay on the device until read with
// vo_detect_read.  status_out == NULL: nothing is read back and the stream is not synchronised (vo_detect_status reads the flag later).
cudaError_t vo_detect_run(VODetect** pd, Profiler* prof, cudaStream_t st, int B, const uint8_t* images, int H, int W, int maxCorners,
                          double quality, double minDistance, int* status_out) {
  const int cell = (int)nearbyint(minDistance) > 0 ? (int)nearbyint(minDistance) : 1;     // cvRound (half to even)
  cudaError_t e = vo_detect_alloc(pd, B, H, W, cell, maxCorners);
  if (e != cudaSuccess) return e;
  VODetect* d = *pd;
  const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
  // this frame goes to the buffer the last one did not use; everything enqueued so far may still read the last one's
  const int j = d->cur < 0 ? 0 : d->cur ^ 1;
  if (d->cur >= 0) {
    e = cudaEventRecord(d->freed[d->cur], st);
    if (e != cudaSuccess) return e;
    d->freedValid[d->cur] = true;
  }
  cudaPointerAttributes pa{};
  const bool pinned_src = cudaPointerGetAttributes(&pa, images) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  (void)cudaGetLastError();
  if (pinned_src) {     // pinned host memory: copy on the copy stream, overlapping the kernels still running for the last frame
    if (d->freedValid[j]) e = cudaStreamWaitEvent(d->copyStream, d->freed[j], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d->imgBuf[j], images, (size_t)B * H * W, cudaMemcpyDefault, d->copyStream);
    if (e == cudaSuccess) e = cudaEventRecord(d->uploaded[j], d->copyStream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, d->uploaded[j], 0);
  } else {              // pageable memory (the driver stages the copy and returns when the source may be reused) or device memory
    e = cudaMemcpyAsync(d->imgBuf[j], images, (size_t)B * H * W, cudaMemcpyDefault, st);      // (whose producer is ordered by this stream)
  }
  d->cur = j; d->img = d->imgBuf[j];
  if (e == cudaSuccess) e = cudaMemsetAsync(d->maxBits, 0, B * sizeof(unsigned), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d->nCand, 0, B * sizeof(int), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d->status, 0, B * sizeof(int), st);
  if (e != cudaSuccess) return e;
  VB_LAUNCH(prof, K_VO_DETECT, st, vo_min_eigen<<<dim3((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH, B), 256, 0, st>>>(d->img, H, W, d->eig, d->maxBits));
  VB_LAUNCH(prof, K_VO_DETECT, st, vo_corner_candidates<<<dim3((W + 31) / 32, (H + 8 * kCandRows - 1) / (8 * kCandRows), B), dim3(32, 8), 0, st>>>(d->eig, H, W, d->maxBits, quality, d->capCand, d->kA, d->vA, d->nCand, d->status));
  VB_LAUNCH(prof, K_VO_DETECT, st, vo_select_corners<<<B, 1024, sizeof(SortSmem), st>>>(H, W, d->capCand, d->cells, gw, gh, cell, (float)(minDistance * minDistance), maxCorners,
                                                                                      d->nCand, d->maxBits, quality, d->kA, d->vA, d->kB, d->vB, d->kC, d->vC, d->state, d->cellStart, d->cellFill,
                                                                                      d->cellItems, d->corners, d->nCorners));
  e = cudaGetLastError();
  if (e != cudaSuccess || !status_out) return e;
  return vo_detect_status(d, st, status_out);
}

// 1 when a stream's candidate list overflowed in the last run (synchronises the stream)
cudaError_t vo_detect_status(VODetect* d, cudaStream_t st, int* status_out) {
  std::vector<int> stt(d->B);
  cudaError_t e = cudaMemcpyAsync(stt.data(), d->status, d->B * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  *status_out = 0;
  if (e == cudaSuccess) for (int b = 0; b < d->B; ++b) *status_out |= stt[b];
  return e;
}

cudaError_t vo_detect_read(VODetect* d, cudaStream_t st, float* corners, int* n) {
  cudaError_t e = cudaMemcpyAsync(corners, d->corners, (size_t)d->B * d->maxCorners * 2 * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(n, d->nCorners, d->B * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

cudaError_t vo_detect_response(VODetect* d, cudaStream_t st, int stream, float* out, size_t pixels) {
  const size_t px = (size_t)d->H * d->W;
  cudaError_t e = cudaMemcpyAsync(out, d->eig + (size_t)stream * px, (pixels < px ? pixels : px) * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

const float* vo_detect_corners_device(const VODetect* d) { return d->corners; }
const int* vo_detect_counts_device(const VODetect* d) { return d->nCorners; }
int vo_detect_height(const VODetect* d) { return d->H; }
int vo_detect_width(const VODetect* d) { return d->W; }
const uint8_t* vo_detect_image_device(const VODetect* d) { return d->img; }
int vo_detect_max_corners(const VODetect* d) { return d->maxCorners; }

}  // namespace vb
