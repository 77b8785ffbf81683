// ORB description on the device (the visual-odometry front end, between detection and matching):
//   ImageUtil::descKeypoints with DescriptorType::ORB (/root/reference/src/visual_odometry/src/image_util.cpp:162-212)
//   = cv::ORB::create()->compute(img, keypoints, descriptors) on the Shi-Tomasi corners of detKeypoints, i.e. on key points of
//   octave 0 and the default angle -1 (image_util.cpp:29-35 sets only pt and size).
//
// What cv::ORB does with such key points, restated in oracle/vo_frontend.py orb_describe and pinned there against cv2 4.13:
//   1. KeyPointsFilter::runByImageBorder(keypoints, size, edgeThreshold = 31): a key point stays when its ROUNDED position lies
//      in [31, cols - 31) x [31, rows - 31); the caller's vector is rewritten to the survivors, in order.
//   2. the level-0 image is blurred: GaussianBlur(7 x 7, sigma 2) on a sub-matrix of the pyramid buffer, which OpenCV routes
//      through sepFilter2D with float taps (not through its 8-bit fixed-point Gaussian): rows (taps 0..6 in order), columns
//      (centre tap, then tap_k * (row[+k] + row[-k])), one round-half-even to 8 bits.  OpenCV's vector loops fuse the
//      multiply-adds, its scalar tail loops do not: in the row pass the columns x >= 32 * (cols / 32) are unfused.
//   3. bit i of the descriptor = blurred(c + a_i) < blurred(c + b_i), c the rounded key point, (a_i, b_i) the i-th pair of the
//      learned pattern (orb_pattern.inc) — steered by the key point's angle, and a rotation by -1 degree moves no sample to
//      another pixel.
//
// vo_orb_describe: one warp per key point.  The samples reach 13 pixels from the centre and the blur 3 more, so the warp
// stages the 33 x 33 patch around its key point in shared memory (always inside the image, by 1.), filters it there (33 x 27
// row sums, 27 x 27 blurred bytes) and evaluates the 256 tests, one descriptor byte per lane: the blurred image is never
// materialised (1024 patches of ~1 KB, L2-resident, instead of a read + write of the whole frame per stream).
// The order-preserving compaction of 1. needs the number of survivors before each key point: every CTA counts them for its
// own eight key points from the (<= 8 KB) key-point list.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/vloam_b200.h"
#include "common.cuh"
#include "internal.h"

namespace vb {
namespace {

constexpr int kOrbEdge = 31;                     // cv::ORB::create(): edgeThreshold
constexpr int kOrbReach = 13;                    // largest |offset| in the pattern
constexpr int kBlurR = 3;                        // 7 x 7
constexpr int kPatch = 2 * (kOrbReach + kBlurR) + 1;     // 33 raw pixels
constexpr int kBlurred = 2 * kOrbReach + 1;               // 27 blurred pixels
constexpr int kPatchStride = 36;
constexpr int kOrbWarps = 8;

// the 256 pairs (x_a, y_a, x_b, y_b)
__device__ const signed char kOrbPattern[256 * 4] = {
#include "orb_pattern.inc"
};

// cv::getGaussianKernel(7, 2, CV_32F): exp(-x^2 / 8) normalised in double, stored as float (taps 3..6; symmetric)
__device__ __forceinline__ float gauss_tap(int i) {
  switch (i) {
    case 0: case 6: return 0x1.1f5f62p-4f;
    case 1: case 5: return 0x1.0c70fcp-3f;
    case 2: case 4: return 0x1.869472p-3f;
    default:        return 0x1.ba95c0p-3f;
  }
}

struct OrbSmem {
  unsigned short offA[256], offB[256];                      // sample positions inside the blurred patch (row * 28 + col)
  unsigned char patch[kOrbWarps][kPatch * kPatchStride];    // raw pixels, later the 27 x 27 blurred ones (stride 28)
  float rows[kOrbWarps][kPatch * kBlurred];                 // row-filtered patch
  int partial[kOrbWarps];
};

}  // namespace

// grid (ceil(maxK / 8), B), block 256.  kp: [B][kpStride][2] key points (cv::KeyPoint::pt), nKp[B].
// out: keptXY [B][maxK][2], keptIdx [B][maxK] (index in the input list), desc [B][maxK][32], nKept [B].
__global__ void __launch_bounds__(kOrbWarps * 32) vo_orb_describe(const uint8_t* __restrict__ imgAll, int H, int W, const float* __restrict__ kpAll,
                                                                   const int* __restrict__ nKp, int kpStride, int maxK,
                                                                   float* __restrict__ keptXY, int* __restrict__ keptIdx,
                                                                   uint8_t* __restrict__ desc, int* __restrict__ nKept) {
  __shared__ OrbSmem sm;
  const int b = blockIdx.y, w = threadIdx.x >> 5, l = lane_id();
  const int n = min(max(nKp[b], 0), min(kpStride, maxK));
  const int first = blockIdx.x * kOrbWarps;
  if (first >= n && !(blockIdx.x == 0 && n == 0)) return;
  const float2* kp = reinterpret_cast<const float2*>(kpAll) + (size_t)b * kpStride;
  auto keeps = [&](float2 p, int& cx, int& cy) {
    cx = __float2int_rn(p.x); cy = __float2int_rn(p.y);      // Rect_<int>::contains(Point(pt)): cvRound, half to even
    return cx >= kOrbEdge && cx < W - kOrbEdge && cy >= kOrbEdge && cy < H - kOrbEdge;
  };
  // survivors before this CTA's key points
  int cnt = 0;
  for (int i = threadIdx.x; i < first; i += blockDim.x) { int cx, cy; cnt += keeps(kp[i], cx, cy) ? 1 : 0; }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (l == 0) sm.partial[w] = cnt;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const signed char* p = kOrbPattern + 4 * i;
    sm.offA[i] = (unsigned short)((p[1] + kOrbReach) * 28 + p[0] + kOrbReach);
    sm.offB[i] = (unsigned short)((p[3] + kOrbReach) * 28 + p[2] + kOrbReach);
  }
  __syncthreads();
  int base = 0;
  for (int i = 0; i < kOrbWarps; ++i) base += sm.partial[i];
  // this CTA's key points: warp w owns first + w
  const int k = first + w;
  int cx = 0, cy = 0;
  float2 pt = make_float2(0.f, 0.f);
  bool keep = false;
  if (k < n) { pt = kp[k]; keep = keeps(pt, cx, cy); }
  unsigned keepMask = 0;                                     // bit i: key point first + i survives
  {
    int ccx, ccy;
    const bool mine = l < kOrbWarps && first + l < n && keeps(kp[first + l], ccx, ccy);
    keepMask = __ballot_sync(0xffffffffu, mine);
  }
  if (nKept && threadIdx.x == 0 && first + kOrbWarps >= n) nKept[b] = base + __popc(keepMask);    // the CTA holding the last key point
  if (!keep) return;                                         // (whole warp: no barrier below)
  const int slot = base + __popc(keepMask & ((1u << w) - 1u));
  const uint8_t* img = imgAll + (size_t)b * H * W;
  unsigned char* patch = sm.patch[w];
  float* rows = sm.rows[w];
  // 33 x 33 raw pixels around (cx, cy)
  const uint8_t* src = img + (size_t)(cy - kOrbReach - kBlurR) * W + (cx - kOrbReach - kBlurR);
  for (int i = l; i < kPatch * kPatch; i += 32) {
    const int r = i / kPatch, c = i - r * kPatch;
    patch[r * kPatchStride + c] = __ldg(src + (size_t)r * W + c);
  }
  __syncwarp();
  // rows: taps 0..6 in order; fused multiply-add where OpenCV's vector loop runs, separate roundings in its scalar tail
  const int xFused = 32 * (W / 32), x0 = cx - kOrbReach;
  for (int i = l; i < kPatch * kBlurred; i += 32) {
    const int r = i / kBlurred, c = i - r * kBlurred;
    const unsigned char* p = patch + r * kPatchStride + c;
    float acc = __fmul_rn((float)p[0], gauss_tap(0));
    if (x0 + c < xFused) {
#pragma unroll
      for (int t = 1; t < 7; ++t) acc = __fmaf_rn((float)p[t], gauss_tap(t), acc);
    } else {
#pragma unroll
      for (int t = 1; t < 7; ++t) acc = __fadd_rn(acc, __fmul_rn((float)p[t], gauss_tap(t)));
    }
    rows[i] = acc;
  }
  __syncwarp();
  // columns: centre tap, then the symmetric pairs (always inside the fused range: x <= cols - 19); cvRound + saturate
  for (int i = l; i < kBlurred * kBlurred; i += 32) {
    const int r = i / kBlurred, c = i - r * kBlurred;
    const float* p = rows + (r + kBlurR) * kBlurred + c;
    float acc = __fmul_rn(p[0], gauss_tap(3));
#pragma unroll
    for (int t = 1; t <= 3; ++t) acc = __fmaf_rn(__fadd_rn(p[t * kBlurred], p[-t * kBlurred]), gauss_tap(3 + t), acc);
    patch[r * 28 + c] = (unsigned char)min(max(__float2int_rn(acc), 0), 255);     // (the raw pixels are dead: rows[] holds what is needed)
  }
  __syncwarp();
  // 256 tests, descriptor byte l = pairs 8 l .. 8 l + 7, first pair in bit 0
  unsigned byte = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) byte |= (patch[sm.offA[8 * l + i]] < patch[sm.offB[8 * l + i]] ? 1u : 0u) << i;
  const size_t o = (size_t)b * maxK + slot;
  desc[o * 32 + l] = (uint8_t)byte;
  if (l == 0) {
    reinterpret_cast<float2*>(keptXY)[o] = pt;
    if (keptIdx) keptIdx[o] = k;
  }
}

void launch_vo_orb_describe(Profiler* prof, cudaStream_t st, int B, const uint8_t* img, int H, int W, const float* kp, const int* nKp,
                            int kpStride, int maxK, float* keptXY, int* keptIdx, uint8_t* desc, int* nKept) {
  const int cap = kpStride < maxK ? kpStride : maxK;
  VB_LAUNCH(prof, K_VO_DESCRIBE, st,
            vo_orb_describe<<<dim3((cap + kOrbWarps - 1) / kOrbWarps, B), kOrbWarps * 32, 0, st>>>(img, H, W, kp, nKp, kpStride, maxK, keptXY, keptIdx, desc, nKept));
}

}  // namespace vb
