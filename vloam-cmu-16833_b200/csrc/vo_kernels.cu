// vloam_b200 — visual odometry depth association + residuals on sm_100a.
//
//   vo_project        VisualOdometry::processPointCloud (visual_odometry.cpp:157-172) + PointCloudUtil::projectPointCloud
//                     (point_cloud_util.cpp:148-174): X~ * cam_T_velo' * rect0_T_cam' * P_rect0' in float with Eigen's
//                     sequential-k accumulation, depth > 0.1 filter, pixel -> 5 px bucket id
//   vo_bucket_sort/fold  PointCloudUtil::downsamplePointCloud (:205-260): the reference's running "mean" divides by the
//                     count *before* the hit and is order dependent, so the points are stably sorted by
//                     bucket and every bucket is folded sequentially in input order
//   vo_query          PointCloudUtil::queryDepth (:302-407): 5 x 5 bucket window, >= 10 occupied, 3 nearest, inverse-
//                     distance weights
//   vo_build_residuals VisualOdometry::solveNlsAll (visual_odometry.cpp:283-416): integer-truncated pixels (Q7), flow gate,
//                     depth lookup, back-projection through P_rect0 (float column-pivoted Householder 3x3)
//   vo_bf_match       ImageUtil::matchDescriptors (image_util.cpp:214-296) in the configuration visual_odometry.cpp:34-37 selects:
//                     brute-force Hamming 2-NN over 256-bit ORB descriptors + the 0.8 ratio test; the train descriptors are
//                     staged in shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier); vo_bf_match_cluster deals a
//                     stream's queries to an 8-CTA thread-block cluster (accepted counts exchanged through DSMEM)
//   vo_solve          CostFunctor32 / CostFunctor22 (ceres_cost_function.h:54-96, 147-185) with analytic Jacobians of
//                     ceres::AngleAxisRotatePoint + ceres::Solve (<= 100 iterations, Huber 0.1, no manifold) in one launch
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/vloam_b200.h"
#include "common.cuh"
#include "cta_sort.cuh"
#include "gn_solver.cuh"
#include "internal.h"
#include "tma_bulk.cuh"

namespace vb {

constexpr int kBucket = 5;                                  // point_cloud_util.h:26,41-42: IMG_WIDTH 1242, IMG_HEIGHT 375
constexpr int kBW = 249, kBH = 75, kBuckets = kBW * kBH;   // ceil(1242 / 5), ceil(375 / 5)

struct VOCalib { float cam_T_velo[16], rect0_T_cam[16], P_rect0[12]; };

struct VOResidual {
  double obs[5];
  int type;  // 0 none, 1 CostFunctor32 (X0(3), x1_bar, y1_bar), 2 CostFunctor22 (x0_bar, y0_bar, x1_bar, y1_bar)
  int pad;
};
struct VOState {
  double x[6];  // angles_0to1, t_0to1
  int counter32, counter22;
  SolveTrace trace;
};

// Counts handed in through device memory (vloam_vo_process_cloud_device, e.g. the lidar handle's input buffer) cannot be
// validated on the host: every kernel clamps them to this handle's capacity, the excess points are ignored.
__device__ __forceinline__ int vo_clamped_count(const int* __restrict__ n_points, int b, int nmax) { return min(max(n_points[b], 0), nmax); }

// grid (ceil(cap / 256), B), block 256
__global__ void __launch_bounds__(256) vo_project(const float* __restrict__ xyz, int stride, size_t slab_floats, const int* __restrict__ n_points,
                                                   VOCalib C, int cap, int nmax, float4* __restrict__ uvd, unsigned* __restrict__ key, unsigned* __restrict__ val) {
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const int n = vo_clamped_count(n_points, b, nmax);
  if (i >= n) return;
  const float* p = xyz + (size_t)b * slab_floats + (size_t)i * stride;
  const float X[4] = {p[0], p[1], p[2], 1.0f};
  float a[4], bb[4], c[3];
#pragma unroll
  for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(X[k], C.cam_T_velo[j * 4 + k])); a[j] = s; }
#pragma unroll
  for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(a[k], C.rect0_T_cam[j * 4 + k])); bb[j] = s; }
#pragma unroll
  for (int j = 0; j < 3; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(bb[k], C.P_rect0[j * 4 + k])); c[j] = s; }
  unsigned k16 = 0xffffu;
  float u = 0.f, v = 0.f;
  if (c[2] > 0.1f) {
    const float inv = __fdiv_rn(1.0f, c[2]);
    u = __fmul_rn(c[0], inv); v = __fmul_rn(c[1], inv);
    const int ix = (int)__fdiv_rn(u, (float)kBucket), iy = (int)__fdiv_rn(v, (float)kBucket);
    if (ix >= 0 && ix < kBW && iy >= 0 && iy < kBH) k16 = (unsigned)(ix * kBH + iy);
  }
  uvd[(size_t)b * cap + i] = make_float4(u, v, c[2], 0.f);
  key[(size_t)b * cap + i] = k16;
  val[(size_t)b * cap + i] = (unsigned)i;
}
// grid (B), block 1024: stable sort of the point indices by bucket id (result left in kA / vA)
__global__ void __launch_bounds__(1024) vo_bucket_sort(const int* __restrict__ n_points, int cap, int nmax, unsigned* kA, unsigned* vA, unsigned* kB, unsigned* vB) {
  __shared__ SortSmem S;
  const int b = blockIdx.x;
  cta_radix_sort(kA + (size_t)b * cap, vA + (size_t)b * cap, kB + (size_t)b * cap, vB + (size_t)b * cap, vo_clamped_count(n_points, b, nmax), 16, S);
}
// grid (ceil(kBuckets / 256), B), block 256: one thread folds one bucket in input order
__global__ void __launch_bounds__(256) vo_bucket_fold(const int* __restrict__ n_points, int cap, int nmax, const unsigned* __restrict__ key,
                                                       const unsigned* __restrict__ val, const float4* __restrict__ uvd, float* __restrict__ bx,
                                                       float* __restrict__ by, float* __restrict__ bd, int* __restrict__ bc) {
  const int b = blockIdx.y, bucket = blockIdx.x * 256 + threadIdx.x;
  if (bucket >= kBuckets) return;
  const int n = vo_clamped_count(n_points, b, nmax);
  const unsigned* k = key + (size_t)b * cap;
  const unsigned* v = val + (size_t)b * cap;
  int lo = 0, hi = n;  // first position with key >= bucket
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (k[mid] < (unsigned)bucket) lo = mid + 1; else hi = mid; }
  float x = 0.f, y = 0.f, d = 0.f;
  int cnt = 0;
  for (int t = lo; t < n && k[t] == (unsigned)bucket; ++t) {
    const float4 p = uvd[(size_t)b * cap + v[t]];
    if (cnt == 0) { x = p.x; y = p.y; d = p.z; }
    else {  // :229-236: bucket += (value - bucket) / count, count = hits before this one
      const float fc = (float)cnt;
      x = __fadd_rn(x, __fdiv_rn(__fsub_rn(p.x, x), fc));
      y = __fadd_rn(y, __fdiv_rn(__fsub_rn(p.y, y), fc));
      d = __fadd_rn(d, __fdiv_rn(__fsub_rn(p.z, d), fc));
    }
    ++cnt;
  }
  const size_t o = (size_t)b * kBuckets + bucket;
  bx[o] = x; by[o] = y; bd[o] = d; bc[o] = cnt;
}

// PointCloudUtil::queryDepth, searching_radius = 2
__device__ float query_depth(const float* __restrict__ bx, const float* __restrict__ by, const float* __restrict__ bd,
                             const int* __restrict__ bc, float x, float y) {
  const int index_x = (int)__fdiv_rn(x, (float)kBucket), index_y = (int)__fdiv_rn(y, (float)kBucket);
  float nd[25], nz[25];
  int cnt = 0;
  for (int ix = index_x - 2; ix <= index_x + 2; ++ix)
    for (int iy = index_y - 2; iy <= index_y + 2; ++iy)
      if (ix >= 0 && ix < kBW && iy >= 0 && iy < kBH && bc[ix * kBH + iy] > 0) {
        const int o = ix * kBH + iy;
        const double dx = (double)__fsub_rn(x, bx[o]), dy = (double)__fsub_rn(y, by[o]);
        nd[cnt] = (float)sqrt(dx * dx + dy * dy);  // std::pow(., 2) in double, sqrt in double, stored as float (:332)
        nz[cnt] = bd[o];
        ++cnt;
      }
  if (cnt < 10) return -1.0f;
  // the three nearest, ties in insertion order (stable)
  int i0 = -1, i1 = -1, i2 = -1;
  for (int r = 0; r < 3; ++r) {
    int best = -1;
    for (int i = 0; i < cnt; ++i) {
      if (i == i0 || i == i1) continue;
      if (best < 0 || nd[i] < nd[best]) best = i;
    }
    if (r == 0) i0 = best; else if (r == 1) i1 = best; else i2 = best;
  }
  const float d0 = nd[i0], d1 = nd[i1], d2 = nd[i2], z0 = nz[i0], z1 = nz[i1], z2 = nz[i2];
  // :382-385, left-to-right float arithmetic
  const float num = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(z0, d1), d2), __fmul_rn(__fmul_rn(z1, d0), d2)), __fmul_rn(__fmul_rn(z2, d0), d1));
  const float den = __fadd_rn(__fadd_rn(__fadd_rn(0.0001f, __fmul_rn(d1, d2)), __fmul_rn(d0, d2)), __fmul_rn(d0, d1));
  return __fdiv_rn(num, den);
}
__global__ void vo_query(const float* __restrict__ bx, const float* __restrict__ by, const float* __restrict__ bd, const int* __restrict__ bc,
                         const float* __restrict__ xy, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = query_depth(bx, by, bd, bc, xy[2 * i], xy[2 * i + 1]);
}

// Eigen MatrixXf(3x3).colPivHouseholderQr().solve(b) in float: the same operations in the same order as
// oracle::colpiv_qr_solve3x3f, written with compile-time indices only (columns are swapped through registers) so that
// nothing is indexed dynamically.
__device__ __forceinline__ void swap_col(float (&A)[3][3], int (&perm)[3], int k, int best) {
  // swap columns k and best (best > k); k is a compile-time constant after unrolling, best is 1 or 2
#pragma unroll
  for (int c = 1; c < 3; ++c)
    if (c == best) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { const float t = A[i][k]; A[i][k] = A[i][c]; A[i][c] = t; }
      const int t = perm[k]; perm[k] = perm[c]; perm[c] = t;
    }
}
__device__ void colpiv_qr_solve3x3f_dev(const float Ain[9], const float bin[3], float x[3]) {
  float A[3][3], b[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    b[i] = bin[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) A[i][c] = Ain[i * 3 + c];
  }
  int perm[3] = {0, 1, 2};
  int rank = 0;
  bool stop = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (stop) continue;
    int best = k;
    float bestn = -1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < k) continue;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) if (i >= k) s = __fadd_rn(s, __fmul_rn(A[i][c], A[i][c]));
      if (s > bestn) { bestn = s; best = c; }
    }
    if (!(bestn > 0.f)) { stop = true; continue; }
    if (best != k) swap_col(A, perm, k, best);
    const float nrm = __fsqrt_rn(bestn);
    const float alpha = A[k][k] > 0 ? -nrm : nrm;
    float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) if (i >= k) v[i] = A[i][k];
    v[k] = __fsub_rn(v[k], alpha);
    float vn = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (i >= k) vn = __fadd_rn(vn, __fmul_rn(v[i], v[i]));
    if (vn > 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < k) continue;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) if (i >= k) s = __fadd_rn(s, __fmul_rn(v[i], A[i][c]));
        s = __fdiv_rn(__fmul_rn(2.0f, s), vn);
#pragma unroll
        for (int i = 0; i < 3; ++i) if (i >= k) A[i][c] = __fsub_rn(A[i][c], __fmul_rn(s, v[i]));
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) if (i >= k) s = __fadd_rn(s, __fmul_rn(v[i], b[i]));
      s = __fdiv_rn(__fmul_rn(2.0f, s), vn);
#pragma unroll
      for (int i = 0; i < 3; ++i) if (i >= k) b[i] = __fsub_rn(b[i], __fmul_rn(s, v[i]));
    }
    ++rank;
  }
  float y[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    if (k >= rank) continue;
    float s = b[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) if (c > k && c < rank) s = __fsub_rn(s, __fmul_rn(A[k][c], y[c]));
    y[k] = __fdiv_rn(s, A[k][k]);
  }
  x[0] = x[1] = x[2] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) if (perm[k] == j) x[j] = y[k];
  }
}

// grid (ceil(maxM / 128), B), block 128: one thread per match
__global__ void __launch_bounds__(128) vo_build_residuals(const float* __restrict__ prev_uv, const float* __restrict__ curr_uv,
                                                           const int* __restrict__ n_matches, int maxM, VOCalib C, int remove_outlier,
                                                           const float* __restrict__ pbx, const float* __restrict__ pby, const float* __restrict__ pbd,
                                                           const int* __restrict__ pbc, VOResidual* __restrict__ res) {
  const int b = blockIdx.y, j = blockIdx.x * 128 + threadIdx.x;
  if (j >= maxM) return;
  VOResidual R;
  R.type = 0; R.pad = 0;
  for (int i = 0; i < 5; ++i) R.obs[i] = 0.0;
  if (j < n_matches[b]) {
    const float* pu = prev_uv + ((size_t)b * maxM + j) * 2;
    const float* cu = curr_uv + ((size_t)b * maxM + j) * 2;
    const int px = (int)pu[0], py = (int)pu[1], cx = (int)cu[0], cy = (int)cu[1];  // Q7: truncation (:291-294)
    bool keep = true;
    if (remove_outlier > 0) {  // :309-314
      const double dx = (double)(px - cx), dy = (double)(py - cy);
      if (dx * dx + dy * dy > (double)(remove_outlier * remove_outlier)) keep = false;
    }
    if (keep) {
      const size_t o = (size_t)b * kBuckets;
      const float depth0 = query_depth(pbx + o, pby + o, pbd + o, pbc + o, (float)px, (float)py);  // :316
      const float K[9] = {C.P_rect0[0], C.P_rect0[1], C.P_rect0[2], C.P_rect0[4], C.P_rect0[5], C.P_rect0[6], C.P_rect0[8], C.P_rect0[9], C.P_rect0[10]};
      float p0[3], p1[3] = {(float)cx, (float)cy, 1.0f}, X0[3], X1[3];
      if (depth0 > 0) {  // :345-367
        p0[0] = __fmul_rn((float)px, depth0); p0[1] = __fmul_rn((float)py, depth0); p0[2] = depth0;
        colpiv_qr_solve3x3f_dev(K, p0, X0);
        colpiv_qr_solve3x3f_dev(K, p1, X1);
        R.type = 1;
        R.obs[0] = (double)X0[0]; R.obs[1] = (double)X0[1]; R.obs[2] = (double)X0[2];
        R.obs[3] = (double)X1[0] / (double)X1[2]; R.obs[4] = (double)X1[1] / (double)X1[2];
      } else {  // :394-414
        p0[0] = (float)px; p0[1] = (float)py; p0[2] = 1.0f;
        colpiv_qr_solve3x3f_dev(K, p0, X0);
        colpiv_qr_solve3x3f_dev(K, p1, X1);
        R.type = 2;
        R.obs[0] = (double)X0[0] / (double)X0[2]; R.obs[1] = (double)X0[1] / (double)X0[2];
        R.obs[2] = (double)X1[0] / (double)X1[2]; R.obs[3] = (double)X1[1] / (double)X1[2];
      }
    }
  }
  res[(size_t)b * maxM + j] = R;
}

// y = R(w) p and M = dy/dw exactly as differentiating ceres::AngleAxisRotatePoint (Rodrigues branch for
// theta^2 > eps: dy/dw = -R [p]x J_r(w); first-order branch otherwise: y = p + w x p, dy/dw = -[p]x).
__device__ void angle_axis_rotate_jac(const double w[3], const double p[3], double y[3], double M[3][3]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double P[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2), c = cos(th), s = sin(th), ti = 1.0 / th;
    const double k[3] = {w[0] * ti, w[1] * ti, w[2] * ti};
    const double kxp[3] = {k[1] * p[2] - k[2] * p[1], k[2] * p[0] - k[0] * p[2], k[0] * p[1] - k[1] * p[0]};
    const double tmp = (k[0] * p[0] + k[1] * p[1] + k[2] * p[2]) * (1.0 - c);
    for (int i = 0; i < 3; ++i) y[i] = p[i] * c + kxp[i] * s + k[i] * tmp;
    // R = c I + s [k]x + (1 - c) k k'
    const double Kx[3][3] = {{0, -k[2], k[1]}, {k[2], 0, -k[0]}, {-k[1], k[0], 0}};
    double R[3][3], Jr[3][3];
    const double hs = sin(0.5 * th);
    const double A = 2.0 * hs * hs / th2;                                   // (1 - cos) / theta^2, cancellation-free
    const double Bc = th < 1e-2 ? (1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0) : (th - s) / (th2 * th);  // (theta - sin) / theta^3
    const double Wx[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        R[i][j] = (i == j ? c : 0.0) + s * Kx[i][j] + (1.0 - c) * k[i] * k[j];
        double w2 = 0.0;
        for (int q = 0; q < 3; ++q) w2 += Wx[i][q] * Wx[q][j];
        Jr[i][j] = (i == j ? 1.0 : 0.0) - A * Wx[i][j] + Bc * w2;
      }
    double RP[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double t = 0; for (int q = 0; q < 3; ++q) t += R[i][q] * P[q][j]; RP[i][j] = t; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double t = 0; for (int q = 0; q < 3; ++q) t += RP[i][q] * Jr[q][j]; M[i][j] = -t; }
  } else {
    y[0] = p[0] + (w[1] * p[2] - w[2] * p[1]); y[1] = p[1] + (w[2] * p[0] - w[0] * p[2]); y[2] = p[2] + (w[0] * p[1] - w[1] * p[0]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = -P[i][j];
  }
}

// grid (B), block 256: VisualOdometry::solveNlsAll's ceres::Solve (visual_odometry.cpp:423)
__global__ void __launch_bounds__(256) vo_solve(VOState* __restrict__ stAll, const VOResidual* __restrict__ res, int maxM, const int* __restrict__ n_matches,
                                                 const double* __restrict__ init, int max_iterations) {
  __shared__ LMShared S;
  __shared__ int s_cnt[2];
  const int b = blockIdx.x;
  VOState& st = stAll[b];
  const VOResidual* rr = res + (size_t)b * maxM;
  const int n = min(n_matches[b], maxM);
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) S.x[i] = init ? init[b * 6 + i] : 0.0;  // :261-281
    S.x[6] = 0.0;
  }
  __syncthreads();
  {
    int a = 0, c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { a += rr[i].type == 1; c += rr[i].type == 2; }
    a = __reduce_add_sync(0xffffffffu, a); c = __reduce_add_sync(0xffffffffu, c);
    if (lane_id() == 0) { atomicAdd(&s_cnt[0], a); atomicAdd(&s_cnt[1], c); }
    __syncthreads();
    if (threadIdx.x == 0) { st.counter32 = s_cnt[0]; st.counter22 = s_cnt[1]; st.trace.n_corner = s_cnt[0]; st.trace.n_plane = s_cnt[1]; }
  }
  auto evaluate = [&](const double* x) {
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    const double w[3] = {x[0], x[1], x[2]}, t[3] = {x[3], x[4], x[5]};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const VOResidual& R = rr[i];
      if (R.type == 1) {  // CostFunctor32
        const double p[3] = {R.obs[0], R.obs[1], R.obs[2]};
        double y[3], M[3][3];
        angle_axis_rotate_jac(w, p, y, M);
        const double X[3] = {y[0] + t[0], y[1] + t[1], y[2] + t[2]};
        const double r0 = X[0] - X[2] * R.obs[3], r1 = X[1] - X[2] * R.obs[4];
        const double wgt = huber_weight(r0 * r0 + r1 * r1, &acc[27]);
        const double A0[3] = {1.0, 0.0, -R.obs[3]}, A1[3] = {0.0, 1.0, -R.obs[4]};
        double J0[6], J1[6];
        for (int c = 0; c < 3; ++c) {
          J0[c] = A0[0] * M[0][c] + A0[1] * M[1][c] + A0[2] * M[2][c];
          J1[c] = A1[0] * M[0][c] + A1[1] * M[1][c] + A1[2] * M[2][c];
          J0[3 + c] = A0[c]; J1[3 + c] = A1[c];
        }
        accum_row(acc, J0, r0, wgt);
        accum_row(acc, J1, r1, wgt);
      } else if (R.type == 2) {  // CostFunctor22: r = x1h . (t x (R x0h))
        const double p[3] = {R.obs[0], R.obs[1], 1.0}, x1[3] = {R.obs[2], R.obs[3], 1.0};
        double y[3], M[3][3];
        angle_axis_rotate_jac(w, p, y, M);
        const double txy[3] = {t[1] * y[2] - t[2] * y[1], t[2] * y[0] - t[0] * y[2], t[0] * y[1] - t[1] * y[0]};
        const double r = x1[0] * txy[0] + x1[1] * txy[1] + x1[2] * txy[2];
        const double wgt = huber_weight(r * r, &acc[27]);
        const double dy[3] = {x1[1] * t[2] - x1[2] * t[1], x1[2] * t[0] - x1[0] * t[2], x1[0] * t[1] - x1[1] * t[0]};  // x1 x t
        const double dt[3] = {y[1] * x1[2] - y[2] * x1[1], y[2] * x1[0] - y[0] * x1[2], y[0] * x1[1] - y[1] * x1[0]};  // y x x1
        double J[6];
        for (int c = 0; c < 3; ++c) { J[c] = dy[0] * M[0][c] + dy[1] * M[1][c] + dy[2] * M[2][c]; J[3 + c] = dt[c]; }
        accum_row(acc, J, r, wgt);
      }
    }
    block_reduce28(acc, S.red, S.scratch);
  };
  lm_solve_block(S, &st.trace, max_iterations, false, evaluate, /*euclid=*/1);
  if (threadIdx.x == 0) for (int i = 0; i < 6; ++i) st.x[i] = S.x[i];
}

}  // namespace vb

// ==================================================================================================================
using namespace vb;

// ---------------------------------------------------------------------------------------------
// VO result -> laser-odometry prior.  VisualOdometry::solveNlsAll stores (angles_0to1, t_0to1) as the tf2 transform
// cam0_curr_T_cam0_last (visual_odometry.cpp:426-430: axis-angle -> quaternion -> basis); VloamTF::VO2VeloAndBase
// turns it into velo_last_VOT_velo_curr = velo_T_cam0 * cam0_curr_T_cam0_last^-1 * velo_T_cam0^-1 (vloam_tf.cpp:59-63),
// which LaserOdometry::solveLO copies into para_q / para_t (laser_odometry.cpp:225-232).  tf2 arithmetic (double)
// restated: Quaternion::setRotation, Matrix3x3::setRotation / getRotation, Transform::operator* / inverse.
struct VOFrame { double R[3][3]; double t[3]; };
__device__ __forceinline__ VOFrame frame_mul(const VOFrame& a, const VOFrame& b) {
  VOFrame o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) o.R[i][j] = a.R[i][0] * b.R[0][j] + a.R[i][1] * b.R[1][j] + a.R[i][2] * b.R[2][j];
    o.t[i] = a.R[i][0] * b.t[0] + a.R[i][1] * b.t[1] + a.R[i][2] * b.t[2] + a.t[i];
  }
  return o;
}
__device__ __forceinline__ VOFrame frame_inv(const VOFrame& a) {
  VOFrame o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) o.R[i][j] = a.R[j][i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) o.t[i] = o.R[i][0] * -a.t[0] + o.R[i][1] * -a.t[1] + o.R[i][2] * -a.t[2];
  return o;
}
__global__ void vo_export_prior(const VOState* __restrict__ st, VOFrame A, int B, double* __restrict__ prior /*[B][7]*/) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* x = st[b].x;
  const double angle = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  double q[4] = {0.0, 0.0, 0.0, 1.0};
  if (angle > 0.0) {
    // tf2::Quaternion::setRotation(axis = angles / angle, angle): s = sin(angle / 2) / |axis|
    const double ax = x[0] / angle, ay = x[1] / angle, az = x[2] / angle;
    const double d = sqrt(ax * ax + ay * ay + az * az);
    const double s = sin(angle * 0.5) / d;
    q[0] = ax * s; q[1] = ay * s; q[2] = az * s; q[3] = cos(angle * 0.5);
  }
  VOFrame T;
  {  // tf2::Matrix3x3::setRotation
    const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double s = 2.0 / d;
    const double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    const double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    T.R[0][0] = 1.0 - (yy + zz); T.R[0][1] = xy - wz; T.R[0][2] = xz + wy;
    T.R[1][0] = xy + wz; T.R[1][1] = 1.0 - (xx + zz); T.R[1][2] = yz - wx;
    T.R[2][0] = xz - wy; T.R[2][1] = yz + wx; T.R[2][2] = 1.0 - (xx + yy);
    T.t[0] = x[3]; T.t[1] = x[4]; T.t[2] = x[5];
  }
  const VOFrame V = frame_mul(frame_mul(A, frame_inv(T)), frame_inv(A));
  // tf2::Matrix3x3::getRotation
  double o[4];
  const double trace = V.R[0][0] + V.R[1][1] + V.R[2][2];
  if (trace > 0.0) {
    double s = sqrt(trace + 1.0);
    o[3] = s * 0.5;
    s = 0.5 / s;
    o[0] = (V.R[2][1] - V.R[1][2]) * s; o[1] = (V.R[0][2] - V.R[2][0]) * s; o[2] = (V.R[1][0] - V.R[0][1]) * s;
  } else {
    const int i = V.R[0][0] < V.R[1][1] ? (V.R[1][1] < V.R[2][2] ? 2 : 1) : (V.R[0][0] < V.R[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(V.R[i][i] - V.R[j][j] - V.R[k][k] + 1.0);
    o[i] = s * 0.5;
    s = 0.5 / s;
    o[3] = (V.R[k][j] - V.R[j][k]) * s; o[j] = (V.R[j][i] + V.R[i][j]) * s; o[k] = (V.R[k][i] + V.R[i][k]) * s;
  }
  double* out = prior + (size_t)b * 7;
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3];
  out[4] = V.t[0]; out[5] = V.t[1]; out[6] = V.t[2];
}

// ---------------------------------------------------------------------------------------------------------------
// vo_bf_match: grid (B), block kMatchThreads, dynamic shared memory kMatchChunk * 32 bytes + an mbarrier.
// cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k = 2) + `m[0].distance < ratio * m[1].distance`
// (image_util.cpp:263-283), for binary descriptors of 32 bytes (ORB, cv::ORB::create()).
//   * a thread owns a query descriptor (8 words in registers) and scans the train descriptors, which the CTA stages
//     in shared memory chunk by chunk with one TMA bulk copy per chunk (all threads then read the same descriptor:
//     a shared-memory broadcast); Hamming distance = 8 x popc(xor);
//   * OpenCV keeps the k best in insertion order — a candidate enters only if strictly closer than the current k-th and
//     stays behind equal distances — i.e. the two smallest (distance, train index) pairs;
//   * accepted matches are compacted in query order (the order of knn_matches), like the push_back loop of :275-281.
constexpr int kMatchThreads = 1024, kMatchChunk = 1024;          // 32 KB of descriptors per chunk
__global__ void __launch_bounds__(kMatchThreads) vo_bf_match(const uint8_t* __restrict__ descQ, const int* __restrict__ nQ,
                                                              const uint8_t* __restrict__ descT, const int* __restrict__ nT, int maxK,
                                                              const float* __restrict__ kpQ, const float* __restrict__ kpT, double ratio,
                                                              int* __restrict__ matches /*[B][maxK][3]*/, int* __restrict__ nMatches,
                                                              float* __restrict__ uvQ, float* __restrict__ uvT /*[B][maxK][2]*/,
                                                              int4* __restrict__ knn /*[B][maxK]: idx0, idx1, d0, d1*/) {
  extern __shared__ __align__(128) unsigned char match_smem[];
  uint4* sT = reinterpret_cast<uint4*>(match_smem);                                         // [kMatchChunk][2]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(match_smem + (size_t)kMatchChunk * 32);
  __shared__ int s_wsum[kMatchThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int nq = min(max(nQ[b], 0), maxK), nt = min(max(nT[b], 0), maxK);
  const uint4* gQ = reinterpret_cast<const uint4*>(descQ + (size_t)b * maxK * 32);
  const unsigned char* gT = descT + (size_t)b * maxK * 32;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_fence_init(); s_base = 0; }
  __syncthreads();
  unsigned phase = 0;
  for (int q0 = 0; q0 < nq; q0 += kMatchThreads) {          // (one round unless a frame has more than 1024 keypoints)
    const int q = q0 + threadIdx.x;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (q < nq) { a0 = gQ[2 * q]; a1 = gQ[2 * q + 1]; }
    int d0 = 0x7fffffff, d1 = 0x7fffffff, i0 = -1, i1 = -1;      // best and second best (distance, train index)
    for (int t0 = 0; t0 < nt; t0 += kMatchChunk) {
      const int m = min(kMatchChunk, nt - t0);
      if (threadIdx.x == 0) {
        fence_proxy_async();                                 // the previous chunk was read through the generic proxy
        mbar_arrive_expect_tx(mbar, (unsigned)m * 32u);
        bulk_g2s(sT, gT + (size_t)t0 * 32, (unsigned)m * 32u, mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1u;
      if (q < nq) {
        for (int t = 0; t < m; ++t) {
          const uint4 b0 = sT[2 * t], b1 = sT[2 * t + 1];
          const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                        __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
          if (d < d1) {                                      // strictly closer than the current second best
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = t0 + t; } else { d1 = d; i1 = t0 + t; }
          }
        }
      }
      __syncthreads();                                       // everyone is done with the chunk before it is refilled
    }
    // ratio test in double, as `float < double * float` evaluates (:277); fewer than two train descriptors: no second
    // neighbour to test against (the reference would index past the end), no match
    if (q < nq) knn[(size_t)b * maxK + q] = make_int4(i0, i1, i0 >= 0 ? d0 : -1, i1 >= 0 ? d1 : -1);
    const bool ok = q < nq && i1 >= 0 && (double)(float)d0 < ratio * (double)(float)d1;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int w = threadIdx.x >> 5, l = lane_id();
    if (l == 0) s_wsum[w] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int i = 0; i < w; ++i) before += s_wsum[i];
    if (ok) {
      const int o = before + __popc(bal & ((1u << l) - 1u));
      int* mo = matches + ((size_t)b * maxK + o) * 3;
      mo[0] = q; mo[1] = i0; mo[2] = d0;
      if (kpQ && kpT) {
        const size_t oq = ((size_t)b * maxK + q) * 2, ot = ((size_t)b * maxK + i0) * 2, oo = ((size_t)b * maxK + o) * 2;
        uvQ[oo] = kpQ[oq]; uvQ[oo + 1] = kpQ[oq + 1];
        uvT[oo] = kpT[ot]; uvT[oo + 1] = kpT[ot + 1];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int i = 0; i < kMatchThreads / 32; ++i) tot += s_wsum[i]; s_base += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) nMatches[b] = s_base;
}

// vo_bf_match_cluster: the same matcher with a stream's queries dealt to the CTAs of a thread-block cluster — grid (kMatchClCtas, B),
// cluster (kMatchClCtas, 1, 1), block kMatchClThreads — so that a small batch still covers the chip (32 streams: 256 CTAs instead of
// 32).  Round by round (kMatchClCtas * kMatchClThreads consecutive queries, CTA `rank` owns a contiguous slice) every CTA scans the
// whole train set from its own TMA-staged copy; the ordered compaction needs the number of accepted matches in the lower-ranked
// CTAs, which each CTA reads from its peers' shared memory (DSMEM) after one cluster barrier per round.
constexpr int kMatchClCtas = 8, kMatchClThreads = 128;
__global__ void __launch_bounds__(kMatchClThreads) vo_bf_match_cluster(const uint8_t* __restrict__ descQ, const int* __restrict__ nQ,
                                                                        const uint8_t* __restrict__ descT, const int* __restrict__ nT, int maxK,
                                                                        const float* __restrict__ kpQ, const float* __restrict__ kpT, double ratio,
                                                                        int* __restrict__ matches, int* __restrict__ nMatches,
                                                                        float* __restrict__ uvQ, float* __restrict__ uvT, int4* __restrict__ knn) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) unsigned char match_smem[];
  uint4* sT = reinterpret_cast<uint4*>(match_smem);                                         // [kMatchChunk][2]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(match_smem + (size_t)kMatchChunk * 32);
  __shared__ int s_wsum[kMatchClThreads / 32];
  __shared__ int s_cnt[2];                           // this CTA's accepted matches of the round, two generations
  const int b = blockIdx.y, rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
  const int nq = min(max(nQ[b], 0), maxK), nt = min(max(nT[b], 0), maxK);
  const uint4* gQ = reinterpret_cast<const uint4*>(descQ + (size_t)b * maxK * 32);
  const unsigned char* gT = descT + (size_t)b * maxK * 32;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); mbar_fence_init(); }
  __syncthreads();
  unsigned phase = 0;
  int accepted = 0, gen = 0;                         // matches accepted in the rounds so far (the same value in every CTA)
  for (int q0 = 0; q0 < nq; q0 += cs * kMatchClThreads) {
    const int q = q0 + rank * kMatchClThreads + (int)threadIdx.x;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (q < nq) { a0 = gQ[2 * q]; a1 = gQ[2 * q + 1]; }
    int d0 = 0x7fffffff, d1 = 0x7fffffff, i0 = -1, i1 = -1;
    for (int t0 = 0; t0 < nt; t0 += kMatchChunk) {
      const int m = min(kMatchChunk, nt - t0);
      if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_arrive_expect_tx(mbar, (unsigned)m * 32u);
        bulk_g2s(sT, gT + (size_t)t0 * 32, (unsigned)m * 32u, mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1u;
      if (q < nq) {
        for (int t = 0; t < m; ++t) {
          const uint4 b0 = sT[2 * t], b1 = sT[2 * t + 1];
          const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                        __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
          if (d < d1) {
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = t0 + t; } else { d1 = d; i1 = t0 + t; }
          }
        }
      }
      __syncthreads();
    }
    if (q < nq) knn[(size_t)b * maxK + q] = make_int4(i0, i1, i0 >= 0 ? d0 : -1, i1 >= 0 ? d1 : -1);
    const bool ok = q < nq && i1 >= 0 && (double)(float)d0 < ratio * (double)(float)d1;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int w = threadIdx.x >> 5, l = lane_id();
    if (l == 0) s_wsum[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int i = 0; i < kMatchClThreads / 32; ++i) tot += s_wsum[i]; s_cnt[gen] = tot; }
    cluster.sync();                                  // (also orders s_cnt within the CTA)
    int before = accepted, total = 0;
    for (int r = 0; r < cs; ++r) {
      const int c = *cluster.map_shared_rank(&s_cnt[gen], r);
      if (r < rank) before += c;
      total += c;
    }
    for (int i = 0; i < w; ++i) before += s_wsum[i];
    if (ok) {
      const int o = before + __popc(bal & ((1u << l) - 1u));
      int* mo = matches + ((size_t)b * maxK + o) * 3;
      mo[0] = q; mo[1] = i0; mo[2] = d0;
      if (kpQ && kpT) {
        const size_t oq = ((size_t)b * maxK + q) * 2, ot = ((size_t)b * maxK + i0) * 2, oo = ((size_t)b * maxK + o) * 2;
        uvQ[oo] = kpQ[oq]; uvQ[oo + 1] = kpQ[oq + 1];
        uvT[oo] = kpT[ot]; uvT[oo + 1] = kpT[ot + 1];
      }
    }
    accepted += total;
    gen ^= 1;
    __syncthreads();                                 // s_wsum is rewritten in the next round
  }
  cluster.sync();                                    // nobody leaves while a peer may still read its count
  if (rank == 0 && threadIdx.x == 0) nMatches[b] = accepted;
}

// The matcher launch: the cluster kernel unless VLOAM_VO_MATCH_CLUSTER=0 asks for the one-CTA-per-stream kernel.
static cudaError_t launch_vo_bf_match(vb::Profiler* prof, cudaStream_t st, int B, const uint8_t* descQ, const int* nQ, const uint8_t* descT, const int* nT,
                                      int maxK, const float* kpQ, const float* kpT, double ratio, int* matches, int* nMatches, float* uvQ, float* uvT,
                                      int4* knn) {
  static const bool use_cluster = [] { const char* e = getenv("VLOAM_VO_MATCH_CLUSTER"); return !(e && atoi(e) == 0); }();
  if (use_cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kMatchClCtas, B); cfg.blockDim = dim3(kMatchClThreads); cfg.dynamicSmemBytes = kMatchChunk * 32 + 16; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kMatchClCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    VB_LAUNCH(prof, K_VO_MATCH, st, e = cudaLaunchKernelEx(&cfg, vo_bf_match_cluster, descQ, nQ, descT, nT, maxK, kpQ, kpT, ratio, matches, nMatches, uvQ, uvT, knn));
    return e;
  }
  VB_LAUNCH(prof, K_VO_MATCH, st, vo_bf_match<<<B, kMatchThreads, kMatchChunk * 32 + 16, st>>>(descQ, nQ, descT, nT, maxK, kpQ, kpT, ratio, matches, nMatches,
                                                                                            uvQ, uvT, knn));
  return cudaGetLastError();
}

struct vloam_vo {
  vloam_ctx* ctx = nullptr;
  int B = 0, cap = 0, maxM = 0;
  long long count = -1;  // VisualOdometry::count (visual_odometry.cpp:28)
  VOCalib calib{};
  bool have_calib = false;
  float* d_in = nullptr; int* d_n = nullptr;
  float4* d_uvd = nullptr;
  unsigned *kA = nullptr, *vA = nullptr, *kB = nullptr, *vB = nullptr;
  float* bx[2] = {nullptr, nullptr}; float* by[2] = {nullptr, nullptr}; float* bd[2] = {nullptr, nullptr}; int* bc[2] = {nullptr, nullptr};
  float* d_prev = nullptr; float* d_curr = nullptr; int* d_nm = nullptr; double* d_init = nullptr;
  VOResidual* d_res = nullptr;
  VOState* d_st = nullptr;
  float* d_q = nullptr; float* d_qo = nullptr;  // query scratch
  // descriptor matching: query / train descriptors [B][maxM][32], keypoints [B][maxM][2], counts, match list [B][maxM][3]
  uint8_t* d_desc[2] = {nullptr, nullptr}; float* d_kp[2] = {nullptr, nullptr}; int* d_nkp[2] = {nullptr, nullptr};
  int* d_matches = nullptr; int* d_nmatch = nullptr; float* d_muv[2] = {nullptr, nullptr}; int4* d_knn = nullptr;
  // ORB description / processImage: per ping-pong slot the key points ORB kept, their descriptors and counts (same layouts),
  // the index each kept key point had in the list given to ORB; upload buffers for host images / key points
  uint8_t* d_fdesc[2] = {nullptr, nullptr}; float* d_fkp[2] = {nullptr, nullptr}; int* d_fn[2] = {nullptr, nullptr}; int* d_fidx = nullptr;
  uint8_t* d_orbimg = nullptr; size_t orbimg_bytes = 0; float* d_orbkp = nullptr; int* d_orbn = nullptr;
  int qcap = 0;
  vb::VODetect* det = nullptr;   // key-point detection (vo_detect.cu), created by the first vloam_vo_detect_corners
  int slot() const { return (int)(count % 2); }
};

namespace {
int vfail(vloam_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) { c->last_error = what; if (e != cudaSuccess) { c->last_error += ": "; c->last_error += cudaGetErrorString(e); } }
  return code;
}
#define VCU(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return vfail((ctx), VLOAM_E_CUDA, #call, e__); } while (0)
}  // namespace

extern "C" {

int vloam_vo_destroy(vloam_vo* h) {
  if (!h) return VLOAM_E_INVALID;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  cudaFree(h->d_in); cudaFree(h->d_n); cudaFree(h->d_uvd); cudaFree(h->kA); cudaFree(h->vA); cudaFree(h->kB); cudaFree(h->vB);
  for (int i = 0; i < 2; ++i) { cudaFree(h->bx[i]); cudaFree(h->by[i]); cudaFree(h->bd[i]); cudaFree(h->bc[i]); }
  cudaFree(h->d_prev); cudaFree(h->d_curr); cudaFree(h->d_nm); cudaFree(h->d_init); cudaFree(h->d_res); cudaFree(h->d_st);
  cudaFree(h->d_q); cudaFree(h->d_qo);
  for (int i = 0; i < 2; ++i) { cudaFree(h->d_desc[i]); cudaFree(h->d_kp[i]); cudaFree(h->d_nkp[i]); cudaFree(h->d_muv[i]); }
  cudaFree(h->d_matches); cudaFree(h->d_nmatch); cudaFree(h->d_knn);
  for (int i = 0; i < 2; ++i) { cudaFree(h->d_fdesc[i]); cudaFree(h->d_fkp[i]); cudaFree(h->d_fn[i]); }
  cudaFree(h->d_fidx); cudaFree(h->d_orbimg); cudaFree(h->d_orbkp); cudaFree(h->d_orbn);
  vb::vo_detect_destroy(h->det);
  delete h;
  return VLOAM_OK;
}

int vloam_vo_create(vloam_ctx* c, int batch, int max_points, int max_matches, vloam_vo** out) {
  if (!c || !out || batch < 1 || max_points < 1 || max_matches < 1) return VLOAM_E_INVALID;
  *out = nullptr;
  VCU(c, cudaSetDevice(c->device));
  vloam_vo* h = new (std::nothrow) vloam_vo();
  if (!h) return VLOAM_E_NOMEM;
  h->ctx = c; h->B = batch; h->cap = (max_points + 255) / 256 * 256; h->maxM = max_matches;
  const size_t B = batch, cap = h->cap, M = max_matches;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); } };
  A((void**)&h->d_in, B * cap * 4 * sizeof(float)); A((void**)&h->d_n, B * sizeof(int));
  A((void**)&h->d_uvd, B * cap * sizeof(float4));
  A((void**)&h->kA, B * cap * 4); A((void**)&h->vA, B * cap * 4); A((void**)&h->kB, B * cap * 4); A((void**)&h->vB, B * cap * 4);
  for (int i = 0; i < 2; ++i) {
    A((void**)&h->bx[i], B * kBuckets * 4); A((void**)&h->by[i], B * kBuckets * 4); A((void**)&h->bd[i], B * kBuckets * 4); A((void**)&h->bc[i], B * kBuckets * 4);
  }
  A((void**)&h->d_prev, B * M * 2 * sizeof(float)); A((void**)&h->d_curr, B * M * 2 * sizeof(float)); A((void**)&h->d_nm, B * sizeof(int));
  A((void**)&h->d_init, B * 6 * sizeof(double)); A((void**)&h->d_res, B * M * sizeof(VOResidual)); A((void**)&h->d_st, B * sizeof(VOState));
  h->qcap = 4096;
  A((void**)&h->d_q, h->qcap * 2 * sizeof(float)); A((void**)&h->d_qo, h->qcap * sizeof(float));
  for (int i = 0; i < 2; ++i) {
    A((void**)&h->d_desc[i], B * M * 32); A((void**)&h->d_kp[i], B * M * 2 * sizeof(float)); A((void**)&h->d_nkp[i], B * sizeof(int));
    A((void**)&h->d_muv[i], B * M * 2 * sizeof(float));
    A((void**)&h->d_fdesc[i], B * M * 32); A((void**)&h->d_fkp[i], B * M * 2 * sizeof(float)); A((void**)&h->d_fn[i], B * sizeof(int));
  }
  A((void**)&h->d_fidx, B * M * sizeof(int)); A((void**)&h->d_orbkp, B * M * 2 * sizeof(float)); A((void**)&h->d_orbn, B * sizeof(int));
  A((void**)&h->d_matches, B * M * 3 * sizeof(int)); A((void**)&h->d_nmatch, B * sizeof(int)); A((void**)&h->d_knn, B * M * sizeof(int4));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(vo_bf_match, cudaFuncAttributeMaxDynamicSharedMemorySize, kMatchChunk * 32 + 16);
  if (e != cudaSuccess) { vloam_vo_destroy(h); return vfail(c, VLOAM_E_CUDA, "vloam_vo_create: allocation", e); }
  *out = h;
  return VLOAM_OK;
}

int vloam_vo_set_calibration(vloam_vo* h, const float* cam_T_velo, const float* rect0_T_cam, const float* P_rect0) {
  if (!h || !cam_T_velo || !rect0_T_cam || !P_rect0) return VLOAM_E_INVALID;
  std::memcpy(h->calib.cam_T_velo, cam_T_velo, 16 * sizeof(float));
  std::memcpy(h->calib.rect0_T_cam, rect0_T_cam, 16 * sizeof(float));
  std::memcpy(h->calib.P_rect0, P_rect0, 12 * sizeof(float));
  h->have_calib = true;
  return VLOAM_OK;
}

int vloam_vo_reset(vloam_vo* h) {  // visual_odometry.cpp:86-90
  if (!h) return VLOAM_E_INVALID;
  ++h->count;
  return VLOAM_OK;
}

static int vo_run_cloud(vloam_vo* h, const float* xyz_dev, const int* n_dev, int stride, size_t slab_points) {
  vloam_ctx* c = h->ctx;
  if (!h->have_calib) return vfail(c, VLOAM_E_STATE, "vloam_vo_process_cloud before vloam_vo_set_calibration");
  if (h->count < 0) return vfail(c, VLOAM_E_STATE, "vloam_vo_process_cloud before vloam_vo_reset");
  const int s = h->slot();
  cudaStream_t st = c->stream;
  (void)cudaGetLastError();   // a stale, unrelated error must not be blamed on the launches below
  const int nmax = (int)(slab_points < (size_t)h->cap ? slab_points : (size_t)h->cap);   // counts are clamped to this in every kernel
  VB_LAUNCH(&c->prof, K_VO_PROJECT, st, vo_project<<<dim3(h->cap / 256, h->B), 256, 0, st>>>(xyz_dev, stride, slab_points * (size_t)stride, n_dev, h->calib, h->cap,
                                                                                           nmax, h->d_uvd, h->kA, h->vA));
  VB_LAUNCH(&c->prof, K_VO_BUCKET, st, vo_bucket_sort<<<h->B, 1024, 0, st>>>(n_dev, h->cap, nmax, h->kA, h->vA, h->kB, h->vB));
  VB_LAUNCH(&c->prof, K_VO_BUCKET, st, vo_bucket_fold<<<dim3((kBuckets + 255) / 256, h->B), 256, 0, st>>>(n_dev, h->cap, nmax, h->kA, h->vA, h->d_uvd, h->bx[s], h->by[s],
                                                                                                        h->bd[s], h->bc[s]));
  VCU(c, cudaGetLastError());
  return VLOAM_OK;
}

int vloam_vo_process_cloud(vloam_vo* h, const float* xyz, const int* n_points, int stride, size_t slab_points) {
  if (!h || !xyz || !n_points || (stride != 3 && stride != 4)) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  for (int b = 0; b < h->B; ++b) {
    if (n_points[b] < 0 || (size_t)n_points[b] > slab_points) return vfail(c, VLOAM_E_INVALID, "n_points[b] exceeds slab_points");
    if (n_points[b] > h->cap) return vfail(c, VLOAM_E_CAPACITY, "cloud larger than max_points");
    if (n_points[b]) VCU(c, cudaMemcpyAsync(h->d_in + (size_t)b * h->cap * stride, xyz + (size_t)b * slab_points * stride,
                                            (size_t)n_points[b] * stride * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  VCU(c, cudaMemcpyAsync(h->d_n, n_points, h->B * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  return vo_run_cloud(h, h->d_in, h->d_n, stride, (size_t)h->cap);
}
int vloam_vo_process_cloud_device(vloam_vo* h, const float* xyz_dev, const int* n_dev, int stride, size_t slab_points) {
  if (!h || !xyz_dev || !n_dev || stride < 3 || stride > 16) return VLOAM_E_INVALID;
  VCU(h->ctx, cudaSetDevice(h->ctx->device));
  return vo_run_cloud(h, xyz_dev, n_dev, stride, slab_points);
}

// slot 0 = current frame, 1 = previous frame
static int vo_abs_slot(const vloam_vo* h, int slot) { return slot == 0 ? h->slot() : 1 - h->slot(); }

int vloam_vo_query_depth(vloam_vo* h, int stream, int slot, const float* xy, int n, float* depth_out) {
  if (!h || stream < 0 || stream >= h->B || (slot != 0 && slot != 1) || n < 0 || (n && (!xy || !depth_out))) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  if (h->count < 0) return vfail(c, VLOAM_E_STATE, "no cloud processed yet");
  const int s = vo_abs_slot(h, slot);
  const size_t o = (size_t)stream * kBuckets;
  for (int base = 0; base < n; base += h->qcap) {
    const int m = n - base < h->qcap ? n - base : h->qcap;
    VCU(c, cudaMemcpyAsync(h->d_q, xy + (size_t)base * 2, (size_t)m * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VB_LAUNCH(&c->prof, K_VO_QUERY, c->stream, vo_query<<<(m + 127) / 128, 128, 0, c->stream>>>(h->bx[s] + o, h->by[s] + o, h->bd[s] + o, h->bc[s] + o, h->d_q, m, h->d_qo));
    VCU(c, cudaMemcpyAsync(depth_out + base, h->d_qo, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VCU(c, cudaStreamSynchronize(c->stream));
  }
  return VLOAM_OK;
}

int vloam_vo_get_buckets(vloam_vo* h, int stream, int slot, float* bx, float* by, float* bd, int* bc) {
  if (!h || stream < 0 || stream >= h->B || (slot != 0 && slot != 1)) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  if (h->count < 0) return vfail(c, VLOAM_E_STATE, "no cloud processed yet");
  const int s = vo_abs_slot(h, slot);
  const size_t o = (size_t)stream * kBuckets;
  if (bx) VCU(c, cudaMemcpyAsync(bx, h->bx[s] + o, kBuckets * 4, cudaMemcpyDeviceToHost, c->stream));
  if (by) VCU(c, cudaMemcpyAsync(by, h->by[s] + o, kBuckets * 4, cudaMemcpyDeviceToHost, c->stream));
  if (bd) VCU(c, cudaMemcpyAsync(bd, h->bd[s] + o, kBuckets * 4, cudaMemcpyDeviceToHost, c->stream));
  if (bc) VCU(c, cudaMemcpyAsync(bc, h->bc[s] + o, kBuckets * 4, cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

static int vo_enqueue_solve(vloam_vo* h, const float* prev_dev, const float* curr_dev, const int* nm_dev, const double* init_dev,
                            int remove_VO_outlier, int max_iterations) {
  vloam_ctx* c = h->ctx;
  cudaStream_t st = c->stream;
  (void)cudaGetLastError();
  const int sp = 1 - h->slot();  // depth of the PREVIOUS frame's cloud is used (depth0, visual_odometry.cpp:316)
  VB_LAUNCH(&c->prof, K_VO_QUERY, st, vo_build_residuals<<<dim3((h->maxM + 127) / 128, h->B), 128, 0, st>>>(prev_dev, curr_dev, nm_dev, h->maxM, h->calib,
                                                                                                          remove_VO_outlier, h->bx[sp], h->by[sp], h->bd[sp],
                                                                                                          h->bc[sp], h->d_res));
  VB_LAUNCH(&c->prof, K_VO_SOLVE, st, vo_solve<<<h->B, 256, 0, st>>>(h->d_st, h->d_res, h->maxM, nm_dev, init_dev, max_iterations));
  VCU(c, cudaGetLastError());
  return VLOAM_OK;
}

int vloam_vo_solve(vloam_vo* h, const float* prev_uv, const float* curr_uv, const int* n_matches, const double* init, int remove_VO_outlier,
                   int max_iterations, double* out) {
  if (!h || !prev_uv || !curr_uv || !n_matches || !out) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  if (h->count < 1) return vfail(c, VLOAM_E_STATE, "vloam_vo_solve needs two processed frames");
  for (int b = 0; b < h->B; ++b) if (n_matches[b] < 0 || n_matches[b] > h->maxM) return vfail(c, VLOAM_E_CAPACITY, "n_matches exceeds max_matches");
  cudaStream_t st = c->stream;
  const size_t M = h->maxM;
  VCU(c, cudaMemcpyAsync(h->d_prev, prev_uv, (size_t)h->B * M * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  VCU(c, cudaMemcpyAsync(h->d_curr, curr_uv, (size_t)h->B * M * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  VCU(c, cudaMemcpyAsync(h->d_nm, n_matches, h->B * sizeof(int), cudaMemcpyHostToDevice, st));
  if (init) VCU(c, cudaMemcpyAsync(h->d_init, init, (size_t)h->B * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  const int r = vo_enqueue_solve(h, h->d_prev, h->d_curr, h->d_nm, init ? h->d_init : nullptr, remove_VO_outlier, max_iterations);
  if (r != VLOAM_OK) return r;
  return vloam_vo_get_result(h, out);
}

int vloam_vo_solve_device_async(vloam_vo* h, const float* prev_uv_dev, const float* curr_uv_dev, const int* n_matches_dev,
                                const double* init_dev, int remove_VO_outlier, int max_iterations) {
  if (!h || !prev_uv_dev || !curr_uv_dev || !n_matches_dev) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  if (h->count < 1) return vfail(c, VLOAM_E_STATE, "vloam_vo_solve needs two processed frames");
  return vo_enqueue_solve(h, prev_uv_dev, curr_uv_dev, n_matches_dev, init_dev, remove_VO_outlier, max_iterations);
}

int vloam_vo_get_result(vloam_vo* h, double* out) {
  if (!h || !out) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  std::vector<VOState> hs(h->B);
  VCU(c, cudaMemcpyAsync(hs.data(), h->d_st, hs.size() * sizeof(VOState), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  for (int b = 0; b < h->B; ++b) {
    for (int i = 0; i < 6; ++i) out[(size_t)b * 8 + i] = hs[b].x[i];
    out[(size_t)b * 8 + 6] = hs[b].counter32; out[(size_t)b * 8 + 7] = hs[b].counter22;
  }
  return VLOAM_OK;
}

int vloam_vo_export_lo_prior(vloam_vo* h, const double* velo_T_cam0, double* prior_dev) {
  if (!h || !velo_T_cam0 || !prior_dev) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  VOFrame A;
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) A.R[i][j] = velo_T_cam0[i * 4 + j]; A.t[i] = velo_T_cam0[i * 4 + 3]; }
  VB_LAUNCH(&c->prof, K_VO_MISC, c->stream, vo_export_prior<<<(h->B + 63) / 64, 64, 0, c->stream>>>(h->d_st, A, h->B, prior_dev));
  VCU(c, cudaGetLastError());
  return VLOAM_OK;
}

int vloam_vo_get_residuals(vloam_vo* h, int stream, int* type, double* obs) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  std::vector<VOResidual> r(h->maxM);
  VCU(c, cudaMemcpyAsync(r.data(), h->d_res + (size_t)stream * h->maxM, r.size() * sizeof(VOResidual), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < h->maxM; ++i) {
    if (type) type[i] = r[i].type;
    if (obs) for (int k = 0; k < 5; ++k) obs[(size_t)i * 5 + k] = r[i].obs[k];
  }
  return VLOAM_OK;
}

// ImageUtil::matchDescriptors (image_util.cpp:214-296): BF + NORM_HAMMING + knnMatch(k = 2) + ratio test
int vloam_vo_match_descriptors(vloam_vo* h, const uint8_t* desc_query, const int* n_query, const uint8_t* desc_train, const int* n_train,
                               const float* kp_query, const float* kp_train, double ratio, int* matches_out, int* n_matches_out) {
  if (!h || !desc_query || !n_query || !desc_train || !n_train || (kp_query == nullptr) != (kp_train == nullptr)) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  const size_t B = h->B, M = h->maxM;
  for (int b = 0; b < h->B; ++b)
    if (n_query[b] < 0 || n_train[b] < 0 || n_query[b] > h->maxM || n_train[b] > h->maxM) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_match_descriptors: more keypoints than max_matches");
  cudaStream_t st = c->stream;
  VCU(c, cudaMemcpyAsync(h->d_desc[0], desc_query, B * M * 32, cudaMemcpyHostToDevice, st));
  VCU(c, cudaMemcpyAsync(h->d_desc[1], desc_train, B * M * 32, cudaMemcpyHostToDevice, st));
  VCU(c, cudaMemcpyAsync(h->d_nkp[0], n_query, B * sizeof(int), cudaMemcpyHostToDevice, st));
  VCU(c, cudaMemcpyAsync(h->d_nkp[1], n_train, B * sizeof(int), cudaMemcpyHostToDevice, st));
  if (kp_query) {
    VCU(c, cudaMemcpyAsync(h->d_kp[0], kp_query, B * M * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    VCU(c, cudaMemcpyAsync(h->d_kp[1], kp_train, B * M * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  (void)cudaGetLastError();
  VCU(c, launch_vo_bf_match(&c->prof, st, h->B, h->d_desc[0], h->d_nkp[0], h->d_desc[1], h->d_nkp[1], h->maxM, kp_query ? h->d_kp[0] : nullptr,
                            kp_query ? h->d_kp[1] : nullptr, ratio, h->d_matches, h->d_nmatch, h->d_muv[0], h->d_muv[1], h->d_knn));
  if (matches_out) VCU(c, cudaMemcpyAsync(matches_out, h->d_matches, B * M * 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_matches_out) VCU(c, cudaMemcpyAsync(n_matches_out, h->d_nmatch, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  VCU(c, cudaStreamSynchronize(st));
  return VLOAM_OK;
}
// Host copy of the matched pixel pairs (the arrays vloam_vo_get_match_buffers names): query_uv / train_uv [batch][max_matches][2].
int vloam_vo_get_match_uv(vloam_vo* h, float* query_uv, float* train_uv) {
  if (!h || !query_uv || !train_uv) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)h->B * h->maxM * 2 * sizeof(float);
  VCU(c, cudaMemcpyAsync(query_uv, h->d_muv[0], bytes, cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaMemcpyAsync(train_uv, h->d_muv[1], bytes, cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}
// Parity read-out of the last vloam_vo_match_descriptors call: knn[batch][max_matches][4] = (trainIdx of the nearest, of the second
// nearest, their Hamming distances) per query row, i.e. the knn_matches of image_util.cpp:263 before the ratio test.
int vloam_vo_get_knn(vloam_vo* h, int* knn) {
  if (!h || !knn) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  VCU(c, cudaMemcpyAsync(knn, h->d_knn, (size_t)h->B * h->maxM * sizeof(int4), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}
// The matched pixel pairs of the last vloam_vo_match_descriptors call with keypoints, in solveNlsAll's layout: feed them to
// vloam_vo_solve_device_async(h, query_uv_dev, train_uv_dev, n_matches_dev, ...) (query = previous frame, train = current frame,
// visual_odometry.cpp:118-119).
int vloam_vo_get_match_buffers(vloam_vo* h, const float** query_uv_dev, const float** train_uv_dev, const int** n_matches_dev) {
  if (!h || !query_uv_dev || !train_uv_dev || !n_matches_dev) return VLOAM_E_INVALID;
  *query_uv_dev = h->d_muv[0]; *train_uv_dev = h->d_muv[1]; *n_matches_dev = h->d_nmatch;
  return VLOAM_OK;
}

int vloam_vo_detect_corners(vloam_vo* h, const uint8_t* images, int height, int width, int max_corners, double quality_level,
                            double min_distance, float* corners_xy, int* n_corners) {
  if (!h || !images || height < 3 || width < 3 || max_corners < 1 || !(quality_level > 0.0) || !(min_distance >= 1.0)) return VLOAM_E_INVALID;
  if ((long long)height * width > (1ll << 28) || height > 65535 || width > 65535) return VLOAM_E_CAPACITY;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  int status = 0;
  VCU(c, vb::vo_detect_run(&h->det, &c->prof, c->stream, h->B, images, height, width, max_corners, quality_level, min_distance, &status));
  if (corners_xy && n_corners) VCU(c, vb::vo_detect_read(h->det, c->stream, corners_xy, n_corners));
  if (status) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_detect_corners: more local maxima than a quarter of the pixels (plateaus of equal response)");
  return VLOAM_OK;
}

int vloam_vo_get_corner_response(vloam_vo* h, int stream, float* out, size_t capacity_pixels) {
  if (!h || !out || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  if (!h->det) return vfail(h->ctx, VLOAM_E_STATE, "vloam_vo_get_corner_response before vloam_vo_detect_corners");
  VCU(h->ctx, cudaSetDevice(h->ctx->device));
  VCU(h->ctx, vb::vo_detect_response(h->det, h->ctx->stream, stream, out, capacity_pixels));
  return VLOAM_OK;
}

int vloam_vo_get_corner_buffers(vloam_vo* h, const float** corners_xy_dev, const int** n_corners_dev) {
  if (!h || !corners_xy_dev || !n_corners_dev) return VLOAM_E_INVALID;
  if (!h->det) return vfail(h->ctx, VLOAM_E_STATE, "vloam_vo_get_corner_buffers before vloam_vo_detect_corners");
  *corners_xy_dev = vb::vo_detect_corners_device(h->det); *n_corners_dev = vb::vo_detect_counts_device(h->det);
  return VLOAM_OK;
}

// ImageUtil::descKeypoints with DescriptorType::ORB (image_util.cpp:162-212): cv::ORB::create()->compute on key points of angle -1
namespace {
// Enqueues the description of `kp_dev` ([B][kp_stride][2], counts n_dev) on `img_dev` into the handle's slot `fs`.
int vo_enqueue_describe(vloam_vo* h, const uint8_t* img_dev, int height, int width, const float* kp_dev, const int* n_dev, int kp_stride, int fs) {
  vloam_ctx* c = h->ctx;
  vb::launch_vo_orb_describe(&c->prof, c->stream, h->B, img_dev, height, width, kp_dev, n_dev, kp_stride, h->maxM, h->d_fkp[fs], h->d_fidx,
                             h->d_fdesc[fs], h->d_fn[fs]);
  VCU(c, cudaGetLastError());
  return VLOAM_OK;
}
int vo_feature_slot(const vloam_vo* h) { return h->count < 0 ? 0 : h->slot(); }
}  // namespace

int vloam_vo_describe_orb(vloam_vo* h, const uint8_t* images, int height, int width, const float* keypoints_xy, const int* n_keypoints,
                          float* kept_xy, int* kept_index, uint8_t* descriptors, int* n_kept) {
  if (!h || (keypoints_xy == nullptr) != (n_keypoints == nullptr)) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const size_t B = h->B, M = h->maxM;
  const uint8_t* img_dev = nullptr;
  if (images) {
    if (height < 1 || width < 1 || (long long)height * width > (1ll << 28)) return VLOAM_E_INVALID;
    const size_t bytes = B * (size_t)height * width;
    if (bytes > h->orbimg_bytes) {
      VCU(c, cudaStreamSynchronize(st));
      cudaFree(h->d_orbimg); h->d_orbimg = nullptr; h->orbimg_bytes = 0;
      VCU(c, cudaMalloc((void**)&h->d_orbimg, bytes));
      h->orbimg_bytes = bytes;
    }
    VCU(c, cudaMemcpyAsync(h->d_orbimg, images, bytes, cudaMemcpyHostToDevice, st));
    img_dev = h->d_orbimg;
  } else {
    if (!h->det) return vfail(c, VLOAM_E_STATE, "vloam_vo_describe_orb: images == NULL needs a previous vloam_vo_detect_corners");
    img_dev = vb::vo_detect_image_device(h->det); height = vb::vo_detect_height(h->det); width = vb::vo_detect_width(h->det);
  }
  const float* kp_dev = nullptr; const int* n_dev = nullptr; int stride = 0;
  if (keypoints_xy) {
    for (int b = 0; b < h->B; ++b) if (n_keypoints[b] < 0 || n_keypoints[b] > h->maxM) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_describe_orb: more keypoints than max_matches");
    VCU(c, cudaMemcpyAsync(h->d_orbkp, keypoints_xy, B * M * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    VCU(c, cudaMemcpyAsync(h->d_orbn, n_keypoints, B * sizeof(int), cudaMemcpyHostToDevice, st));
    kp_dev = h->d_orbkp; n_dev = h->d_orbn; stride = h->maxM;
  } else {
    if (!h->det) return vfail(c, VLOAM_E_STATE, "vloam_vo_describe_orb: keypoints == NULL needs a previous vloam_vo_detect_corners");
    kp_dev = vb::vo_detect_corners_device(h->det); n_dev = vb::vo_detect_counts_device(h->det); stride = vb::vo_detect_max_corners(h->det);
    if (stride > h->maxM) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_describe_orb: the detection's max_corners exceeds max_matches");
  }
  const int fs = vo_feature_slot(h);
  if (int rc = vo_enqueue_describe(h, img_dev, height, width, kp_dev, n_dev, stride, fs)) return rc;
  if (kept_xy) VCU(c, cudaMemcpyAsync(kept_xy, h->d_fkp[fs], B * M * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (kept_index) VCU(c, cudaMemcpyAsync(kept_index, h->d_fidx, B * M * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (descriptors) VCU(c, cudaMemcpyAsync(descriptors, h->d_fdesc[fs], B * M * 32, cudaMemcpyDeviceToHost, st));
  if (n_kept) VCU(c, cudaMemcpyAsync(n_kept, h->d_fn[fs], B * sizeof(int), cudaMemcpyDeviceToHost, st));
  VCU(c, cudaStreamSynchronize(st));
  return VLOAM_OK;
}

// VisualOdometry::processImage (visual_odometry.cpp:92-130) with the reference's selections (ShiTomasi + ORB + BF / kNN 0.8):
// images -> keypoints[i], descriptors[i] -> (count > 0) matches of descriptors[1 - i] (query, previous frame) in descriptors[i] (train).
int vloam_vo_process_image(vloam_vo* h, const uint8_t* images, int height, int width, int* n_keypoints, int* n_matches) {
  if (!h || !images || height < 3 || width < 3) return VLOAM_E_INVALID;
  if ((long long)height * width > (1ll << 28) || height > 65535 || width > 65535) return VLOAM_E_CAPACITY;
  vloam_ctx* c = h->ctx;
  if (h->count < 0) return vfail(c, VLOAM_E_STATE, "vloam_vo_process_image before vloam_vo_reset");
  if (h->maxM < 1024) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_process_image: max_matches must hold the detector's 1024 corners");
  VCU(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const bool fetch = n_keypoints || n_matches;
  int status = 0;
  VCU(c, vb::vo_detect_run(&h->det, &c->prof, st, h->B, images, height, width, 1024, 0.03, 7.5, fetch ? &status : nullptr));      // image_util.cpp:13-26
  if (status) return vfail(c, VLOAM_E_CAPACITY, "vloam_vo_process_image: more local maxima than a quarter of the pixels (plateaus of equal response)");
  const int i = h->slot();
  if (int rc = vo_enqueue_describe(h, vb::vo_detect_image_device(h->det), height, width, vb::vo_detect_corners_device(h->det),
                                   vb::vo_detect_counts_device(h->det), 1024, i)) return rc;
  if (h->count > 0) {
    VCU(c, launch_vo_bf_match(&c->prof, st, h->B, h->d_fdesc[1 - i], h->d_fn[1 - i], h->d_fdesc[i], h->d_fn[i], h->maxM, h->d_fkp[1 - i], h->d_fkp[i], 0.8,
                              h->d_matches, h->d_nmatch, h->d_muv[0], h->d_muv[1], h->d_knn));
  } else {
    VCU(c, cudaMemsetAsync(h->d_nmatch, 0, h->B * sizeof(int), st));
  }
  if (n_keypoints) VCU(c, cudaMemcpyAsync(n_keypoints, h->d_fn[i], h->B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_matches) VCU(c, cudaMemcpyAsync(n_matches, h->d_nmatch, h->B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (fetch) VCU(c, cudaStreamSynchronize(st));
  return VLOAM_OK;
}

// The capacity flag of the last detection (vloam_vo_detect_corners / vloam_vo_process_image), for callers that did not synchronise:
// *overflowed = 1 when a stream's response map had more local maxima than a quarter of its pixels (the corners of that frame are
// then those of a truncated candidate list).  Synchronises the stream.
int vloam_vo_get_detect_status(vloam_vo* h, int* overflowed) {
  if (!h || !overflowed) return VLOAM_E_INVALID;
  if (!h->det) return vfail(h->ctx, VLOAM_E_STATE, "vloam_vo_get_detect_status before a detection");
  VCU(h->ctx, cudaSetDevice(h->ctx->device));
  VCU(h->ctx, vb::vo_detect_status(h->det, h->ctx->stream, overflowed));
  return VLOAM_OK;
}

// keypoints[slot] / descriptors[slot] of the processImage chain (slot 0 = current frame, 1 = previous frame), host copies.
int vloam_vo_get_frame_features(vloam_vo* h, int slot, float* keypoints_xy, uint8_t* descriptors, int* n_keypoints) {
  if (!h || slot < 0 || slot > 1) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  const int fs = slot == 0 ? vo_feature_slot(h) : 1 - vo_feature_slot(h);
  const size_t B = h->B, M = h->maxM;
  if (keypoints_xy) VCU(c, cudaMemcpyAsync(keypoints_xy, h->d_fkp[fs], B * M * 2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (descriptors) VCU(c, cudaMemcpyAsync(descriptors, h->d_fdesc[fs], B * M * 32, cudaMemcpyDeviceToHost, c->stream));
  if (n_keypoints) VCU(c, cudaMemcpyAsync(n_keypoints, h->d_fn[fs], B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

// The match list of the last vloam_vo_match_descriptors / vloam_vo_process_image: matches[batch][max_matches][3], n_matches[batch].
int vloam_vo_get_matches(vloam_vo* h, int* matches, int* n_matches) {
  if (!h || !matches || !n_matches) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  VCU(c, cudaMemcpyAsync(matches, h->d_matches, (size_t)h->B * h->maxM * 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaMemcpyAsync(n_matches, h->d_nmatch, h->B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  return VLOAM_OK;
}

int vloam_vo_get_trace(vloam_vo* h, int stream, double* records, int* info, double* para) {
  if (!h || stream < 0 || stream >= h->B) return VLOAM_E_INVALID;
  vloam_ctx* c = h->ctx;
  VCU(c, cudaSetDevice(c->device));
  VOState st;
  VCU(c, cudaMemcpyAsync(&st, h->d_st + stream, sizeof(VOState), cudaMemcpyDeviceToHost, c->stream));
  VCU(c, cudaStreamSynchronize(c->stream));
  const SolveTrace& t = st.trace;
  if (info) { info[0] = t.n_records; info[1] = t.termination; info[2] = t.n_corner; info[3] = t.n_plane; }
  if (para) for (int i = 0; i < 7; ++i) para[i] = t.para[i];
  if (records) for (int i = 0; i < kMaxLMRecords; ++i) {
    const LMRecord& r = t.rec[i];
    double* o = records + (size_t)i * 7;
    o[0] = r.cost; o[1] = r.candidate_cost; o[2] = r.model_cost_change; o[3] = r.relative_decrease; o[4] = r.radius;
    o[5] = r.step_is_valid; o[6] = r.step_is_successful;
  }
  return VLOAM_OK;
}

}  // extern "C"
