// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// visual_odometry.hpp: restatement of the in-scope part of visual odometry
// (SURVEY.md §8a rows D1-D7):
//   PointCloudUtil::projectPointCloud   src/visual_odometry/src/point_cloud_util.cpp:148-174
//   PointCloudUtil::downsamplePointCloud                                      :205-260
//   PointCloudUtil::queryDepth                                               :302-407
//   VisualOdometry::processPointCloud   src/visual_odometry/src/visual_odometry.cpp:157-186
//   VisualOdometry::solveNlsAll                                              :254-450
//   CostFunctor32 / CostFunctor22       include/visual_odometry/ceres_cost_function.h:54-96,147-185
// Third party restated: ceres::AngleAxisRotatePoint (ceres/rotation.h), Eigen float
// matrix products (sequential k accumulation) and colPivHouseholderQr 3x3 solve.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

#include "ceres_lm.hpp"
#include "jet.hpp"
#include "types.hpp"

namespace oracle {

struct PointCloudUtil {
  static constexpr int IMG_HEIGHT = 375, IMG_WIDTH = 1242;  // point_cloud_util.h:41-42
  int downsample_grid_size = 5;                              // point_cloud_util.h:26
  float cam_T_velo[16] = {0}, rect0_T_cam[16] = {0}, P_rect0[12] = {0};  // row-major
  std::vector<float> point_cloud_2d;  // M x 3 (u, v, depth)
  int new_width = 0, new_height = 0;
  std::vector<float> bucket_x, bucket_y, bucket_depth;  // [ix * new_height + iy]
  std::vector<int> bucket_count;

  void projectPointCloud(const float* xyz, int n, int stride) {  // :148-174 (+ visual_odometry.cpp:163-170)
    point_cloud_2d.clear();
    for (int i = 0; i < n; ++i) {
      const float X[4] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2], 1.0f};
      float a[4], b[4], c[3];
      for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += X[k] * cam_T_velo[j * 4 + k]; a[j] = s; }
      for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += a[k] * rect0_T_cam[j * 4 + k]; b[j] = s; }
      for (int j = 0; j < 3; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += b[k] * P_rect0[j * 4 + k]; c[j] = s; }
      if (c[2] > 0.1f) {  // Eigen `array() > 0.1` converts the literal to the array's Scalar (float)
        const float inv = 1.0f / c[2];  // Eigen::inverse(col(2)) then colwise product
        point_cloud_2d.push_back(c[0] * inv);
        point_cloud_2d.push_back(c[1] * inv);
        point_cloud_2d.push_back(c[2]);
      }
    }
  }

  void downsamplePointCloud() {  // :205-260
    new_width = (int)std::ceil(static_cast<float>(IMG_WIDTH) / static_cast<float>(downsample_grid_size));
    new_height = (int)std::ceil(static_cast<float>(IMG_HEIGHT) / static_cast<float>(downsample_grid_size));
    const size_t nb = (size_t)new_width * new_height;
    bucket_x.assign(nb, 0.f); bucket_y.assign(nb, 0.f); bucket_depth.assign(nb, 0.f); bucket_count.assign(nb, 0);
    const int m = (int)(point_cloud_2d.size() / 3);
    for (int i = 0; i < m; ++i) {
      const float u = point_cloud_2d[3 * i], v = point_cloud_2d[3 * i + 1], d = point_cloud_2d[3 * i + 2];
      const int index_x = static_cast<int>(u / downsample_grid_size);
      const int index_y = static_cast<int>(v / downsample_grid_size);
      if (index_x >= 0 && index_x < new_width && index_y >= 0 && index_y < new_height) {
        const size_t b = (size_t)index_x * new_height + index_y;
        if (bucket_count[b] == 0) {
          bucket_x[b] = u; bucket_y[b] = v; bucket_depth[b] = d;
        } else {  // Q6: divisor is the count *before* this hit
          bucket_x[b] += (u - bucket_x[b]) / bucket_count[b];
          bucket_y[b] += (v - bucket_y[b]) / bucket_count[b];
          bucket_depth[b] += (d - bucket_depth[b]) / bucket_count[b];
        }
        ++bucket_count[b];
      }
    }
  }

  float queryDepth(const float x, const float y, const int searching_radius = 2) const {  // :302-407
    int index_x = static_cast<int>(x / downsample_grid_size);
    int index_y = static_cast<int>(y / downsample_grid_size);
    struct Nb { float x, y, d, dist; };
    Nb nb[32];
    int cnt = 0;
    for (int ix = index_x - searching_radius; ix <= index_x + searching_radius; ++ix)
      for (int iy = index_y - searching_radius; iy <= index_y + searching_radius; ++iy)
        if (ix >= 0 && ix < new_width && iy >= 0 && iy < new_height && bucket_count[(size_t)ix * new_height + iy] > 0) {
          const size_t b = (size_t)ix * new_height + iy;
          Nb n{bucket_x[b], bucket_y[b], bucket_depth[b], 0.f};
          n.dist = static_cast<float>(std::sqrt(std::pow(static_cast<double>(x - n.x), 2) + std::pow(static_cast<double>(y - n.y), 2)));
          nb[cnt++] = n;
        }
    if (cnt < 10) return -1.0f;
    std::stable_sort(nb, nb + cnt, [](const Nb& a, const Nb& b) { return a.dist < b.dist; });
    float z = (nb[0].d * nb[1].dist * nb[2].dist + nb[1].d * nb[0].dist * nb[2].dist + nb[2].d * nb[0].dist * nb[1].dist) /
              (0.0001f + nb[1].dist * nb[2].dist + nb[0].dist * nb[2].dist + nb[0].dist * nb[1].dist);
    return z;
  }
};

// ceres::AngleAxisRotatePoint (ceres/rotation.h)
template <typename T>
inline void AngleAxisRotatePoint(const T angle_axis[3], const T pt[3], T result[3]) {
  const T theta2 = angle_axis[0] * angle_axis[0] + angle_axis[1] * angle_axis[1] + angle_axis[2] * angle_axis[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {angle_axis[0] * theta_inverse, angle_axis[1] * theta_inverse, angle_axis[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {angle_axis[1] * pt[2] - angle_axis[2] * pt[1], angle_axis[2] * pt[0] - angle_axis[0] * pt[2],
                             angle_axis[0] * pt[1] - angle_axis[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

struct CostFunctor32 {  // ceres_cost_function.h:54-96
  double observed_x0, observed_y0, observed_z0, observed_x1_bar, observed_y1_bar;
  template <typename T>
  bool operator()(const T* const angles, const T* const t, T* residuals) const {
    T X0[3] = {T(observed_x0), T(observed_y0), T(observed_z0)};
    T observed_x1_bar_T = T(observed_x1_bar);
    T observed_y1_bar_T = T(observed_y1_bar);
    T R_dot_X0[3];
    AngleAxisRotatePoint(angles, X0, R_dot_X0);
    R_dot_X0[0] = R_dot_X0[0] + t[0];
    R_dot_X0[1] = R_dot_X0[1] + t[1];
    R_dot_X0[2] = R_dot_X0[2] + t[2];
    residuals[0] = R_dot_X0[0] - R_dot_X0[2] * observed_x1_bar_T;
    residuals[1] = R_dot_X0[1] - R_dot_X0[2] * observed_y1_bar_T;
    return true;
  }
};

struct CostFunctor22 {  // ceres_cost_function.h:147-185
  double observed_x0_bar, observed_y0_bar, observed_x1_bar, observed_y1_bar;
  template <typename T>
  bool operator()(const T* const angles, const T* const t, T* residuals) const {
    T observed_X0_bar_T[3] = {T(observed_x0_bar), T(observed_y0_bar), T(1.0)};
    T observed_X1_bar_T[3] = {T(observed_x1_bar), T(observed_y1_bar), T(1.0)};
    T to1[3];
    AngleAxisRotatePoint(angles, observed_X0_bar_T, to1);
    T c[3] = {t[1] * to1[2] - t[2] * to1[1], t[2] * to1[0] - t[0] * to1[2], t[0] * to1[1] - t[1] * to1[0]};  // ceres::CrossProduct
    residuals[0] = observed_X1_bar_T[0] * c[0] + observed_X1_bar_T[1] * c[1] + observed_X1_bar_T[2] * c[2];  // ceres::DotProduct
    return true;
  }
};

// ceres::AutoDiffCostFunction<Functor, NR, 3, 3>
template <typename Functor, int NR>
struct AutoDiffBlock33 : CostBlock {
  Functor f;
  explicit AutoDiffBlock33(const Functor& f_) : f(f_) {}
  int num_residuals() const override { return NR; }
  void evaluate(const double* x, double* r, double* J) const override {
    if (!J) { f(x, x + 3, r); return; }
    typedef Jet<6> JT;
    JT a[3], t[3], res[NR];
    for (int i = 0; i < 3; ++i) { a[i] = JT(x[i], i); t[i] = JT(x[3 + i], 3 + i); }
    f(a, t, res);
    for (int i = 0; i < NR; ++i) { r[i] = res[i].a; for (int c = 0; c < 6; ++c) J[i * 6 + c] = res[i].v[c]; }
  }
};

// Eigen MatrixXf(3x3).colPivHouseholderQr().solve(b) in float.
inline void colpiv_qr_solve3x3f(const float Ain[9], const float bin[3], float x[3]) {
  float A[3][3], b[3];
  for (int i = 0; i < 3; ++i) { b[i] = bin[i]; for (int c = 0; c < 3; ++c) A[i][c] = Ain[i * 3 + c]; }
  int perm[3] = {0, 1, 2};
  int rank = 0;
  for (int k = 0; k < 3; ++k) {
    int best = k; float bestn = -1.f;
    for (int c = k; c < 3; ++c) { float s = 0; for (int i = k; i < 3; ++i) s += A[i][c] * A[i][c]; if (s > bestn) { bestn = s; best = c; } }
    if (!(bestn > 0.f)) break;
    if (best != k) { for (int i = 0; i < 3; ++i) std::swap(A[i][k], A[i][best]); std::swap(perm[k], perm[best]); }
    const float nrm = std::sqrt(bestn);
    const float alpha = A[k][k] > 0 ? -nrm : nrm;
    float v[3] = {0, 0, 0};
    for (int i = k; i < 3; ++i) v[i] = A[i][k];
    v[k] -= alpha;
    float vn = 0; for (int i = k; i < 3; ++i) vn += v[i] * v[i];
    if (vn > 0) {
      for (int c = k; c < 3; ++c) {
        float s = 0; for (int i = k; i < 3; ++i) s += v[i] * A[i][c];
        s = 2.0f * s / vn;
        for (int i = k; i < 3; ++i) A[i][c] -= s * v[i];
      }
      float s = 0; for (int i = k; i < 3; ++i) s += v[i] * b[i];
      s = 2.0f * s / vn;
      for (int i = k; i < 3; ++i) b[i] -= s * v[i];
    }
    ++rank;
  }
  float y[3] = {0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    float s = b[k];
    for (int c = k + 1; c < rank; ++c) s -= A[k][c] * y[c];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

struct VisualOdometry {
  PointCloudUtil point_cloud_utils[2];
  int count = -1, i = 0;        // visual_odometry.cpp:28,86-90
  int remove_VO_outlier = 100;  // vloam_main.launch:6
  int max_num_iterations = 100; // visual_odometry.cpp:67
  double angles_0to1[3] = {0, 0, 0}, t_0to1[3] = {0, 0, 0};
  int counter32 = 0, counter22 = 0;
  LMSummary summary;
  std::vector<int> res_type;       // per match: 0 skipped, 1 CostFunctor32, 2 CostFunctor22 (parity read-out)
  std::vector<double> res_obs;     // 5 per match

  void reset() { ++count; i = count % 2; }
  void setCalibration(const float* cam_T_velo, const float* rect0_T_cam, const float* P_rect0) {
    for (int s = 0; s < 2; ++s) {
      for (int k = 0; k < 16; ++k) { point_cloud_utils[s].cam_T_velo[k] = cam_T_velo[k]; point_cloud_utils[s].rect0_T_cam[k] = rect0_T_cam[k]; }
      for (int k = 0; k < 12; ++k) point_cloud_utils[s].P_rect0[k] = P_rect0[k];
    }
  }
  void processPointCloud(const float* xyz, int n, int stride) {  // :157-186
    point_cloud_utils[i].projectPointCloud(xyz, n, stride);
    point_cloud_utils[i].downsamplePointCloud();
  }

  // prev_uv / curr_uv: m matched keypoint pixel coordinates (float, as cv::KeyPoint::pt).  init_*: LO prior
  // (cam0_curr_LOT_cam0_prev) or null for reset_VO_to_identity.
  void solveNlsAll(const float* prev_uv, const float* curr_uv, int m, const double* init_aa, const double* init_t) {  // :254-450
    for (int j = 0; j < 3; ++j) { angles_0to1[j] = init_aa ? init_aa[j] : 0.0; t_0to1[j] = init_t ? init_t[j] : 0.0; }
    counter32 = counter22 = 0;
    res_type.assign(m, 0); res_obs.assign((size_t)m * 5, 0.0);
    std::vector<CostBlock*> owned;
    const PointCloudUtil& pc_prev = point_cloud_utils[1 - i];
    const PointCloudUtil& pc_curr = point_cloud_utils[i];
    const float K0[9] = {pc_prev.P_rect0[0], pc_prev.P_rect0[1], pc_prev.P_rect0[2], pc_prev.P_rect0[4], pc_prev.P_rect0[5],
                         pc_prev.P_rect0[6], pc_prev.P_rect0[8], pc_prev.P_rect0[9], pc_prev.P_rect0[10]};
    const float K1[9] = {pc_curr.P_rect0[0], pc_curr.P_rect0[1], pc_curr.P_rect0[2], pc_curr.P_rect0[4], pc_curr.P_rect0[5],
                         pc_curr.P_rect0[6], pc_curr.P_rect0[8], pc_curr.P_rect0[9], pc_curr.P_rect0[10]};
    for (int j = 0; j < m; ++j) {
      int prev_pt_x = prev_uv[2 * j], prev_pt_y = prev_uv[2 * j + 1];  // Q7: truncation to int
      int curr_pt_x = curr_uv[2 * j], curr_pt_y = curr_uv[2 * j + 1];
      if (remove_VO_outlier > 0) {
        if (std::pow(prev_pt_x - curr_pt_x, 2) + std::pow(prev_pt_y - curr_pt_y, 2) > remove_VO_outlier * remove_VO_outlier) continue;
      }
      float depth0 = pc_prev.queryDepth(prev_pt_x, prev_pt_y);
      float p0[3], p1[3], X0[3], X1[3];
      if (depth0 > 0) {
        p0[0] = prev_pt_x * depth0; p0[1] = prev_pt_y * depth0; p0[2] = depth0;
        p1[0] = curr_pt_x; p1[1] = curr_pt_y; p1[2] = 1.0f;
        colpiv_qr_solve3x3f(K0, p0, X0);
        colpiv_qr_solve3x3f(K1, p1, X1);
        CostFunctor32 f{static_cast<double>(X0[0]), static_cast<double>(X0[1]), static_cast<double>(X0[2]),
                        static_cast<double>(X1[0]) / static_cast<double>(X1[2]), static_cast<double>(X1[1]) / static_cast<double>(X1[2])};
        owned.push_back(new AutoDiffBlock33<CostFunctor32, 2>(f));
        res_type[j] = 1; res_obs[j * 5] = f.observed_x0; res_obs[j * 5 + 1] = f.observed_y0; res_obs[j * 5 + 2] = f.observed_z0;
        res_obs[j * 5 + 3] = f.observed_x1_bar; res_obs[j * 5 + 4] = f.observed_y1_bar;
        ++counter32;
      } else {
        p0[0] = prev_pt_x; p0[1] = prev_pt_y; p0[2] = 1.0f;
        p1[0] = curr_pt_x; p1[1] = curr_pt_y; p1[2] = 1.0f;
        colpiv_qr_solve3x3f(K0, p0, X0);
        colpiv_qr_solve3x3f(K1, p1, X1);
        CostFunctor22 f{static_cast<double>(X0[0]) / static_cast<double>(X0[2]), static_cast<double>(X0[1]) / static_cast<double>(X0[2]),
                        static_cast<double>(X1[0]) / static_cast<double>(X1[2]), static_cast<double>(X1[1]) / static_cast<double>(X1[2])};
        owned.push_back(new AutoDiffBlock33<CostFunctor22, 1>(f));
        res_type[j] = 2; res_obs[j * 5] = f.observed_x0_bar; res_obs[j * 5 + 1] = f.observed_y0_bar;
        res_obs[j * 5 + 2] = f.observed_x1_bar; res_obs[j * 5 + 3] = f.observed_y1_bar;
        ++counter22;
      }
    }
    LMOptions opt;
    opt.max_num_iterations = max_num_iterations;
    opt.quaternion_manifold = false;
    opt.use_huber = true;
    opt.huber_a = 0.1;
    double x[6] = {angles_0to1[0], angles_0to1[1], angles_0to1[2], t_0to1[0], t_0to1[1], t_0to1[2]};
    std::vector<const CostBlock*> blocks(owned.begin(), owned.end());
    lm_solve(blocks, opt, x, &summary);
    for (int j = 0; j < 3; ++j) { angles_0to1[j] = x[j]; t_0to1[j] = x[3 + j]; }
    for (CostBlock* b : owned) delete b;
  }
};

}  // namespace oracle
