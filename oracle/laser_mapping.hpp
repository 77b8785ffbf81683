// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// laser_mapping.hpp: restatement of vloam::LaserMapping (reference
// src/lidar_odometry_mapping/src/laser_mapping.cpp: init :40-125, reset :127-131,
// transformUpdate :140-144, pointAssociateToMap :146-155, input :167-196,
// solveMapping :198-708; constants include/lidar_odometry_mapping/laser_mapping.h:76-78,
// 110-117).  SURVEY.md §8a rows C1-C12.
#pragma once
#include <utility>
#include <vector>

#include "ceres_lm.hpp"
#include "kdtree.hpp"
#include "lidar_factors.hpp"
#include "small_linalg.hpp"
#include "types.hpp"
#include "voxel_grid.hpp"

namespace oracle {

struct LMPassTrace {
  int corner_num = 0, surf_num = 0;
  std::vector<int> corner_query;  // indices into the down-sampled corner stack that produced a factor
  std::vector<int> surf_query;
  LMSummary summary;
  double parameters[7];
};

class LaserMapping {
 public:
  static constexpr int laserCloudWidth = 21, laserCloudHeight = 21, laserCloudDepth = 11;
  static constexpr int laserCloudNum = laserCloudWidth * laserCloudHeight * laserCloudDepth;  // 4851

  double lineRes = 0.4, planeRes = 0.8;  // mapping_line_resolution / mapping_plane_resolution
  int lm_max_iterations = 4;             // laser_mapping.cpp:612
  int num_outer_passes = 2;              // laser_mapping.cpp:458

  int laserCloudCenWidth = 10, laserCloudCenHeight = 10, laserCloudCenDepth = 5;
  double parameters[7] = {0, 0, 0, 1, 0, 0, 0};  // q_w_curr (x,y,z,w), t_w_curr
  Quat q_wmap_wodom, q_wodom_curr, q_w_curr_highfreq;
  Vec3 t_wmap_wodom, t_wodom_curr, t_w_curr_highfreq;
  bool skip_frame = false;
  std::vector<Cloud> laserCloudCornerArray, laserCloudSurfArray;
  int laserCloudValidInd[125], laserCloudSurroundInd[125];
  int laserCloudValidNum = 0, laserCloudSurroundNum = 0;
  Cloud laserCloudCornerLast, laserCloudSurfLast, laserCloudFullRes;
  Cloud laserCloudCornerFromMap, laserCloudSurfFromMap;
  Cloud laserCloudCornerStack, laserCloudSurfStack;  // down-sampled scan (exposed for parity)
  int frameCount = 0;
  std::vector<LMPassTrace> trace;

  LaserMapping() : laserCloudCornerArray(laserCloudNum), laserCloudSurfArray(laserCloudNum) {}

  void reset() { laserCloudValidNum = 0; laserCloudSurroundNum = 0; }  // :127-131

  Quat q_w_curr() const { return {parameters[0], parameters[1], parameters[2], parameters[3]}; }
  Vec3 t_w_curr() const { return {parameters[4], parameters[5], parameters[6]}; }

  void transformUpdate() {  // :140-144
    q_wmap_wodom = q_w_curr() * inverse(q_wodom_curr);
    t_wmap_wodom = t_w_curr() - rotate(q_wmap_wodom, t_wodom_curr);
  }
  void pointAssociateToMap(const PointXYZI& pi, PointXYZI* po) const {  // :146-155
    Vec3 point_w = rotate(q_w_curr(), Vec3{pi.x, pi.y, pi.z}) + t_w_curr();
    po->x = static_cast<float>(point_w.x);
    po->y = static_cast<float>(point_w.y);
    po->z = static_cast<float>(point_w.z);
    po->intensity = pi.intensity;
  }

  void input(const Cloud& cornerLast, const Cloud& surfLast, const Cloud& fullRes, const Quat& q_odom,
             const Vec3& t_odom, bool skip_frame_ = false) {  // :167-196
    skip_frame = skip_frame_;
    if (!skip_frame) {  // :175-180
      laserCloudCornerLast = cornerLast;
      laserCloudSurfLast = surfLast;
      laserCloudFullRes = fullRes;
    }
    q_wodom_curr = q_odom;  // :182-183
    t_wodom_curr = t_odom;
    Quat q = q_wmap_wodom * q_wodom_curr;
    Vec3 t = rotate(q_wmap_wodom, t_wodom_curr) + t_wmap_wodom;
    if (skip_frame) {  // :186-190: only the high-frequency pose (what publish() sends for such a frame, :742-756)
      q_w_curr_highfreq = q;
      t_w_curr_highfreq = t;
    } else {  // :191-195
      parameters[0] = q.x; parameters[1] = q.y; parameters[2] = q.z; parameters[3] = q.w;
      parameters[4] = t.x; parameters[5] = t.y; parameters[6] = t.z;
    }
  }
  // The cloud part of LaserMapping::publish.  :797-801 moves laserCloudFullRes into the map frame IN PLACE (the object's own
  // deep copy, :175-180) before it goes out on /velodyne_cloud_registered — so a skipped frame, which keeps the previous
  // copy, transforms it a second time; restated literally.  Call once per frame, like publish().
  void publish_registered() {
    for (PointXYZI& p : laserCloudFullRes) pointAssociateToMap(p, &p);
  }
  // /laser_cloud_map (:778-785): every cube, corner points then surf points
  Cloud map_cloud() const {
    Cloud all;
    for (int i = 0; i < laserCloudNum; ++i) {
      all.insert(all.end(), laserCloudCornerArray[i].begin(), laserCloudCornerArray[i].end());
      all.insert(all.end(), laserCloudSurfArray[i].begin(), laserCloudSurfArray[i].end());
    }
    return all;
  }
  // The pose LaserMapping::publish puts into /aft_mapped_to_init for the last input() (:720-756)
  void published_pose(double out[7]) const {
    if (skip_frame) {
      out[0] = q_w_curr_highfreq.x; out[1] = q_w_curr_highfreq.y; out[2] = q_w_curr_highfreq.z; out[3] = q_w_curr_highfreq.w;
      out[4] = t_w_curr_highfreq.x; out[5] = t_w_curr_highfreq.y; out[6] = t_w_curr_highfreq.z;
    } else {
      for (int i = 0; i < 7; ++i) out[i] = parameters[i];
    }
  }

  static int cube_coord(double v, int cen) {  // :207-216 / :643-652
    int c = int((v + 25.0) / 50.0) + cen;
    if (v + 25.0 < 0) c--;
    return c;
  }
  int cidx(int i, int j, int k) const { return i + laserCloudWidth * j + laserCloudWidth * laserCloudHeight * k; }

  void shift_grid(int* centerCubeI, int* centerCubeJ, int* centerCubeK) {  // :218-402
    auto roll = [&](std::vector<Cloud>& arr, int a0, int stride, int n, bool toward_high) {
      if (toward_high) {  // contents move to higher index, the highest wraps to 0 and is cleared
        for (int i = n - 1; i >= 1; --i) std::swap(arr[a0 + i * stride], arr[a0 + (i - 1) * stride]);
        arr[a0].clear();
      } else {
        for (int i = 0; i < n - 1; ++i) std::swap(arr[a0 + i * stride], arr[a0 + (i + 1) * stride]);
        arr[a0 + (n - 1) * stride].clear();
      }
    };
    const int W = laserCloudWidth, H = laserCloudHeight, D = laserCloudDepth;
    while (*centerCubeI < 3) {
      for (int j = 0; j < H; j++) for (int k = 0; k < D; k++) {
        roll(laserCloudCornerArray, cidx(0, j, k), 1, W, true); roll(laserCloudSurfArray, cidx(0, j, k), 1, W, true);
      }
      (*centerCubeI)++; laserCloudCenWidth++;
    }
    while (*centerCubeI >= W - 3) {
      for (int j = 0; j < H; j++) for (int k = 0; k < D; k++) {
        roll(laserCloudCornerArray, cidx(0, j, k), 1, W, false); roll(laserCloudSurfArray, cidx(0, j, k), 1, W, false);
      }
      (*centerCubeI)--; laserCloudCenWidth--;
    }
    while (*centerCubeJ < 3) {
      for (int i = 0; i < W; i++) for (int k = 0; k < D; k++) {
        roll(laserCloudCornerArray, cidx(i, 0, k), W, H, true); roll(laserCloudSurfArray, cidx(i, 0, k), W, H, true);
      }
      (*centerCubeJ)++; laserCloudCenHeight++;
    }
    while (*centerCubeJ >= H - 3) {
      for (int i = 0; i < W; i++) for (int k = 0; k < D; k++) {
        roll(laserCloudCornerArray, cidx(i, 0, k), W, H, false); roll(laserCloudSurfArray, cidx(i, 0, k), W, H, false);
      }
      (*centerCubeJ)--; laserCloudCenHeight--;
    }
    while (*centerCubeK < 3) {
      for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
        roll(laserCloudCornerArray, cidx(i, j, 0), W * H, D, true); roll(laserCloudSurfArray, cidx(i, j, 0), W * H, D, true);
      }
      (*centerCubeK)++; laserCloudCenDepth++;
    }
    while (*centerCubeK >= D - 3) {
      for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
        roll(laserCloudCornerArray, cidx(i, j, 0), W * H, D, false); roll(laserCloudSurfArray, cidx(i, j, 0), W * H, D, false);
      }
      (*centerCubeK)--; laserCloudCenDepth--;
    }
  }

  void solveMapping() {  // :198-708
    trace.clear();
    Vec3 t = t_w_curr();
    int centerCubeI = cube_coord(t.x, laserCloudCenWidth);
    int centerCubeJ = cube_coord(t.y, laserCloudCenHeight);
    int centerCubeK = cube_coord(t.z, laserCloudCenDepth);
    shift_grid(&centerCubeI, &centerCubeJ, &centerCubeK);

    for (int i = centerCubeI - 2; i <= centerCubeI + 2; i++)
      for (int j = centerCubeJ - 2; j <= centerCubeJ + 2; j++)
        for (int k = centerCubeK - 1; k <= centerCubeK + 1; k++)
          if (i >= 0 && i < laserCloudWidth && j >= 0 && j < laserCloudHeight && k >= 0 && k < laserCloudDepth) {
            laserCloudValidInd[laserCloudValidNum++] = cidx(i, j, k);
            laserCloudSurroundInd[laserCloudSurroundNum++] = cidx(i, j, k);
          }

    laserCloudCornerFromMap.clear();
    laserCloudSurfFromMap.clear();
    for (int i = 0; i < laserCloudValidNum; i++) {
      const Cloud& c = laserCloudCornerArray[laserCloudValidInd[i]];
      const Cloud& s = laserCloudSurfArray[laserCloudValidInd[i]];
      laserCloudCornerFromMap.insert(laserCloudCornerFromMap.end(), c.begin(), c.end());
      laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), s.begin(), s.end());
    }
    const int laserCloudCornerFromMapNum = static_cast<int>(laserCloudCornerFromMap.size());
    const int laserCloudSurfFromMapNum = static_cast<int>(laserCloudSurfFromMap.size());

    voxel_grid_filter(laserCloudCornerLast, static_cast<float>(lineRes), &laserCloudCornerStack);  // :432-435
    voxel_grid_filter(laserCloudSurfLast, static_cast<float>(planeRes), &laserCloudSurfStack);     // :437-440
    const int laserCloudCornerStackNum = static_cast<int>(laserCloudCornerStack.size());
    const int laserCloudSurfStackNum = static_cast<int>(laserCloudSurfStack.size());

    if (laserCloudCornerFromMapNum > 10 && laserCloudSurfFromMapNum > 50) {  // :448
      kdtreeCornerFromMap.set_input(&laserCloudCornerFromMap);
      kdtreeSurfFromMap.set_input(&laserCloudSurfFromMap);
      for (int iterCount = 0; iterCount < num_outer_passes; iterCount++) {
        LMPassTrace tr;
        std::vector<CostBlock*> owned;
        PointXYZI pointOri, pointSel;
        int pointSearchInd[5];
        float pointSearchSqDis[5];
        for (int i = 0; i < laserCloudCornerStackNum; i++) {  // :472-517
          pointOri = laserCloudCornerStack[i];
          pointAssociateToMap(pointOri, &pointSel);
          if (kdtreeCornerFromMap.nearest_k(pointSel, 5, pointSearchInd, pointSearchSqDis) < 5) continue;  // Q16
          if (pointSearchSqDis[4] < 1.0) {
            Vec3 nearCorners[5];
            Vec3 center{0, 0, 0};
            for (int j = 0; j < 5; j++) {
              const PointXYZI& p = laserCloudCornerFromMap[pointSearchInd[j]];
              Vec3 tmp{p.x, p.y, p.z};
              center = center + tmp;
              nearCorners[j] = tmp;
            }
            center = {center.x / 5.0, center.y / 5.0, center.z / 5.0};
            double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < 5; j++) {
              Vec3 d = nearCorners[j] - center;
              const double dv[3] = {d.x, d.y, d.z};
              for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[r * 3 + c] += dv[r] * dv[c];
            }
            double evals[3], evecs[3][3];
            sym_eig3(cov, evals, evecs);
            Vec3 unit_direction{evecs[2][0], evecs[2][1], evecs[2][2]};
            if (evals[2] > 3 * evals[1]) {
              LidarEdgeFunctor f;
              f.curr_point = {pointOri.x, pointOri.y, pointOri.z};
              f.last_point_a = 0.1 * unit_direction + center;
              f.last_point_b = -0.1 * unit_direction + center;
              f.s = 1.0;
              owned.push_back(new AutoDiffBlock43<LidarEdgeFunctor, 3>(f));
              tr.corner_query.push_back(i);
              tr.corner_num++;
            }
          }
        }
        for (int i = 0; i < laserCloudSurfStackNum; i++) {  // :538-581
          pointOri = laserCloudSurfStack[i];
          pointAssociateToMap(pointOri, &pointSel);
          if (kdtreeSurfFromMap.nearest_k(pointSel, 5, pointSearchInd, pointSearchSqDis) < 5) continue;  // Q16
          if (pointSearchSqDis[4] < 1.0) {
            double matA0[15], matB0[5] = {-1, -1, -1, -1, -1};
            for (int j = 0; j < 5; j++) {
              const PointXYZI& p = laserCloudSurfFromMap[pointSearchInd[j]];
              matA0[j * 3 + 0] = p.x; matA0[j * 3 + 1] = p.y; matA0[j * 3 + 2] = p.z;
            }
            double nrm[3];
            colpiv_qr_solve3<5>(matA0, matB0, nrm);
            const double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            const double negative_OA_dot_norm = 1 / nn;
            nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn;
            bool planeValid = true;
            for (int j = 0; j < 5; j++) {
              const PointXYZI& p = laserCloudSurfFromMap[pointSearchInd[j]];
              if (std::fabs(nrm[0] * p.x + nrm[1] * p.y + nrm[2] * p.z + negative_OA_dot_norm) > 0.2) {
                planeValid = false;
                break;
              }
            }
            if (planeValid) {
              LidarPlaneNormFunctor f;
              f.curr_point = {pointOri.x, pointOri.y, pointOri.z};
              f.plane_unit_norm = {nrm[0], nrm[1], nrm[2]};
              f.negative_OA_dot_norm = negative_OA_dot_norm;
              owned.push_back(new AutoDiffBlock43<LidarPlaneNormFunctor, 1>(f));
              tr.surf_query.push_back(i);
              tr.surf_num++;
            }
          }
        }
        LMOptions opt;  // :609-617
        opt.max_num_iterations = lm_max_iterations;
        opt.quaternion_manifold = true;
        opt.use_huber = true;
        opt.huber_a = 0.1;
        std::vector<const CostBlock*> blocks(owned.begin(), owned.end());
        lm_solve(blocks, opt, parameters, &tr.summary);
        for (CostBlock* b : owned) delete b;
        for (int i = 0; i < 7; ++i) tr.parameters[i] = parameters[i];
        trace.push_back(std::move(tr));
      }
    }
    transformUpdate();  // :636

    auto insert = [&](const Cloud& stack, std::vector<Cloud>& arr) {  // :639-683
      PointXYZI pointSel;
      for (const PointXYZI& p : stack) {
        pointAssociateToMap(p, &pointSel);
        int cubeI = int((pointSel.x + 25.0) / 50.0) + laserCloudCenWidth;
        int cubeJ = int((pointSel.y + 25.0) / 50.0) + laserCloudCenHeight;
        int cubeK = int((pointSel.z + 25.0) / 50.0) + laserCloudCenDepth;
        if (pointSel.x + 25.0 < 0) cubeI--;
        if (pointSel.y + 25.0 < 0) cubeJ--;
        if (pointSel.z + 25.0 < 0) cubeK--;
        if (cubeI >= 0 && cubeI < laserCloudWidth && cubeJ >= 0 && cubeJ < laserCloudHeight && cubeK >= 0 &&
            cubeK < laserCloudDepth)
          arr[cidx(cubeI, cubeJ, cubeK)].push_back(pointSel);
      }
    };
    insert(laserCloudCornerStack, laserCloudCornerArray);
    insert(laserCloudSurfStack, laserCloudSurfArray);

    for (int i = 0; i < laserCloudValidNum; i++) {  // :689-702
      int ind = laserCloudValidInd[i];
      Cloud tmpCorner, tmpSurf;
      voxel_grid_filter(laserCloudCornerArray[ind], static_cast<float>(lineRes), &tmpCorner);
      laserCloudCornerArray[ind] = std::move(tmpCorner);
      voxel_grid_filter(laserCloudSurfArray[ind], static_cast<float>(planeRes), &tmpSurf);
      laserCloudSurfArray[ind] = std::move(tmpSurf);
    }
    frameCount++;
  }

 private:
  KdTree kdtreeCornerFromMap, kdtreeSurfFromMap;
};

}  // namespace oracle
