// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// laser_odometry.hpp: restatement of vloam::LaserOdometry
// (reference src/lidar_odometry_mapping/src/laser_odometry.cpp: init :41-117,
// TransformToStart :149-167, solveLO :187-536, output :610-629; constants
// include/lidar_odometry_mapping/laser_odometry.h:90-95).  SURVEY.md §8a B1-B9.
#pragma once
#include <vector>

#include "ceres_lm.hpp"
#include "kdtree.hpp"
#include "lidar_factors.hpp"
#include "types.hpp"

namespace oracle {

struct LOPassTrace {
  std::vector<int> corner;  // triples (query i, closestPointInd, minPointInd2) of accepted correspondences
  std::vector<int> plane;   // quads (query i, closestPointInd, minPointInd2, minPointInd3)
  LMSummary summary;
  double para_q[4], para_t[3];  // after this pass
};

class LaserOdometry {
 public:
  // laser_odometry.h:90-95
  bool DISTORTION = false;   // laser_odometry.h:90 (a compile-time constant of the reference, false as shipped; a run-time switch here)
  static constexpr double SCAN_PERIOD = 0.1;
  static constexpr double DISTANCE_SQ_THRESHOLD = 25;
  static constexpr double NEARBY_SCAN = 2.5;

  bool detach_VO_LO = true;
  int mapping_skip_frame = 1;
  int lm_max_iterations = 4;   // laser_odometry.cpp:460
  int num_outer_passes = 2;    // laser_odometry.cpp:211

  bool systemInited = false;
  double para_q[4] = {0, 0, 0, 1};  // q_last_curr (x,y,z,w)    :84-87
  double para_t[3] = {0, 0, 0};     // t_last_curr              :88-90
  Quat q_w_curr;                    // :80
  Vec3 t_w_curr;                    // :81
  Cloud laserCloudCornerLast, laserCloudSurfLast, laserCloudFullRes;
  int corner_correspondence = 0, plane_correspondence = 0;
  int frameCount = 0;
  std::vector<LOPassTrace> trace;  // of the latest solve

  void init() { *this = LaserOdometry(); }

  // interpolation ratio of a point inside the sweep, :152-156 / :329-335 / :425-431
  double ratio(const PointXYZI& pi) const {
    if (DISTORTION) return (pi.intensity - int(pi.intensity)) / SCAN_PERIOD;   // float - int -> float, / double
    return 1.0;
  }
  // TransformToStart :149-167
  void TransformToStart(const PointXYZI& pi, PointXYZI* po) const {
    const double s = ratio(pi);
    // q_point_last = Identity().slerp(s, q_last_curr) (Eigen::QuaternionBase::slerp, restated in lidar_factors.hpp); with
    // s == 1 that is +-q_last_curr up to rounding — the shipped configuration, kept on its exact short path
    Quat q{para_q[0], para_q[1], para_q[2], para_q[3]};
    if (DISTORTION) {
      const Q4<double> qs = slerp_from_identity(s, Q4<double>{para_q[3], para_q[0], para_q[1], para_q[2]});
      q = Quat{qs.x, qs.y, qs.z, qs.w};
    }
    Vec3 point{pi.x, pi.y, pi.z};
    Vec3 un_point = rotate(q, point) + s * Vec3{para_t[0], para_t[1], para_t[2]};
    po->x = static_cast<float>(un_point.x);
    po->y = static_cast<float>(un_point.y);
    po->z = static_cast<float>(un_point.z);
    po->intensity = pi.intensity;
  }

  static double sqdist(const PointXYZI& a, const PointXYZI& sel) {
    // :289-292 — float arithmetic, widened to double on assignment
    return (a.x - sel.x) * (a.x - sel.x) + (a.y - sel.y) * (a.y - sel.y) + (a.z - sel.z) * (a.z - sel.z);
  }

  // solveLO :187-536.  prior_* = vloam_tf->velo_last_VOT_velo_curr, used only when !detach_VO_LO.
  void solveLO(const Cloud& laserCloud, const Cloud& cornerPointsSharp, const Cloud& cornerPointsLessSharp,
               const Cloud& surfPointsFlat, const Cloud& surfPointsLessFlat, const double* prior_q,
               const double* prior_t) {
    trace.clear();
    if (!systemInited) {
      systemInited = true;
    } else {
      const int cornerPointsSharpNum = static_cast<int>(cornerPointsSharp.size());
      const int surfPointsFlatNum = static_cast<int>(surfPointsFlat.size());
      for (int opti_counter = 0; opti_counter < num_outer_passes; ++opti_counter) {
        corner_correspondence = 0;
        plane_correspondence = 0;
        LOPassTrace tr;
        if (!detach_VO_LO && prior_q && prior_t) {  // :223-236 (both passes: quirk Q1)
          for (int i = 0; i < 4; ++i) para_q[i] = prior_q[i];
          for (int i = 0; i < 3; ++i) para_t[i] = prior_t[i];
        }
        std::vector<CostBlock*> owned;
        PointXYZI pointSel;
        int pointSearchInd[1];
        float pointSearchSqDis[1];

        // corner features :266-350
        for (int i = 0; i < cornerPointsSharpNum; ++i) {
          TransformToStart(cornerPointsSharp[i], &pointSel);
          if (kdtreeCornerLast.nearest_k(pointSel, 1, pointSearchInd, pointSearchSqDis) < 1) continue;  // Q16
          int closestPointInd = -1, minPointInd2 = -1;
          if (pointSearchSqDis[0] < DISTANCE_SQ_THRESHOLD) {
            closestPointInd = pointSearchInd[0];
            int closestPointScanID = int(laserCloudCornerLast[closestPointInd].intensity);
            double minPointSqDis2 = DISTANCE_SQ_THRESHOLD;
            for (int j = closestPointInd + 1; j < (int)laserCloudCornerLast.size(); ++j) {
              if (int(laserCloudCornerLast[j].intensity) <= closestPointScanID) continue;
              if (int(laserCloudCornerLast[j].intensity) > (closestPointScanID + NEARBY_SCAN)) break;
              double pointSqDis = sqdist(laserCloudCornerLast[j], pointSel);
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
            for (int j = closestPointInd - 1; j >= 0; --j) {
              if (int(laserCloudCornerLast[j].intensity) >= closestPointScanID) continue;
              if (int(laserCloudCornerLast[j].intensity) < (closestPointScanID - NEARBY_SCAN)) break;
              double pointSqDis = sqdist(laserCloudCornerLast[j], pointSel);
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
          if (minPointInd2 >= 0) {
            LidarEdgeFunctor f;
            f.curr_point = {cornerPointsSharp[i].x, cornerPointsSharp[i].y, cornerPointsSharp[i].z};
            f.last_point_a = {laserCloudCornerLast[closestPointInd].x, laserCloudCornerLast[closestPointInd].y,
                              laserCloudCornerLast[closestPointInd].z};
            f.last_point_b = {laserCloudCornerLast[minPointInd2].x, laserCloudCornerLast[minPointInd2].y,
                              laserCloudCornerLast[minPointInd2].z};
            f.s = ratio(cornerPointsSharp[i]);   // :329-335
            owned.push_back(new AutoDiffBlock43<LidarEdgeFunctor, 3>(f));
            tr.corner.push_back(i); tr.corner.push_back(closestPointInd); tr.corner.push_back(minPointInd2);
            corner_correspondence++;
          }
        }

        // plane features :353-444
        for (int i = 0; i < surfPointsFlatNum; ++i) {
          TransformToStart(surfPointsFlat[i], &pointSel);
          if (kdtreeSurfLast.nearest_k(pointSel, 1, pointSearchInd, pointSearchSqDis) < 1) continue;  // Q16
          int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
          if (pointSearchSqDis[0] < DISTANCE_SQ_THRESHOLD) {
            closestPointInd = pointSearchInd[0];
            int closestPointScanID = int(laserCloudSurfLast[closestPointInd].intensity);
            double minPointSqDis2 = DISTANCE_SQ_THRESHOLD, minPointSqDis3 = DISTANCE_SQ_THRESHOLD;
            for (int j = closestPointInd + 1; j < (int)laserCloudSurfLast.size(); ++j) {
              if (int(laserCloudSurfLast[j].intensity) > (closestPointScanID + NEARBY_SCAN)) break;
              double pointSqDis = sqdist(laserCloudSurfLast[j], pointSel);
              if (int(laserCloudSurfLast[j].intensity) <= closestPointScanID && pointSqDis < minPointSqDis2) {
                minPointSqDis2 = pointSqDis; minPointInd2 = j;
              } else if (int(laserCloudSurfLast[j].intensity) > closestPointScanID && pointSqDis < minPointSqDis3) {
                minPointSqDis3 = pointSqDis; minPointInd3 = j;
              }
            }
            for (int j = closestPointInd - 1; j >= 0; --j) {
              if (int(laserCloudSurfLast[j].intensity) < (closestPointScanID - NEARBY_SCAN)) break;
              double pointSqDis = sqdist(laserCloudSurfLast[j], pointSel);
              if (int(laserCloudSurfLast[j].intensity) >= closestPointScanID && pointSqDis < minPointSqDis2) {
                minPointSqDis2 = pointSqDis; minPointInd2 = j;
              } else if (int(laserCloudSurfLast[j].intensity) < closestPointScanID && pointSqDis < minPointSqDis3) {
                minPointSqDis3 = pointSqDis; minPointInd3 = j;
              }
            }
            if (minPointInd2 >= 0 && minPointInd3 >= 0) {
              Vec3 c{surfPointsFlat[i].x, surfPointsFlat[i].y, surfPointsFlat[i].z};
              Vec3 a{laserCloudSurfLast[closestPointInd].x, laserCloudSurfLast[closestPointInd].y,
                     laserCloudSurfLast[closestPointInd].z};
              Vec3 b{laserCloudSurfLast[minPointInd2].x, laserCloudSurfLast[minPointInd2].y,
                     laserCloudSurfLast[minPointInd2].z};
              Vec3 cc{laserCloudSurfLast[minPointInd3].x, laserCloudSurfLast[minPointInd3].y,
                      laserCloudSurfLast[minPointInd3].z};
              owned.push_back(new AutoDiffBlock43<LidarPlaneFunctor, 1>(LidarPlaneFunctor(c, a, b, cc, ratio(surfPointsFlat[i]))));
              tr.plane.push_back(i); tr.plane.push_back(closestPointInd);
              tr.plane.push_back(minPointInd2); tr.plane.push_back(minPointInd3);
              plane_correspondence++;
            }
          }
        }

        // :457-463
        LMOptions opt;
        opt.max_num_iterations = lm_max_iterations;
        opt.quaternion_manifold = true;
        opt.use_huber = true;
        opt.huber_a = 0.1;
        double x[7] = {para_q[0], para_q[1], para_q[2], para_q[3], para_t[0], para_t[1], para_t[2]};
        std::vector<const CostBlock*> blocks(owned.begin(), owned.end());
        lm_solve(blocks, opt, x, &tr.summary);
        for (int i = 0; i < 4; ++i) para_q[i] = x[i];
        for (int i = 0; i < 3; ++i) para_t[i] = x[4 + i];
        for (CostBlock* b : owned) delete b;
        for (int i = 0; i < 4; ++i) tr.para_q[i] = para_q[i];
        for (int i = 0; i < 3; ++i) tr.para_t[i] = para_t[i];
        trace.push_back(std::move(tr));
      }
      // :477-478
      Quat q_last_curr{para_q[0], para_q[1], para_q[2], para_q[3]};
      Vec3 t_last_curr{para_t[0], para_t[1], para_t[2]};
      t_w_curr = t_w_curr + rotate(q_w_curr, t_last_curr);
      q_w_curr = q_w_curr * q_last_curr;
    }
    // :511-526 (swap == copy of the current less-sharp / less-flat into "last"), kd-tree rebuild
    laserCloudCornerLast = cornerPointsLessSharp;
    laserCloudSurfLast = surfPointsLessFlat;
    laserCloudFullRes = laserCloud;
    kdtreeCornerLast.set_input(&laserCloudCornerLast);
    kdtreeSurfLast.set_input(&laserCloudSurfLast);
    frameCount++;
  }

  bool skip_frame() const { return !(frameCount % mapping_skip_frame == 0); }  // :618-628

 private:
  KdTree kdtreeCornerLast, kdtreeSurfLast;
};

}  // namespace oracle
