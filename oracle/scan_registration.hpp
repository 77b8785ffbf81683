// ORACLE — TEST INFRASTRUCTURE ONLY (see types.hpp header).  PARITY UNPINNED.
//
// scan_registration.hpp: restatement of vloam::ScanRegistration::input
// (reference src/lidar_odometry_mapping/src/scan_registration.cpp:131-449,
// helper removeClosedPointCloud :100-129).  SURVEY.md §8a rows A1-A8.
//
// Conventions fixed where the reference is toolchain-dependent:
//   Q10  std::sort on curvature only is unstable -> total order (curvature, index).
//   Q11  unqualified atan/sqrt at :192 -> evaluated in double, assigned to float.
//   Q18  no FMA contraction: built with plain -O3 on x86-64 (no -march), like the
//        reference (src/lidar_odometry_mapping/CMakeLists.txt:5-6).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "types.hpp"
#include "voxel_grid.hpp"

namespace oracle {

struct ScanRegistrationOutput {
  Cloud laserCloud;             // ring-major, intensity = ring + 0.1*relTime   (:264-266, :276-281)
  Cloud cornerPointsSharp;      // (:338)
  Cloud cornerPointsLessSharp;  // (:339, :344)
  Cloud surfPointsFlat;         // (:388)
  Cloud surfPointsLessFlat;     // (:424-439)
  // parity/debug views of the reference's scratch arrays (scan_registration.h:90-93)
  std::vector<float> curvature;  // cloudCurvature, defined on [5, size-5)
  std::vector<int> label;        // cloudLabel
  std::vector<int> picked;       // cloudNeighborPicked at the end of the scan
  std::vector<int> scanStartInd, scanEndInd;
  std::vector<int> sharpInd, lessSharpInd, flatInd;  // indices into laserCloud, in push order
  std::vector<int> ringLessFlatCount;                // per-ring size of the down-sampled less-flat cloud
  int status = 0;  // 0 ok, 1 = no point survived the filters (the reference would index points[0])
};

inline bool ring_id(int n_scans, float angle, int* scan_id) {
  int id = 0;
  if (n_scans == 16) {
    id = int((angle + 15) / 2 + 0.5);
    if (id > (n_scans - 1) || id < 0) return false;
  } else if (n_scans == 32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
    if (id > (n_scans - 1) || id < 0) return false;
  } else {  // 64
    if (angle >= -8.83)
      id = int((2 - angle) * 3.0 + 0.5);
    else
      id = n_scans / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 50 || id < 0) return false;
  }
  *scan_id = id;
  return true;
}

// xyz: n points, `stride` floats apart (3 for packed xyz, 4 for pcl::PointXYZ).
inline void scan_registration(const float* xyz, int n, int stride, int n_scans, double minimum_range,
                              ScanRegistrationOutput* out, bool voxel_literal_unstable = false) {
  *out = ScanRegistrationOutput();
  const double scanPeriod = 0.1;
  // :157 removeNaNFromPointCloud, :158 removeClosedPointCloud (order preserving)
  struct P3 { float x, y, z; };
  std::vector<P3> in;
  in.reserve(n);
  const float thres = static_cast<float>(minimum_range);
  for (int i = 0; i < n; ++i) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    if (x * x + y * y + z * z < thres * thres) continue;
    in.push_back({x, y, z});
  }
  int cloudSize = static_cast<int>(in.size());
  out->scanStartInd.assign(n_scans, 0);
  out->scanEndInd.assign(n_scans, 0);
  out->ringLessFlatCount.assign(n_scans, 0);
  if (cloudSize == 0) { out->status = 1; return; }

  // :166-176
  float startOri = -std::atan2(in[0].y, in[0].x);
  float endOri = -std::atan2(in[cloudSize - 1].y, in[cloudSize - 1].x) + 2 * M_PI;
  if (endOri - startOri > 3 * M_PI) {
    endOri -= 2 * M_PI;
  } else if (endOri - startOri < M_PI) {
    endOri += 2 * M_PI;
  }

  // :183-267
  bool halfPassed = false;
  int count = cloudSize;
  std::vector<Cloud> laserCloudScans(n_scans);
  for (int i = 0; i < cloudSize; i++) {
    PointXYZI point;
    point.x = in[i].x; point.y = in[i].y; point.z = in[i].z;
    // Q11: double evaluation of atan(z / sqrt(x*x+y*y)) * 180 / M_PI, assigned to float
    const float xy2 = point.x * point.x + point.y * point.y;
    float angle = std::atan(static_cast<double>(point.z) / std::sqrt(static_cast<double>(xy2))) * 180 / M_PI;
    int scanID = 0;
    if (!ring_id(n_scans, angle, &scanID)) { count--; continue; }

    float ori = -std::atan2(point.y, point.x);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) {
        ori += 2 * M_PI;
      } else if (ori > startOri + M_PI * 3 / 2) {
        ori -= 2 * M_PI;
      }
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < endOri - M_PI * 3 / 2) {
        ori += 2 * M_PI;
      } else if (ori > endOri + M_PI / 2) {
        ori -= 2 * M_PI;
      }
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    point.intensity = scanID + scanPeriod * relTime;
    laserCloudScans[scanID].push_back(point);
  }
  cloudSize = count;

  // :276-281
  Cloud& laserCloud = out->laserCloud;
  for (int i = 0; i < n_scans; i++) {
    out->scanStartInd[i] = static_cast<int>(laserCloud.size()) + 5;
    laserCloud.insert(laserCloud.end(), laserCloudScans[i].begin(), laserCloudScans[i].end());
    out->scanEndInd[i] = static_cast<int>(laserCloud.size()) - 6;
  }

  // :288-307
  std::vector<float>& cloudCurvature = out->curvature;
  std::vector<int> cloudSortInd(cloudSize, 0);
  std::vector<int>& cloudNeighborPicked = out->picked;
  std::vector<int>& cloudLabel = out->label;
  cloudCurvature.assign(cloudSize, 0.f);
  cloudNeighborPicked.assign(cloudSize, 0);
  cloudLabel.assign(cloudSize, 0);
  for (int i = 5; i < cloudSize - 5; i++) {
    float diffX = laserCloud[i - 5].x + laserCloud[i - 4].x + laserCloud[i - 3].x + laserCloud[i - 2].x +
                  laserCloud[i - 1].x - 10 * laserCloud[i].x + laserCloud[i + 1].x + laserCloud[i + 2].x +
                  laserCloud[i + 3].x + laserCloud[i + 4].x + laserCloud[i + 5].x;
    float diffY = laserCloud[i - 5].y + laserCloud[i - 4].y + laserCloud[i - 3].y + laserCloud[i - 2].y +
                  laserCloud[i - 1].y - 10 * laserCloud[i].y + laserCloud[i + 1].y + laserCloud[i + 2].y +
                  laserCloud[i + 3].y + laserCloud[i + 4].y + laserCloud[i + 5].y;
    float diffZ = laserCloud[i - 5].z + laserCloud[i - 4].z + laserCloud[i - 3].z + laserCloud[i - 2].z +
                  laserCloud[i - 1].z - 10 * laserCloud[i].z + laserCloud[i + 1].z + laserCloud[i + 2].z +
                  laserCloud[i + 3].z + laserCloud[i + 4].z + laserCloud[i + 5].z;
    cloudCurvature[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
    cloudSortInd[i] = i;
    cloudNeighborPicked[i] = 0;
    cloudLabel[i] = 0;
  }

  auto gap2 = [&](int a, int b) {
    float diffX = laserCloud[a].x - laserCloud[b].x;
    float diffY = laserCloud[a].y - laserCloud[b].y;
    float diffZ = laserCloud[a].z - laserCloud[b].z;
    return diffX * diffX + diffY * diffY + diffZ * diffZ;
  };
  auto mark_neighbours = [&](int ind) {
    for (int l = 1; l <= 5; l++) {
      if (gap2(ind + l, ind + l - 1) > 0.05) break;
      cloudNeighborPicked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      if (gap2(ind + l, ind + l + 1) > 0.05) break;
      cloudNeighborPicked[ind + l] = 1;
    }
  };

  // :312-440
  for (int i = 0; i < n_scans; i++) {
    if (out->scanEndInd[i] - out->scanStartInd[i] < 6) continue;
    Cloud surfPointsLessFlatScan;
    for (int j = 0; j < 6; j++) {
      int sp = out->scanStartInd[i] + (out->scanEndInd[i] - out->scanStartInd[i]) * j / 6;
      int ep = out->scanStartInd[i] + (out->scanEndInd[i] - out->scanStartInd[i]) * (j + 1) / 6 - 1;

      // Q10: (curvature, index) total order
      std::sort(cloudSortInd.begin() + sp, cloudSortInd.begin() + ep + 1, [&](const int& a, const int& b) {
        return cloudCurvature[a] != cloudCurvature[b] ? cloudCurvature[a] < cloudCurvature[b] : a < b;
      });

      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > 0.1) {
          largestPickedNum++;
          if (largestPickedNum <= 2) {
            cloudLabel[ind] = 2;
            out->cornerPointsSharp.push_back(laserCloud[ind]);
            out->cornerPointsLessSharp.push_back(laserCloud[ind]);
            out->sharpInd.push_back(ind);
            out->lessSharpInd.push_back(ind);
          } else if (largestPickedNum <= 20) {
            cloudLabel[ind] = 1;
            out->cornerPointsLessSharp.push_back(laserCloud[ind]);
            out->lessSharpInd.push_back(ind);
          } else {
            break;
          }
          cloudNeighborPicked[ind] = 1;
          mark_neighbours(ind);
        }
      }

      int smallestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < 0.1) {
          cloudLabel[ind] = -1;
          out->surfPointsFlat.push_back(laserCloud[ind]);
          out->flatInd.push_back(ind);
          smallestPickedNum++;
          if (smallestPickedNum >= 4) break;  // Q2: before marking
          cloudNeighborPicked[ind] = 1;
          mark_neighbours(ind);
        }
      }

      for (int k = sp; k <= ep; k++) {
        if (cloudLabel[k] <= 0) surfPointsLessFlatScan.push_back(laserCloud[k]);  // Q3: by position
      }
    }
    Cloud surfPointsLessFlatScanDS;
    voxel_grid_filter(surfPointsLessFlatScan, 0.2f, &surfPointsLessFlatScanDS, voxel_literal_unstable);
    out->ringLessFlatCount[i] = static_cast<int>(surfPointsLessFlatScanDS.size());
    out->surfPointsLessFlat.insert(out->surfPointsLessFlat.end(), surfPointsLessFlatScanDS.begin(),
                                   surfPointsLessFlatScanDS.end());
  }
}

}  // namespace oracle
