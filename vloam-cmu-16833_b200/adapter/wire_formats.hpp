// vloam_b200 — the data formats on either side of the hot path, ROS-free C++.
//
//   load_kitti_bin        KITTI velodyne .bin -> (x, y, z, reflectance) records
//                         reference: PointCloudUtil::loadPointCloud, src/visual_odometry/src/point_cloud_util.cpp:118-146
//   PointCloud2View       sensor_msgs/PointCloud2 payload -> the (pointer, count, stride) triple vloam_scan_registration
//                         takes, WITHOUT the pcl::fromROSMsg copy of src/vloam_main/src/vloam_main_node.cpp:148 when x, y, z
//                         are consecutive float32 fields (they are in KITTI bags); a packed copy otherwise
//   Cam0StartFrameWriter  the KITTI-format pose dump: VloamTF::{VO,LO,MO}2Cam0StartFrame,
//                         src/vloam_tf/src/vloam_tf.cpp:77-153 ("%f" x 12 per line, float-cast 3 x 4 matrix)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace vloam_b200 {

// point_cloud_util.cpp:118-146: at most 1 000 000 floats are read (the reference's fixed buffer), 4 floats per point.
inline int load_kitti_bin(const std::string& path, std::vector<float>& xyzi) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return -1;
  xyzi.resize(1000000);
  const size_t got = std::fread(xyzi.data(), sizeof(float), xyzi.size(), f);
  std::fclose(f);
  const int n = (int)(got / 4);
  xyzi.resize((size_t)n * 4);
  return n;
}

// One PointField of sensor_msgs/PointCloud2 (datatype 7 = FLOAT32).
struct PointFieldDesc { std::string name; uint32_t offset; uint8_t datatype; };

struct PointCloud2View {
  const float* points = nullptr;   // first x
  int n = 0;
  int stride_floats = 0;           // floats between consecutive points
  std::vector<float> packed;       // owns the data when a copy was unavoidable
  bool zero_copy = false;

  // data: msg.data.data(); n_points = msg.width * msg.height; point_step = msg.point_step.
  // Returns false when the message has no float32 x / y / z fields.
  bool bind(const uint8_t* data, size_t n_points, uint32_t point_step, const std::vector<PointFieldDesc>& fields,
            bool is_bigendian = false) {
    int ox = -1, oy = -1, oz = -1;
    for (const auto& f : fields) {
      if (f.datatype != 7) continue;
      if (f.name == "x") ox = (int)f.offset;
      else if (f.name == "y") oy = (int)f.offset;
      else if (f.name == "z") oz = (int)f.offset;
    }
    if (ox < 0 || oy < 0 || oz < 0 || is_bigendian) return false;
    n = (int)n_points;
    const bool contiguous = oy == ox + 4 && oz == ox + 8 && point_step % 4 == 0 && ox % 4 == 0 &&
                            reinterpret_cast<uintptr_t>(data + ox) % alignof(float) == 0;
    if (contiguous) {
      points = reinterpret_cast<const float*>(data + ox);
      stride_floats = (int)(point_step / 4);
      zero_copy = true;
      return true;
    }
    packed.resize(n_points * 3);
    for (size_t i = 0; i < n_points; ++i) {
      const uint8_t* p = data + i * point_step;
      std::memcpy(&packed[3 * i + 0], p + ox, 4);
      std::memcpy(&packed[3 * i + 1], p + oy, 4);
      std::memcpy(&packed[3 * i + 2], p + oz, 4);
    }
    points = packed.data();
    stride_floats = 3;
    zero_copy = false;
    return true;
  }
};

// Rigid transform as a row-major 4 x 4 double matrix.
struct Mat4 {
  double m[16];
  static Mat4 identity() { Mat4 r{}; for (int i = 0; i < 4; ++i) r.m[5 * i] = 1.0; return r; }
  // q = (x, y, z, w), t
  static Mat4 from_qt(const double q[4], const double t[3]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    Mat4 r = identity();
    r.m[0] = 1 - 2 * (y * y + z * z); r.m[1] = 2 * (x * y - z * w);     r.m[2] = 2 * (x * z + y * w);     r.m[3] = t[0];
    r.m[4] = 2 * (x * y + z * w);     r.m[5] = 1 - 2 * (x * x + z * z); r.m[6] = 2 * (y * z - x * w);     r.m[7] = t[1];
    r.m[8] = 2 * (x * z - y * w);     r.m[9] = 2 * (y * z + x * w);     r.m[10] = 1 - 2 * (x * x + y * y); r.m[11] = t[2];
    return r;
  }
  Mat4 operator*(const Mat4& o) const {
    Mat4 r{};
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += m[4 * i + k] * o.m[4 * k + j]; r.m[4 * i + j] = s; }
    return r;
  }
  Mat4 rigid_inverse() const {  // [R t; 0 1]^-1 = [R' -R't; 0 1]
    Mat4 r = identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[4 * i + j] = m[4 * j + i];
    for (int i = 0; i < 3; ++i) r.m[4 * i + 3] = -(r.m[4 * i] * m[3] + r.m[4 * i + 1] * m[7] + r.m[4 * i + 2] * m[11]);
    return r;
  }
};

// vloam_tf.cpp:77-153: cam0_init_T_cam0_last = base_T_cam0^-1 * world_T_base_last * base_T_cam0; the first dumped frame
// becomes the origin; the 3 x 4 matrix is cast to float and printed with "%f".
class Cam0StartFrameWriter {
 public:
  explicit Cam0StartFrameWriter(const Mat4& base_T_cam0) : base_T_cam0_(base_T_cam0), cam0_T_base_(base_T_cam0.rigid_inverse()) {}
  // count = frame index relative to start_frame (vloam_main_node.cpp:171-176); negative counts are ignored.  Returns the line.
  std::string write(FILE* fp, int count, const Mat4& world_T_base_last) {
    if (count < 0) return std::string();
    const Mat4 init_T_last = cam0_T_base_ * world_T_base_last * base_T_cam0_;
    if (count == 0) start_T_init_ = init_T_last.rigid_inverse();
    const Mat4 start_T_last = start_T_init_ * init_T_last;
    char buf[512];
    int len = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        len += std::snprintf(buf + len, sizeof(buf) - (size_t)len, (i == 2 && j == 3) ? "%f\n" : "%f ", (double)(float)start_T_last.m[4 * i + j]);
    if (fp) std::fputs(buf, fp);
    return std::string(buf);
  }

 private:
  Mat4 base_T_cam0_, cam0_T_base_, start_T_init_ = Mat4::identity();
};

}  // namespace vloam_b200
