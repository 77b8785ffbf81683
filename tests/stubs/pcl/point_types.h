// stub of <pcl/point_types.h>: the two point records with PCL's memory layout (16 bytes, SSE padded)
#pragma once
namespace pcl {
struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, pad = 1.f; };
struct alignas(16) PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };   // (real PCL pads this one to 32 bytes; the adapters copy field-wise)
}  // namespace pcl
