// oracle/stubs -- TEST INFRASTRUCTURE ONLY. 32-byte PCL-compatible point (x,y,z,pad | intensity,pad x3).
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
  float x, y, z, _pad0;
  float intensity, _pad1, _pad2, _pad3;
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");
}  // namespace pcl
