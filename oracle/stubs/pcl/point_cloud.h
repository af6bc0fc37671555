// oracle/stubs -- TEST INFRASTRUCTURE ONLY. The reference reads only `.points` (SC.cpp:155,166-168).
#pragma once
#include <vector>
namespace pcl {
template <class PointT>
struct PointCloud {
  std::vector<PointT> points;
};
}  // namespace pcl
