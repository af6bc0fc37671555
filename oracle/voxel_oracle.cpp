// oracle/voxel_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
//
// CPU restatement of pcl::VoxelGrid<pcl::PointXYZI>::applyFilter, the step the reference runs immediately in
// front of the Scan Context path: mapOptmization.cpp:264 (leaf 0.5 m), 1235-1237 (filter of the raw scan),
// 1628-1630 (the filtered cloud goes into makeAndSaveScancontextAndKeys).
//
// PARITY UNPINNED.  PCL is a third-party dependency that is NOT under /root/reference (CMakeLists.txt:26
// `find_package(PCL REQUIRED QUIET)`, no version pin, no lock file; ROS Melodic ships PCL 1.8.1, Noetic 1.10) and is
// not installed in this image, so this file restates the published algorithm of
// pcl/filters/impl/voxel_grid.hpp (1.8.1 .. 1.10: getMinMax3D, leaf index, std::sort by leaf index, CentroidPoint
// accumulation) and pcl/common/impl/accumulators.hpp (AccumulatorXYZ: Eigen::Vector3f running sum, divided by n;
// AccumulatorIntensity: float running sum, divided by n) from the library's documentation of that algorithm; the
// reference has no test, golden vector or fixture at this boundary.  What IS pinned here: std::sort is the
// libstdc++ introsort of this toolchain, so the (unstable) order in which the points of one voxel are summed is
// exactly the one PCL gets when built with the same compiler.
//
//   step 1  min / max of x, y, z over the finite points                           (getMinMax3D)
//   step 2  min_b = (int) floor(min * inv_leaf), max_b likewise, div_b = max_b - min_b + 1,
//           refuse when (dx * dy * dz) overflows int32                            ("Leaf size is too small")
//   step 3  per finite point: ijk = (int)(floor(p * inv_leaf) - (float) min_b),  idx = i + j*div0 + k*div0*div1
//   step 4  std::sort of (idx, point index) by idx
//   step 5  per run of equal idx, in sorted order: xyz += p.xyz (float), intensity += p.intensity (float);
//           centroid = sum / (float) n; output voxels in ascending idx order
// inv_leaf = 1.0f / leaf (Eigen::Array4f::Ones() / leaf_size_.array()).  All arithmetic FP32 as in PCL.
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace {
struct IdxPt {
  unsigned idx;
  unsigned pt;
  bool operator<(const IdxPt& o) const { return idx < o.idx; }  // cloud_point_index_idx::operator<
};
}  // namespace

extern "C" {

// pts: n records of stride_floats floats (x, y, z, intensity first).  out_xyzi: capacity cap points (4 floats each),
// out_idx / out_count: leaf index and number of points of each output voxel.  min_b / div_b: 3 ints each (optional).
// Returns the number of output voxels, -1 when PCL would refuse (index overflow), -2 when cap is too small.
long vox_downsample(const float* pts, size_t n, size_t stride_floats, float leaf, float* out_xyzi, unsigned* out_idx,
                    unsigned* out_count, size_t cap, int* min_b_out, int* div_b_out) {
  const float inv = 1.0f / leaf;
  float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
  float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
  for (size_t i = 0; i < n; ++i) {
    const float* p = pts + i * stride_floats;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    for (int a = 0; a < 3; ++a) {
      if (p[a] < mn[a]) mn[a] = p[a];
      if (p[a] > mx[a]) mx[a] = p[a];
    }
  }
  int min_b[3], div_b[3];
  if (mn[0] > mx[0]) {  // no finite point
    if (min_b_out) min_b_out[0] = min_b_out[1] = min_b_out[2] = 0;
    if (div_b_out) div_b_out[0] = div_b_out[1] = div_b_out[2] = 0;
    return 0;
  }
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1,
                dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) return -1;
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)floorf(mn[a] * inv);
    const int max_b = (int)floorf(mx[a] * inv);
    div_b[a] = max_b - min_b[a] + 1;
    if (min_b_out) min_b_out[a] = min_b[a];
    if (div_b_out) div_b_out[a] = div_b[a];
  }
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  std::vector<IdxPt> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    const float* p = pts + i * stride_floats;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    const int i0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
    const int i1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
    const int i2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
    IdxPt e;
    e.idx = (unsigned)(i0 + i1 * mul1 + i2 * mul2);
    e.pt = (unsigned)i;
    v.push_back(e);
  }
  std::sort(v.begin(), v.end(), std::less<IdxPt>());
  size_t out = 0, a = 0;
  while (a < v.size()) {
    size_t b = a + 1;
    while (b < v.size() && v[b].idx == v[a].idx) ++b;
    if (out >= cap) return -2;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t k = a; k < b; ++k) {
      const float* p = pts + (size_t)v[k].pt * stride_floats;
      s[0] += p[0];
      s[1] += p[1];
      s[2] += p[2];
      s[3] += stride_floats > 3 ? p[3] : 0.f;
    }
    const float cnt = (float)(b - a);
    out_xyzi[4 * out] = s[0] / cnt;
    out_xyzi[4 * out + 1] = s[1] / cnt;
    out_xyzi[4 * out + 2] = s[2] / cnt;
    out_xyzi[4 * out + 3] = s[3] / cnt;
    if (out_idx) out_idx[out] = v[a].idx;
    if (out_count) out_count[out] = (unsigned)(b - a);
    ++out;
    a = b;
  }
  return (long)out;
}

}  // extern "C"
