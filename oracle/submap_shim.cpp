// oracle/submap_shim.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
//
// Two things behind one C interface:
//  * SUBMAP_REF defined: the reference's OWN transformPointCloud(cloud, PointTypePose*) -- the text of
//    mapOptmization.cpp:598-627, cut out of the file where it lies under /root/reference by the Makefile (awk) into the
//    git-ignored oracle/_ref/submap_fn.inc and compiled here against three stub types (the file as a whole needs ROS, PCL
//    and GTSAM; this member function uses none of them beyond `points`, `resize` and the pose fields).  -> oracle/_ref/libsubmapref.so
//  * otherwise: a plain restatement of the same arithmetic (the "port") -> oracle/_build/libsubmaporacle.so
// The point arithmetic is FP32 throughout: utility.h:49 `using namespace std` + <cmath> make cos(float) / sin(float) the
// float overloads, products and sums of floats stay float, and the reference build has no FMA (no -march, CMakeLists.txt:4-5).
#include <math.h>

#include <cmath>
#include <cstddef>
#include <memory>
#include <vector>

using namespace std;  // utility.h:49

struct PointType {  // pcl::PointXYZI (utility.h:51): the four floats the function touches
  float x, y, z, intensity;
};
struct PointTypePose {  // PointXYZIRPYT (utility.h:175-184)
  float x, y, z, intensity, roll, pitch, yaw;
  double time;
};
namespace pcl {
template <class T>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;
  std::vector<T> points;
  void resize(size_t n) { points.resize(n); }
};
}  // namespace pcl

#ifdef SUBMAP_REF
#include "submap_fn.inc"  // pcl::PointCloud<PointType>::Ptr transformPointCloud(pcl::PointCloud<PointType>::Ptr cloudIn, PointTypePose* transformIn)
#else
static pcl::PointCloud<PointType>::Ptr transformPointCloud(pcl::PointCloud<PointType>::Ptr in, PointTypePose* t) {  // mapOptmization.cpp:598-627
  pcl::PointCloud<PointType>::Ptr out(new pcl::PointCloud<PointType>());
  out->resize(in->points.size());
  const float cy = cosf(t->yaw), sy = sinf(t->yaw), cr = cosf(t->roll), sr = sinf(t->roll), cp = cosf(t->pitch), sp = sinf(t->pitch);
  for (size_t i = 0; i < in->points.size(); ++i) {
    const PointType& p = in->points[i];
    const float x1 = cy * p.x - sy * p.y, y1 = sy * p.x + cy * p.y, z1 = p.z;  // about z
    const float x2 = x1, y2 = cr * y1 - sr * z1, z2 = sr * y1 + cr * z1;      // about x
    PointType o;
    o.x = cp * x2 + sp * z2 + t->x;                                            // about y, then the translation
    o.y = y2 + t->y;
    o.z = -sp * x2 + cp * z2 + t->z;
    o.intensity = p.intensity;
    out->points[i] = o;
  }
  return out;
}
#endif

// in: n points of `stride_floats` floats (x, y, z first; intensity at float index intensity_at, < 0 = none);
// pose6 = x, y, z, roll, pitch, yaw; out: n x 4 floats
extern "C" void submap_transform(const float* in, size_t n, size_t stride_floats, int intensity_at, const float* pose6, float* out) {
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
  c->resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float* f = in + i * stride_floats;
    c->points[i] = PointType{f[0], f[1], f[2], intensity_at >= 0 ? f[intensity_at] : 0.f};
  }
  PointTypePose ps{};
  ps.x = pose6[0], ps.y = pose6[1], ps.z = pose6[2], ps.roll = pose6[3], ps.pitch = pose6[4], ps.yaw = pose6[5];
  pcl::PointCloud<PointType>::Ptr r = transformPointCloud(c, &ps);
  for (size_t i = 0; i < n; ++i) {
    out[4 * i] = r->points[i].x, out[4 * i + 1] = r->points[i].y, out[4 * i + 2] = r->points[i].z, out[4 * i + 3] = r->points[i].intensity;
  }
}

// (int)intensity >= 0 as the reference's filter of the query cloud evaluates it (mapOptmization.cpp:932-939)
extern "C" int submap_keeps(float intensity) { return (int)intensity >= 0; }
