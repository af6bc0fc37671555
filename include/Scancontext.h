// Scancontext.h -- drop-in replacement of SC-LeGO-LOAM's header of the same name.
//
// mapOptmization.cpp includes "Scancontext.h", default-constructs one `SCManager scManager;` member
// (mapOptmization.cpp:234) and calls exactly two methods on it: makeAndSaveScancontextAndKeys(cloud)
// (mapOptmization.cpp:1630/1633) and detectLoopClosureID() (mapOptmization.cpp:916).  Put this directory in
// front of the reference's include path, drop Scancontext.cpp from the target, link libscgpu.so, and the
// node compiles unchanged (INTEGRATION.md shows the three CMake lines).
//
// What is here: the class surface of the reference (Scancontext.h:58-108: the eight public methods, the
// public hyper-parameters under their original names, the free helper declarations of lines 49-55, and the
// global `using` directives of lines 30-39 that mapOptmization.cpp is compiled under).  What is NOT here:
// arithmetic.  Every method marshals its arguments into the C ABI of scgpu.h, where hand-written sm_100a
// kernels do the work on a B200; a failing CUDA call is fatal (message + abort) because there is no CPU
// path to fall back to.  The descriptor database lives in GPU memory, so the reference's public std::vector
// members (Scancontext.h:99-106; never touched by the caller) do not exist; size() replaces
// polarcontexts_.size().
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <utility>
#include <vector>

#include <Eigen/Dense>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

// The reference header pulls in nanoflann and its vector-of-vectors adaptor (Scancontext.h:25-26) and opens the namespace;
// with the reference's include directory behind this one they are found and the same names stay visible to the caller.
// (The KD-tree itself is not used: retrieval is exact brute force on the device.)
#if defined(__has_include)
#if __has_include("nanoflann.hpp") && __has_include("KDTreeVectorOfVectorsAdaptor.h")
#include "nanoflann.hpp"
#include "KDTreeVectorOfVectorsAdaptor.h"
#define SCGPU_HAVE_NANOFLANN 1
#endif
#endif

#include "scgpu.h"

// the global using-directives and -declarations mapOptmization.cpp is compiled under (Scancontext.h:30-39)
using namespace Eigen;
#ifdef SCGPU_HAVE_NANOFLANN
using namespace nanoflann;
#endif

using std::cout;
using std::endl;
using std::make_pair;

using std::atan2;
using std::cos;
using std::sin;

using SCPointType = pcl::PointXYZI;  // x, y, z are read; stride = sizeof(SCPointType)
using KeyMat = std::vector<std::vector<float> >;  // Scancontext.h:42
#ifdef SCGPU_HAVE_NANOFLANN
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;  // Scancontext.h:43 (type kept; no tree is ever built)
#endif

// -- helpers the reference declares at namespace scope (Scancontext.h:49-55) ---------------------
inline void coreImportTest(void) { cout << "scancontext lib is successfully imported (scgpu, sm_100a)." << endl; }

inline float xy2theta(const float& _x, const float& _y) {
  float deg = 0.f;
  if (scgpu_xy2theta(_x, _y, &deg) != SCGPU_OK) {
    std::fprintf(stderr, "[scgpu] xy2theta: %s\n", scgpu_last_error());
    std::abort();
  }
  return deg;
}

// pure data movement (column rotation to the right / double -> float copy): no arithmetic to offload
inline MatrixXd circshift(MatrixXd& _mat, int _num_shift) {
  const int cols = static_cast<int>(_mat.cols());
  MatrixXd out(_mat.rows(), _mat.cols());
  for (int c = 0; c < cols; ++c) {
    const int dst = (c + _num_shift) % cols;
    for (int r = 0; r < static_cast<int>(_mat.rows()); ++r) out(r, dst) = _mat(r, c);
  }
  return out;
}
inline std::vector<float> eig2stdvec(MatrixXd _eigmat) {
  return std::vector<float>(_eigmat.data(), _eigmat.data() + _eigmat.size());
}

class SCManager {
 public:
  SCManager() = default;
  SCManager(const SCManager&) = delete;
  SCManager& operator=(const SCManager&) = delete;
  ~SCManager() {
    if (handle_) scgpu_destroy(handle_);
  }

  // ---- the reference's methods (Scancontext.h:63-73) ------------------------------------------------
  Eigen::MatrixXd makeScancontext(pcl::PointCloud<SCPointType>& _scan_down) {
    Eigen::MatrixXd desc(PC_NUM_RING, PC_NUM_SECTOR);
    ok(scgpu_make_sc(h(), points(_scan_down), _scan_down.points.size(), sizeof(SCPointType), desc.data()), "makeScancontext");
    return desc;
  }
  Eigen::MatrixXd makeRingkeyFromScancontext(Eigen::MatrixXd& _desc) {
    Eigen::MatrixXd key(_desc.rows(), 1);
    ok(scgpu_ringkey(h(), _desc.data(), key.data()), "makeRingkeyFromScancontext");
    return key;
  }
  Eigen::MatrixXd makeSectorkeyFromScancontext(Eigen::MatrixXd& _desc) {
    Eigen::MatrixXd key(1, _desc.cols());
    ok(scgpu_sectorkey(h(), _desc.data(), key.data()), "makeSectorkeyFromScancontext");
    return key;
  }
  int fastAlignUsingVkey(MatrixXd& _vkey1, MatrixXd& _vkey2) {
    int shift = 0;
    ok(scgpu_fast_align(h(), _vkey1.data(), _vkey2.data(), &shift), "fastAlignUsingVkey");
    return shift;
  }
  double distDirectSC(MatrixXd& _sc1, MatrixXd& _sc2) {
    double d = 0;
    ok(scgpu_dist_direct(h(), _sc1.data(), _sc2.data(), &d), "distDirectSC");
    return d;
  }
  std::pair<double, int> distanceBtnScanContext(MatrixXd& _sc1, MatrixXd& _sc2) {
    double d = 0;
    int shift = 0;
    ok(scgpu_distance(h(), _sc1.data(), _sc2.data(), &d, &shift), "distanceBtnScanContext");
    return make_pair(d, shift);
  }

  // User-side API
  void makeAndSaveScancontextAndKeys(pcl::PointCloud<SCPointType>& _scan_down) {
    ok(scgpu_append_scan(h(), points(_scan_down), _scan_down.points.size(), sizeof(SCPointType)), "makeAndSaveScancontextAndKeys");
  }

  // Extension (not in the reference class): take over the voxel-grid filter the caller runs in front of this class --
  // downSizeFilterScancontext.setLeafSize(0.5, 0.5, 0.5) / .filter() (mapOptmization.cpp:264, 1235-1237).  With a leaf
  // set, makeAndSaveScancontextAndKeys / makeScancontext expect the RAW scan and filter it on the device.
  void setDownsampleLeaf(float leaf) { ok(scgpu_set_downsample_leaf(h(), leaf), "setDownsampleLeaf"); }

  std::pair<int, float> detectLoopClosureID(void) {  // int: nearest node index, float: relative yaw
    int loop_id = -1, nn_idx = 0, nn_align = 0;
    float yaw = 0.f;
    double min_dist = 0;
    const unsigned long long before = size();
    ok(scgpu_detect(h(), &loop_id, &yaw, &min_dist, &nn_idx, &nn_align), "detectLoopClosureID");
    // the reference's log lines (Scancontext.cpp:317-330), including its sticky precision(3) in the
    // not-loop branch; nothing is printed on the early return (database smaller than NUM_EXCLUDE_RECENT+1)
    if (before >= static_cast<unsigned long long>(NUM_EXCLUDE_RECENT) + 1 && !quiet()) {
      if (min_dist < SC_DIST_THRES) {
        cout << "[Loop found] Nearest distance: " << min_dist << " btn " << before - 1 << " and " << nn_idx << "." << endl;
        cout << "[Loop found] yaw diff: " << nn_align * PC_UNIT_SECTORANGLE << " deg." << endl;
      } else {
        std::cout.precision(3);
        cout << "[Not loop] Nearest distance: " << min_dist << " btn " << before - 1 << " and " << nn_idx << "." << endl;
        cout << "[Not loop] yaw diff: " << nn_align * PC_UNIT_SECTORANGLE << " deg." << endl;
      }
    }
    return std::pair<int, float>{loop_id, yaw};
  }

  // polarcontexts_.size() of the reference
  unsigned long long size() {
    uint64_t n = 0;
    ok(scgpu_size(h(), &n), "size");
    return n;
  }
  scgpu_handle* native_handle() { return h(); }

 public:
  // hyper parameters: same names and values as Scancontext.h:77-96
  const double LIDAR_HEIGHT = 2.0;

  const int PC_NUM_RING = 20;
  const int PC_NUM_SECTOR = 60;
  const double PC_MAX_RADIUS = 80.0;
  const double PC_UNIT_SECTORANGLE = 360.0 / double(PC_NUM_SECTOR);
  const double PC_UNIT_RINGGAP = PC_MAX_RADIUS / double(PC_NUM_RING);

  const int NUM_EXCLUDE_RECENT = 50;
  const int NUM_CANDIDATES_FROM_TREE = 10;

  const double SEARCH_RATIO = 0.1;
  const double SC_DIST_THRES = 0.5;

  const int TREE_MAKING_PERIOD_ = 10;

 private:
  scgpu_handle* handle_ = nullptr;

  // the handle is created on first use from the constants above.  $SCGPU_DEVICES=0,1,... : the database is sharded over
  // those GPUs (entry i on the (i % n)-th; one process drives them all, NVLink peer access between the shards);
  // otherwise the single device $SCGPU_DEVICE (default 0).  $SCGPU_CAPACITY: keyframes to reserve (a sharded database has
  // a fixed capacity; default 65536).
  scgpu_handle* h() {
    if (!handle_) {
      scgpu_config c;
      scgpu_default_config(&c);
      c.num_ring = PC_NUM_RING;
      c.num_sector = PC_NUM_SECTOR;
      c.lidar_height = LIDAR_HEIGHT;
      c.max_radius = PC_MAX_RADIUS;
      c.exclude_recent = NUM_EXCLUDE_RECENT;
      c.num_candidates = NUM_CANDIDATES_FROM_TREE;
      c.search_ratio = SEARCH_RATIO;
      c.dist_thres = SC_DIST_THRES;
      c.tree_period = TREE_MAKING_PERIOD_;
      if (const char* d = std::getenv("SCGPU_DEVICE")) c.device = std::atoi(d);
      if (const char* cap = std::getenv("SCGPU_CAPACITY")) c.capacity_hint = std::strtoull(cap, nullptr, 10);
      if (const char* list = std::getenv("SCGPU_DEVICES")) {
        int n = 0;
        for (const char* p = list; *p && n < SCGPU_MAX_DEVICES;) {
          char* end = nullptr;
          const long v = std::strtol(p, &end, 10);
          if (end == p) break;
          c.devices[n++] = static_cast<int>(v);
          p = (*end == ',') ? end + 1 : end;
        }
        c.n_devices = n;
        if (n > 1 && !std::getenv("SCGPU_CAPACITY")) c.capacity_hint = 65536;
      }
      ok(scgpu_create(&c, &handle_), "SCManager");
    }
    return handle_;
  }
  static bool quiet() {
    const char* q = std::getenv("SCGPU_QUIET");
    return q && *q && *q != '0';
  }
  static const void* points(pcl::PointCloud<SCPointType>& cloud) {
    return cloud.points.empty() ? nullptr : static_cast<const void*>(&cloud.points[0]);
  }
  static void ok(int rc, const char* what) {
    if (rc == SCGPU_OK) return;
    std::fprintf(stderr, "[scgpu] %s failed (%d): %s -- no CPU fallback, aborting\n", what, rc, scgpu_last_error());
    std::abort();
  }
};
