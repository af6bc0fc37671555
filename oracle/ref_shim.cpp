// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never on the product path).
//
// A C ABI around the reference's OWN SCManager.  This file contains no Scan Context arithmetic:
// it includes the reference header from /root/reference (via -I, see oracle/Makefile) and is
// linked against the reference's Scancontext.cpp compiled verbatim from where it lies.  The
// resulting oracle/_ref/libscref*.so is what tests call "the reference" and what
// `bench.py --impl reference` times.
//
// Everything that peeks inside a detectLoopClosureID() call does so through the reference's
// public data members (Scancontext.h:99-106) AFTER the call, re-issuing the same kNN query on the
// tree the call left behind (Scancontext.cpp:283-289) and the same per-candidate
// distanceBtnScanContext (Scancontext.cpp:296-299).
#include "Scancontext.h"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

namespace {

struct NullBuf : std::streambuf {
  int overflow(int c) override { return c; }
};

struct Ref {
  SCManager m;
  std::string last_log;
};

void fill_cloud(pcl::PointCloud<SCPointType>& cloud, const void* pts, size_t n, size_t stride) {
  cloud.points.resize(n);
  const unsigned char* p = static_cast<const unsigned char*>(pts);
  for (size_t i = 0; i < n; ++i) {
    float xyz[3];
    std::memcpy(xyz, p + i * stride, sizeof xyz);
    SCPointType q{};
    q.x = xyz[0];
    q.y = xyz[1];
    q.z = xyz[2];
    q.intensity = 0.f;
    if (stride >= 20) std::memcpy(&q.intensity, p + i * stride + 16, sizeof(float));  // pcl::PointXYZI layout: intensity at byte 16
    cloud.points[i] = q;
  }
}

MatrixXd to_mat(const double* a, int r, int c) {
  MatrixXd m(r, c);
  std::memcpy(m.data(), a, sizeof(double) * size_t(r) * size_t(c));
  return m;
}

// Runs f() with std::cout captured; returns what was printed.
template <class F>
std::string with_cout_captured(F f) {
  std::ostringstream os;
  std::streambuf* old = std::cout.rdbuf(os.rdbuf());
  f();
  std::cout.rdbuf(old);
  return os.str();
}

}  // namespace

extern "C" {

void* scref_create() { return new Ref(); }
void scref_destroy(void* h) { delete static_cast<Ref*>(h); }

// Compile-time constants of THIS build of the reference (Scancontext.h:77-96).
void scref_params(void* h, int* R, int* S, int* K, int* exclude_recent, int* tree_period,
                  double* lidar_height, double* max_radius, double* search_ratio, double* dist_thres) {
  SCManager& m = static_cast<Ref*>(h)->m;
  *R = m.PC_NUM_RING;
  *S = m.PC_NUM_SECTOR;
  *K = m.NUM_CANDIDATES_FROM_TREE;
  *exclude_recent = m.NUM_EXCLUDE_RECENT;
  *tree_period = m.TREE_MAKING_PERIOD_;
  *lidar_height = m.LIDAR_HEIGHT;
  *max_radius = m.PC_MAX_RADIUS;
  *search_ratio = m.SEARCH_RATIO;
  *dist_thres = m.SC_DIST_THRES;
}

float scref_xy2theta(float x, float y) { return xy2theta(x, y); }

void scref_make_sc(void* h, const void* pts, size_t n, size_t stride, double* out) {
  SCManager& m = static_cast<Ref*>(h)->m;
  pcl::PointCloud<SCPointType> cloud;
  fill_cloud(cloud, pts, n, stride);
  MatrixXd sc = m.makeScancontext(cloud);
  std::memcpy(out, sc.data(), sizeof(double) * size_t(sc.size()));
}

void scref_ringkey(void* h, const double* sc, double* out_R) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd d = to_mat(sc, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  MatrixXd k = m.makeRingkeyFromScancontext(d);
  std::memcpy(out_R, k.data(), sizeof(double) * size_t(k.size()));
}

void scref_sectorkey(void* h, const double* sc, double* out_S) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd d = to_mat(sc, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  MatrixXd k = m.makeSectorkeyFromScancontext(d);
  std::memcpy(out_S, k.data(), sizeof(double) * size_t(k.size()));
}

int scref_fast_align(void* h, const double* vkey1, const double* vkey2) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd a = to_mat(vkey1, 1, m.PC_NUM_SECTOR), b = to_mat(vkey2, 1, m.PC_NUM_SECTOR);
  return m.fastAlignUsingVkey(a, b);
}

double scref_dist_direct(void* h, const double* sc1, const double* sc2) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd a = to_mat(sc1, m.PC_NUM_RING, m.PC_NUM_SECTOR), b = to_mat(sc2, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  return m.distDirectSC(a, b);
}

void scref_distance(void* h, const double* sc1, const double* sc2, double* dist, int* shift) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd a = to_mat(sc1, m.PC_NUM_RING, m.PC_NUM_SECTOR), b = to_mat(sc2, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  std::pair<double, int> r = m.distanceBtnScanContext(a, b);
  *dist = r.first;
  *shift = r.second;
}

void scref_append_scan(void* h, const void* pts, size_t n, size_t stride) {
  SCManager& m = static_cast<Ref*>(h)->m;
  pcl::PointCloud<SCPointType> cloud;
  fill_cloud(cloud, pts, n, stride);
  m.makeAndSaveScancontextAndKeys(cloud);
}

// Appends an already-built descriptor: the same four push_backs as Scancontext.cpp:233-240,
// with the keys produced by the reference's own key functions.
void scref_append_desc(void* h, const double* sc_in) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd sc = to_mat(sc_in, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  MatrixXd ringkey = m.makeRingkeyFromScancontext(sc);
  MatrixXd sectorkey = m.makeSectorkeyFromScancontext(sc);
  std::vector<float> v = eig2stdvec(ringkey);
  m.polarcontexts_.push_back(sc);
  m.polarcontext_invkeys_.push_back(ringkey);
  m.polarcontext_vkeys_.push_back(sectorkey);
  m.polarcontext_invkeys_mat_.push_back(v);
}

size_t scref_size(void* h) { return static_cast<Ref*>(h)->m.polarcontexts_.size(); }

void scref_get_entry(void* h, size_t i, double* sc, double* ringkey, double* sectorkey, float* ringkey_f) {
  SCManager& m = static_cast<Ref*>(h)->m;
  if (sc) std::memcpy(sc, m.polarcontexts_[i].data(), sizeof(double) * size_t(m.polarcontexts_[i].size()));
  if (ringkey) std::memcpy(ringkey, m.polarcontext_invkeys_[i].data(), sizeof(double) * size_t(m.PC_NUM_RING));
  if (sectorkey) std::memcpy(sectorkey, m.polarcontext_vkeys_[i].data(), sizeof(double) * size_t(m.PC_NUM_SECTOR));
  if (ringkey_f) std::memcpy(ringkey_f, m.polarcontext_invkeys_mat_[i].data(), sizeof(float) * size_t(m.PC_NUM_RING));
}

// detectLoopClosureID() with its stdout lines captured into log (NUL-terminated, truncated to cap).
void scref_detect(void* h, int* loop_id, float* yaw, char* log, size_t cap) {
  Ref* r = static_cast<Ref*>(h);
  std::pair<int, float> res;
  r->last_log = with_cout_captured([&] { res = r->m.detectLoopClosureID(); });
  *loop_id = res.first;
  *yaw = res.second;
  if (log && cap) {
    size_t n = r->last_log.size() < cap - 1 ? r->last_log.size() : cap - 1;
    std::memcpy(log, r->last_log.data(), n);
    log[n] = 0;
  }
}

// What the LAST detect call saw: the K tree candidates in tree order with their squared ring-key
// distances, the per-candidate (SC distance, shift), and the number of keys in the tree.
// Returns K, or 0 when the last call took the early return (DB < NUM_EXCLUDE_RECENT+1).
int scref_last_candidates(void* h, uint64_t* idx, float* d2, double* sc_dist, int* sc_shift, uint64_t* n_tree) {
  SCManager& m = static_cast<Ref*>(h)->m;
  if (m.polarcontext_invkeys_mat_.size() < size_t(m.NUM_EXCLUDE_RECENT) + 1 || !m.polarcontext_tree_) return 0;
  const int K = m.NUM_CANDIDATES_FROM_TREE;
  std::vector<float> q = m.polarcontext_invkeys_mat_.back();
  MatrixXd qd = m.polarcontexts_.back();
  std::vector<size_t> ci(K);
  std::vector<float> cd(K);
  nanoflann::KNNResultSet<float> rs(K);
  rs.init(&ci[0], &cd[0]);
  m.polarcontext_tree_->index->findNeighbors(rs, &q[0], nanoflann::SearchParams(10));
  for (int i = 0; i < K; ++i) {
    idx[i] = ci[i];
    d2[i] = cd[i];
    MatrixXd cand = m.polarcontexts_[ci[i]];
    std::pair<double, int> r = m.distanceBtnScanContext(qd, cand);
    sc_dist[i] = r.first;
    sc_shift[i] = r.second;
  }
  *n_tree = m.polarcontext_invkeys_to_search_.size();
  return K;
}

// kNN only, against an explicit key set (builds a fresh tree exactly as Scancontext.cpp:272 does).
// Used to pin brute-force retrieval against nanoflann at sizes where full detects are slow.
int scref_knn(void* h, const float* keys, size_t n, const float* query, uint64_t* idx, float* d2) {
  SCManager& m = static_cast<Ref*>(h)->m;
  const int R = m.PC_NUM_RING, K = m.NUM_CANDIDATES_FROM_TREE;
  KeyMat mat(n, std::vector<float>(R));
  for (size_t i = 0; i < n; ++i) std::memcpy(mat[i].data(), keys + i * R, sizeof(float) * R);
  InvKeyTree tree(R, mat, 10);
  std::vector<size_t> ci(K);
  std::vector<float> cd(K);
  nanoflann::KNNResultSet<float> rs(K);
  rs.init(&ci[0], &cd[0]);
  tree.index->findNeighbors(rs, query, nanoflann::SearchParams(10));
  for (int i = 0; i < K; ++i) {
    idx[i] = ci[i];
    d2[i] = cd[i];
  }
  return int(rs.size());
}

// CPU baseline: per scan makeAndSaveScancontextAndKeys + detectLoopClosureID on one thread, stdout
// to a null sink.  Returns seconds for the build calls and the detect calls separately.
void scref_time_run(void* h, const void* scans, size_t n_scans, size_t pts_per_scan, size_t stride,
                    double* sec_build, double* sec_detect, int* loop_ids, float* yaws) {
  SCManager& m = static_cast<Ref*>(h)->m;
  NullBuf nb;
  std::streambuf* old = std::cout.rdbuf(&nb);
  const unsigned char* base = static_cast<const unsigned char*>(scans);
  double tb = 0, td = 0;
  pcl::PointCloud<SCPointType> cloud;
  for (size_t s = 0; s < n_scans; ++s) {
    fill_cloud(cloud, base + s * pts_per_scan * stride, pts_per_scan, stride);  // untimed: caller-side container
    auto t0 = std::chrono::steady_clock::now();
    m.makeAndSaveScancontextAndKeys(cloud);
    auto t1 = std::chrono::steady_clock::now();
    std::pair<int, float> r = m.detectLoopClosureID();
    auto t2 = std::chrono::steady_clock::now();
    tb += std::chrono::duration<double>(t1 - t0).count();
    td += std::chrono::duration<double>(t2 - t1).count();
    if (loop_ids) loop_ids[s] = r.first;
    if (yaws) yaws[s] = r.second;
  }
  std::cout.rdbuf(old);
  *sec_build = tb;
  *sec_detect = td;
}

// CPU baseline for the exhaustive config: score ONE query descriptor against db entries [0, n) with
// the reference's distanceBtnScanContext; strict-min in index order.  Returns seconds.
double scref_time_exhaustive(void* h, const double* query_sc, size_t n, double* best_dist, int* best_shift,
                             int64_t* best_idx) {
  SCManager& m = static_cast<Ref*>(h)->m;
  MatrixXd q = to_mat(query_sc, m.PC_NUM_RING, m.PC_NUM_SECTOR);
  auto t0 = std::chrono::steady_clock::now();
  double bd = 10000000;
  int bs = 0;
  int64_t bi = 0;
  for (size_t i = 0; i < n && i < m.polarcontexts_.size(); ++i) {
    std::pair<double, int> r = m.distanceBtnScanContext(q, m.polarcontexts_[i]);
    if (r.first < bd) {
      bd = r.first;
      bs = r.second;
      bi = int64_t(i);
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  *best_dist = bd;
  *best_shift = bs;
  *best_idx = bi;
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
