// tests/dropin_driver.cpp -- a caller shaped like mapOptmization.cpp's use of SCManager.
// Compiled TWICE from this one source: against the reference's own Scancontext.h/.cpp (oracle/Makefile ->
// oracle/_ref/dropin_ref) and against include/Scancontext.h + libscgpu.so (tests/test_gpu_dropin.py).
// Both binaries must print byte-identical output, including SCManager's own "[Loop found]" / "[Not loop]" lines.
#include "Scancontext.h"

#include <cstdint>
#include <cstdio>
#include <cstring>

extern "C" {
struct scangen_cfg {
  uint64_t seed;
  int n_beams, n_azim, n_places;
  float sensor_h, max_range, jitter, range_sigma;
};
void scangen_default_cfg(scangen_cfg* c, int hdl64);
void scangen_scan(const scangen_cfg* c, uint64_t scan_index, void* out, size_t stride);
}

static void print_bits(const char* tag, double v) {
  uint64_t u;
  std::memcpy(&u, &v, 8);
  std::printf("%s %016llx\n", tag, (unsigned long long)u);
}

struct Node {  // mapOptmization.cpp:234: a default-constructed member
  SCManager scManager;
};

int main(int argc, char** argv) {
  const int n_scans = argc > 1 ? std::atoi(argv[1]) : 120;
  Node node;
  scangen_cfg cfg;
  scangen_default_cfg(&cfg, 1);
  cfg.seed = 4711;
  cfg.n_places = (n_scans * 2) / 3;
  cfg.n_azim = 300;
  pcl::PointCloud<SCPointType> cloud, first;
  for (int i = 0; i < n_scans; ++i) {
    cloud.points.resize((size_t)cfg.n_beams * cfg.n_azim);
    scangen_scan(&cfg, (uint64_t)i, cloud.points.data(), sizeof(SCPointType));
    if (i == 0) first = cloud;
    node.scManager.makeAndSaveScancontextAndKeys(cloud);          // mapOptmization.cpp:1630
    auto detectResult = node.scManager.detectLoopClosureID();     // mapOptmization.cpp:916
    std::fflush(stdout);
    std::cout.flush();
    uint32_t ybits;
    std::memcpy(&ybits, &detectResult.second, 4);
    std::printf("kf %d -> %d %08x\n", i, detectResult.first, ybits);
  }
  // the other public methods, on real data
  Eigen::MatrixXd a = node.scManager.makeScancontext(first);
  Eigen::MatrixXd b = node.scManager.makeScancontext(cloud);
  Eigen::MatrixXd rk = node.scManager.makeRingkeyFromScancontext(a);
  Eigen::MatrixXd sa = node.scManager.makeSectorkeyFromScancontext(a), sb = node.scManager.makeSectorkeyFromScancontext(b);
  for (int i = 0; i < (int)a.size(); i += 97) print_bits("sc", a.data()[i]);
  for (int i = 0; i < (int)rk.size(); ++i) print_bits("rk", rk.data()[i]);
  for (int i = 0; i < (int)sa.size(); i += 7) print_bits("sk", sa.data()[i]);
  std::printf("align %d\n", node.scManager.fastAlignUsingVkey(sa, sb));
  print_bits("direct", node.scManager.distDirectSC(a, b));
  std::pair<double, int> d = node.scManager.distanceBtnScanContext(a, b);
  print_bits("dist", d.first);
  std::printf("shift %d\n", d.second);
  Eigen::MatrixXd sh = circshift(a, 5);
  print_bits("circ", sh(3, 7));
  std::printf("theta %.9g %.9g\n", (double)xy2theta(1.0f, 2.0f), (double)xy2theta(-3.0f, -0.5f));
  std::printf("consts %d %d %d %g\n", node.scManager.PC_NUM_RING, node.scManager.PC_NUM_SECTOR, node.scManager.NUM_CANDIDATES_FROM_TREE,
              node.scManager.SC_DIST_THRES);
  return 0;
}
