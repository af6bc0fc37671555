// scgpu.cu -- host side of libscgpu.so: the C ABI of include/scgpu.h over the sm_100a kernels in
// scgpu_kernels.cuh.  No CPU implementation of any stage lives here: every entry point either launches the
// kernels or fails.
#include "scgpu.h"

#include <cuda_runtime.h>
#include <float.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "scgpu_exhaustive.cuh"
#include "scgpu_kernels.cuh"
#include "scgpu_voxel.cuh"

using namespace scgpu;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(SCGPU_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
  } while (0)
#define RET(call)              \
  do {                         \
    int r_ = (call);           \
    if (r_ != SCGPU_OK) return r_; \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t want, bool zero = false, cudaStream_t st = 0) {
    if (want <= bytes) return SCGPU_OK;
    if (p) CK(cudaFree(p));
    p = nullptr;
    bytes = 0;
    size_t cap = want + want / 4 + 256;
    CK(cudaMalloc(&p, cap));
    if (zero) CK(cudaMemsetAsync(p, 0, cap, st));
    bytes = cap;
    return SCGPU_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t want) {
    if (want <= bytes) return SCGPU_OK;
    if (p) CK(cudaFreeHost(p));
    p = nullptr;
    bytes = 0;
    CK(cudaHostAlloc(&p, want + want / 4 + 256, cudaHostAllocDefault));
    bytes = want + want / 4 + 256;
    return SCGPU_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace

struct scgpu_handle {
  scgpu_config cfg;
  Layout L;
  int K, radius, W, slots;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_pin[2] = {nullptr, nullptr};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_t2 = nullptr;  // call start / builds done / call end
  cudaEvent_t ev_nl = nullptr;
  bool timing_valid = false;
  float voxel_leaf = 0.f;  // > 0: scans are voxel-grid filtered in front of the descriptor build (k_build_voxel)
  DevBuf vox_info, vox_pts, vox_idx, vox_in, vox_keys, vox_hint;
  // database shard
  Db db{};
  uint64_t n_global = 0;
  // screening copy for the exhaustive search (only for configurations k_exh_screen is instantiated for)
  bool exh = false;
  int exh_cfg = 0;  // 1: 20x60 radius 3, 2: 40x120 radius 6
  float* x_sc_hat = nullptr;
  unsigned char* x_vk = nullptr;  // [cap] ExhVkRec: float sector key + aux
  DevBuf x_query, x_d32, x_keys, x_pd, x_ps, x_small, x_best;
  int sm_count = 148;
  unsigned last_exh_rescored = 0;
  uint64_t x_upto = 0;  // local entries [0, x_upto) have an up-to-date screening copy
  DevBuf c_d32, c_list, c_count;
  bool last_screened = false;  // the last pipeline scored only the candidates that could win exactly
  const void* last_qrec = nullptr;
  // build workspace
  DevBuf gbins, btickets;
  size_t build_cap = 0;
  // point staging
  DevBuf d_pts[2];
  PinBuf h_pts[2];
  int pin_turn = 0;
  // query workspace
  DevBuf records, rec_single, nsearch, keys, partial, ttickets, pair_dist, pair_shift, best, o_loop, o_yaw, o_dist, o_idx, o_shift;
  DevBuf api_in, api_out;
  PinBuf h_out, h_ns;
  size_t ttickets_cap = 0;
  // snapshot state (SC.h:96, SC.cpp:264-276)
  long long counter = 0;
  uint64_t n_tree = 0;
  // last pipeline (candidate dumps)
  size_t last_nq = 0;
  std::vector<uint64_t> last_nsearch;
  uint64_t launches = 0;
};

namespace {

uint64_t local_count(const scgpu_handle* h, uint64_t n_global) {
  const uint64_t G = (uint64_t)h->cfg.shard_count, r = (uint64_t)h->cfg.shard_rank;
  return n_global > r ? (n_global - 1 - r) / G + 1 : 0;
}

int db_reserve(scgpu_handle* h, uint64_t want_local) {
  if (want_local <= h->db.cap) return SCGPU_OK;
  CK(cudaDeviceSynchronize());  // rare: the shard moves to a larger allocation; nothing may still be reading the old one
  uint64_t cap = h->db.cap ? h->db.cap * 2 : 1024;
  while (cap < want_local) cap *= 2;
  const Layout& L = h->L;
  Db nd = h->db;
  nd.cap = cap;
  CK(cudaMalloc(&nd.sc, cap * L.RS * sizeof(float)));
  CK(cudaMalloc(&nd.ringT, cap * L.R * sizeof(float)));
  CK(cudaMalloc(&nd.sector, cap * L.S * sizeof(double)));
  CK(cudaMalloc(&nd.colnorm, cap * L.S * sizeof(double)));
  float* n_hat = nullptr;
  unsigned char* n_vk = nullptr;
  if (h->exh) {
    CK(cudaMalloc(&n_hat, cap * L.RS * sizeof(float)));
    CK(cudaMalloc(&n_vk, (cap + 8) * exh_vk_bytes(L.S)));
  }
  const uint64_t n = local_count(h, h->n_global);
  if (h->x_upto > n) h->x_upto = n;
  if (h->exh && h->db.cap && h->x_upto) {
    const uint64_t n = h->x_upto;
    CK(cudaMemcpyAsync(n_hat, h->x_sc_hat, n * L.RS * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(n_vk, h->x_vk, n * exh_vk_bytes(L.S), cudaMemcpyDeviceToDevice, h->stream));
  }
  if (h->db.cap && n) {
    CK(cudaMemcpyAsync(nd.sc, h->db.sc, n * L.RS * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpy2DAsync(nd.ringT, cap * sizeof(float), h->db.ringT, h->db.cap * sizeof(float), n * sizeof(float), L.R,
                         cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(nd.sector, h->db.sector, n * L.S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(nd.colnorm, h->db.colnorm, n * L.S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  if (h->db.cap) {
    cudaFree(h->db.sc);
    cudaFree(h->db.ringT);
    cudaFree(h->db.sector);
    cudaFree(h->db.colnorm);
    if (h->exh) {
      cudaFree(h->x_sc_hat);
      cudaFree(h->x_vk);
    }
  }
  h->db = nd;
  h->x_sc_hat = n_hat;
  h->x_vk = n_vk;
  return SCGPU_OK;
}

int build_reserve(scgpu_handle* h, size_t n_scans, cudaStream_t st) {
  if (n_scans <= h->build_cap) return SCGPU_OK;
  CK(cudaStreamSynchronize(st));
  const size_t cap = n_scans + n_scans / 2 + 8;
  h->build_cap = 0;
  RET(h->gbins.reserve(cap * h->L.RS * sizeof(int)));
  RET(h->btickets.reserve(cap * sizeof(unsigned)));
  const size_t n = h->gbins.bytes / sizeof(int);
  k_fill_int<<<256, 256, 0, st>>>(h->gbins.as<int>(), n, SCGPU_ENC_NOPOINT);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(h->btickets.p, 0, h->btickets.bytes, st));
  h->build_cap = cap;
  return SCGPU_OK;
}

// Voxel-grid filter (pcl::VoxelGrid, mapOpt.cpp:264,1235-1237) fused with stage 1+2; device-resident points.
// d_records may be null (downsample only); out_* optional device buffers [n_scans][out_cap].
int launch_build_voxel(scgpu_handle* h, const void* d_pts, size_t n_scans, size_t pts_per_scan, size_t stride, float leaf, void* d_records,
                       VoxInfo* d_info, float4* d_out_pts, unsigned* d_out_idx, unsigned out_cap, cudaStream_t st) {
  if (n_scans == 0) return SCGPU_OK;
  if (!(leaf > 0.f) || !(leaf < 1e30f)) return fail(SCGPU_E_INVALID, "voxel leaf size must be positive and finite");
  if (stride < 12 || (stride & 3) || ((uintptr_t)d_pts & 3)) return fail(SCGPU_E_INVALID, "points must be 4-byte aligned, stride >= 12 and a multiple of 4");
  if (pts_per_scan > 0xfffffff0ull) return fail(SCGPU_E_INVALID, "scan too large");
  VoxelBuildParams p;
  p.pts = static_cast<const unsigned char*>(d_pts);
  p.scan_pitch = (unsigned long long)pts_per_scan * stride;
  p.n_pts = (unsigned)pts_per_scan;
  p.stride = (unsigned)stride;
  p.inv_leaf = 1.0f / leaf;
  p.bc = make_bin_const(h->L.R, h->L.S, h->cfg.lidar_height, h->cfg.max_radius, 1);
  p.L = h->L;
  p.out_cap = out_cap;
  // scratch for the leaf indices: launches are cut so that it stays below ~512 MB
  size_t per_launch = pts_per_scan ? (size_t)(512u << 20) / (pts_per_scan * 8) : 65535;
  if (per_launch < 1) per_launch = 1;
  if (per_launch > 65535) per_launch = 65535;  // gridDim.y limit
  if (per_launch > n_scans) per_launch = n_scans;
  RET(h->vox_keys.reserve(per_launch * pts_per_scan * 8 + 16));
  if (!h->vox_hint.p) {
    RET(h->vox_hint.reserve(16));
    CK(cudaMemsetAsync(h->vox_hint.p, 0, 16, st));
  }
  p.keys = h->vox_keys.as<unsigned>();
  p.passes_hint = h->vox_hint.as<int>();
  const size_t smem = vox_smem_bytes(h->L.RS);
  const bool al16 = (((uintptr_t)d_pts & 15) == 0);
  const int sk = (stride == 16 && al16) ? 16 : ((stride == 32 && al16) ? 32 : 0);
  for (size_t s0 = 0; s0 < n_scans; s0 += per_launch) {
    const size_t ns = n_scans - s0 < per_launch ? n_scans - s0 : per_launch;
    VoxelBuildParams q = p;
    q.pts = p.pts + s0 * p.scan_pitch;
    q.records = d_records ? static_cast<unsigned char*>(d_records) + s0 * h->L.rec_bytes : nullptr;
    q.info = d_info ? d_info + s0 : nullptr;
    q.out_pts = d_out_pts ? d_out_pts + s0 * out_cap : nullptr;
    q.out_idx = d_out_idx ? d_out_idx + s0 * out_cap : nullptr;
    dim3 grid(VOX_CLUSTER, (unsigned)ns);
    if (sk == 16) k_build_voxel<16><<<grid, VOX_THREADS, smem, st>>>(q);
    else if (sk == 32) k_build_voxel<32><<<grid, VOX_THREADS, smem, st>>>(q);
    else k_build_voxel<0><<<grid, VOX_THREADS, smem, st>>>(q);
    h->launches++;
    CK(cudaGetLastError());
  }
  return SCGPU_OK;
}

// Stage 1+2 on device-resident points.
int launch_build(scgpu_handle* h, const void* d_pts, size_t n_scans, size_t pts_per_scan, size_t stride, void* d_records,
                 cudaStream_t st) {
  if (n_scans == 0) return SCGPU_OK;
  if (h->voxel_leaf > 0.f && pts_per_scan > 0)
    return launch_build_voxel(h, d_pts, n_scans, pts_per_scan, stride, h->voxel_leaf, d_records, nullptr, nullptr, nullptr, 0, st);
  if (stride < 12 || (stride & 3) || ((uintptr_t)d_pts & 3)) return fail(SCGPU_E_INVALID, "points must be 4-byte aligned, stride >= 12 and a multiple of 4");
  if (pts_per_scan > 0xfffffff0ull || n_scans > 65535ull * 1024) return fail(SCGPU_E_INVALID, "scan too large");
  RET(build_reserve(h, n_scans, st));
  BuildParams p;
  p.pts = static_cast<const unsigned char*>(d_pts);
  p.scan_pitch = (unsigned long long)pts_per_scan * stride;
  p.n_pts = (unsigned)pts_per_scan;
  p.stride = (unsigned)stride;
  // tile: enough blocks for two full waves of the GPU (4 resident blocks per SM), otherwise as large as possible -- a
  // scan binned by ONE block needs no global merge (atomics, fences, ticket) and starts / drains its TMA ring once
  // (4,541 HDL-64 scans: 16k-point tiles 1.71 ms, whole-scan tiles 1.56 ms)
  unsigned ppb = 256;
  if (pts_per_scan) {
    const uint64_t want = (uint64_t)h->sm_count * 8;
    uint64_t tiles = (want + n_scans - 1) / n_scans;
    const uint64_t max_tiles = (pts_per_scan + 2047) / 2048;
    if (tiles > max_tiles) tiles = max_tiles;
    if (tiles < 1) tiles = 1;
    ppb = (unsigned)(((pts_per_scan + tiles - 1) / tiles + 1023) / 1024 * 1024);
  }
  if (const char* e = getenv("SCGPU_BUILD_TILE")) {  // experiments: points per block (a multiple of 1024)
    const long v = atol(e);
    if (v >= 1024 && v % 1024 == 0) ppb = (unsigned)v;
  }
  p.pts_per_block = ppb;
  p.bc = make_bin_const(h->L.R, h->L.S, h->cfg.lidar_height, h->cfg.max_radius, !(h->cfg.flags & SCGPU_FLAG_EXACT_BINNING));
  p.L = h->L;
  p.gbins = h->gbins.as<int>();
  p.tickets = h->btickets.as<unsigned>();
  p.records = static_cast<unsigned char*>(d_records);
  const unsigned tiles = pts_per_scan ? (unsigned)((pts_per_scan + ppb - 1) / ppb) : 1;
  const size_t smem = (size_t)h->L.RS * sizeof(int);
  for (size_t s0 = 0; s0 < n_scans; s0 += 65535) {  // gridDim.y limit
    const size_t ns = n_scans - s0 < 65535 ? n_scans - s0 : 65535;
    BuildParams q = p;
    q.pts = p.pts + s0 * p.scan_pitch;
    q.records = p.records + s0 * h->L.rec_bytes;
    q.gbins = p.gbins;  // per-launch scan index restarts at 0: workspace rows [0, ns)
    dim3 grid(tiles, (unsigned)ns);
    const bool al16 = (((uintptr_t)q.pts & 15) == 0);
    const int sk = (stride == 16 && al16) ? 16 : ((stride == 32 && al16) ? 32 : 0);
    if (sk && q.bc.fast && !(h->cfg.flags & SCGPU_FLAG_NO_TMA_BUILD)) {
      // TMA-staged variant (tile start offsets are multiples of 16 bytes because pts_per_block * stride is)
      const size_t sm = sk == 16 ? build_tma_smem<16>(h->L.RS) : build_tma_smem<32>(h->L.RS);
      if (sk == 16 && q.bc.lh_is_float) k_build_tma<16, true, true><<<grid, 256, sm, st>>>(q);
      else if (sk == 16) k_build_tma<16, true, false><<<grid, 256, sm, st>>>(q);
      else if (q.bc.lh_is_float) k_build_tma<32, true, true><<<grid, 256, sm, st>>>(q);
      else k_build_tma<32, true, false><<<grid, 256, sm, st>>>(q);
      h->launches++;
      CK(cudaGetLastError());
      continue;
    }
    const int variant = sk * 4 + (q.bc.fast ? 2 : 0) + (q.bc.lh_is_float ? 1 : 0);
#define SCGPU_BUILD_CASE(SK, F, L) \
  case (SK) * 4 + ((F) ? 2 : 0) + ((L) ? 1 : 0): k_build<SK, F, L><<<grid, 256, smem, st>>>(q); break;
    switch (variant) {
      SCGPU_BUILD_CASE(16, true, true)
      SCGPU_BUILD_CASE(16, true, false)
      SCGPU_BUILD_CASE(32, true, true)
      SCGPU_BUILD_CASE(32, true, false)
      SCGPU_BUILD_CASE(0, true, true)
      SCGPU_BUILD_CASE(0, true, false)
      default:  // exact binning: one instantiation per stride kind
        if (sk == 16) k_build<16, false, false><<<grid, 256, smem, st>>>(q);
        else if (sk == 32) k_build<32, false, false><<<grid, 256, smem, st>>>(q);
        else k_build<0, false, false><<<grid, 256, smem, st>>>(q);
    }
#undef SCGPU_BUILD_CASE
    h->launches++;
    CK(cudaGetLastError());
  }
  return SCGPU_OK;
}

int launch_append(scgpu_handle* h, const void* d_records, uint64_t first_global, uint64_t step, size_t n, cudaStream_t st) {
  if (n == 0) return SCGPU_OK;
  if (step < 1) return fail(SCGPU_E_INVALID, "global_step must be >= 1");
  const uint64_t new_size = first_global + (n - 1) * step + 1;
  if (new_size > 0xffffffffull) return fail(SCGPU_E_INVALID, "database index space is 32 bits");
  RET(db_reserve(h, local_count(h, new_size)));
  k_append<<<(unsigned)n, 128, 0, st>>>(static_cast<const unsigned char*>(d_records), h->L, h->db, first_global, step);
  h->launches++;
  CK(cudaGetLastError());
  if (new_size > h->n_global) h->n_global = new_size;
  return SCGPU_OK;
}

void choose_chunks(uint64_t n_local, size_t nq, unsigned& chunk, unsigned& chunks) {
  if (n_local == 0) {
    chunk = 256;
    chunks = 1;
    return;
  }
  uint64_t by_size = (n_local + 255) / 256;
  uint64_t by_grid = (592 + nq - 1) / nq;
  uint64_t c = by_size < by_grid ? by_size : by_grid;
  if (c < 1) c = 1;
  uint64_t ch = (n_local + c - 1) / c;
  ch = (ch + 255) / 256 * 256;
  chunk = (unsigned)ch;
  chunks = (unsigned)((n_local + ch - 1) / ch);
}

int query_reserve(scgpu_handle* h, size_t nq, unsigned chunks, cudaStream_t st) {
  const int K = h->K;
  RET(h->keys.reserve(nq * K * sizeof(uint64_t)));
  RET(h->partial.reserve(nq * (size_t)chunks * K * sizeof(uint64_t)));
  if (nq > h->ttickets_cap) {
    CK(cudaStreamSynchronize(st));
    h->ttickets_cap = 0;
    RET(h->ttickets.reserve(nq * sizeof(unsigned)));
    CK(cudaMemsetAsync(h->ttickets.p, 0, h->ttickets.bytes, st));
    h->ttickets_cap = h->ttickets.bytes / sizeof(unsigned);
  }
  RET(h->pair_dist.reserve(nq * K * sizeof(double)));
  RET(h->pair_shift.reserve(nq * K * sizeof(int)));
  RET(h->best.reserve(nq * sizeof(Best)));
  RET(h->o_loop.reserve(nq * sizeof(int)));
  RET(h->o_yaw.reserve(nq * sizeof(float)));
  RET(h->o_dist.reserve(nq * sizeof(double)));
  RET(h->o_idx.reserve(nq * sizeof(int)));
  RET(h->o_shift.reserve(nq * sizeof(int)));
  return SCGPU_OK;
}

int launch_topk(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_ns, uint64_t* d_keys_out, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  if (nq > 65535) return fail(SCGPU_E_INVALID, "at most 65535 queries per call");
  unsigned chunk, chunks;
  const uint64_t n_local = local_count(h, h->n_global);
  // large batches over a large shard: groups of TOPK_QT queries share every key they load (k_topk_tile) -- only when each
  // warp's stream stays long (>= 8k keys), see the kernel's header.  SCGPU_TOPK_TILE=1/0 forces / forbids it (tests).
  bool tile = nq >= 16 && h->slots <= 2 && (h->L.R == 20 || h->L.R == 40);
  const size_t groups = (nq + TOPK_QT - 1) / TOPK_QT;
  if (tile) {
    choose_chunks(n_local, groups, chunk, chunks);
    static const char* force = getenv("SCGPU_TOPK_TILE");
    tile = force ? atoi(force) != 0 : (uint64_t)chunk / TOPK_TILE_WARPS >= 8192;
  }
  const size_t units = tile ? groups : nq;
  choose_chunks(n_local, units, chunk, chunks);
  RET(query_reserve(h, nq, chunks, st));
  TopkParams p;
  p.qrecords = static_cast<const unsigned char*>(d_qrec);
  p.L = h->L;
  p.db = h->db;
  p.n_search = reinterpret_cast<const unsigned long long*>(d_ns);
  p.n_local = n_local;
  p.chunk = chunk;
  p.K = h->K;
  p.partial = h->partial.as<unsigned long long>();
  p.tickets = h->ttickets.as<unsigned>();
  p.keys_out = reinterpret_cast<unsigned long long*>(d_keys_out);
  if (tile) {
    dim3 tgrid(chunks, (unsigned)units);
    const unsigned n = (unsigned)nq;
    if (h->slots == 1 && h->L.R == 20) k_topk_tile<1, 20><<<tgrid, TOPK_TILE_WARPS * 32, 0, st>>>(p, n);
    else if (h->slots == 1) k_topk_tile<1, 40><<<tgrid, TOPK_TILE_WARPS * 32, 0, st>>>(p, n);
    else if (h->L.R == 20) k_topk_tile<2, 20><<<tgrid, TOPK_TILE_WARPS * 32, 0, st>>>(p, n);
    else k_topk_tile<2, 40><<<tgrid, TOPK_TILE_WARPS * 32, 0, st>>>(p, n);
    h->launches++;
    CK(cudaGetLastError());
    return SCGPU_OK;
  }
  dim3 grid(chunks, (unsigned)nq);
  const bool many = (uint64_t)chunks * nq >= 4096;  // enough blocks to fill the GPU with 2-warp blocks
  if (h->slots == 1) {
    if (many) k_topk<1, 2><<<grid, 64, 0, st>>>(p);
    else k_topk<1, 8><<<grid, 256, 0, st>>>(p);
  } else if (h->slots == 2) {
    if (many) k_topk<2, 2><<<grid, 64, 0, st>>>(p);
    else k_topk<2, 8><<<grid, 256, 0, st>>>(p);
  } else {
    if (many) k_topk<4, 2><<<grid, 64, 0, st>>>(p);
    else k_topk<4, 8><<<grid, 256, 0, st>>>(p);
  }
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int launch_merge(scgpu_handle* h, const uint64_t* d_parts, int parts, size_t nq, uint64_t* d_out, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  const unsigned blocks = (unsigned)((nq + 3) / 4);
  const unsigned long long* in = reinterpret_cast<const unsigned long long*>(d_parts);
  unsigned long long* out = reinterpret_cast<unsigned long long*>(d_out);
  if (h->slots == 1) k_merge<1><<<blocks, 128, 0, st>>>(in, parts, (unsigned)nq, h->K, out);
  else if (h->slots == 2) k_merge<2><<<blocks, 128, 0, st>>>(in, parts, (unsigned)nq, h->K, out);
  else k_merge<4><<<blocks, 128, 0, st>>>(in, parts, (unsigned)nq, h->K, out);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

// candidates -> per (query, slot) distance/shift in h->pair_*; K_eff slots per query
int launch_score(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_keys, const uint64_t* d_ns, int K_eff,
                 double* d_pair_dist, int* d_pair_shift, int flip, cudaStream_t st) {
  if (nq == 0 || K_eff == 0) return SCGPU_OK;
  ScoreParams p;
  p.qrecords = static_cast<const unsigned char*>(d_qrec);
  p.L = h->L;
  p.db = h->db;
  p.keys = reinterpret_cast<const unsigned long long*>(d_keys);
  p.n_search = reinterpret_cast<const unsigned long long*>(d_ns);
  p.K = K_eff;
  p.radius = h->radius;
  p.pair_dist = d_pair_dist;
  p.pair_shift = d_pair_shift;
  p.flip = flip;
  p.active = nullptr;
  const size_t smem = pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float));
  dim3 grid((unsigned)K_eff, (unsigned)nq);
  k_score<<<grid, 128, smem, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

// bring the screening copy of the shard up to date (entries appended since the last search that used it)
int exh_sync(scgpu_handle* h, cudaStream_t st) {
  const uint64_t have = local_count(h, h->n_global);
  if (have > h->x_upto) {
    k_exh_append<<<(unsigned)(have - h->x_upto), 128, 0, st>>>(h->L, h->db, h->x_sc_hat, h->x_vk, h->x_upto);
    h->launches++;
    CK(cudaGetLastError());
    h->x_upto = have;
  }
  return SCGPU_OK;
}

// Stage 4 for the top-K path: FP32 screening of all K candidates, exact FP64 scoring of those that can be the minimum.
int launch_score_screened(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_keys, const uint64_t* d_ns, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  RET(exh_sync(h, st));
  RET(h->c_d32.reserve(nq * h->K * sizeof(float)));
  RET(h->c_list.reserve(nq * h->K * sizeof(uint64_t)));
  RET(h->c_count.reserve(16));
  CK(cudaMemsetAsync(h->c_count.p, 0, 4, st));
  CandScreenParams cp;
  cp.qrecords = static_cast<const unsigned char*>(d_qrec);
  cp.L = h->L;
  cp.db = h->db;
  cp.xdb.sc_hat = h->x_sc_hat;
  cp.xdb.vk = h->x_vk;
  cp.keys = reinterpret_cast<const unsigned long long*>(d_keys);
  cp.n_search = reinterpret_cast<const unsigned long long*>(d_ns);
  cp.K = h->K;
  cp.d32 = h->c_d32.as<float>();
  // warps per block, staging slots per warp.  One slot: three blocks fit an SM and cover each other's fetch latency
  // (two slots = one block per SM measured slower at K = 50).
  if (h->exh_cfg == 1) k_cand_screen<20, 60, 3, 1, 10, 1><<<(unsigned)nq, 10 * 32, cand_smem_bytes<20, 60, 3, 10, 1>(), st>>>(cp);
  else k_cand_screen<40, 120, 6, 2, 5, 1><<<(unsigned)nq, 5 * 32, cand_smem_bytes<40, 120, 6, 5, 1>(), st>>>(cp);
  k_cand_select<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(h->c_d32.as<float>(), (unsigned)nq, h->K, h->c_list.as<unsigned long long>(),
                                                             h->c_count.as<unsigned>(), h->pair_dist.as<double>(), h->pair_shift.as<int>());
  ScoreParams p;
  p.qrecords = static_cast<const unsigned char*>(d_qrec);
  p.L = h->L;
  p.db = h->db;
  p.keys = reinterpret_cast<const unsigned long long*>(d_keys);
  p.n_search = reinterpret_cast<const unsigned long long*>(d_ns);
  p.K = h->K;
  p.radius = h->radius;
  p.pair_dist = h->pair_dist.as<double>();
  p.pair_shift = h->pair_shift.as<int>();
  p.flip = 0;
  p.active = nullptr;
  k_score_pairs<<<(unsigned)h->sm_count * 8, 128, pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float)), st>>>(p, h->c_list.as<unsigned long long>(),
                                                                                                            h->c_count.as<unsigned>());
  h->launches += 3;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int launch_best(scgpu_handle* h, size_t nq, const uint64_t* d_keys, Best* d_best, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  k_best<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(h->pair_dist.as<double>(), h->pair_shift.as<int>(),
                                                       reinterpret_cast<const unsigned long long*>(d_keys), (unsigned)nq, h->K, d_best);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int launch_finalize(scgpu_handle* h, const Best* d_parts, int parts, size_t nq, const uint64_t* d_ns, int* d_loop, float* d_yaw,
                    double* d_dist, int* d_idx, int* d_shift, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  k_finalize<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(d_parts, parts, (unsigned)nq, reinterpret_cast<const unsigned long long*>(d_ns),
                                                           h->K, h->L.S, h->cfg.dist_thres, d_loop, d_yaw, d_dist, d_idx, d_shift);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

// n_search plan on the host; advances the snapshot state exactly like SC.cpp:257-276
void plan(scgpu_handle* h, uint64_t first_size, size_t n, uint64_t* out) {
  const uint64_t excl = (uint64_t)h->cfg.exclude_recent;
  const bool fresh = (h->cfg.flags & SCGPU_FLAG_FRESH_TREE) || h->cfg.tree_period <= 0;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t size = first_size + i;
    if (size < excl + 1) {
      out[i] = 0;
      continue;
    }
    if (fresh || (h->counter % h->cfg.tree_period) == 0) h->n_tree = size - excl;
    h->counter++;
    out[i] = h->n_tree;
  }
}

// full single-shard query pipeline for nq query records already on the device; results land in h->o_*.
int run_pipeline(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* h_ns) {
  cudaStream_t st = h->stream;
  RET(h->nsearch.reserve(nq * sizeof(uint64_t)));
  RET(h->h_ns.reserve(nq * sizeof(uint64_t)));
  memcpy(h->h_ns.p, h_ns, nq * sizeof(uint64_t));
  CK(cudaMemcpyAsync(h->nsearch.p, h->h_ns.p, nq * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  unsigned chunk, chunks;
  choose_chunks(local_count(h, h->n_global), nq, chunk, chunks);
  RET(query_reserve(h, nq, chunks, st));
  const uint64_t* d_ns = h->nsearch.as<uint64_t>();
  RET(launch_topk(h, d_qrec, nq, d_ns, h->keys.as<uint64_t>(), st));
  h->last_screened = h->exh && !(h->cfg.flags & SCGPU_FLAG_NO_SCREENING);
  h->last_qrec = d_qrec;
  if (h->last_screened) RET(launch_score_screened(h, d_qrec, nq, h->keys.as<uint64_t>(), d_ns, st));
  else RET(launch_score(h, d_qrec, nq, h->keys.as<uint64_t>(), d_ns, h->K, h->pair_dist.as<double>(), h->pair_shift.as<int>(), 0, st));
  RET(launch_best(h, nq, h->keys.as<uint64_t>(), h->best.as<Best>(), st));
  RET(launch_finalize(h, h->best.as<Best>(), 1, nq, d_ns, h->o_loop.as<int>(), h->o_yaw.as<float>(), h->o_dist.as<double>(),
                      h->o_idx.as<int>(), h->o_shift.as<int>(), st));
  h->last_nq = nq;
  h->last_nsearch.assign(h_ns, h_ns + nq);
  return SCGPU_OK;
}

// device results -> caller arrays (one packed D2H through pinned memory)
int fetch_results(scgpu_handle* h, size_t nq, int* loop_id, float* yaw, double* dist, int* idx, int* shift) {
  const size_t per = sizeof(int) + sizeof(float) + sizeof(double) + 2 * sizeof(int);
  RET(h->h_out.reserve(nq * per + 64));
  unsigned char* b = static_cast<unsigned char*>(h->h_out.p);
  double* hd = reinterpret_cast<double*>(b);
  int* hl = reinterpret_cast<int*>(b + nq * 8);
  float* hy = reinterpret_cast<float*>(b + nq * 12);
  int* hi = reinterpret_cast<int*>(b + nq * 16);
  int* hs = reinterpret_cast<int*>(b + nq * 20);
  cudaStream_t st = h->stream;
  CK(cudaMemcpyAsync(hl, h->o_loop.p, nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hy, h->o_yaw.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (dist) CK(cudaMemcpyAsync(hd, h->o_dist.p, nq * 8, cudaMemcpyDeviceToHost, st));
  if (idx) CK(cudaMemcpyAsync(hi, h->o_idx.p, nq * 4, cudaMemcpyDeviceToHost, st));
  if (shift) CK(cudaMemcpyAsync(hs, h->o_shift.p, nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (loop_id) memcpy(loop_id, hl, nq * 4);
  if (yaw) memcpy(yaw, hy, nq * 4);
  if (dist) memcpy(dist, hd, nq * 8);
  if (idx) memcpy(idx, hi, nq * 4);
  if (shift) memcpy(shift, hs, nq * 4);
  return SCGPU_OK;
}

bool is_pinned_or_device(const void* p, int* is_device) {
  cudaPointerAttributes a;
  *is_device = 0;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) {
    *is_device = 1;
    return true;
  }
  return a.type == cudaMemoryTypeHost;
}

// Host points -> records, double-buffered H2D on the copy stream overlapped with k_build on the compute stream.
int build_from_host(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, void* d_records) {
  const size_t scan_bytes = pts_per_scan * stride;
  if (scan_bytes == 0) return launch_build(h, h->d_pts[0].p ? h->d_pts[0].p : d_records, n_scans, 0, stride ? stride : 16, d_records, h->stream);
  size_t per_chunk = (size_t)(32u << 20) / scan_bytes;
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > n_scans) per_chunk = n_scans;
  RET(h->d_pts[0].reserve(per_chunk * scan_bytes));
  if (per_chunk < n_scans) RET(h->d_pts[1].reserve(per_chunk * scan_bytes));
  int dev;
  const bool pinned = is_pinned_or_device(pts, &dev) && !dev;
  const unsigned char* src = static_cast<const unsigned char*>(pts);
  int turn = 0;
  for (size_t s0 = 0; s0 < n_scans; s0 += per_chunk, turn ^= 1) {
    const size_t ns = n_scans - s0 < per_chunk ? n_scans - s0 : per_chunk;
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[turn], 0));
    if (pinned) {
      CK(cudaMemcpyAsync(h->d_pts[turn].p, src + s0 * scan_bytes, ns * scan_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    } else {
      // pageable source: stage through our own pinned buffer so the copy is truly asynchronous
      RET(h->h_pts[turn].reserve(per_chunk * scan_bytes));
      CK(cudaEventSynchronize(h->ev_pin[turn]));
      memcpy(h->h_pts[turn].p, src + s0 * scan_bytes, ns * scan_bytes);
      CK(cudaMemcpyAsync(h->d_pts[turn].p, h->h_pts[turn].p, ns * scan_bytes, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaEventRecord(h->ev_pin[turn], h->copy_stream));
    }
    CK(cudaEventRecord(h->ev_copied[turn], h->copy_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_copied[turn], 0));
    RET(launch_build(h, h->d_pts[turn].p, ns, pts_per_scan, stride, static_cast<unsigned char*>(d_records) + s0 * h->L.rec_bytes, h->stream));
    CK(cudaEventRecord(h->ev_consumed[turn], h->stream));
  }
  return SCGPU_OK;
}

int build_any(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location, void* d_records) {
  if (location == 1) return launch_build(h, pts, n_scans, pts_per_scan, stride, d_records, h->stream);
  return build_from_host(h, pts, n_scans, pts_per_scan, stride, d_records);
}

int pair_api(scgpu_handle* h, const double* a, size_t na, const double* b, size_t nb, int mode, double* out_d, size_t nd, int* out_i) {
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  RET(h->api_in.reserve((na + nb) * sizeof(double)));
  RET(h->api_out.reserve((nd + 2) * sizeof(double) + 16));
  double* d_a = h->api_in.as<double>();
  double* d_b = d_a + na;
  CK(cudaMemcpyAsync(d_a, a, na * sizeof(double), cudaMemcpyHostToDevice, st));
  if (nb) CK(cudaMemcpyAsync(d_b, b, nb * sizeof(double), cudaMemcpyHostToDevice, st));
  double* d_od = h->api_out.as<double>();
  int* d_oi = reinterpret_cast<int*>(d_od + nd + 1);
  const int W = (mode == 0) ? h->W : 1;
  const size_t smem = pair_smem_bytes(h->L.R, h->L.S, W, sizeof(double));
  k_pair_api<<<1, 128, smem, st>>>(d_a, d_b, h->L.R, h->L.S, h->radius, mode, d_od, d_oi);
  h->launches++;
  CK(cudaGetLastError());
  std::vector<double> hd(nd + 2);
  CK(cudaMemcpyAsync(hd.data(), d_od, (nd + 2) * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (out_d) memcpy(out_d, hd.data(), nd * sizeof(double));
  if (out_i) memcpy(out_i, reinterpret_cast<int*>(hd.data() + nd + 1), sizeof(int));
  return SCGPU_OK;
}

constexpr unsigned EXH_CAND_CAP = 65536;   // rescoring list of one batch
constexpr size_t EXH_MAX_BATCH = 64;       // queries per screening launch

// Screen + rescore for nq (<= EXH_MAX_BATCH) query records on the device; result q (Best: dist, rank = #rescored in the
// batch, shift, global idx) goes to d_best_out[q].  Everything is enqueued on st; no host synchronisation.
int launch_exhaustive_fast(scgpu_handle* h, const unsigned char* d_qrecs, size_t nq, const uint64_t* h_n_search, Best* d_best_out,
                           cudaStream_t st, cudaEvent_t ev_screen0, cudaEvent_t ev_screen1, int flipped = 0) {
  if (nq == 0) return SCGPU_OK;
  const size_t rows = nq * (flipped ? 2 : 1);  // screening rows: forward (+ column-reversed) pass per query
  if (rows > EXH_MAX_BATCH) return fail(SCGPU_E_INVALID, "at most %zu screening rows per exhaustive batch", EXH_MAX_BATCH);
  uint64_t n_max = 0;
  std::vector<unsigned long long> nl(nq);
  for (size_t i = 0; i < nq; ++i) {
    nl[i] = local_count(h, h_n_search[i]);
    if (nl[i] > n_max) n_max = nl[i];
  }
  const uint64_t pitch = (n_max + 15) & ~15ull;
  RET(exh_sync(h, st));
  RET(h->x_query.reserve(EXH_MAX_BATCH * sizeof(ExhQuery)));
  RET(h->x_d32.reserve((rows * pitch + 16) * sizeof(float)));
  RET(h->x_keys.reserve(EXH_CAND_CAP * sizeof(uint64_t)));
  RET(h->x_pd.reserve(EXH_CAND_CAP * sizeof(double)));
  RET(h->x_ps.reserve(EXH_CAND_CAP * sizeof(int)));
  if (!h->x_small.p) {
    RET(h->x_small.reserve(4096));
    CK(cudaMemsetAsync(h->x_small.p, 0xff, 4096, st));  // includes the constant non-zero "n_search" cell k_score_list reads
  }
  // x_small: [0] count | [2..3] non-zero u64 | [16 .. 16+64) min_bits | [256 ..] n_local (u64 x 64)
  unsigned* d_count = h->x_small.as<unsigned>();
  unsigned long long* d_one = reinterpret_cast<unsigned long long*>(d_count + 2);
  unsigned* d_min = d_count + 16;
  unsigned long long* d_nl = reinterpret_cast<unsigned long long*>(h->x_small.as<unsigned char>() + 1024);
  RET(h->h_ns.reserve(EXH_MAX_BATCH * 8 * 2));
  CK(cudaEventSynchronize(h->ev_nl));  // the pinned staging cell of the previous batch has been consumed
  memcpy(h->h_ns.p, nl.data(), nq * 8);
  CK(cudaMemcpyAsync(d_nl, h->h_ns.p, nq * 8, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(h->ev_nl, st));
  k_exh_prep<<<(unsigned)nq, 256, 0, st>>>(d_qrecs, h->L, h->x_query.as<ExhQuery>(), d_min, d_count);
  h->launches += 1;
  if (n_max) {
    ExhScreenParams sp;
    sp.db.sc_hat = h->x_sc_hat;
    sp.db.vk = h->x_vk;
    sp.q = h->x_query.as<ExhQuery>();
    sp.n_local = d_nl;
    sp.d32_pitch = pitch;
    sp.d32 = h->x_d32.as<float>();
    sp.min_bits = d_min;
    sp.flip_mode = flipped ? 1 : 0;
    const uint64_t ew = h->exh_cfg == 1 ? 20 : 4;  // consumer warps per block of the instantiation
    const uint64_t groups = (n_max + ew - 1) / ew;
    const unsigned grid = (unsigned)(groups < (uint64_t)h->sm_count ? groups : (uint64_t)h->sm_count);
    if (ev_screen0) CK(cudaEventRecord(ev_screen0, st));
    if (h->exh_cfg == 1)
      k_exh_screen<20, 60, 3, 1, 20><<<dim3(grid, (unsigned)rows), 20 * 32, exh_smem_bytes<20, 60, 3, 20>(), st>>>(sp);
    else
      k_exh_screen<40, 120, 6, 1, 4, 2><<<dim3(grid, (unsigned)rows), 4 * 2 * 32, exh_smem_bytes<40, 120, 6, 4, 2>(), st>>>(sp);
    if (ev_screen1) CK(cudaEventRecord(ev_screen1, st));
    CK(cudaGetLastError());
    const unsigned rb = (unsigned)((n_max + 1023) / 1024 < 296 ? (n_max + 1023) / 1024 : 296);
    k_exh_compact<<<dim3(rb, (unsigned)rows), 256, 0, st>>>(sp.d32, pitch, d_nl, d_min, h->db.rank, h->db.G, sp.flip_mode,
                                                           h->x_keys.as<unsigned long long>(), d_count, EXH_CAND_CAP);
    ScoreParams p;
    p.qrecords = d_qrecs;
    p.L = h->L;
    p.db = h->db;
    p.keys = h->x_keys.as<unsigned long long>();
    p.n_search = d_one;
    p.K = (int)EXH_CAND_CAP;
    p.radius = h->radius;
    p.pair_dist = h->x_pd.as<double>();
    p.pair_shift = h->x_ps.as<int>();
    p.flip = 0;
    p.active = d_count;
    k_score_list<<<dim3((unsigned)h->sm_count * 4, 1), 128, pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float)), st>>>(p);
    h->launches += 3;
    CK(cudaGetLastError());
  }
  k_exh_final<<<(unsigned)nq, 256, 0, st>>>(h->x_pd.as<double>(), h->x_ps.as<int>(), h->x_keys.as<unsigned long long>(), d_count, EXH_CAND_CAP, d_best_out);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* scgpu_last_error(void) { return g_err; }
const char* scgpu_version(void) {
  return "scgpu 0.1 (sm_100a; kernels: k_build k_build_tma k_build_voxel k_append k_gather k_topk k_topk_tile k_merge k_cand_screen "
         "k_cand_select k_score k_score_pairs k_score_list k_best k_finalize k_pair_api k_exh_prep k_exh_append k_exh_screen k_exh_compact "
         "k_exh_final)";
}

int scgpu_default_config(scgpu_config* c) {
  if (!c) return fail(SCGPU_E_INVALID, "null config");
  memset(c, 0, sizeof *c);
  c->num_ring = 20;
  c->num_sector = 60;
  c->lidar_height = 2.0;
  c->max_radius = 80.0;
  c->exclude_recent = 50;
  c->num_candidates = 10;
  c->search_ratio = 0.1;
  c->dist_thres = 0.5;
  c->tree_period = 10;
  c->device = 0;
  c->shard_rank = 0;
  c->shard_count = 1;
  c->capacity_hint = 8192;
  return SCGPU_OK;
}

int scgpu_create(const scgpu_config* cfg, scgpu_handle** out) {
  if (!cfg || !out) return fail(SCGPU_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->num_ring < 1 || cfg->num_ring > 64 || cfg->num_sector < 1 || cfg->num_sector > 1024)
    return fail(SCGPU_E_INVALID, "num_ring must be in [1,64], num_sector in [1,1024]");
  if (cfg->num_candidates < 1 || cfg->num_candidates > 128) return fail(SCGPU_E_INVALID, "num_candidates must be in [1,128]");
  if (cfg->exclude_recent < 0 || cfg->shard_count < 1 || cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count)
    return fail(SCGPU_E_INVALID, "bad exclude_recent / shard placement");
  if (!(cfg->search_ratio >= 0.0) || !(cfg->search_ratio <= 2.0)) return fail(SCGPU_E_INVALID, "search_ratio must be in [0,2]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SCGPU_E_NODEVICE, "no CUDA device: libscgpu has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(SCGPU_E_INVALID, "device %d out of range (%d devices)", cfg->device, ndev);
  CK(cudaSetDevice(cfg->device));
  scgpu_handle* h = new scgpu_handle();
  h->cfg = *cfg;
  h->L = make_layout(cfg->num_ring, cfg->num_sector);
  h->K = cfg->num_candidates;
  h->slots = h->K <= 32 ? 1 : (h->K <= 64 ? 2 : 4);
  h->radius = (int)llround(0.5 * cfg->search_ratio * cfg->num_sector);  // SC.cpp:123
  h->W = 2 * h->radius + 1;
  h->db.rank = cfg->shard_rank;
  h->db.G = cfg->shard_count;
  // FP32 screening kernels are instantiated for the reference's 20x60 (radius 3) and BASELINE's 40x120 (radius 6)
  h->exh_cfg = (h->L.R == 20 && h->L.S == 60 && h->radius == 3) ? 1 : ((h->L.R == 40 && h->L.S == 120 && h->radius == 6) ? 2 : 0);
  h->exh = h->exh_cfg != 0 && !(cfg->flags & SCGPU_FLAG_NO_SCREENING);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
  const size_t smem_f = pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float));
  const size_t smem_d = pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(double));
  if (smem_d > 220 * 1024) {
    delete h;
    return fail(SCGPU_E_INVALID, "descriptor %dx%d does not fit shared memory", cfg->num_ring, cfg->num_sector);
  }
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t1);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t2);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_nl, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<16>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<16>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<32, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<32>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<32, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<32>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess && h->exh)
    e = h->exh_cfg == 1 ? cudaFuncSetAttribute(k_exh_screen<20, 60, 3, 1, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)exh_smem_bytes<20, 60, 3, 20>())
                        : cudaFuncSetAttribute(k_exh_screen<40, 120, 6, 1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)exh_smem_bytes<40, 120, 6, 4, 2>());
  if (e == cudaSuccess && h->exh && h->exh_cfg == 1)
    e = cudaFuncSetAttribute(k_cand_screen<20, 60, 3, 1, 10, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cand_smem_bytes<20, 60, 3, 10, 1>());
  if (e == cudaSuccess && h->exh && h->exh_cfg == 2)
    e = cudaFuncSetAttribute(k_cand_screen<40, 120, 6, 2, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cand_smem_bytes<40, 120, 6, 5, 1>());
  if (e == cudaSuccess && smem_f > 48 * 1024) e = cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
  if (e == cudaSuccess && smem_f > 48 * 1024) e = cudaFuncSetAttribute(k_score_list, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
  if (e == cudaSuccess && smem_f > 48 * 1024) e = cudaFuncSetAttribute(k_score_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
  if (e == cudaSuccess && smem_d > 48 * 1024) e = cudaFuncSetAttribute(k_pair_api, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
  if (e != cudaSuccess) {
    delete h;
    return fail(SCGPU_E_CUDA, "handle setup: %s", cudaGetErrorString(e));
  }
  uint64_t cap = cfg->capacity_hint / (uint64_t)cfg->shard_count + 1;
  if (cap < 1024) cap = 1024;
  int r = db_reserve(h, cap);
  if (r == SCGPU_OK) r = h->rec_single.reserve(h->L.rec_bytes);
  if (r != SCGPU_OK) {
    delete h;
    return r;
  }
  *out = h;
  return SCGPU_OK;
}

int scgpu_destroy(scgpu_handle* h) {
  if (!h) return SCGPU_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  DevBuf* bufs[] = {&h->gbins, &h->btickets, &h->d_pts[0], &h->d_pts[1], &h->records, &h->rec_single, &h->nsearch, &h->keys, &h->partial,
                    &h->ttickets, &h->pair_dist, &h->pair_shift, &h->best, &h->o_loop, &h->o_yaw, &h->o_dist, &h->o_idx, &h->o_shift,
                    &h->api_in, &h->api_out, &h->vox_info, &h->vox_pts, &h->vox_idx, &h->vox_in, &h->vox_keys, &h->vox_hint};
  for (DevBuf* b : bufs) b->release();
  h->h_pts[0].release();
  h->h_pts[1].release();
  h->h_out.release();
  h->h_ns.release();
  if (h->db.cap) {
    cudaFree(h->db.sc);
    cudaFree(h->db.ringT);
    cudaFree(h->db.sector);
    cudaFree(h->db.colnorm);
    if (h->exh) {
      cudaFree(h->x_sc_hat);
      cudaFree(h->x_vk);
    }
  }
  DevBuf* xb[] = {&h->x_query, &h->x_d32, &h->x_keys, &h->x_pd, &h->x_ps, &h->x_small, &h->x_best, &h->c_d32, &h->c_list, &h->c_count};
  for (DevBuf* b : xb) b->release();
  for (int i = 0; i < 2; ++i) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
    if (h->ev_pin[i]) cudaEventDestroy(h->ev_pin[i]);
  }
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->ev_t2) cudaEventDestroy(h->ev_t2);
  if (h->ev_nl) cudaEventDestroy(h->ev_nl);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
  return SCGPU_OK;
}

int scgpu_size(scgpu_handle* h, uint64_t* out_n) {
  if (!h || !out_n) return fail(SCGPU_E_INVALID, "null argument");
  *out_n = h->n_global;
  return SCGPU_OK;
}

int scgpu_launch_count(scgpu_handle* h, uint64_t* out) {
  if (!h || !out) return fail(SCGPU_E_INVALID, "null argument");
  *out = h->launches;
  return SCGPU_OK;
}

int scgpu_get_timing(scgpu_handle* h, double* ms_total, double* ms_build, double* ms_query) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->timing_valid) return fail(SCGPU_E_INVALID, "no timed call yet");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventSynchronize(h->ev_t2));
  float a = 0, b = 0, c = 0;
  CK(cudaEventElapsedTime(&a, h->ev_t0, h->ev_t2));
  CK(cudaEventElapsedTime(&b, h->ev_t0, h->ev_t1));
  CK(cudaEventElapsedTime(&c, h->ev_t1, h->ev_t2));
  if (ms_total) *ms_total = a;
  if (ms_build) *ms_build = b;
  if (ms_query) *ms_query = c;
  return SCGPU_OK;
}

int scgpu_record_bytes(scgpu_handle* h, size_t* out) {
  if (!h || !out) return fail(SCGPU_E_INVALID, "null argument");
  *out = h->L.rec_bytes;
  return SCGPU_OK;
}

// ---- reference surface ----------------------------------------------------------------------------

int scgpu_make_sc(scgpu_handle* h, const void* pts, size_t n, size_t stride, double* out_sc) {
  if (!h || !out_sc || (!pts && n)) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  RET(build_from_host(h, pts, 1, n, stride ? stride : 16, h->rec_single.p));
  std::vector<float> sc(h->L.RS);
  CK(cudaMemcpyAsync(sc.data(), h->rec_single.p, sizeof(float) * h->L.RS, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->L.RS; ++i) out_sc[i] = (double)sc[i];
  return SCGPU_OK;
}

int scgpu_ringkey(scgpu_handle* h, const double* sc, double* out_ring) {
  if (!h || !sc || !out_ring) return fail(SCGPU_E_INVALID, "null argument");
  std::vector<double> o(h->L.R + h->L.S);
  RET(pair_api(h, sc, h->L.RS, nullptr, 0, 3, o.data(), o.size(), nullptr));
  memcpy(out_ring, o.data(), sizeof(double) * h->L.R);
  return SCGPU_OK;
}

int scgpu_sectorkey(scgpu_handle* h, const double* sc, double* out_sector) {
  if (!h || !sc || !out_sector) return fail(SCGPU_E_INVALID, "null argument");
  std::vector<double> o(h->L.R + h->L.S);
  RET(pair_api(h, sc, h->L.RS, nullptr, 0, 3, o.data(), o.size(), nullptr));
  memcpy(out_sector, o.data() + h->L.R, sizeof(double) * h->L.S);
  return SCGPU_OK;
}

int scgpu_fast_align(scgpu_handle* h, const double* v1, const double* v2, int* out_shift) {
  if (!h || !v1 || !v2 || !out_shift) return fail(SCGPU_E_INVALID, "null argument");
  return pair_api(h, v1, h->L.S, v2, h->L.S, 2, nullptr, 1, out_shift);
}

int scgpu_dist_direct(scgpu_handle* h, const double* sc1, const double* sc2, double* out_dist) {
  if (!h || !sc1 || !sc2 || !out_dist) return fail(SCGPU_E_INVALID, "null argument");
  return pair_api(h, sc1, h->L.RS, sc2, h->L.RS, 1, out_dist, 1, nullptr);
}

int scgpu_distance(scgpu_handle* h, const double* sc1, const double* sc2, double* out_dist, int* out_shift) {
  if (!h || !sc1 || !sc2 || !out_dist || !out_shift) return fail(SCGPU_E_INVALID, "null argument");
  return pair_api(h, sc1, h->L.RS, sc2, h->L.RS, 0, out_dist, 1, out_shift);
}

int scgpu_append_scans_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location) {
  if (!h || (!pts && n_scans && pts_per_scan)) return fail(SCGPU_E_INVALID, "null argument");
  if (n_scans == 0) return SCGPU_OK;
  CK(cudaSetDevice(h->cfg.device));
  RET(h->records.reserve(n_scans * h->L.rec_bytes));
  CK(cudaEventRecord(h->ev_t0, h->stream));
  RET(build_any(h, pts, n_scans, pts_per_scan, stride, location, h->records.p));
  CK(cudaEventRecord(h->ev_t1, h->stream));
  RET(launch_append(h, h->records.p, h->n_global, 1, n_scans, h->stream));
  CK(cudaEventRecord(h->ev_t2, h->stream));
  h->timing_valid = true;
  if (location == 0) CK(cudaStreamSynchronize(h->copy_stream));  // the caller's buffer is not retained
  return SCGPU_OK;
}

int scgpu_append_scan(scgpu_handle* h, const void* pts, size_t n, size_t stride) {
  return scgpu_append_scans_batched(h, pts, 1, n, stride, 0);
}

int scgpu_append_descs(scgpu_handle* h, const float* sc, size_t n) {
  if (!h || (!sc && n)) return fail(SCGPU_E_INVALID, "null argument");
  if (n == 0) return SCGPU_OK;
  CK(cudaSetDevice(h->cfg.device));
  const size_t batch = 8192;
  DevBuf tmp;
  RET(tmp.reserve(batch * h->L.RS * sizeof(float)));
  int rc = SCGPU_OK;
  for (size_t s0 = 0; s0 < n && rc == SCGPU_OK; s0 += batch) {
    const size_t m = n - s0 < batch ? n - s0 : batch;
    rc = h->records.reserve(m * h->L.rec_bytes);
    if (rc != SCGPU_OK) break;
    cudaError_t e = cudaMemcpyAsync(tmp.p, sc + s0 * h->L.RS, m * h->L.RS * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) {
      rc = fail(SCGPU_E_CUDA, "H2D descriptors: %s", cudaGetErrorString(e));
      break;
    }
    k_records_from_sc<<<(unsigned)m, 128, h->L.RS * sizeof(float), h->stream>>>(tmp.as<float>(), h->L, h->records.as<unsigned char>());
    h->launches++;
    rc = launch_append(h, h->records.p, h->n_global, 1, m, h->stream);
    if (rc == SCGPU_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(SCGPU_E_CUDA, "append_descs sync failed");
  }
  tmp.release();
  return rc;
}

int scgpu_detect(scgpu_handle* h, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift) {
  if (!h || !loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "scgpu_detect needs the whole database on one device; use the staged API for shards");
  if (h->n_global == 0) return fail(SCGPU_E_EMPTY, "detect on an empty database");
  CK(cudaSetDevice(h->cfg.device));
  uint64_t ns;
  plan(h, h->n_global, 1, &ns);
  if (ns == 0) {  // SC.cpp:257-261
    *loop_id = -1;
    *yaw = 0.0f;
    if (nearest_dist) *nearest_dist = 10000000.0;
    if (nearest_idx) *nearest_idx = 0;
    if (nearest_shift) *nearest_shift = 0;
    h->last_nq = 0;
    return SCGPU_OK;
  }
  k_gather<<<1, 128, 0, h->stream>>>(h->rec_single.as<unsigned char>(), h->L, h->db, h->n_global - 1);
  h->launches++;
  CK(cudaGetLastError());
  RET(run_pipeline(h, h->rec_single.p, 1, &ns));
  return fetch_results(h, 1, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

int scgpu_replay_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location,
                         int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift) {
  if (!h || (!pts && n_scans && pts_per_scan) || !loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "scgpu_replay_batched is single-shard; use the staged API for shards");
  if (n_scans == 0) return SCGPU_OK;
  if (n_scans > 65535) return fail(SCGPU_E_INVALID, "at most 65535 scans per call");
  CK(cudaSetDevice(h->cfg.device));
  const uint64_t first = h->n_global;
  std::vector<uint64_t> ns(n_scans);
  plan(h, first + 1, n_scans, ns.data());
  RET(h->records.reserve(n_scans * h->L.rec_bytes));
  CK(cudaEventRecord(h->ev_t0, h->stream));
  RET(build_any(h, pts, n_scans, pts_per_scan, stride, location, h->records.p));
  CK(cudaEventRecord(h->ev_t1, h->stream));
  RET(launch_append(h, h->records.p, first, 1, n_scans, h->stream));
  RET(run_pipeline(h, h->records.p, n_scans, ns.data()));
  CK(cudaEventRecord(h->ev_t2, h->stream));
  h->timing_valid = true;
  RET(fetch_results(h, n_scans, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift));
  if (location == 0) CK(cudaStreamSynchronize(h->copy_stream));
  return SCGPU_OK;
}

int scgpu_query_batched(scgpu_handle* h, uint64_t first, size_t nq, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx,
                        int* nearest_shift) {
  if (!h || !loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "scgpu_query_batched is single-shard; use the staged API for shards");
  if (nq == 0) return SCGPU_OK;
  if (first + nq > h->n_global || nq > 65535) return fail(SCGPU_E_INVALID, "query range outside the database");
  CK(cudaSetDevice(h->cfg.device));
  std::vector<uint64_t> ns(nq);
  const uint64_t excl = (uint64_t)h->cfg.exclude_recent;
  for (size_t i = 0; i < nq; ++i) {
    const uint64_t size = first + i + 1;
    ns[i] = size >= excl + 1 ? size - excl : 0;
  }
  RET(h->records.reserve(nq * h->L.rec_bytes));
  CK(cudaEventRecord(h->ev_t0, h->stream));
  CK(cudaEventRecord(h->ev_t1, h->stream));
  k_gather<<<(unsigned)nq, 128, 0, h->stream>>>(h->records.as<unsigned char>(), h->L, h->db, first);
  h->launches++;
  CK(cudaGetLastError());
  RET(run_pipeline(h, h->records.p, nq, ns.data()));
  CK(cudaEventRecord(h->ev_t2, h->stream));
  h->timing_valid = true;
  return fetch_results(h, nq, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

int scgpu_get_batch_candidates(scgpu_handle* h, size_t q, uint64_t* cand_idx, float* cand_d2, double* cand_dist, int* cand_shift,
                               uint64_t* n_search) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (q >= h->last_nq) return fail(SCGPU_E_INVALID, "no such query in the last call");
  CK(cudaSetDevice(h->cfg.device));
  if (h->last_screened) {
    // the pipeline scored exactly only the candidates that could win; complete the table for the dump
    RET(launch_score(h, h->last_qrec, h->last_nq, h->keys.as<uint64_t>(), h->nsearch.as<uint64_t>(), h->K, h->pair_dist.as<double>(),
                     h->pair_shift.as<int>(), 0, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->last_screened = false;
  }
  const int K = h->K;
  std::vector<uint64_t> keys(K);
  std::vector<double> pd(K);
  std::vector<int> ps(K);
  CK(cudaMemcpy(keys.data(), h->keys.as<uint64_t>() + q * K, K * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pd.data(), h->pair_dist.as<double>() + q * K, K * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ps.data(), h->pair_shift.as<int>() + q * K, K * sizeof(int), cudaMemcpyDeviceToHost));
  for (int k = 0; k < K; ++k) {
    const bool none = keys[k] == ~0ull;
    // unfilled slots keep the reference's initial state: index 0, distance 0, last slot FLT_MAX (SC.cpp:283-284, nf.hpp:159-165)
    if (cand_idx) cand_idx[k] = none ? 0 : (keys[k] & 0xffffffffull);
    if (cand_d2) {
      uint32_t bits = (uint32_t)(keys[k] >> 32);
      float f;
      memcpy(&f, &bits, 4);
      cand_d2[k] = none ? (k == K - 1 ? FLT_MAX : 0.0f) : f;
    }
    if (cand_dist) cand_dist[k] = pd[k];
    if (cand_shift) cand_shift[k] = ps[k];
  }
  if (n_search) *n_search = h->last_nsearch[q];
  return SCGPU_OK;
}

int scgpu_get_candidates(scgpu_handle* h, uint64_t* cand_idx, float* cand_d2, double* cand_dist, int* cand_shift, uint64_t* n_search) {
  return scgpu_get_batch_candidates(h, 0, cand_idx, cand_d2, cand_dist, cand_shift, n_search);
}

int scgpu_get_entry(scgpu_handle* h, uint64_t i, float* sc, float* ring, double* sector) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (i >= h->n_global) return fail(SCGPU_E_INVALID, "entry out of range");
  if ((int)(i % (uint64_t)h->cfg.shard_count) != h->cfg.shard_rank) return fail(SCGPU_E_INVALID, "entry lives on another shard");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  const uint64_t l = i / (uint64_t)h->cfg.shard_count;
  if (sc) CK(cudaMemcpy(sc, h->db.sc + l * h->L.RS, sizeof(float) * h->L.RS, cudaMemcpyDeviceToHost));
  if (ring) CK(cudaMemcpy2D(ring, sizeof(float), h->db.ringT + l, h->db.cap * sizeof(float), sizeof(float), h->L.R, cudaMemcpyDeviceToHost));
  if (sector) CK(cudaMemcpy(sector, h->db.sector + l * h->L.S, sizeof(double) * h->L.S, cudaMemcpyDeviceToHost));
  return SCGPU_OK;
}

int scgpu_truncate(scgpu_handle* h, uint64_t n) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (n < h->n_global) h->n_global = n;
  if (h->x_upto > local_count(h, h->n_global)) h->x_upto = local_count(h, h->n_global);
  // the snapshot may reference forgotten entries: the next detect takes a fresh one (counter % period == 0)
  h->counter = 0;
  h->n_tree = 0;
  return SCGPU_OK;
}

int scgpu_plan_n_search(scgpu_handle* h, uint64_t first_size, size_t n, uint64_t* out) {
  if (!h || (!out && n)) return fail(SCGPU_E_INVALID, "null argument");
  plan(h, first_size, n, out);
  return SCGPU_OK;
}

// Exhaustive scoring with the exact FP64 pair kernel (every entry is a "candidate").
int scgpu_exhaustive(scgpu_handle* h, uint64_t q, uint64_t n_search, int flipped, double* best_dist, int* best_shift, int64_t* best_idx,
                     int* best_flip) {
  if (!h || !best_dist || !best_shift || !best_idx) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "scgpu_exhaustive is single-shard; use the staged API for shards");
  if (q >= h->n_global || n_search > h->n_global) return fail(SCGPU_E_INVALID, "range outside the database");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  *best_dist = 10000000.0;
  *best_shift = 0;
  *best_idx = 0;
  if (best_flip) *best_flip = 0;
  if (n_search == 0) return SCGPU_OK;
  k_gather<<<1, 128, 0, st>>>(h->rec_single.as<unsigned char>(), h->L, h->db, q);
  h->launches++;
  if (h->exh) {
    RET(h->x_best.reserve(sizeof(Best)));
    CK(cudaEventRecord(h->ev_t0, st));
    RET(launch_exhaustive_fast(h, h->rec_single.as<unsigned char>(), 1, &n_search, h->x_best.as<Best>(), st, h->ev_t0, h->ev_t1, flipped));
    CK(cudaEventRecord(h->ev_t2, st));
    h->timing_valid = true;
    Best b;
    CK(cudaMemcpyAsync(&b, h->x_best.p, sizeof b, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if ((unsigned)b.rank <= EXH_CAND_CAP) {
      *best_dist = b.dist;
      *best_shift = b.shift & 0x3fffffff;
      *best_idx = b.idx;
      if (best_flip) *best_flip = (b.shift >> 30) & 1;
      h->last_exh_rescored = (unsigned)b.rank;
      return SCGPU_OK;
    }
    // more near-ties than the rescoring list holds (e.g. a database of duplicates): score everything exactly
  }
  h->last_exh_rescored = (unsigned)n_search;
  const size_t n = (size_t)n_search;
  DevBuf keys, pd, ps, ns;
  RET(keys.reserve(n * 8));
  RET(pd.reserve(n * 8 * 2));
  RET(ps.reserve(n * 4 * 2));
  RET(ns.reserve(8));
  std::vector<uint64_t> hk(n);
  for (size_t i = 0; i < n; ++i) hk[i] = i;
  uint64_t one = n_search;
  CK(cudaMemcpyAsync(keys.p, hk.data(), n * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ns.p, &one, 8, cudaMemcpyHostToDevice, st));
  int rc = SCGPU_OK;
  const size_t slab = 1u << 20;
  for (int f = 0; f <= (flipped ? 1 : 0) && rc == SCGPU_OK; ++f)
    for (size_t s0 = 0; s0 < n && rc == SCGPU_OK; s0 += slab) {
      const size_t m = n - s0 < slab ? n - s0 : slab;
      // one "query" with m candidate slots per launch (grid.x = slots)
      ScoreParams p;
      p.qrecords = h->rec_single.as<unsigned char>();
      p.L = h->L;
      p.db = h->db;
      p.keys = keys.as<unsigned long long>() + s0;
      p.n_search = ns.as<unsigned long long>();
      p.K = (int)m;
      p.radius = h->radius;
      p.pair_dist = pd.as<double>() + f * n + s0;
      p.pair_shift = ps.as<int>() + f * n + s0;
      p.flip = f;
      p.active = nullptr;
      k_score<<<dim3((unsigned)m, 1), 128, pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float)), st>>>(p);
      h->launches++;
      if (cudaGetLastError() != cudaSuccess) rc = fail(SCGPU_E_CUDA, "k_score launch failed");
    }
  std::vector<double> hd(n * (flipped ? 2 : 1));
  std::vector<int> hs(n * (flipped ? 2 : 1));
  if (rc == SCGPU_OK && cudaMemcpyAsync(hd.data(), pd.p, hd.size() * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(SCGPU_E_CUDA, "D2H failed");
  if (rc == SCGPU_OK && cudaMemcpyAsync(hs.data(), ps.p, hs.size() * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = fail(SCGPU_E_CUDA, "D2H failed");
  if (rc == SCGPU_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(SCGPU_E_CUDA, "exhaustive sync failed");
  keys.release();
  pd.release();
  ps.release();
  ns.release();
  if (rc != SCGPU_OK) return rc;
  // strict-min in (index, forward-before-flipped) order: the argmin reduction of SC.cpp:296-311 over all entries
  for (size_t i = 0; i < n; ++i)
    for (int f = 0; f <= (flipped ? 1 : 0); ++f) {
      const double d = hd[f * n + i];
      if (d < *best_dist) {
        *best_dist = d;
        *best_shift = hs[f * n + i];
        *best_idx = (int64_t)i;
        if (best_flip) *best_flip = f;
      }
    }
  return SCGPU_OK;
}

int scgpu_exhaustive_batched(scgpu_handle* h, const uint64_t* q, const uint64_t* n_search, size_t nq, double* best_dist, int* best_shift,
                             int64_t* best_idx) {
  if (!h || ((!q || !n_search || !best_dist || !best_shift || !best_idx) && nq)) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "scgpu_exhaustive_batched is single-shard; use the staged API for shards");
  for (size_t i = 0; i < nq; ++i)
    if (q[i] >= h->n_global || n_search[i] > h->n_global) return fail(SCGPU_E_INVALID, "range outside the database");
  if (nq == 0) return SCGPU_OK;
  if (!h->exh) {
    for (size_t i = 0; i < nq; ++i) RET(scgpu_exhaustive(h, q[i], n_search[i], 0, best_dist + i, best_shift + i, best_idx + i, nullptr));
    return SCGPU_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  RET(h->records.reserve(nq * h->L.rec_bytes));
  RET(h->x_best.reserve(nq * sizeof(Best)));
  CK(cudaEventRecord(h->ev_t0, st));
  CK(cudaEventRecord(h->ev_t1, st));
  for (size_t i = 0; i < nq; ++i) {
    unsigned char* rec = h->records.as<unsigned char>() + i * h->L.rec_bytes;
    k_gather<<<1, 128, 0, st>>>(rec, h->L, h->db, q[i]);
    h->launches++;
  }
  for (size_t i0 = 0; i0 < nq; i0 += EXH_MAX_BATCH) {
    const size_t m = nq - i0 < EXH_MAX_BATCH ? nq - i0 : EXH_MAX_BATCH;
    RET(launch_exhaustive_fast(h, h->records.as<unsigned char>() + i0 * h->L.rec_bytes, m, n_search + i0, h->x_best.as<Best>() + i0, st, nullptr,
                               nullptr));
  }
  CK(cudaEventRecord(h->ev_t2, st));
  h->timing_valid = true;
  std::vector<Best> b(nq);
  CK(cudaMemcpyAsync(b.data(), h->x_best.p, nq * sizeof(Best), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (size_t i = 0; i < nq; ++i) {
    if (n_search[i] == 0 || (unsigned)b[i].rank > EXH_CAND_CAP) {
      RET(scgpu_exhaustive(h, q[i], n_search[i], 0, best_dist + i, best_shift + i, best_idx + i, nullptr));
      continue;
    }
    best_dist[i] = b[i].dist;
    best_shift[i] = b[i].shift & 0x3fffffff;
    best_idx[i] = b[i].idx;
    h->last_exh_rescored = (unsigned)b[i].rank;
  }
  return SCGPU_OK;
}

int scgpu_exhaustive_stats(scgpu_handle* h, uint64_t* rescored) {
  if (!h || !rescored) return fail(SCGPU_E_INVALID, "null argument");
  *rescored = h->last_exh_rescored;
  return SCGPU_OK;
}

int scgpu_save(scgpu_handle* h, const char* path) {
  if (!h || !path) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "save is single-shard");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  const uint64_t n = h->n_global;
  std::vector<float> sc((size_t)n * h->L.RS);
  if (n) CK(cudaMemcpy(sc.data(), h->db.sc, sc.size() * sizeof(float), cudaMemcpyDeviceToHost));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(SCGPU_E_IO, "cannot open %s for writing", path);
  const char magic[8] = {'S', 'C', 'G', 'P', 'U', 'D', 'B', '1'};
  uint64_t hdr[4] = {(uint64_t)h->L.R, (uint64_t)h->L.S, n, 0};
  bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(hdr, 8, 4, f) == 4 && fwrite(&h->cfg, sizeof h->cfg, 1, f) == 1 &&
            (sc.empty() || fwrite(sc.data(), sizeof(float), sc.size(), f) == sc.size());
  ok = (fclose(f) == 0) && ok;
  return ok ? SCGPU_OK : fail(SCGPU_E_IO, "short write to %s", path);
}

int scgpu_load(scgpu_handle* h, const char* path) {
  if (!h || !path) return fail(SCGPU_E_INVALID, "null argument");
  if (h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "load is single-shard");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(SCGPU_E_IO, "cannot open %s", path);
  char magic[8];
  uint64_t hdr[4];
  scgpu_config saved;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SCGPUDB1", 8) != 0 || fread(hdr, 8, 4, f) != 4 || fread(&saved, sizeof saved, 1, f) != 1) {
    fclose(f);
    return fail(SCGPU_E_IO, "%s is not a scgpu database", path);
  }
  if ((int)hdr[0] != h->L.R || (int)hdr[1] != h->L.S) {
    fclose(f);
    return fail(SCGPU_E_INVALID, "database is %llux%llu, handle is %dx%d", (unsigned long long)hdr[0], (unsigned long long)hdr[1], h->L.R, h->L.S);
  }
  std::vector<float> sc((size_t)hdr[2] * h->L.RS);
  const bool ok = sc.empty() || fread(sc.data(), sizeof(float), sc.size(), f) == sc.size();
  fclose(f);
  if (!ok) return fail(SCGPU_E_IO, "short read from %s", path);
  return scgpu_append_descs(h, sc.data(), (size_t)hdr[2]);
}

int scgpu_probe_atanf(const float* x, size_t n, float* out) {
  if ((!x || !out) && n) return fail(SCGPU_E_INVALID, "null argument");
  if (n == 0) return SCGPU_OK;
  float *dx = nullptr, *dy = nullptr;
  CK(cudaMalloc(&dx, n * 4));
  CK(cudaMalloc(&dy, n * 4));
  CK(cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice));
  k_probe_atanf<<<(unsigned)((n + 255) / 256), 256>>>(dx, dy, n);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, dy, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dy);
  if (e != cudaSuccess) return fail(SCGPU_E_CUDA, "probe_atanf: %s", cudaGetErrorString(e));
  return SCGPU_OK;
}

int scgpu_probe_bins(scgpu_handle* h, const float* xyz, size_t n, int32_t* bin, float* height, float* theta) {
  if (!h || ((!xyz || !bin || !height || !theta) && n)) return fail(SCGPU_E_INVALID, "null argument");
  if (n == 0) return SCGPU_OK;
  CK(cudaSetDevice(h->cfg.device));
  float *dx = nullptr, *dh = nullptr, *dt = nullptr;
  int* db = nullptr;
  CK(cudaMalloc(&dx, n * 12));
  CK(cudaMalloc(&dh, n * 4));
  CK(cudaMalloc(&dt, n * 4));
  CK(cudaMalloc(&db, n * 4));
  CK(cudaMemcpy(dx, xyz, n * 12, cudaMemcpyHostToDevice));
  const BinConst bc = make_bin_const(h->L.R, h->L.S, h->cfg.lidar_height, h->cfg.max_radius, !(h->cfg.flags & SCGPU_FLAG_EXACT_BINNING));
  k_probe_bins<<<(unsigned)((n + 255) / 256), 256>>>(dx, n, bc, db, dh, dt);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(bin, db, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(height, dh, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(theta, dt, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dh);
  cudaFree(dt);
  cudaFree(db);
  if (e != cudaSuccess) return fail(SCGPU_E_CUDA, "probe_bins: %s", cudaGetErrorString(e));
  return SCGPU_OK;
}

int scgpu_probe_selfcheck(scgpu_handle* h, uint64_t n, uint64_t seed, int mode, uint64_t* mismatches, uint64_t* fallbacks, float* first_bad) {
  if (!h || !mismatches || !fallbacks) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  unsigned long long* d = nullptr;
  float* db = nullptr;
  CK(cudaMalloc(&d, 16));
  CK(cudaMalloc(&db, 5 * sizeof(float)));
  CK(cudaMemset(d, 0, 16));
  CK(cudaMemset(db, 0, 5 * sizeof(float)));
  const BinConst bc = make_bin_const(h->L.R, h->L.S, h->cfg.lidar_height, h->cfg.max_radius, 1);
  cudaError_t e = cudaSuccess;
  const uint64_t slab = 1ull << 28;
  for (uint64_t s0 = 0; s0 < n && e == cudaSuccess; s0 += slab) {
    const uint64_t m = n - s0 < slab ? n - s0 : slab;
    k_selfcheck<<<(unsigned)((m + 255) / 256), 256>>>(m, seed + s0, mode, bc, d, d + 1, db);
    h->launches++;
    e = cudaGetLastError();
  }
  unsigned long long out[2] = {0, 0};
  float bad[5] = {0, 0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(bad, db, sizeof bad, cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaFree(db);
  if (e != cudaSuccess) return fail(SCGPU_E_CUDA, "selfcheck: %s", cudaGetErrorString(e));
  *mismatches = out[0];
  *fallbacks = out[1];
  if (first_bad) memcpy(first_bad, bad, sizeof bad);
  return SCGPU_OK;
}

int scgpu_xy2theta(float x, float y, float* out_deg) {
  if (!out_deg) return fail(SCGPU_E_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SCGPU_E_NODEVICE, "no CUDA device: libscgpu has no CPU path");
  }
  float xyz[3] = {x, y, 0.f}, *dx = nullptr, *dh = nullptr, *dt = nullptr;
  int* db = nullptr;
  CK(cudaMalloc(&dx, 12));
  CK(cudaMalloc(&dh, 4));
  CK(cudaMalloc(&dt, 4));
  CK(cudaMalloc(&db, 4));
  CK(cudaMemcpy(dx, xyz, 12, cudaMemcpyHostToDevice));
  const BinConst bc = make_bin_const(20, 60, 2.0, 80.0, 0);
  k_probe_bins<<<1, 32>>>(dx, 1, bc, db, dh, dt);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_deg, dt, 4, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dh);
  cudaFree(dt);
  cudaFree(db);
  if (e != cudaSuccess) return fail(SCGPU_E_CUDA, "xy2theta: %s", cudaGetErrorString(e));
  return SCGPU_OK;
}

// ---- staged API -------------------------------------------------------------------------------------

// the caller's stream exactly as given: NULL is the CUDA legacy default stream (what torch uses unless told otherwise),
// NOT the handle's private stream -- the staged calls must be ordered with the caller's own work
#define ST(s) (static_cast<cudaStream_t>(s))

int scgpu_stage_build(scgpu_handle* h, const void* d_pts, size_t n_scans, size_t pts_per_scan, size_t stride, void* d_records, void* stream) {
  if (!h || !d_records) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_build(h, d_pts, n_scans, pts_per_scan, stride, d_records, ST(stream));
}

int scgpu_set_downsample_leaf(scgpu_handle* h, float leaf) {
  if (!h) return fail(SCGPU_E_INVALID, "null handle");
  if (!(leaf >= 0.f) || !(leaf < 1e30f)) return fail(SCGPU_E_INVALID, "leaf size must be >= 0 (0 = no downsampling) and finite");
  if (leaf > 0.f && vox_smem_bytes(h->L.RS) > 227 * 1024) return fail(SCGPU_E_INVALID, "descriptor too large for the voxel kernel's shared memory");
  h->voxel_leaf = leaf;
  return SCGPU_OK;
}

int scgpu_voxel_downsample(scgpu_handle* h, const void* pts, size_t n, size_t stride_bytes, float leaf, float* out_xyzn, uint32_t* out_idx,
                           size_t cap, size_t* out_n, int32_t* min_b, int32_t* div_b, int32_t* status) {
  if (!h || !out_n || (n && !pts)) return fail(SCGPU_E_INVALID, "null argument");
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(SCGPU_E_INVALID, "stride must be >= 12 and a multiple of 4");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  const unsigned ocap = (unsigned)(n ? n : 1);  // a voxel per point at most
  RET(h->vox_in.reserve(n * stride_bytes + 16));
  RET(h->vox_info.reserve(sizeof(VoxInfo)));
  RET(h->vox_pts.reserve((size_t)ocap * sizeof(float4)));
  RET(h->vox_idx.reserve((size_t)ocap * sizeof(unsigned)));
  CK(cudaMemsetAsync(h->vox_info.p, 0, sizeof(VoxInfo), st));
  VoxInfo vi;
  memset(&vi, 0, sizeof vi);
  if (n) {
    CK(cudaMemcpyAsync(h->vox_in.p, pts, n * stride_bytes, cudaMemcpyHostToDevice, st));
    RET(launch_build_voxel(h, h->vox_in.p, 1, n, stride_bytes, leaf, nullptr, h->vox_info.as<VoxInfo>(), h->vox_pts.as<float4>(),
                           h->vox_idx.as<unsigned>(), ocap, st));
    CK(cudaMemcpyAsync(&vi, h->vox_info.p, sizeof vi, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  *out_n = vi.n_out;
  if (min_b) memcpy(min_b, vi.min_b, sizeof vi.min_b);
  if (div_b) memcpy(div_b, vi.div_b, sizeof vi.div_b);
  if (status) *status = vi.status | (vi.passes << 8);
  if (vi.n_out > cap) return fail(SCGPU_E_INVALID, "output capacity %zu < %u voxels", cap, vi.n_out);
  if (vi.n_out) {
    if (out_xyzn) CK(cudaMemcpyAsync(out_xyzn, h->vox_pts.p, (size_t)vi.n_out * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (out_idx) CK(cudaMemcpyAsync(out_idx, h->vox_idx.p, (size_t)vi.n_out * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return SCGPU_OK;
}

int scgpu_stage_append(scgpu_handle* h, const void* d_records, uint64_t first_global, uint64_t global_step, size_t n, void* stream) {
  if (!h || !d_records) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_append(h, d_records, first_global, global_step, n, ST(stream));
}

int scgpu_stage_set_size(scgpu_handle* h, uint64_t n_global) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (local_count(h, n_global) > h->db.cap) return fail(SCGPU_E_INVALID, "size beyond this shard's stored entries");
  h->n_global = n_global;
  return SCGPU_OK;
}

int scgpu_stage_topk(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_ns, uint64_t* d_keys_out, void* stream) {
  if (!h || !d_qrec || !d_ns || !d_keys_out) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_topk(h, d_qrec, nq, d_ns, d_keys_out, ST(stream));
}

int scgpu_stage_merge(scgpu_handle* h, const uint64_t* d_parts, int parts, size_t nq, uint64_t* d_out, void* stream) {
  if (!h || !d_parts || !d_out || parts < 1) return fail(SCGPU_E_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_merge(h, d_parts, parts, nq, d_out, ST(stream));
}

int scgpu_stage_score(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_keys, const uint64_t* d_ns, void* d_best_out,
                      void* stream) {
  if (!h || !d_qrec || !d_keys || !d_ns || !d_best_out) return fail(SCGPU_E_INVALID, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = ST(stream);
  RET(query_reserve(h, nq, 1, st));
  if (h->exh && !(h->cfg.flags & SCGPU_FLAG_NO_SCREENING)) RET(launch_score_screened(h, d_qrec, nq, d_keys, d_ns, st));
  else RET(launch_score(h, d_qrec, nq, d_keys, d_ns, h->K, h->pair_dist.as<double>(), h->pair_shift.as<int>(), 0, st));
  return launch_best(h, nq, d_keys, static_cast<Best*>(d_best_out), st);
}

int scgpu_stage_gather(scgpu_handle* h, uint64_t global_idx, void* d_record, void* stream) {
  if (!h || !d_record) return fail(SCGPU_E_INVALID, "null argument");
  if (global_idx >= h->n_global) return fail(SCGPU_E_INVALID, "entry out of range");
  if ((int)(global_idx % (uint64_t)h->cfg.shard_count) != h->cfg.shard_rank) return fail(SCGPU_E_INVALID, "entry lives on another shard");
  CK(cudaSetDevice(h->cfg.device));
  k_gather<<<1, 128, 0, ST(stream)>>>(static_cast<unsigned char*>(d_record), h->L, h->db, global_idx / (uint64_t)h->cfg.shard_count);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int scgpu_stage_exhaustive(scgpu_handle* h, const void* d_qrecords, size_t nq, const uint64_t* n_search, void* d_best_out, void* stream) {
  if (!h || !d_qrecords || !d_best_out || (!n_search && nq)) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->exh) return fail(SCGPU_E_INVALID, "screening kernels are instantiated for 20x60 (radius 3) and 40x120 (radius 6) only");
  CK(cudaSetDevice(h->cfg.device));
  for (size_t i0 = 0; i0 < nq; i0 += EXH_MAX_BATCH) {
    const size_t m = nq - i0 < EXH_MAX_BATCH ? nq - i0 : EXH_MAX_BATCH;
    RET(launch_exhaustive_fast(h, static_cast<const unsigned char*>(d_qrecords) + i0 * h->L.rec_bytes, m, n_search + i0,
                               static_cast<Best*>(d_best_out) + i0, ST(stream), nullptr, nullptr));
  }
  return SCGPU_OK;
}

int scgpu_stage_finalize(scgpu_handle* h, const void* d_best_parts, int parts, size_t nq, const uint64_t* d_ns, int32_t* d_loop_id,
                         float* d_yaw, double* d_nearest_dist, int32_t* d_nearest_idx, int32_t* d_nearest_shift, void* stream) {
  if (!h || !d_best_parts || !d_ns || !d_loop_id || !d_yaw || parts < 1) return fail(SCGPU_E_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_finalize(h, static_cast<const Best*>(d_best_parts), parts, nq, d_ns, d_loop_id, d_yaw, d_nearest_dist, d_nearest_idx,
                         d_nearest_shift, ST(stream));
}

}  // extern "C"
