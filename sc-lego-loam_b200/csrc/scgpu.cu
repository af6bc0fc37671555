// scgpu.cu -- host side of libscgpu.so: the C ABI of include/scgpu.h over the sm_100a kernels in
// scgpu_kernels.cuh.  No CPU implementation of any stage lives here: every entry point either launches the
// kernels or fails.
#include "scgpu.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <float.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "scgpu_exhaustive.cuh"
#include "scgpu_icp.cuh"
#include "scgpu_kernels.cuh"
#include "scgpu_tc.cuh"
#include "scgpu_voxel.cuh"

using namespace scgpu;

namespace {

// NVTX range around a public call (tracing).  Header-only nvtx3: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define SCGPU_TRACE_CALL() NvtxRange nvtx_range_(__func__)

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

// A device array that GROWS IN PLACE: a virtual address range reserved once (cuMemAddressReserve), physical memory mapped
// behind it chunk by chunk (cuMemCreate / cuMemMap / cuMemSetAccess).  The base pointer never changes, so growing a shard
// needs no device synchronisation, no copy and no pointer patching in work that is already enqueued (round 1 grew by
// cudaDeviceSynchronize + cudaMalloc + copy + cudaFree).  The driver entry points come through cudaGetDriverEntryPoint
// (no -lcuda); if they are missing, or the device cannot do virtual memory management, ok() is false and the caller
// keeps the copy path.
struct VmApi {
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*afree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*gran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
  static const VmApi& get() {
    static VmApi api = [] {
      VmApi a;
      if (getenv("SCGPU_NO_VMM") && atoi(getenv("SCGPU_NO_VMM")) != 0) return a;
      auto sym = [](const char* name) -> void* {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return fn;
      };
      a.reserve = reinterpret_cast<decltype(a.reserve)>(sym("cuMemAddressReserve"));
      a.afree = reinterpret_cast<decltype(a.afree)>(sym("cuMemAddressFree"));
      a.create = reinterpret_cast<decltype(a.create)>(sym("cuMemCreate"));
      a.release = reinterpret_cast<decltype(a.release)>(sym("cuMemRelease"));
      a.map = reinterpret_cast<decltype(a.map)>(sym("cuMemMap"));
      a.unmap = reinterpret_cast<decltype(a.unmap)>(sym("cuMemUnmap"));
      a.access = reinterpret_cast<decltype(a.access)>(sym("cuMemSetAccess"));
      a.gran = reinterpret_cast<decltype(a.gran)>(sym("cuMemGetAllocationGranularity"));
      a.ok = a.reserve && a.afree && a.create && a.release && a.map && a.unmap && a.access && a.gran;
      cudaGetLastError();
      return a;
    }();
    return api;
  }
};

struct VmBuf {
  CUdeviceptr va = 0;
  size_t va_bytes = 0, mapped = 0, gran = 0;
  int device = 0;
  std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;
  CUmemAllocationProp prop() const {
    CUmemAllocationProp pr = {};
    pr.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    pr.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    pr.location.id = device;
    return pr;
  }
  // reserve the address range (once); false = virtual memory management is not usable here
  bool init(int dev, size_t max_bytes) {
    const VmApi& a = VmApi::get();
    if (!a.ok) return false;
    device = dev;
    const CUmemAllocationProp pr = prop();
    if (a.gran(&gran, &pr, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
    va_bytes = (max_bytes + gran - 1) / gran * gran;
    if (a.reserve(&va, va_bytes, 0, 0, 0) != CUDA_SUCCESS) {
      va = 0;
      return false;
    }
    return true;
  }
  // make [0, want) usable; what was there stays where it is.  SCGPU_OK, or E_CUDA (out of memory / range exhausted)
  int grow(size_t want) {
    if (want <= mapped) return SCGPU_OK;
    const VmApi& a = VmApi::get();
    size_t add = (want - mapped + gran - 1) / gran * gran;
    if (mapped + add > va_bytes) return fail(SCGPU_E_CUDA, "database larger than the reserved address range (%zu bytes)", va_bytes);
    const CUmemAllocationProp pr = prop();
    CUmemGenericAllocationHandle hd;
    CUresult r = a.create(&hd, add, &pr, 0);
    if (r != CUDA_SUCCESS) return fail(SCGPU_E_CUDA, "cuMemCreate(%zu bytes) failed (%d)", add, (int)r);
    r = a.map(va + mapped, add, 0, hd, 0);
    if (r != CUDA_SUCCESS) {
      a.release(hd);
      return fail(SCGPU_E_CUDA, "cuMemMap failed (%d)", (int)r);
    }
    CUmemAccessDesc ad = {};
    ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ad.location.id = device;
    ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = a.access(va + mapped, add, &ad, 1);
    if (r != CUDA_SUCCESS) {
      a.unmap(va + mapped, add);
      a.release(hd);
      return fail(SCGPU_E_CUDA, "cuMemSetAccess failed (%d)", (int)r);
    }
    chunks.emplace_back(hd, add);
    mapped += add;
    return SCGPU_OK;
  }
  void release() {
    if (!va) return;
    const VmApi& a = VmApi::get();
    size_t off = 0;
    for (auto& c : chunks) {
      a.unmap(va + off, c.second);
      a.release(c.first);
      off += c.second;
    }
    chunks.clear();
    a.afree(va, va_bytes);
    va = 0;
    va_bytes = mapped = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(va); }
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
  int exh_cfg = 0;  // 1: 20x60 radius 3, 2: 40x120 radius 6, 3: 20x60 full-shift search (tensor-core screening, scgpu_tc.cuh)
  // full-shift search on the tensor cores: hi / lo split of the screening copy, the queries' circulant expansion, TMA maps
  DevBuf tc_e_hi, tc_e_lo, tc_q_hi, tc_q_lo, tc_qaux, tc_shift;
  DevBuf tc_ev_hi, tc_ev_lo, tc_qv_hi, tc_qv_lo, tc_align;  // sector keys (windowed search on the tensor cores: alignment GEMM)
  DevBuf icp_src, icp_tgt, icp_state, icp_part;  // loop verification (scgpu_icp.cuh)
  uint64_t tc_rows = 0;    // rows the E buffers (and their tensor maps) were sized for
  uint64_t tc_upto = 0;    // local entries [0, tc_upto) are split
  CUtensorMap tc_maps[12];  // E_hi, E_lo, Q_hi, Q_lo, EV_hi, EV_lo, QV_hi, QV_lo | the first four again with 64-byte rows (k_tc_fullshift2)
  bool tc_want_shifts = false;
  VmBuf vm[6];            // sc, ring keys, sector keys, column norms, screening copy, screening sector-key records: grown in place
  int vm_state = 0;       // 0 = not tried yet, 1 = arrays live in vm[], -1 = virtual memory management unavailable: cudaMalloc + copy
  unsigned n_grow_inplace = 0, n_grow_copy = 0;
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
  // "leave room" mode of the binning kernel: its blocks ask for a little more shared memory than they need, so that THREE fit an
  // SM instead of four and a quarter of every SM (16K registers, ~47 KB of shared memory) stays free for the latency-bound query
  // kernels of the previous step to run underneath (asynchronous replays).
  bool build_room = false;
  // SCGPU_TRACE=1: per-step events of the replay engine (build start / end, query start / end), dumped by scgpu_timer_stop
  cudaEvent_t tr_ev[16][4] = {};
  int tr_n = 0;
  bool trace = false;
  uint64_t mutation = 0, last_mutation = 0;  // candidate dumps are valid only while nothing has changed since the pipeline ran
  DevBuf records2[2];      // record buffers of asynchronous replays (alternating: the next build overlaps this query stage)
  int rec_turn = 0;
  DevBuf res_buf;          // result block of replays on a handle without a slab
  ResultBlock res_rb{};
  size_t replay_nq = 0;    // results held by the last asynchronous replay
  bool replay_ipc = false;
  uint64_t n_written = 0;  // local slots [0, n_written) have been stored at some point (scgpu_stage_set_size validates against it)
  // ---- peer-sharded mode (SCGPU_FLAG_PEER, or a shard of a device-list handle) ---------------------------------------
  bool peer = false;      // one fixed-capacity slab holds every array the other shards read or write
  bool attached = false;  // the other shards' slabs are mapped
  void* slab = nullptr;
  size_t slab_bytes = 0;
  struct SlabOff {
    size_t sc, sector, colnorm, hat, vk, ring, flags, results, total;
  } so{};
  PeerTab peers{};                          // every shard's arrays as seen from this device
  float* ring_of[MAX_SHARDS] = {};          // every shard's ring-key replica
  unsigned* flags_of[MAX_SHARDS] = {};      // every shard's barrier cells
  void* results_of[MAX_SHARDS] = {};        // every shard's result block
  void* ipc_base[MAX_SHARDS] = {};          // slabs opened with cudaIpcOpenMemHandle (closed in scgpu_destroy)
  ResultBlock rb{};
  unsigned epoch[2] = {0, 0};               // barrier generations: channel 0 = "appended", 1 = "queries done"
  cudaStream_t qstream = nullptr;           // query stage of chunk c runs here while chunk c+1 is binned on `stream`
  cudaEvent_t ev_app = nullptr, ev_qdone = nullptr, ev_side = nullptr;
  cudaEvent_t ev_b0 = nullptr, ev_b1 = nullptr;  // scgpu_timer_start / stop
  PinBuf h_ns_ring[4];                      // n_search staging of asynchronous replays (rotating, guarded by events)
  cudaEvent_t ev_ns_ring[4] = {nullptr, nullptr, nullptr, nullptr};
  int ns_turn = 0;
  DevBuf q_idx;
  // ---- device-list handle (cfg.n_devices > 1): this object only routes to its shards ------------------------------------
  std::vector<scgpu_handle*> shards;
  bool is_group = false;
  scgpu_handle* parent = nullptr;
};

extern "C" {
static int create_one(const scgpu_config* cfg, scgpu_handle** out);
}

namespace {

// a public call on a device-list handle that only needs "a device": use the first shard
#define GROUP_FIRST(h) ((h)->is_group ? (h)->shards[0] : (h))

uint64_t local_count(const scgpu_handle* h, uint64_t n_global) {
  const uint64_t G = (uint64_t)h->cfg.shard_count, r = (uint64_t)h->cfg.shard_rank;
  return n_global > r ? (n_global - 1 - r) / G + 1 : 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr size_t BUILD_ROOM_SMEM = 57 * 1024 + 512;  // > 227 KB / 4: at most three binning blocks per SM
constexpr uint64_t PEER_RESULT_CAP = 65536;  // queries per replay batch whose results a shard's slab can hold

// Peer-sharded mode: ONE allocation per shard (so that one cudaIpcMemHandle maps everything the other shards touch) laid out
// identically on every shard: sc | sector | colnorm | sc_hat | vk | ring-key replica [R][cap * G] | barrier cells | results.
int slab_alloc(scgpu_handle* h, uint64_t cap) {
  const Layout& L = h->L;
  const uint64_t G = (uint64_t)h->cfg.shard_count;
  scgpu_handle::SlabOff o{};
  size_t at = 0;
  auto take = [&](size_t bytes) {
    const size_t r = at;
    at = align_up(at + bytes, 256);
    return r;
  };
  o.sc = take(cap * L.RS * sizeof(float));
  o.sector = take(cap * L.S * sizeof(double));
  o.colnorm = take(cap * L.S * sizeof(double));
  o.hat = take(h->exh ? cap * L.RS * sizeof(float) : 0);
  o.vk = take(h->exh ? (cap + 8) * exh_vk_bytes(L.S) : 0);
  o.ring = take((size_t)L.R * cap * G * sizeof(float));
  o.flags = take(1024);
  h->rb.cap_q = PEER_RESULT_CAP;
  o.results = take(h->rb.bytes());
  o.total = at;
  CK(cudaMalloc(&h->slab, o.total));
  CK(cudaMemsetAsync(h->slab, 0, o.total, h->stream));  // barrier cells start at generation 0
  CK(cudaStreamSynchronize(h->stream));
  h->slab_bytes = o.total;
  h->so = o;
  unsigned char* b = static_cast<unsigned char*>(h->slab);
  h->db.sc = reinterpret_cast<float*>(b + o.sc);
  h->db.sector = reinterpret_cast<double*>(b + o.sector);
  h->db.colnorm = reinterpret_cast<double*>(b + o.colnorm);
  h->db.ringT = reinterpret_cast<float*>(b + o.ring);
  h->db.cap = cap;
  h->db.ring_cap = cap * G;
  h->db.ring_global = 1;
  h->x_sc_hat = h->exh ? reinterpret_cast<float*>(b + o.hat) : nullptr;
  h->x_vk = h->exh ? b + o.vk : nullptr;
  return SCGPU_OK;
}

// the arrays of shard s as seen from h's device, given the base of s's slab in h's address space
void peer_fill(scgpu_handle* h, int s, void* base) {
  unsigned char* b = static_cast<unsigned char*>(base);
  h->peers.sc[s] = reinterpret_cast<const float*>(b + h->so.sc);
  h->peers.sector[s] = reinterpret_cast<const double*>(b + h->so.sector);
  h->peers.colnorm[s] = reinterpret_cast<const double*>(b + h->so.colnorm);
  h->peers.sc_hat[s] = reinterpret_cast<const float*>(b + h->so.hat);
  h->peers.vk[s] = b + h->so.vk;
  h->ring_of[s] = reinterpret_cast<float*>(b + h->so.ring);
  h->flags_of[s] = reinterpret_cast<unsigned*>(b + h->so.flags);
  h->results_of[s] = b + h->so.results;
}

int db_reserve(scgpu_handle* h, uint64_t want_local) {
  if (want_local <= h->db.cap) return SCGPU_OK;
  if (h->peer)
    return fail(SCGPU_E_INVALID, "a peer-sharded database has a fixed capacity (capacity_hint = %llu entries): %llu local slots wanted, %llu there",
                (unsigned long long)h->cfg.capacity_hint, (unsigned long long)want_local, (unsigned long long)h->db.cap);
  uint64_t cap = h->db.cap ? h->db.cap * 2 : 1024;
  while (cap < want_local) cap *= 2;
  const Layout& L = h->L;
  // ---- in place: the arrays live behind reserved address ranges; map more physical memory, nothing moves ---------------
  if (h->vm_state == 0) {
    const uint64_t max_cap = (uint64_t)1 << 26;  // address space only: 64 Mi keyframes per shard (20 x 60: 322 GB of descriptors)
    const size_t per[6] = {L.RS * sizeof(float), L.R * sizeof(float), L.S * sizeof(double), L.S * sizeof(double), L.RS * sizeof(float),
                           exh_vk_bytes(L.S)};
    bool ok = h->db.cap == 0;
    for (int a = 0; a < 6 && ok; ++a) ok = h->vm[a].init(h->cfg.device, (max_cap + 64) * per[a]);
    if (!ok)
      for (VmBuf& b : h->vm) b.release();
    h->vm_state = ok ? 1 : -1;
  }
  if (h->vm_state == 1) {
    const size_t want[6] = {cap * L.RS * sizeof(float), cap * L.R * sizeof(float), cap * L.S * sizeof(double), cap * L.S * sizeof(double),
                            h->exh ? cap * L.RS * sizeof(float) : 0, h->exh ? (cap + 8) * exh_vk_bytes(L.S) : 0};
    for (int a = 0; a < 6; ++a) RET(h->vm[a].grow(want[a]));
    h->db.sc = h->vm[0].as<float>();
    h->db.ringT = h->vm[1].as<float>();
    h->db.sector = h->vm[2].as<double>();
    h->db.colnorm = h->vm[3].as<double>();
    h->x_sc_hat = h->exh ? h->vm[4].as<float>() : nullptr;
    h->x_vk = h->exh ? h->vm[5].as<unsigned char>() : nullptr;
    h->db.cap = cap;
    h->db.ring_cap = cap;
    h->db.ring_global = 0;
    h->n_grow_inplace++;
    return SCGPU_OK;
  }
  // ---- fallback: a larger allocation and a copy
  CK(cudaDeviceSynchronize());  // rare: the shard moves to a larger allocation; nothing may still be reading the old one
  h->n_grow_copy++;
  Db nd = h->db;
  nd.cap = cap;
  nd.ring_cap = cap;
  nd.ring_global = 0;
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
    // (the tiled ring-key layout does not depend on the capacity: the stored tiles move as one block)
    CK(cudaMemcpyAsync(nd.ringT, h->db.ringT, ((n + 31) / 32) * 32 * L.R * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
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
                 cudaStream_t st, size_t scan_pitch = 0) {  // scan_pitch: bytes between scans (0 = contiguous)
  if (n_scans == 0) return SCGPU_OK;
  if (scan_pitch == 0) scan_pitch = pts_per_scan * stride;
  if (h->voxel_leaf > 0.f && pts_per_scan > 0) {
    if (h->cfg.flags & SCGPU_FLAG_INTENSITY) return fail(SCGPU_E_INVALID, "the voxel-grid path carries no intensity");
    if (scan_pitch != pts_per_scan * stride) return fail(SCGPU_E_INVALID, "the voxel-grid path takes contiguous scans");
    return launch_build_voxel(h, d_pts, n_scans, pts_per_scan, stride, h->voxel_leaf, d_records, nullptr, nullptr, nullptr, 0, st);
  }
  if (stride < 12 || (stride & 3) || ((uintptr_t)d_pts & 3)) return fail(SCGPU_E_INVALID, "points must be 4-byte aligned, stride >= 12 and a multiple of 4");
  if (pts_per_scan > 0xfffffff0ull || n_scans > 65535ull * 1024) return fail(SCGPU_E_INVALID, "scan too large");
  RET(build_reserve(h, n_scans, st));
  BuildParams p;
  p.pts = static_cast<const unsigned char*>(d_pts);
  p.scan_pitch = (unsigned long long)scan_pitch;
  p.n_pts = (unsigned)pts_per_scan;
  p.stride = (unsigned)stride;
  // intensity descriptor (SCGPU_FLAG_INTENSITY): bin the float at byte 16 of a pcl::PointXYZI record.  Packed 12-byte points
  // come from the host packer, which has put the intensity into the third slot already.
  const bool intensity = (h->cfg.flags & SCGPU_FLAG_INTENSITY) != 0;
  p.val_off = (intensity && stride != 12) ? 16u : 8u;
  if (intensity && stride != 12 && stride < 20) return fail(SCGPU_E_INVALID, "the intensity descriptor needs pcl::PointXYZI-like records (stride >= 20 bytes)");
  // tile: enough blocks for two full waves of the GPU (4 resident blocks per SM), otherwise as large as possible -- a
  // scan binned by ONE block needs no global merge (atomics, fences, ticket) and starts / drains its TMA ring once
  // (4,541 HDL-64 scans: 16k-point tiles 1.71 ms, whole-scan tiles 1.56 ms).  From ~3/4 of one wave of scans on, whole-scan
  // tiles also win over splitting (568 scans per GPU of the 8-GPU run: three tiles per scan 0.72 of roofline).
  unsigned ppb = 256;
  if (pts_per_scan) {
    const uint64_t want = (uint64_t)h->sm_count * 8;
    uint64_t tiles = (want + n_scans - 1) / n_scans;
    if (n_scans >= (uint64_t)h->sm_count * 3) tiles = 1;
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
  p.bc = make_bin_const(h->L.R, h->L.S, intensity ? 0.0 : h->cfg.lidar_height, h->cfg.max_radius, !(h->cfg.flags & SCGPU_FLAG_EXACT_BINNING));
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
    const bool al16 = (((uintptr_t)q.pts & 15) == 0) && (scan_pitch & 15) == 0;
    int sk = (stride == 16 && al16) ? 16 : ((stride == 32 && al16) ? 32 : 0);
    if (p.val_off != 8 && stride == 16) sk = 0;  // (a 20..28-byte or odd layout goes through the generic loads)
    // packed xyz (12 bytes per point, what the host packer ships over PCIe): TMA copies are multiples of 16 bytes, so
    // every tile must hold a multiple of 4 points
    const bool tma12 = stride == 12 && al16 && (pts_per_scan & 3) == 0;
    if ((sk || tma12) && q.bc.fast && !(h->cfg.flags & SCGPU_FLAG_NO_TMA_BUILD)) {
      // TMA-staged variant (tile start offsets are multiples of 16 bytes because pts_per_block * stride is)
      if (tma12) {
        const size_t sm = build_tma_smem<12>(h->L.RS);
        if (q.bc.lh_is_float) k_build_tma<12, true, true><<<grid, 256, sm, st>>>(q);
        else k_build_tma<12, true, false><<<grid, 256, sm, st>>>(q);
      } else {
        size_t sm = sk == 16 ? build_tma_smem<16>(h->L.RS) : build_tma_smem<32>(h->L.RS);
        if (sk == 16 && h->build_room && sm < BUILD_ROOM_SMEM) sm = BUILD_ROOM_SMEM;  // three blocks per SM instead of four: see build_room
        if (sk == 16 && q.bc.lh_is_float) k_build_tma<16, true, true><<<grid, 256, sm, st>>>(q);
        else if (sk == 16) k_build_tma<16, true, false><<<grid, 256, sm, st>>>(q);
        else if (q.bc.lh_is_float) k_build_tma<32, true, true><<<grid, 256, sm, st>>>(q);
        else k_build_tma<32, true, false><<<grid, 256, sm, st>>>(q);
      }
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

int exh_sync(scgpu_handle* h, cudaStream_t st);

// push: (peer-sharded) also store the records' ring keys into the other shards' replicas -- peer stores over NVLink
int launch_append(scgpu_handle* h, const void* d_records, uint64_t first_global, uint64_t step, size_t n, cudaStream_t st, bool push = false) {
  if (n == 0) return SCGPU_OK;
  if (step < 1) return fail(SCGPU_E_INVALID, "global_step must be >= 1");
  const uint64_t new_size = first_global + (n - 1) * step + 1;
  if (new_size > 0xffffffffull) return fail(SCGPU_E_INVALID, "database index space is 32 bits");
  RET(db_reserve(h, local_count(h, new_size)));
  if (h->peer && new_size > h->db.ring_cap) return fail(SCGPU_E_INVALID, "peer-sharded database is full (%llu entries)", (unsigned long long)h->db.ring_cap);
  PushList pl{};
  if (push) {
    if (!h->attached) return fail(SCGPU_E_INVALID, "peer-sharded handle is not attached to its peers (scgpu_peer_attach)");
    for (int s = 0; s < h->cfg.shard_count; ++s)
      if (s != h->cfg.shard_rank) pl.dst[pl.n++] = h->ring_of[s];
  }
  // entries below the current size are being overwritten: their screening copies are stale
  const uint64_t first_local = local_count(h, first_global);  // this shard's entries below first_global = the first slot written
  if (first_global < h->n_global && h->x_upto > first_local) h->x_upto = first_local;
  // with the screening kernels in use, the screening copy of the new entries is written by the same launch -- provided the
  // entries before them already have theirs (otherwise exh_sync catches up lazily, before the first search that needs it)
  const bool fused_hat = h->exh && (step == 1 || step == (uint64_t)h->cfg.shard_count) && h->x_upto >= first_local;
  if (fused_hat)
    k_append_hat<<<(unsigned)n, 128, 0, st>>>(static_cast<const unsigned char*>(d_records), h->L, h->db, first_global, step, pl, h->x_sc_hat, h->x_vk);
  else
    k_append<<<(unsigned)n, 128, 0, st>>>(static_cast<const unsigned char*>(d_records), h->L, h->db, first_global, step, pl);
  h->launches++;
  CK(cudaGetLastError());
  if (new_size > h->n_global) h->n_global = new_size;
  const uint64_t have = local_count(h, h->n_global);
  if (have > h->n_written) h->n_written = have;
  if (fused_hat && have > h->x_upto) h->x_upto = have;
  if (h->peer && h->exh) RET(exh_sync(h, st));  // the other shards fetch screening rows from here: keep them current
  return SCGPU_OK;
}

// Launch shape of k_topk: `warps` (2, 4 or 8) per block share one (query, chunk).  Every warp keeps its own sorted list, which
// costs ~K(1 + ln(n_warp / K)) warm-up insertions plus a merge per warp -- so a warp's stream should be long (>= ~1k keys) --
// while the launch as a whole wants ~16 warps per SM to cover the L2 latency of the key loads.  (A 568-query step over 4,541
// keys: 2 chunks x 8 warps = 288 keys per warp took 59 us; 1 chunk x 4 warps = 1,135 keys per warp is the shape chosen now.)
void choose_topk_shape(uint64_t n_local, size_t nq, int sm_count, unsigned& chunk, unsigned& chunks, int& warps) {
  if (n_local == 0 || nq == 0) {
    chunk = 256;
    chunks = 1;
    warps = 2;
    return;
  }
  uint64_t wq = ((uint64_t)sm_count * 16 + nq - 1) / nq;  // warps per query the grid would like
  const uint64_t by_stream = n_local / 1024;              // ... and what keeps every stream >= 1,024 keys
  if (wq > by_stream) wq = by_stream;
  if (wq < 2) wq = 2;
  warps = wq <= 2 ? 2 : (wq <= 4 ? 4 : 8);
  uint64_t c = (wq + warps - 1) / warps;
  if (c < 1) c = 1;
  uint64_t ch = (n_local + c - 1) / c;
  ch = (ch + 255) / 256 * 256;
  chunk = (unsigned)ch;
  chunks = (unsigned)((n_local + ch - 1) / ch);
}

// k_topk_tile's shape (groups of TOPK_QT queries per block, TOPK_TILE_WARPS warps): as many chunks as fill the GPU
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
  // peer-sharded: retrieval runs over this shard's replica of ALL ring keys (global indices): one device, no merge
  const uint64_t n_local = h->peer ? h->n_global : local_count(h, h->n_global);
  // large batches over a large shard: groups of TOPK_QT queries share every key they load (k_topk_tile) -- only when each
  // warp's stream stays long (>= 8k keys), see the kernel's header.  SCGPU_TOPK_TILE=1/0 forces / forbids it (tests).
  bool tile = nq >= 16 && h->slots <= 2 && (h->L.R == 20 || h->L.R == 40);
  const size_t groups = (nq + TOPK_QT - 1) / TOPK_QT;
  if (tile) {
    choose_chunks(n_local, groups, chunk, chunks);
    static const char* force = getenv("SCGPU_TOPK_TILE");
    tile = force ? atoi(force) != 0 : (uint64_t)chunk / TOPK_TILE_WARPS >= 8192;
  }
  int warps = 8;
  if (tile) choose_chunks(n_local, groups, chunk, chunks);
  else choose_topk_shape(n_local, nq, h->sm_count, chunk, chunks, warps);
  const size_t units = tile ? groups : nq;
  RET(query_reserve(h, nq, chunks, st));
  TopkParams p;
  p.qrecords = static_cast<const unsigned char*>(d_qrec);
  p.L = h->L;
  p.db = h->db;
  if (h->peer) {
    p.db.cap = h->db.ring_cap;
    p.db.rank = 0;
    p.db.G = 1;
  }
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
  // batches of hundreds of queries over a few thousand keys with K <= 32: one warp per query, keys staged once per block in
  // shared memory (k_topk_qs).  SCGPU_TOPK_QS=1/0 forces / forbids it (tests, A/B).
  {
    static const int qs_env = getenv("SCGPU_TOPK_QS") ? atoi(getenv("SCGPU_TOPK_QS")) : -1;
    const bool can = h->slots == 1 && (h->L.R == 20 || h->L.R == 40);
    const int qpb = nq >= (size_t)h->sm_count * 4 ? 8 : 4;
    // measured (async replay steps of 568 / 1,135 / 2,270 / 4,541 queries over 4,541 keys): -12 % / -2 % / +3 % / +2 % on the step --
    // its 42 KB of shared memory keep it from running under the next step's binning, which k_topk does
    const bool want = qs_env >= 0 ? qs_env != 0 : (nq >= 2048 && n_local <= 65536);
    if (can && want) {
      const unsigned blocks = (unsigned)((nq + qpb - 1) / qpb);
      if (h->L.R == 20 && qpb == 8) k_topk_qs<20, 8><<<blocks, 8 * 32, 0, st>>>(p, (unsigned)nq);
      else if (h->L.R == 20) k_topk_qs<20, 4><<<blocks, 4 * 32, 0, st>>>(p, (unsigned)nq);
      else if (qpb == 8) k_topk_qs<40, 8><<<blocks, 8 * 32, 0, st>>>(p, (unsigned)nq);
      else k_topk_qs<40, 4><<<blocks, 4 * 32, 0, st>>>(p, (unsigned)nq);
      h->launches++;
      CK(cudaGetLastError());
      return SCGPU_OK;
    }
  }
  dim3 grid(chunks, (unsigned)nq);
  // software-pipelined key loads for small grids (few resident warps per SM: the L2 latency of the loads is exposed)
  static const int pipe_env = getenv("SCGPU_TOPK_PIPE") ? atoi(getenv("SCGPU_TOPK_PIPE")) : -1;
  // (K <= 32 only: with two or four list slots per lane the pipelined form needs 117-240 registers and loses -- BASELINE config 3,
  // K = 50 over 40k keys, 64 queries per rank: 545k queries/s without, 439k with, on 2 GPUs)
  const bool pipe = pipe_env >= 0 ? pipe_env != 0 : (h->slots == 1 && (uint64_t)chunks * nq * warps < (uint64_t)h->sm_count * 32);
  const int rc = h->L.R == 20 ? 20 : (h->L.R == 40 ? 40 : 0);
#define SCGPU_TOPK_LAUNCH(SL, W, RC_, PIPE_) k_topk<SL, W, RC_, PIPE_><<<grid, (W) * 32, 0, st>>>(p)
#define SCGPU_TOPK_RC(SL, W)                                              \
  if (rc == 20 && pipe) SCGPU_TOPK_LAUNCH(SL, W, 20, true);               \
  else if (rc == 20) SCGPU_TOPK_LAUNCH(SL, W, 20, false);                 \
  else if (rc == 40 && pipe) SCGPU_TOPK_LAUNCH(SL, W, 40, true);          \
  else if (rc == 40) SCGPU_TOPK_LAUNCH(SL, W, 40, false);                 \
  else SCGPU_TOPK_LAUNCH(SL, W, 0, false);
#define SCGPU_TOPK_CASE(SL)          \
  if (warps == 2) {                  \
    SCGPU_TOPK_RC(SL, 2)             \
  } else if (warps == 4) {           \
    SCGPU_TOPK_RC(SL, 4)             \
  } else {                           \
    SCGPU_TOPK_RC(SL, 8)             \
  }
  if (h->slots == 1) {
    SCGPU_TOPK_CASE(1)
  } else if (h->slots == 2) {
    SCGPU_TOPK_CASE(2)
  } else {
    SCGPU_TOPK_CASE(4)
  }
#undef SCGPU_TOPK_RC
#undef SCGPU_TOPK_LAUNCH
#undef SCGPU_TOPK_CASE
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
  p.peers = h->peers;
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
// The rescoring list's counter is armed (zero) on entry: zeroed at allocation, re-armed by k_best / k_best_finalize.
int launch_score_screened(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_keys, const uint64_t* d_ns, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  RET(exh_sync(h, st));
  RET(h->c_d32.reserve(nq * h->K * sizeof(float)));
  RET(h->c_list.reserve(nq * h->K * sizeof(uint64_t)));
  if (!h->c_count.p) {
    RET(h->c_count.reserve(16));
    CK(cudaMemsetAsync(h->c_count.p, 0, 16, st));
  }
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
  cp.peers = h->peers;
  cp.list = h->c_list.as<unsigned long long>();
  cp.count = h->c_count.as<unsigned>();
  cp.pair_dist = h->pair_dist.as<double>();
  cp.pair_shift = h->pair_shift.as<int>();
  // warps per block, staging slots per warp.  One slot: three blocks fit an SM and cover each other's fetch latency
  // (two slots = one block per SM measured slower at K = 50).
  static const int cw_env = getenv("SCGPU_CAND_CW") ? atoi(getenv("SCGPU_CAND_CW")) : 0;
  const bool cw5 = cw_env ? cw_env == 5 : (h->build_room || h->K <= 10);
  if (h->exh_cfg == 1 && cw5) k_cand_screen<20, 60, 3, 1, 5, 1><<<(unsigned)nq, 5 * 32, cand_smem_bytes<20, 60, 3, 5, 1>(), st>>>(cp);
  else if (h->exh_cfg == 1) k_cand_screen<20, 60, 3, 1, 10, 1><<<(unsigned)nq, 10 * 32, cand_smem_bytes<20, 60, 3, 10, 1>(), st>>>(cp);
  else k_cand_screen<40, 120, 6, 2, 5, 1><<<(unsigned)nq, 5 * 32, cand_smem_bytes<40, 120, 6, 5, 1>(), st>>>(cp);
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
  p.peers = h->peers;
  k_score_pairs<<<(unsigned)h->sm_count * 8, 128, pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float)), st>>>(p, h->c_list.as<unsigned long long>(),
                                                                                                            h->c_count.as<unsigned>());
  h->launches += 2;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int launch_best(scgpu_handle* h, size_t nq, const uint64_t* d_keys, Best* d_best, cudaStream_t st) {
  if (nq == 0) return SCGPU_OK;
  k_best<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(h->pair_dist.as<double>(), h->pair_shift.as<int>(),
                                                       reinterpret_cast<const unsigned long long*>(d_keys), (unsigned)nq, h->K, d_best,
                                                       h->c_count.as<unsigned>());
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int launch_finalize(scgpu_handle* h, const Best* d_parts, int parts, size_t nq, const uint64_t* d_ns, int* d_loop, float* d_yaw,
                    double* d_dist, int* d_idx, int* d_shift, cudaStream_t st, unsigned out_off = 0, unsigned out_step = 1,
                    const PushList* push = nullptr) {
  if (nq == 0) return SCGPU_OK;
  PushList pl{};
  if (push) pl = *push;
  k_finalize<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(d_parts, parts, (unsigned)nq, reinterpret_cast<const unsigned long long*>(d_ns),
                                                           h->K, h->L.S, h->cfg.dist_thres, d_loop, d_yaw, d_dist, d_idx, d_shift, out_off,
                                                           out_step, h->rb, pl);
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
// Query stage (stages 3 + 4 + decision) for nq query records on the device, enqueued on st.  Result q goes to slot
// out_off + q * out_step of the given arrays (and of the result blocks in `push`).
int query_stage(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* d_ns, cudaStream_t st, int* d_loop, float* d_yaw, double* d_dist,
                int* d_idx, int* d_shift, unsigned out_off, unsigned out_step, const PushList* push) {
  if (nq == 0) return SCGPU_OK;
  if (h->peer && !h->attached) return fail(SCGPU_E_INVALID, "peer-sharded handle is not attached to its peers (scgpu_peer_attach)");
  unsigned chunk, chunks;
  choose_chunks(h->peer ? h->n_global : local_count(h, h->n_global), nq, chunk, chunks);
  RET(query_reserve(h, nq, chunks, st));
  RET(launch_topk(h, d_qrec, nq, d_ns, h->keys.as<uint64_t>(), st));
  h->last_screened = h->exh && h->exh_cfg != 3 && !(h->cfg.flags & SCGPU_FLAG_NO_SCREENING);
  h->last_qrec = d_qrec;
  if (h->last_screened) RET(launch_score_screened(h, d_qrec, nq, h->keys.as<uint64_t>(), d_ns, st));
  else RET(launch_score(h, d_qrec, nq, h->keys.as<uint64_t>(), d_ns, h->K, h->pair_dist.as<double>(), h->pair_shift.as<int>(), 0, st));
  // decision: strict-min in retrieval order, threshold, yaw -- one launch (every candidate's score is on this device)
  PushList pl{};
  if (push) pl = *push;
  k_best_finalize<<<(unsigned)((nq + 3) / 4), 128, 0, st>>>(h->pair_dist.as<double>(), h->pair_shift.as<int>(), h->keys.as<unsigned long long>(),
                                                            (unsigned)nq, h->K, reinterpret_cast<const unsigned long long*>(d_ns), h->L.S,
                                                            h->cfg.dist_thres, d_loop, d_yaw, d_dist, d_idx, d_shift, out_off, out_step, h->rb, pl,
                                                            h->c_count.as<unsigned>());
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

// a public call is about to use the per-handle query workspace on h->stream: order it after an asynchronous replay's
// query stage (which runs on h->qstream)
int join_replay(scgpu_handle* h) {
  CK(cudaStreamWaitEvent(h->stream, h->ev_qdone, 0));
  return SCGPU_OK;
}

// full query pipeline for nq query records already on the device (the shard's own stream); results land in h->o_*.
int run_pipeline(scgpu_handle* h, const void* d_qrec, size_t nq, const uint64_t* h_ns) {
  cudaStream_t st = h->stream;
  RET(join_replay(h));
  RET(h->nsearch.reserve(nq * sizeof(uint64_t)));
  RET(h->h_ns.reserve(nq * sizeof(uint64_t)));
  CK(cudaEventSynchronize(h->ev_nl));  // the previous copy out of the pinned staging cell has finished
  memcpy(h->h_ns.p, h_ns, nq * sizeof(uint64_t));
  CK(cudaMemcpyAsync(h->nsearch.p, h->h_ns.p, nq * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(h->ev_nl, st));
  unsigned chunk, chunks;
  choose_chunks(h->peer ? h->n_global : local_count(h, h->n_global), nq, chunk, chunks);
  RET(query_reserve(h, nq, chunks, st));
  RET(query_stage(h, d_qrec, nq, h->nsearch.as<uint64_t>(), st, h->o_loop.as<int>(), h->o_yaw.as<float>(), h->o_dist.as<double>(),
                  h->o_idx.as<int>(), h->o_shift.as<int>(), 0, 1, nullptr));
  h->last_nq = nq;
  h->last_nsearch.assign(h_ns, h_ns + nq);
  h->last_mutation = h->mutation;
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

// ---- host thread pool: packs x, y, z of host scans into pinned staging (12 bytes per point) ---------------------------
// The caller's clouds are pcl::PointXYZI records (32 bytes; mapOpt.cpp:1628-1630 passes a pageable cloud) or float4: only
// 12 of those bytes are read by the path, and the PCIe link is what bounds the end-to-end rate -- so pageable sources,
// which have to be staged through pinned memory anyway, are packed on the way (2.67x / 1.33x fewer bytes over the link),
// by a few host threads instead of one memcpy.  Pinned sources go straight to the DMA engine unless SCGPU_PACK_PINNED=1.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool p;
    return p;
  }
  int threads() const { return (int)workers_.size() + 1; }
  int local_world() const { return local_world_; }
  // fn(i) for i in [0, n); the calling thread takes part; returns when every item is done
  void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    if (workers_.empty() || n == 1) {
      for (size_t i = 0; i < n; ++i) fn(i);
      return;
    }
    auto job = std::make_shared<Job>();
    job->fn = &fn;
    job->n = n;
    job->left = n;
    {
      std::lock_guard<std::mutex> g(mu_);
      job_ = job;
      ++gen_;
    }
    cv_.notify_all();
    run(*job);
    // the caller has nothing else to do: poll (a sleeping caller would add a wake-up latency to every chunk)
    for (int spins = 0;; ++spins) {
      {
        std::lock_guard<std::mutex> g(job->mu);
        if (job->left == 0) break;
      }
      if (spins > 2000) std::this_thread::yield();
    }
  }

 private:
  struct Job {  // workers hold a reference to the job they joined: a late worker never touches a newer job's counters
    const std::function<void(size_t)>* fn = nullptr;
    size_t n = 0, left = 0;
    std::atomic<size_t> next{0};
    std::mutex mu;
    std::condition_variable done;
  };
  HostPool() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    int local_world = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local_world = atoi(e) > 0 ? atoi(e) : 1;  // one process per GPU: share the cores
    local_world_ = local_world;
    int n = (int)(hw / (unsigned)local_world);
    if (n > 16) n = 16;
    if (const char* e = getenv("SCGPU_HOST_THREADS")) n = atoi(e);
    if (n < 1) n = 1;
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  static void run(Job& j) {
    size_t did = 0;
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.n) break;
      (*j.fn)(i);  // valid: the caller of parallel_for cannot return before `left` reaches 0
      ++did;
    }
    if (did) {
      std::lock_guard<std::mutex> g(j.mu);
      j.left -= did;
      if (j.left == 0) j.done.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        job = job_;
      }
      run(*job);
    }
  }
  std::vector<std::thread> workers_;
  int local_world_ = 1;
  std::mutex mu_;
  std::condition_variable cv_;
  std::shared_ptr<Job> job_;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// x, y and the binned value (z, or the intensity at byte 16) of points [0, n) at `stride` bytes -> 12 bytes per point
void pack_xyz(const unsigned char* src, size_t stride, size_t n, float* dst, size_t val_off = 8) {
  if (stride == 12 && val_off == 8) {
    memcpy(dst, src, n * 12);
    return;
  }
  const size_t vi = val_off / 4;
  for (size_t i = 0; i < n; ++i) {
    const float* p = reinterpret_cast<const float*>(src + i * stride);
    dst[3 * i] = p[0];
    dst[3 * i + 1] = p[1];
    dst[3 * i + 2] = p[vi];
  }
}

// Host points -> records.  Scans at `src_pitch` bytes (0 = contiguous).  Chunks of scans are staged double-buffered: while
// chunk c is on the link (copy stream) and chunk c-1 is being binned (compute stream), the host threads pack chunk c+1.
int build_from_host(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, void* d_records, size_t src_pitch = 0) {
  const size_t scan_bytes = pts_per_scan * stride;
  if (src_pitch == 0) src_pitch = scan_bytes;
  if (scan_bytes == 0) return launch_build(h, h->d_pts[0].p ? h->d_pts[0].p : d_records, n_scans, 0, stride ? stride : 16, d_records, h->stream);
  if (stride < 12 || (stride & 3) || ((uintptr_t)pts & 3) || (src_pitch & 3))
    return fail(SCGPU_E_INVALID, "points must be 4-byte aligned, stride >= 12 and a multiple of 4");
  int dev;
  const bool pinned = is_pinned_or_device(pts, &dev) && !dev;
  // Pinned sources: DMA straight from the caller's buffer costs no host work but ships the padding (16 or 32 bytes per point);
  // packing first ships 12.  Measured on the B200 host (4,541 x 120k float4 points per step, PCIe at 55.5 GB/s): direct 159 ms,
  // packed 122 ms with 16 threads (the link rate for 12 B/point) but 159 ms with 8 -- so pack only when this process has the
  // cores to outrun the link (SCGPU_PACK_PINNED=0/1 overrides; one process per GPU divides the cores, see HostPool).
  // With several ranks on one host the packers of all ranks share the memory system: 2 ranks x 16 threads packed at 40k keyframes/s
  // in total where plain DMA from the pinned buffers gives 57k -- so only a lone process packs pinned sources.
  static const bool pack_pinned = getenv("SCGPU_PACK_PINNED") ? atoi(getenv("SCGPU_PACK_PINNED")) != 0
                                                               : (HostPool::get().threads() >= 12 && HostPool::get().local_world() == 1);
  static const bool no_pack = getenv("SCGPU_NO_PACK") && atoi(getenv("SCGPU_NO_PACK")) != 0;
  // the voxel path wants its input as given (it carries no restriction on stride, but keeps the code path of round 1)
  const bool intensity = (h->cfg.flags & SCGPU_FLAG_INTENSITY) != 0;
  if (intensity && stride < 20) return fail(SCGPU_E_INVALID, "the intensity descriptor needs pcl::PointXYZI-like records (stride >= 20 bytes)");
  const bool pack = !no_pack && (!pinned || (pack_pinned && n_scans * scan_bytes >= ((size_t)64 << 20))) && !(h->voxel_leaf > 0.f);
  const size_t out_stride = pack ? 12 : stride;
  // staged bytes per scan, padded so that every scan starts 16-byte aligned on the device (TMA)
  const size_t out_scan = h->voxel_leaf > 0.f ? pts_per_scan * out_stride : ((pts_per_scan * out_stride + 15) & ~(size_t)15);
  size_t per_chunk = (size_t)(24u << 20) / out_scan;
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > n_scans) per_chunk = n_scans;
  RET(h->d_pts[0].reserve(per_chunk * out_scan));
  if (per_chunk < n_scans) RET(h->d_pts[1].reserve(per_chunk * out_scan));
  const unsigned char* src = static_cast<const unsigned char*>(pts);
  const bool direct = pinned && !pack;
  int turn = 0;
  for (size_t s0 = 0; s0 < n_scans; s0 += per_chunk, turn ^= 1) {
    const size_t ns = n_scans - s0 < per_chunk ? n_scans - s0 : per_chunk;
    if (direct) {
      CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[turn], 0));
      if (src_pitch == scan_bytes && out_scan == scan_bytes)
        CK(cudaMemcpyAsync(h->d_pts[turn].p, src + s0 * src_pitch, ns * scan_bytes, cudaMemcpyHostToDevice, h->copy_stream));
      else
        CK(cudaMemcpy2DAsync(h->d_pts[turn].p, out_scan, src + s0 * src_pitch, src_pitch, scan_bytes, ns, cudaMemcpyHostToDevice, h->copy_stream));
    } else {
      // pageable (or packed) source: through our own pinned buffer so that the copy is truly asynchronous
      RET(h->h_pts[turn].reserve(per_chunk * out_scan));
      CK(cudaEventSynchronize(h->ev_pin[turn]));  // the previous copy out of this staging buffer has finished
      unsigned char* stage = static_cast<unsigned char*>(h->h_pts[turn].p);
      // work items: slices of scans, so that one scan (the online call) also spreads over the threads
      const size_t slice = 16384;
      const size_t slices_per_scan = (pts_per_scan + slice - 1) / slice;
      // One online scan (a few MB) is packed by the calling thread alone: waking sleeping workers costs more than they return
      // (measured on the B200 host, 120k-point PointXYZI scan: 200 us with 1 thread, 240 / 310 / 370 us with 2 / 4 / 16 --
      // idle cores take ~100 us to come back); batches go to the pool (16 threads: 100 GB/s of host reads).
      const bool pooled = ns * pts_per_scan * stride >= ((size_t)16 << 20);
      auto each = [&](size_t n_items, const std::function<void(size_t)>& fn) {
        if (pooled) HostPool::get().parallel_for(n_items, fn);
        else
          for (size_t w = 0; w < n_items; ++w) fn(w);
      };
      each(ns * slices_per_scan, [&](size_t w) {
        const size_t sc = w / slices_per_scan, p0 = (w % slices_per_scan) * slice;
        const size_t np = pts_per_scan - p0 < slice ? pts_per_scan - p0 : slice;
        const unsigned char* from = src + (s0 + sc) * src_pitch + p0 * stride;
        unsigned char* to = stage + sc * out_scan + p0 * out_stride;
        if (pack) pack_xyz(from, stride, np, reinterpret_cast<float*>(to), intensity ? 16 : 8);
        else memcpy(to, from, np * stride);
      });
      CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[turn], 0));
      CK(cudaMemcpyAsync(h->d_pts[turn].p, stage, ns * out_scan, cudaMemcpyHostToDevice, h->copy_stream));
      CK(cudaEventRecord(h->ev_pin[turn], h->copy_stream));
    }
    CK(cudaEventRecord(h->ev_copied[turn], h->copy_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_copied[turn], 0));
    RET(launch_build(h, h->d_pts[turn].p, ns, pts_per_scan, out_stride, static_cast<unsigned char*>(d_records) + s0 * h->L.rec_bytes, h->stream,
                     out_scan));
    CK(cudaEventRecord(h->ev_consumed[turn], h->stream));
  }
  return SCGPU_OK;
}

int build_any(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location, void* d_records,
              size_t src_pitch = 0) {
  if (location == 1) return launch_build(h, pts, n_scans, pts_per_scan, stride, d_records, h->stream, src_pitch);
  return build_from_host(h, pts, n_scans, pts_per_scan, stride, d_records, src_pitch);
}

int pair_api(scgpu_handle* h, const double* a, size_t na, const double* b, size_t nb, int mode, double* out_d, size_t nd, int* out_i) {
  h = GROUP_FIRST(h);
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

constexpr size_t EXH_MAX_BATCH_TC = 64;      // queries per tensor-core screening launch (16 groups of 4)
// ---- tensor-core full-shift screening: host side ------------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D FP32 tensor [rows][TC_K], box = box_rows x TC_BK floats, 128-byte swizzle (what the UMMA shared-memory descriptors expect)
int tc_make_map(CUtensorMap* map, void* base, uint64_t rows, uint32_t box_rows, uint64_t k_extent = TC_K, uint32_t box_k = TC_BK) {
  static TensorMapEncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(SCGPU_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<TensorMapEncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k_extent * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_k, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            box_k * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SCGPU_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SCGPU_OK;
}

// hi / lo split of the screening copy up to date with the shard; (re)creates the buffers and maps when the capacity changed
int tc_sync(scgpu_handle* h, cudaStream_t st) {
  const uint64_t have = local_count(h, h->n_global);
  if (h->tc_rows != h->db.cap) {
    CK(cudaStreamSynchronize(st));
    h->tc_e_hi.release();
    h->tc_e_lo.release();
    const size_t bytes = (size_t)h->db.cap * TC_K * sizeof(float);
    RET(h->tc_e_hi.reserve(bytes));
    RET(h->tc_e_lo.reserve(bytes));
    RET(h->tc_q_hi.reserve((size_t)EXH_MAX_BATCH_TC * TC_S * TC_K * sizeof(float), true, st));
    RET(h->tc_q_lo.reserve((size_t)EXH_MAX_BATCH_TC * TC_S * TC_K * sizeof(float), true, st));
    RET(h->tc_qaux.reserve((size_t)EXH_MAX_BATCH_TC * sizeof(TcQueryAux), true, st));
    RET(tc_make_map(&h->tc_maps[0], h->tc_e_hi.p, h->db.cap, TC_M));
    RET(tc_make_map(&h->tc_maps[1], h->tc_e_lo.p, h->db.cap, TC_M));
    RET(tc_make_map(&h->tc_maps[2], h->tc_q_hi.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N));
    RET(tc_make_map(&h->tc_maps[3], h->tc_q_lo.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N));
    h->tc_ev_hi.release();
    h->tc_ev_lo.release();
    RET(h->tc_ev_hi.reserve((size_t)h->db.cap * TC_VK * sizeof(float)));
    RET(h->tc_ev_lo.reserve((size_t)h->db.cap * TC_VK * sizeof(float)));
    RET(h->tc_qv_hi.reserve((size_t)EXH_MAX_BATCH_TC * TC_S * TC_VK * sizeof(float), true, st));
    RET(h->tc_qv_lo.reserve((size_t)EXH_MAX_BATCH_TC * TC_S * TC_VK * sizeof(float), true, st));
    RET(tc_make_map(&h->tc_maps[4], h->tc_ev_hi.p, h->db.cap, TC_M, TC_VK));
    RET(tc_make_map(&h->tc_maps[5], h->tc_ev_lo.p, h->db.cap, TC_M, TC_VK));
    RET(tc_make_map(&h->tc_maps[6], h->tc_qv_hi.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N, TC_VK));
    RET(tc_make_map(&h->tc_maps[7], h->tc_qv_lo.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N, TC_VK));
    RET(tc_make_map(&h->tc_maps[8], h->tc_e_hi.p, h->db.cap, TC2_M, TC_K, TC2_BK));
    RET(tc_make_map(&h->tc_maps[9], h->tc_e_lo.p, h->db.cap, TC2_M, TC_K, TC2_BK));
    RET(tc_make_map(&h->tc_maps[10], h->tc_q_hi.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N, TC_K, TC2_BK));
    RET(tc_make_map(&h->tc_maps[11], h->tc_q_lo.p, (uint64_t)EXH_MAX_BATCH_TC * TC_S, TC_N, TC_K, TC2_BK));
    h->tc_rows = h->db.cap;
    h->tc_upto = 0;
  }
  if (h->tc_upto > h->x_upto) h->tc_upto = h->x_upto;
  if (have > h->tc_upto) {
    k_tc_split_db<<<(unsigned)(have - h->tc_upto), 256, 0, st>>>(h->x_sc_hat, h->tc_e_hi.as<float>(), h->tc_e_lo.as<float>(), h->tc_upto);
    k_tc_split_vk<<<(unsigned)(have - h->tc_upto), 64, 0, st>>>(h->x_vk, h->tc_ev_hi.as<float>(), h->tc_ev_lo.as<float>(), h->tc_upto);
    h->launches += 2;
    CK(cudaGetLastError());
    h->tc_upto = have;
  }
  return SCGPU_OK;
}

constexpr unsigned EXH_CAND_CAP = 65536;   // rescoring list of one batch
constexpr size_t EXH_MAX_BATCH = 64;       // queries per screening launch

size_t tc_batch_threshold() {  // windowed batches of at least this many queries are screened on the tensor cores (0 = never)
  static const long v = getenv("SCGPU_EXH_TC_BATCH") ? atol(getenv("SCGPU_EXH_TC_BATCH")) : 8;
  return v <= 0 ? (size_t)1 << 30 : (size_t)v;
}

int launch_tc_screen(scgpu_handle* h, size_t nq, uint64_t n_max, uint64_t pitch, const unsigned long long* d_nl, unsigned* d_min, unsigned* d_shift,
                     int windowed, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
  k_tc_prep_queries<<<dim3(TC_S, (unsigned)nq), 128, 0, st>>>(h->x_query.as<ExhQuery>(), h->tc_q_hi.as<float>(), h->tc_q_lo.as<float>(),
                                                              h->tc_qaux.as<TcQueryAux>(), windowed ? h->tc_qv_hi.as<float>() : nullptr,
                                                              windowed ? h->tc_qv_lo.as<float>() : nullptr);
  h->launches++;
  TcParams tp;
  tp.vk = h->x_vk;
  tp.qaux = h->tc_qaux.as<TcQueryAux>();
  tp.n_local = d_nl;
  tp.nq = (unsigned)nq;
  tp.n_groups = (unsigned)((nq + TC_QG - 1) / TC_QG);
  tp.n_tiles = (unsigned)((n_max + TC_M - 1) / TC_M);
  tp.d32_pitch = pitch;
  tp.d32 = h->x_d32.as<float>();
  tp.min_bits = d_min;
  tp.shift_out = d_shift;
  tp.windowed = windowed;
  tp.radius = h->radius;
  tp.align_out = nullptr;
  tp.align_in = nullptr;
  // batches: 256 x 240 tiles (k_tc_fullshift2: 1.48x fewer operand bytes per FLOP); SCGPU_TC_TILE=128 keeps the first kernel.
  // Windowed batches take two passes then: the alignment GEMM alone (k_tc_fullshift, 128-row tiles), then the big-tile kernel.
  static const bool big_tile = !(getenv("SCGPU_TC_TILE") && atoi(getenv("SCGPU_TC_TILE")) == 128);
  const bool use2 = big_tile && nq >= TC_QG;
  if (ev0) CK(cudaEventRecord(ev0, st));
  if (use2 && windowed) {
    RET(h->tc_align.reserve((nq * pitch + 16) * sizeof(unsigned)));
    TcParams ta = tp;
    ta.align_out = h->tc_align.as<unsigned>();
    const uint64_t items_a = (uint64_t)ta.n_groups * ta.n_tiles;
    k_tc_fullshift<<<(unsigned)std::min<uint64_t>(items_a, (uint64_t)h->sm_count), TC_THREADS, tc_smem_bytes(), st>>>(
        h->tc_maps[0], h->tc_maps[1], h->tc_maps[2], h->tc_maps[3], h->tc_maps[4], h->tc_maps[5], h->tc_maps[6], h->tc_maps[7], ta);
    h->launches++;
    tp.align_in = h->tc_align.as<unsigned>();
    tp.windowed = 0;
  }
  if (use2) tp.n_tiles = (unsigned)((n_max + TC2_M - 1) / TC2_M);
  const uint64_t items = (uint64_t)tp.n_groups * tp.n_tiles;
  const unsigned gx = (unsigned)std::min<uint64_t>(items, (uint64_t)h->sm_count);
  if (use2)
    k_tc_fullshift2<<<gx, TC_THREADS, tc2_smem_bytes(), st>>>(h->tc_maps[8], h->tc_maps[9], h->tc_maps[10], h->tc_maps[11], tp);
  else
    k_tc_fullshift<<<gx, TC_THREADS, tc_smem_bytes(), st>>>(h->tc_maps[0], h->tc_maps[1], h->tc_maps[2], h->tc_maps[3], h->tc_maps[4], h->tc_maps[5],
                                                            h->tc_maps[6], h->tc_maps[7], tp);
  if (ev1) CK(cudaEventRecord(ev1, st));
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

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
    // One block per SM in total: with several screening rows (a batch of queries, or forward + column-reversed) every row gets
    // sm_count / rows blocks, all resident at once and all walking the database in the same order -- an entry fetched from HBM
    // for one row is an L2 hit for the others (rows x sm_count blocks would run the rows one after the other, each streaming
    // the whole database from HBM again).
    uint64_t per_row = (uint64_t)h->sm_count / rows;
    if (per_row < 1) per_row = 1;
    if (const char* e = getenv("SCGPU_EXH_BLOCKS_PER_ROW")) per_row = (uint64_t)std::max(1, atoi(e));
    const unsigned grid = (unsigned)(groups < per_row ? groups : per_row);
    float eps = EXH_EPS;
    if (h->exh_cfg == 3) {
      // every shift is searched: the distance table of a (query, entry) pair is a GEMM tile -- tensor cores (scgpu_tc.cuh);
      // SCGPU_FULLSHIFT_SIMT=1 runs the FFMA2 counterpart instead (A/B)
      static const bool simt = getenv("SCGPU_FULLSHIFT_SIMT") && atoi(getenv("SCGPU_FULLSHIFT_SIMT")) != 0;
      if (flipped) return fail(SCGPU_E_INVALID, "the column-reversed pass is not instantiated for the full-shift search");
      RET(tc_sync(h, st));
      unsigned* d_shift = nullptr;
      if (h->tc_want_shifts) {
        RET(h->tc_shift.reserve((rows * pitch + 16) * sizeof(unsigned)));
        d_shift = h->tc_shift.as<unsigned>();
      }
      if (simt) {
        if (ev_screen0) CK(cudaEventRecord(ev_screen0, st));
        const unsigned gx = (unsigned)std::min<uint64_t>((n_max + 7) / 8, (uint64_t)h->sm_count * 8);
        k_fullshift_simt<TC_R, TC_S><<<dim3(gx, (unsigned)nq), 256, qtab_bytes<TC_R, TC_S, 15>(), st>>>(
            h->x_sc_hat, h->x_vk, h->x_query.as<ExhQuery>(), d_nl, pitch, h->x_d32.as<float>(), d_min, d_shift);
        if (ev_screen1) CK(cudaEventRecord(ev_screen1, st));
      } else {
        RET(launch_tc_screen(h, nq, n_max, pitch, d_nl, d_min, d_shift, 0, st, ev_screen0, ev_screen1));
        eps = TC_EPS;
      }
    } else if (h->exh_cfg == 1 && !flipped && nq >= tc_batch_threshold() && h->L.R == TC_R && h->L.S == TC_S) {
      // batches of the WINDOWED search: the tensor-core kernel computes all 60 shifts plus the alignment (a second small GEMM) and
      // is still faster than the FFMA2 kernel that computes only the 7 it needs (measured: DESIGN.md section 4)
      RET(tc_sync(h, st));
      RET(launch_tc_screen(h, nq, n_max, pitch, d_nl, d_min, nullptr, 1, st, ev_screen0, ev_screen1));
      eps = TC_EPS;
    } else {
      if (ev_screen0) CK(cudaEventRecord(ev_screen0, st));
      if (h->exh_cfg == 1)
        k_exh_screen<20, 60, 3, 1, 20><<<dim3(grid, (unsigned)rows), 20 * 32, exh_smem_bytes<20, 60, 3, 20>(), st>>>(sp);
      else
        k_exh_screen<40, 120, 6, 1, 4, 2><<<dim3(grid, (unsigned)rows), 4 * 2 * 32, exh_smem_bytes<40, 120, 6, 4, 2>(), st>>>(sp);
      if (ev_screen1) CK(cudaEventRecord(ev_screen1, st));
    }
    CK(cudaGetLastError());
    const unsigned rb = (unsigned)((n_max + 1023) / 1024 < 296 ? (n_max + 1023) / 1024 : 296);
    k_exh_compact<<<dim3(rb, (unsigned)rows), 256, 0, st>>>(sp.d32, pitch, d_nl, d_min, h->db.rank, h->db.G, sp.flip_mode,
                                                           h->x_keys.as<unsigned long long>(), d_count, EXH_CAND_CAP, eps);
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
    p.peers = PeerTab{};  // every shard rescores its own entries
    k_score_list<<<dim3((unsigned)h->sm_count * 4, 1), 128, pair_smem_bytes(h->L.R, h->L.S, h->W, sizeof(float)), st>>>(p);
    h->launches += 3;
    CK(cudaGetLastError());
  }
  k_exh_final<<<(unsigned)nq, 256, 0, st>>>(h->x_pd.as<double>(), h->x_ps.as<int>(), h->x_keys.as<unsigned long long>(), d_count, EXH_CAND_CAP, d_best_out);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}


// ======================================================================================================================
// Replay engine: "append n scans, detect after each" for one shard, for the G shards of a device-list handle (one host
// process), or for one shard of a database whose other shards are driven by other processes (torchrun, one per GPU).
//
//   build stream (h->stream)    : [H2D] -> k_build(chunk c) -> k_append (+ ring keys pushed into every shard's replica)
//                                  -> barrier "appended" -> event
//   query stream (h->qstream)   : wait event -> k_topk over the replica -> k_cand_screen / k_score_pairs (candidate rows read
//                                  from their owner shards) -> k_best -> k_finalize (results pushed to every shard)
//                                  ... after the last chunk: barrier "queries done"
// Queries are PARTITIONED: shard r scores the scans it binned itself (global entries first + j*G + r), so per-query work
// divides by G, no record leaves its device, and nothing is merged.  The build of chunk c+1 -- and of the NEXT replay call --
// runs while the query stage of chunk c is in flight; the next call's append waits for "queries done" (its writes would
// race with peers still reading).
// ======================================================================================================================
struct ReplayShard {
  scgpu_handle* h;
  const void* pts;  // this shard's scans: scan j at pts + j * pitch
  size_t pitch;     // bytes between this shard's consecutive scans (0 = contiguous)
};

int barrier_timeout_flag(scgpu_handle* h, unsigned* out) {
  *out = 0;
  if (!h->peer) return SCGPU_OK;
  CK(cudaMemcpyAsync(out, h->flags_of[h->cfg.shard_rank] + 2 * MAX_SHARDS, sizeof(unsigned), cudaMemcpyDeviceToHost, h->qstream));
  return SCGPU_OK;
}

int launch_peer_barrier(scgpu_handle* h, int channel, cudaStream_t st) {
  BarrierCells bc{};
  for (int s = 0; s < h->cfg.shard_count; ++s) bc.cells[s] = h->flags_of[s];
  const unsigned e = ++h->epoch[channel];
  // generous bound: ranks reach their first steps seconds apart (allocations, page-ins); $SCGPU_BARRIER_TIMEOUT_MS overrides
  static const unsigned long long timeout_ns =
      (getenv("SCGPU_BARRIER_TIMEOUT_MS") ? strtoull(getenv("SCGPU_BARRIER_TIMEOUT_MS"), nullptr, 10) : 20000ull) * 1000000ull;
  k_peer_barrier<<<1, 32, 0, st>>>(bc, h->cfg.shard_count, h->cfg.shard_rank, channel, e, timeout_ns);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

void* result_base(scgpu_handle* h) { return h->peer ? h->results_of[h->cfg.shard_rank] : h->res_buf.p; }
const ResultBlock& result_rb(scgpu_handle* h) { return h->peer ? h->rb : h->res_rb; }

// sh: the shards this process drives (all G of a device-list handle, or the one shard of this rank).  The batch is n_total
// scans; scan i becomes global entry first + i and belongs to shard (first + i) % G; x.pts is the shard's FIRST scan of the
// batch, its following ones x.pitch bytes apart.  ipc: the other shards live in other processes -> cross-device flag
// barriers instead of events (every process makes the same sequence of barrier calls).
int replay_enqueue(std::vector<ReplayShard>& sh, uint64_t first, size_t n_total, size_t P, size_t stride, int location, bool ipc) {
  scgpu_handle* h0 = sh[0].h;
  const uint64_t G = h0->peer ? (uint64_t)h0->cfg.shard_count : 1;
  if (n_total == 0) return SCGPU_OK;
  if (n_total > 65535 || (h0->peer && n_total > PEER_RESULT_CAP)) return fail(SCGPU_E_INVALID, "at most 65535 scans per replay call");
  if (!ipc && sh.size() != G) return fail(SCGPU_E_INVALID, "internal: shard list does not cover the database");
  const size_t B = (n_total + G - 1) / G;  // scans per shard (the last ones of some shards may be missing)
  // chunks (SCGPU_REPLAY_CHUNKS, default 1): the query stage of chunk c may run while chunk c+1 is binned.  Measured on the
  // 4,541-keyframe run (B200): 1.88 / 2.01 / 2.15 / 2.01 ms per step with 1 / 2 / 3 / 4 chunks -- the binning kernel fills every
  // SM (4 blocks x 256 threads x 64 registers), so the query kernels only get slots in its tail and the smaller launches lose
  // more to partial waves than the overlap returns.  One chunk it is.
  size_t C = 1;
  if (const char* e = getenv("SCGPU_REPLAY_CHUNKS")) C = (size_t)std::max(1, atoi(e));
  if (C > B) C = B;
  auto shard_off = [&](scgpu_handle* h) {  // index within the batch of the shard's first scan (= scgpu_peer_partition)
    const uint64_t r = h->peer ? (uint64_t)h->cfg.shard_rank : 0;
    return (size_t)((r + G - first % G) % G);
  };
  auto shard_cnt = [&](size_t off) { return n_total > off ? (n_total - 1 - off) / G + 1 : 0; };
  std::vector<uint64_t> plan_all(n_total);
  for (auto& x : sh) {
    scgpu_handle* h = x.h;
    CK(cudaSetDevice(h->cfg.device));
    plan(h, first + 1, n_total, plan_all.data());  // every shard advances the same snapshot state
    h->mutation++;
    h->rec_turn ^= 1;
    RET(h->records2[h->rec_turn].reserve(B * h->L.rec_bytes));
    RET(h->nsearch.reserve(B * sizeof(uint64_t)));
    if (!h->peer && h->res_rb.cap_q < n_total) {
      CK(cudaStreamSynchronize(h->qstream));
      RET(h->res_buf.reserve((n_total + n_total / 4 + 64) * 24));
      h->res_rb.cap_q = h->res_buf.bytes / 24;
    }
    // n_search of this shard's queries through a rotating pinned cell (several replays may be in flight)
    const int slot = h->ns_turn++ & 3;
    RET(h->h_ns_ring[slot].reserve(B * sizeof(uint64_t)));
    CK(cudaEventSynchronize(h->ev_ns_ring[slot]));
    uint64_t* hn = static_cast<uint64_t*>(h->h_ns_ring[slot].p);
    const size_t off = shard_off(h), Bx = shard_cnt(off);
    for (size_t j = 0; j < Bx; ++j) hn[j] = plan_all[off + j * G];
    if (Bx) CK(cudaMemcpyAsync(h->nsearch.p, hn, Bx * sizeof(uint64_t), cudaMemcpyHostToDevice, h->qstream));
    CK(cudaEventRecord(h->ev_ns_ring[slot], h->qstream));
    h->last_nsearch.assign(hn, hn + Bx);
  }
  for (size_t c = 0; c < C; ++c) {
    for (auto& x : sh) {
      scgpu_handle* h = x.h;
      CK(cudaSetDevice(h->cfg.device));
      const size_t off = shard_off(h), Bx = shard_cnt(off);
      const size_t j0 = std::min(Bx, B * c / C), j1 = std::min(Bx, B * (c + 1) / C), nj = j1 - j0;
      unsigned char* rec = h->records2[h->rec_turn].as<unsigned char>() + j0 * h->L.rec_bytes;
      const size_t pitch = x.pitch ? x.pitch : P * stride;
      if (c == 0) CK(cudaEventRecord(h->ev_t0, h->stream));
      if (h->trace && c == 0) {
        const int k = h->tr_n & 15;
        for (int e = 0; e < 4; ++e)
          if (!h->tr_ev[k][e]) CK(cudaEventCreate(&h->tr_ev[k][e]));
        CK(cudaEventRecord(h->tr_ev[k][0], h->stream));
      }
      if (nj) RET(build_any(h, static_cast<const unsigned char*>(x.pts) + j0 * pitch, nj, P, stride, location, rec, x.pitch));
      if (c == C - 1) CK(cudaEventRecord(h->ev_t1, h->stream));
      if (h->trace && c == C - 1) CK(cudaEventRecord(h->tr_ev[h->tr_n & 15][1], h->stream));
      if (c == 0)  // the previous replay's query stage (on every shard) may still be reading what the append overwrites
        for (auto& y : sh) CK(cudaStreamWaitEvent(h->stream, y.h->ev_qdone, 0));
      if (nj) RET(launch_append(h, rec, first + off + j0 * G, G, nj, h->stream, G > 1));
      if (ipc && G > 1) RET(launch_peer_barrier(h, 0, h->stream));
      CK(cudaEventRecord(h->ev_app, h->stream));
    }
    const uint64_t present = first + std::min<uint64_t>(n_total, (uint64_t)(B * (c + 1) / C) * G);  // entries in place after this chunk
    for (auto& x : sh) {
      scgpu_handle* h = x.h;
      CK(cudaSetDevice(h->cfg.device));
      const size_t off = shard_off(h), Bx = shard_cnt(off);
      const size_t j0 = std::min(Bx, B * c / C), j1 = std::min(Bx, B * (c + 1) / C), nj = j1 - j0;
      for (auto& y : sh) CK(cudaStreamWaitEvent(h->qstream, y.h->ev_app, 0));
      if (h->trace && c == 0) CK(cudaEventRecord(h->tr_ev[h->tr_n & 15][2], h->qstream));
      h->n_global = present;
      const uint64_t have = local_count(h, h->n_global);
      if (have > h->n_written) h->n_written = have;
      PushList pl{};
      if (G > 1)
        for (int s2 = 0; s2 < (int)G; ++s2)
          if (s2 != h->cfg.shard_rank) pl.dst[pl.n++] = h->results_of[s2];
      void* rbase = result_base(h);
      const ResultBlock& rb = result_rb(h);
      const unsigned char* rec = h->records2[h->rec_turn].as<unsigned char>() + j0 * h->L.rec_bytes;
      if (nj)
        RET(query_stage(h, rec, nj, h->nsearch.as<uint64_t>() + j0, h->qstream, rb.loop(rbase), rb.yaw(rbase), rb.dist(rbase), rb.idx(rbase),
                        rb.shift(rbase), (unsigned)(off + j0 * G), (unsigned)G, &pl));
    }
  }
  for (auto& x : sh) {
    scgpu_handle* h = x.h;
    CK(cudaSetDevice(h->cfg.device));
    if (ipc && G > 1) RET(launch_peer_barrier(h, 1, h->qstream));
    CK(cudaEventRecord(h->ev_qdone, h->qstream));
    CK(cudaEventRecord(h->ev_t2, h->qstream));
    if (h->trace) CK(cudaEventRecord(h->tr_ev[h->tr_n++ & 15][3], h->qstream));
    h->timing_valid = true;
    h->replay_nq = n_total;
    h->replay_ipc = ipc;
    // candidate dumps (scgpu_get_batch_candidates) need the whole batch in the query workspace: one chunk, one shard
    h->last_nq = (C == 1 && !h->peer) ? n_total : 0;
    h->last_mutation = h->mutation;
  }
  return SCGPU_OK;
}

// results of the last replay_enqueue -> caller arrays.  Device-list handle: wait for every shard, read shard 0's block.
int replay_fetch(std::vector<scgpu_handle*>& hs, size_t nq, int* loop_id, float* yaw, double* dist, int* idx, int* shift) {
  for (scgpu_handle* h : hs) {
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaEventSynchronize(h->ev_qdone));
    CK(cudaStreamSynchronize(h->copy_stream));
  }
  scgpu_handle* h = hs[0];
  CK(cudaSetDevice(h->cfg.device));
  if (nq > h->replay_nq) return fail(SCGPU_E_INVALID, "the last replay produced %zu results", h->replay_nq);
  RET(h->h_out.reserve(nq * 24 + 64));
  unsigned char* b = static_cast<unsigned char*>(h->h_out.p);
  void* rbase = result_base(h);
  const ResultBlock& rb = result_rb(h);
  cudaStream_t st = h->qstream;
  CK(cudaMemcpyAsync(b, rb.dist(rbase), nq * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(b + nq * 8, rb.loop(rbase), nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(b + nq * 12, rb.yaw(rbase), nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(b + nq * 16, rb.idx(rbase), nq * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(b + nq * 20, rb.shift(rbase), nq * 4, cudaMemcpyDeviceToHost, st));
  unsigned timed_out = 0;
  if (h->replay_ipc) RET(barrier_timeout_flag(h, &timed_out));
  CK(cudaStreamSynchronize(st));
  if (timed_out) return fail(SCGPU_E_CUDA, "peer barrier timed out: a shard of the database did not reach the step");
  if (dist) memcpy(dist, b, nq * 8);
  if (loop_id) memcpy(loop_id, b + nq * 8, nq * 4);
  if (yaw) memcpy(yaw, b + nq * 12, nq * 4);
  if (idx) memcpy(idx, b + nq * 16, nq * 4);
  if (shift) memcpy(shift, b + nq * 20, nq * 4);
  return SCGPU_OK;
}

// ---- device-list handle: one host process, G shards -----------------------------------------------------------------
int group_attach(std::vector<scgpu_handle*>& hs) {
  const int G = (int)hs.size();
  for (int a = 0; a < G; ++a) {
    scgpu_handle* h = hs[a];
    CK(cudaSetDevice(h->cfg.device));
    for (int b = 0; b < G; ++b) {
      const int db = hs[b]->cfg.device;
      if (db != h->cfg.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->cfg.device, db));
        if (!can) return fail(SCGPU_E_CUDA, "device %d cannot access device %d's memory (no peer path)", h->cfg.device, db);
        const cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SCGPU_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
      peer_fill(h, b, hs[b]->slab);
    }
    h->peers.G = G;
    h->peers.rank = a;
    h->peers.cap = h->db.cap;
    h->attached = true;
  }
  return SCGPU_OK;
}

// every shard's appends so far are visible to `h`'s stream `st` (device-list handle: events instead of flag barriers)
int group_wait_appends(scgpu_handle* g, scgpu_handle* h, cudaStream_t st) {
  for (scgpu_handle* s : g->shards) CK(cudaStreamWaitEvent(st, s->ev_app, 0));
  return SCGPU_OK;
}

void group_set_size(scgpu_handle* g, uint64_t n) {
  g->n_global = n;
  for (scgpu_handle* s : g->shards) {
    s->n_global = n;
    const uint64_t have = local_count(s, n);
    if (have > s->n_written) s->n_written = have;
  }
}

}  // namespace

static int group_create(const scgpu_config* cfg, scgpu_handle** out) {
  const int G = cfg->n_devices;
  *out = nullptr;
  if (G > SCGPU_MAX_DEVICES || G > MAX_SHARDS) return fail(SCGPU_E_INVALID, "at most %d devices", SCGPU_MAX_DEVICES);
  if (cfg->shard_count != 1) return fail(SCGPU_E_INVALID, "a device-list handle is the whole database: shard_count must be 1");
  scgpu_handle* g = new scgpu_handle();
  g->is_group = true;
  g->cfg = *cfg;
  g->L = make_layout(cfg->num_ring, cfg->num_sector);
  g->K = cfg->num_candidates;
  for (int r = 0; r < G; ++r) {
    scgpu_config c = *cfg;
    c.n_devices = 0;
    c.device = cfg->devices[r];
    c.shard_rank = r;
    c.shard_count = G;
    c.flags |= SCGPU_FLAG_PEER;
    scgpu_handle* s = nullptr;
    const int rc = create_one(&c, &s);
    if (rc != SCGPU_OK) {
      scgpu_destroy(g);
      return rc;
    }
    s->parent = g;
    g->shards.push_back(s);
  }
  const int rc = group_attach(g->shards);
  if (rc != SCGPU_OK) {
    scgpu_destroy(g);
    return rc;
  }
  g->radius = g->shards[0]->radius;
  g->W = g->shards[0]->W;
  g->exh = g->shards[0]->exh;
  *out = g;
  return SCGPU_OK;
}

// =================================================================================================
extern "C" {

const char* scgpu_last_error(void) { return g_err; }
const char* scgpu_version(void) {
  return "scgpu 0.2 (sm_100a; kernels: k_build k_build_tma k_build_voxel k_append k_append_hat k_gather k_topk k_topk_tile k_merge "
         "k_cand_screen k_score k_score_pairs k_score_list k_best k_finalize k_best_finalize k_pair_api k_exh_prep k_exh_append k_exh_screen "
         "k_exh_compact k_exh_final k_peer_barrier)";
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
  if (cfg->n_devices > 1) return group_create(cfg, out);
  if (cfg->n_devices == 1) {
    scgpu_config c = *cfg;
    c.device = cfg->devices[0];
    c.n_devices = 0;
    return create_one(&c, out);
  }
  return create_one(cfg, out);
}

static int create_one(const scgpu_config* cfg, scgpu_handle** out) {
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
  h->peer = (cfg->flags & SCGPU_FLAG_PEER) && cfg->shard_count > 1;
  if (h->peer && cfg->shard_count > MAX_SHARDS) {
    delete h;
    return fail(SCGPU_E_INVALID, "at most %d peer shards", MAX_SHARDS);
  }
  // FP32 screening kernels are instantiated for the reference's 20x60 (radius 3) and BASELINE's 40x120 (radius 6)
  h->exh_cfg = (h->L.R == 20 && h->L.S == 60 && h->radius == 3) ? 1 : ((h->L.R == 40 && h->L.S == 120 && h->radius == 6) ? 2 : 0);
  if (h->L.R == TC_R && h->L.S == TC_S && 2 * h->radius >= h->L.S) h->exh_cfg = 3;  // every shift is searched: tensor-core screening
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
  {  // the query stage of a pipelined replay: latency-bound kernels that should win SM slots whenever the binning frees any
    int lo = 0, hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->qstream, cudaStreamNonBlocking, hi);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_app, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_qdone, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&h->ev_ns_ring[i], cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pin[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t1);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev_t2);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_nl, cudaEventDisableTiming);
  const int sm16 = (int)std::max(build_tma_smem<16>(h->L.RS), BUILD_ROOM_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm16);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm16);
  h->build_room = getenv("SCGPU_BUILD_ROOM") && atoi(getenv("SCGPU_BUILD_ROOM")) != 0;
  h->trace = getenv("SCGPU_TRACE") && atoi(getenv("SCGPU_TRACE")) != 0;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<32, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<32>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<32, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<32>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<12, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<12>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_tma<12, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)build_tma_smem<12>(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_build_voxel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vox_smem_bytes(h->L.RS));
  if (e == cudaSuccess && h->exh && h->exh_cfg != 3)
    e = h->exh_cfg == 1 ? cudaFuncSetAttribute(k_exh_screen<20, 60, 3, 1, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)exh_smem_bytes<20, 60, 3, 20>())
                        : cudaFuncSetAttribute(k_exh_screen<40, 120, 6, 1, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)exh_smem_bytes<40, 120, 6, 4, 2>());
  if (e == cudaSuccess && (h->exh_cfg == 3 || h->exh_cfg == 1)) {
    e = cudaFuncSetAttribute(k_tc_fullshift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tc_fullshift2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes());
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k_fullshift_simt<TC_R, TC_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qtab_bytes<TC_R, TC_S, 15>());
  }
  if (e == cudaSuccess && h->exh && h->exh_cfg == 1)
    e = cudaFuncSetAttribute(k_cand_screen<20, 60, 3, 1, 10, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cand_smem_bytes<20, 60, 3, 10, 1>());
  if (e == cudaSuccess && h->exh && h->exh_cfg == 1)
    e = cudaFuncSetAttribute(k_cand_screen<20, 60, 3, 1, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cand_smem_bytes<20, 60, 3, 5, 1>());
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
  cap = (cap + 63) & ~63ull;
  int r = h->peer ? slab_alloc(h, cap) : db_reserve(h, cap);
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
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) {
      cudaSetDevice(s->cfg.device);
      cudaDeviceSynchronize();
    }
    for (scgpu_handle* s : h->shards) scgpu_destroy(s);
    delete h;
    return SCGPU_OK;
  }
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->qstream) cudaStreamSynchronize(h->qstream);
  for (int s = 0; s < MAX_SHARDS; ++s)
    if (h->ipc_base[s]) cudaIpcCloseMemHandle(h->ipc_base[s]);
  DevBuf* bufs[] = {&h->gbins, &h->btickets, &h->d_pts[0], &h->d_pts[1], &h->records, &h->rec_single, &h->nsearch, &h->keys, &h->partial,
                    &h->ttickets, &h->pair_dist, &h->pair_shift, &h->best, &h->o_loop, &h->o_yaw, &h->o_dist, &h->o_idx, &h->o_shift,
                    &h->api_in, &h->api_out, &h->vox_info, &h->vox_pts, &h->vox_idx, &h->vox_in, &h->vox_keys, &h->vox_hint};
  for (DevBuf* b : bufs) b->release();
  h->h_pts[0].release();
  h->h_pts[1].release();
  h->h_out.release();
  h->h_ns.release();
  h->q_idx.release();
  for (int i = 0; i < 4; ++i) {
    h->h_ns_ring[i].release();
    if (h->ev_ns_ring[i]) cudaEventDestroy(h->ev_ns_ring[i]);
  }
  if (h->ev_app) cudaEventDestroy(h->ev_app);
  if (h->ev_qdone) cudaEventDestroy(h->ev_qdone);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  for (auto& row : h->tr_ev)
    for (cudaEvent_t e : row)
      if (e) cudaEventDestroy(e);
  if (h->ev_b0) cudaEventDestroy(h->ev_b0);
  if (h->ev_b1) cudaEventDestroy(h->ev_b1);
  if (h->qstream) cudaStreamDestroy(h->qstream);
  if (h->peer) {
    if (h->slab) cudaFree(h->slab);
  } else if (h->vm_state == 1) {
    for (VmBuf& b : h->vm) b.release();
  } else if (h->db.cap) {
    cudaFree(h->db.sc);
    cudaFree(h->db.ringT);
    cudaFree(h->db.sector);
    cudaFree(h->db.colnorm);
    if (h->exh) {
      cudaFree(h->x_sc_hat);
      cudaFree(h->x_vk);
    }
  }
  DevBuf* xb[] = {&h->x_query, &h->x_d32, &h->x_keys, &h->x_pd,    &h->x_ps,    &h->x_small, &h->x_best,  &h->c_d32,
                  &h->c_list,  &h->c_count, &h->tc_e_hi, &h->tc_e_lo, &h->tc_q_hi, &h->tc_q_lo, &h->tc_qaux, &h->tc_shift,
                  &h->tc_ev_hi, &h->tc_ev_lo, &h->tc_qv_hi, &h->tc_qv_lo, &h->tc_align,
                  &h->records2[0], &h->records2[1], &h->res_buf, &h->icp_src, &h->icp_tgt, &h->icp_state, &h->icp_part};
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
  for (scgpu_handle* s : h->shards) *out += s->launches;
  return SCGPU_OK;
}

int scgpu_get_timing(scgpu_handle* h, double* ms_total, double* ms_build, double* ms_query) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) {  // the slowest shard of each phase
    double t[3] = {0, 0, 0};
    for (scgpu_handle* s : h->shards) {
      double a, b, c;
      RET(scgpu_get_timing(s, &a, &b, &c));
      t[0] = std::max(t[0], a);
      t[1] = std::max(t[1], b);
      t[2] = std::max(t[2], c);
    }
    if (ms_total) *ms_total = t[0];
    if (ms_build) *ms_build = t[1];
    if (ms_query) *ms_query = t[2];
    return SCGPU_OK;
  }
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

// Device-side stopwatch over a sequence of (asynchronous) calls: start is an event on the build stream behind everything
// enqueued so far, stop an event behind everything enqueued on both streams; device-list handle: the slowest shard.
int scgpu_timer_start(scgpu_handle* h) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) RET(scgpu_timer_start(s));
    return SCGPU_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  if (!h->ev_b0) {
    CK(cudaEventCreate(&h->ev_b0));
    CK(cudaEventCreate(&h->ev_b1));
  }
  CK(cudaEventRecord(h->ev_side, h->qstream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_side, 0));
  CK(cudaEventRecord(h->ev_b0, h->stream));
  return SCGPU_OK;
}

int scgpu_timer_stop(scgpu_handle* h, double* ms) {
  if (!h || !ms) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) {
    *ms = 0;
    for (scgpu_handle* s : h->shards) {
      double t;
      RET(scgpu_timer_stop(s, &t));
      *ms = std::max(*ms, t);
    }
    return SCGPU_OK;
  }
  if (!h->ev_b0) return fail(SCGPU_E_INVALID, "scgpu_timer_start first");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(h->ev_side, h->stream));
  CK(cudaStreamWaitEvent(h->qstream, h->ev_side, 0));
  CK(cudaEventRecord(h->ev_b1, h->qstream));
  CK(cudaEventSynchronize(h->ev_b1));
  float t = 0;
  CK(cudaEventElapsedTime(&t, h->ev_b0, h->ev_b1));
  *ms = t;
  if (h->trace) {  // timeline of the last (up to 16) steps relative to the stopwatch start
    const int n = h->tr_n < 16 ? h->tr_n : 16;
    for (int i = h->tr_n - n; i < h->tr_n; ++i) {
      float v[4] = {0, 0, 0, 0};
      for (int e = 0; e < 4; ++e)
        if (cudaEventElapsedTime(&v[e], h->ev_b0, h->tr_ev[i & 15][e]) != cudaSuccess) cudaGetLastError();
      fprintf(stderr, "[scgpu trace] step %d: build %.3f .. %.3f ms | query %.3f .. %.3f ms\n", i, v[0], v[1], v[2], v[3]);
    }
    h->tr_n = 0;
  }
  return SCGPU_OK;
}

// Host side of the H2D path: threads of the packing pool and whether pinned batches are packed to 12 bytes per point
int scgpu_host_info(int* pool_threads, int* packs_pinned) {
  if (pool_threads) *pool_threads = HostPool::get().threads();
  if (packs_pinned)
    *packs_pinned = getenv("SCGPU_PACK_PINNED") ? atoi(getenv("SCGPU_PACK_PINNED")) != 0
                                                 : (HostPool::get().threads() >= 12 && HostPool::get().local_world() == 1);
  return SCGPU_OK;
}

// How the shard's arrays have grown so far: in place (more physical memory mapped behind a reserved address range: no
// synchronisation, no copy) or by reallocation + copy (virtual memory management unavailable, or SCGPU_NO_VMM=1)
int scgpu_growth_stats(scgpu_handle* h, unsigned* in_place, unsigned* by_copy, uint64_t* capacity) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  h = GROUP_FIRST(h);
  if (in_place) *in_place = h->n_grow_inplace;
  if (by_copy) *by_copy = h->n_grow_copy;
  if (capacity) *capacity = h->db.cap;
  return SCGPU_OK;
}

int scgpu_record_bytes(scgpu_handle* h, size_t* out) {
  if (!h || !out) return fail(SCGPU_E_INVALID, "null argument");
  *out = h->L.rec_bytes;
  return SCGPU_OK;
}

// ---- reference surface ----------------------------------------------------------------------------

int scgpu_make_sc(scgpu_handle* h, const void* pts, size_t n, size_t stride, double* out_sc) {
  SCGPU_TRACE_CALL();
  if (!h || !out_sc || (!pts && n)) return fail(SCGPU_E_INVALID, "null argument");
  h = GROUP_FIRST(h);
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
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

// scans [0, n) of this shard's share of a batch: scan j is global entry first_global + j * step, its points at pts + j * pitch
static int append_scans_one(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location,
                            uint64_t first_global, uint64_t step, size_t pitch, bool push) {
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
  CK(cudaEventRecord(h->ev_t0, h->stream));
  if (n_scans) {
    RET(h->records.reserve(n_scans * h->L.rec_bytes));
    RET(build_any(h, pts, n_scans, pts_per_scan, stride, location, h->records.p, pitch));
  }
  CK(cudaEventRecord(h->ev_t1, h->stream));
  if (n_scans) RET(launch_append(h, h->records.p, first_global, step, n_scans, h->stream, push));
  CK(cudaEventRecord(h->ev_t2, h->stream));
  CK(cudaEventRecord(h->ev_app, h->stream));
  h->timing_valid = true;
  return SCGPU_OK;
}

int scgpu_append_scans_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location) {
  SCGPU_TRACE_CALL();
  if (!h || (!pts && n_scans && pts_per_scan)) return fail(SCGPU_E_INVALID, "null argument");
  if (n_scans == 0) return SCGPU_OK;
  if (h->is_group) {  // scan i -> shard (size + i) % G, binned on that shard's device; ring keys pushed to every replica
    const uint64_t G = h->shards.size(), first = h->n_global;
    const size_t scan_bytes = pts_per_scan * stride;
    for (scgpu_handle* s : h->shards) {
      const size_t off = (size_t)(((uint64_t)s->cfg.shard_rank + G - first % G) % G);
      const size_t cnt = n_scans > off ? (n_scans - 1 - off) / G + 1 : 0;
      RET(append_scans_one(s, static_cast<const unsigned char*>(pts) + off * scan_bytes, cnt, pts_per_scan, stride, location, first + off, G,
                           G * scan_bytes, true));
    }
    if (location == 0)
      for (scgpu_handle* s : h->shards) {
        CK(cudaSetDevice(s->cfg.device));
        CK(cudaStreamSynchronize(s->copy_stream));
      }
    group_set_size(h, first + n_scans);
    return SCGPU_OK;
  }
  if (h->peer) return fail(SCGPU_E_INVALID, "a shard of a peer-sharded database appends through scgpu_peer_replay_async / the staged API");
  RET(append_scans_one(h, pts, n_scans, pts_per_scan, stride, location, h->n_global, 1, 0, false));
  if (location == 0) CK(cudaStreamSynchronize(h->copy_stream));  // the caller's buffer is not retained
  return SCGPU_OK;
}

int scgpu_append_scan(scgpu_handle* h, const void* pts, size_t n, size_t stride) {
  return scgpu_append_scans_batched(h, pts, 1, n, stride, 0);
}

// Ready-made descriptors.  A shard of a sharded database is given ALL n descriptors (every rank / shard the same ones): it
// keeps the entries it owns and, peer-sharded, the ring key of every entry in its replica -- nothing crosses devices.
static int append_descs_one(scgpu_handle* h, const float* sc, size_t n) {
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
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
  if (rc == SCGPU_OK && cudaEventRecord(h->ev_app, h->stream) != cudaSuccess) rc = fail(SCGPU_E_CUDA, "event record failed");
  tmp.release();
  return rc;
}

int scgpu_append_descs(scgpu_handle* h, const float* sc, size_t n) {
  if (!h || (!sc && n)) return fail(SCGPU_E_INVALID, "null argument");
  if (n == 0) return SCGPU_OK;
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) RET(append_descs_one(s, sc, n));
    group_set_size(h, h->n_global + n);
    return SCGPU_OK;
  }
  return append_descs_one(h, sc, n);
}

// detect for the stored entries [first, first + nq) of a device-list handle: shard r gathers and scores the queries it
// owns (candidate rows come from whichever shard holds them), results land in shard 0's block
static int group_query_range(scgpu_handle* g, uint64_t first, size_t nq, const uint64_t* ns, int* loop_id, float* yaw, double* nearest_dist,
                             int* nearest_idx, int* nearest_shift) {
  const uint64_t G = g->shards.size();
  if (nq > PEER_RESULT_CAP) return fail(SCGPU_E_INVALID, "at most %llu queries per call", (unsigned long long)PEER_RESULT_CAP);
  for (scgpu_handle* h : g->shards) {
    CK(cudaSetDevice(h->cfg.device));
    RET(join_replay(h));
    const size_t off = (size_t)(((uint64_t)h->cfg.shard_rank + G - first % G) % G);
    const size_t cnt = nq > off ? (nq - 1 - off) / G + 1 : 0;
    CK(cudaEventRecord(h->ev_t0, h->stream));
    CK(cudaEventRecord(h->ev_t1, h->stream));
    if (cnt) {
      RET(group_wait_appends(g, h, h->stream));
      RET(h->records.reserve(cnt * h->L.rec_bytes));
      RET(h->q_idx.reserve(cnt * sizeof(uint64_t)));
      RET(h->nsearch.reserve(cnt * sizeof(uint64_t)));
      RET(h->h_ns.reserve(2 * cnt * sizeof(uint64_t)));
      CK(cudaEventSynchronize(h->ev_nl));
      uint64_t* hq = static_cast<uint64_t*>(h->h_ns.p);
      for (size_t j = 0; j < cnt; ++j) {
        hq[j] = first + off + j * G;
        hq[cnt + j] = ns[off + j * G];
      }
      CK(cudaMemcpyAsync(h->q_idx.p, hq, cnt * 8, cudaMemcpyHostToDevice, h->stream));
      CK(cudaMemcpyAsync(h->nsearch.p, hq + cnt, cnt * 8, cudaMemcpyHostToDevice, h->stream));
      CK(cudaEventRecord(h->ev_nl, h->stream));
      k_gather<<<(unsigned)cnt, 128, 0, h->stream>>>(h->records.as<unsigned char>(), h->L, h->db, h->peers, 0, h->q_idx.as<unsigned long long>());
      h->launches++;
      CK(cudaGetLastError());
      PushList pl{};
      if (h->cfg.shard_rank != 0) pl.dst[pl.n++] = h->results_of[0];
      void* rbase = result_base(h);
      RET(query_stage(h, h->records.p, cnt, h->nsearch.as<uint64_t>(), h->stream, h->rb.loop(rbase), h->rb.yaw(rbase), h->rb.dist(rbase),
                      h->rb.idx(rbase), h->rb.shift(rbase), (unsigned)off, (unsigned)G, &pl));
      h->last_nq = cnt;
      h->last_nsearch.assign(hq + cnt, hq + 2 * cnt);
      h->last_mutation = h->mutation;
    }
    CK(cudaEventRecord(h->ev_t2, h->stream));
    CK(cudaEventRecord(h->ev_qdone, h->stream));
    h->timing_valid = true;
    h->replay_nq = nq;
    h->replay_ipc = false;
  }
  std::vector<scgpu_handle*> hs(g->shards);
  for (scgpu_handle* h : hs) {
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
  }
  return replay_fetch(hs, nq, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

int scgpu_detect(scgpu_handle* h, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift) {
  SCGPU_TRACE_CALL();
  if (!h || !loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->is_group && h->cfg.shard_count != 1)
    return fail(SCGPU_E_INVALID, "scgpu_detect needs the whole database behind the handle (one device, or a device list); use the staged API for shards");
  if (h->n_global == 0) return fail(SCGPU_E_EMPTY, "detect on an empty database");
  uint64_t ns;
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) plan(s, h->n_global, 1, &ns);  // every shard carries the same snapshot state
  } else {
    CK(cudaSetDevice(h->cfg.device));
    plan(h, h->n_global, 1, &ns);
  }
  if (ns == 0) {  // SC.cpp:257-261
    *loop_id = -1;
    *yaw = 0.0f;
    if (nearest_dist) *nearest_dist = 10000000.0;
    if (nearest_idx) *nearest_idx = 0;
    if (nearest_shift) *nearest_shift = 0;
    h->last_nq = 0;
    for (scgpu_handle* s : h->shards) s->last_nq = 0;
    return SCGPU_OK;
  }
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) s->last_nq = 0;
    return group_query_range(h, h->n_global - 1, 1, &ns, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
  }
  RET(join_replay(h));
  k_gather<<<1, 128, 0, h->stream>>>(h->rec_single.as<unsigned char>(), h->L, h->db, h->peers, h->n_global - 1, nullptr);
  h->launches++;
  CK(cudaGetLastError());
  RET(run_pipeline(h, h->rec_single.p, 1, &ns));
  return fetch_results(h, 1, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

// the shards (and their first scans) of a replay over host / device scans laid out contiguously
static void replay_shards(scgpu_handle* h, const void* pts, size_t scan_bytes, uint64_t first, std::vector<ReplayShard>& sh) {
  if (!h->is_group) {
    sh.push_back({h, pts, 0});
    return;
  }
  const uint64_t G = h->shards.size();
  for (scgpu_handle* s : h->shards) {
    const size_t off = (size_t)(((uint64_t)s->cfg.shard_rank + G - first % G) % G);
    sh.push_back({s, static_cast<const unsigned char*>(pts) + off * scan_bytes, G * scan_bytes});
  }
}

int scgpu_replay_async(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location) {
  SCGPU_TRACE_CALL();
  if (!h || (!pts && n_scans && pts_per_scan)) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->is_group && h->cfg.shard_count != 1)
    return fail(SCGPU_E_INVALID, "scgpu_replay_* needs the whole database behind the handle; shards of other processes use scgpu_peer_replay_async");
  if (n_scans == 0) return SCGPU_OK;
  if (n_scans > 65535) return fail(SCGPU_E_INVALID, "at most 65535 scans per call");
  const uint64_t first = h->n_global;
  std::vector<ReplayShard> sh;
  replay_shards(h, pts, pts_per_scan * stride, first, sh);
  RET(replay_enqueue(sh, first, n_scans, pts_per_scan, stride, location, false));
  if (h->is_group) group_set_size(h, first + n_scans);
  h->replay_nq = n_scans;
  return SCGPU_OK;
}

int scgpu_replay_results(scgpu_handle* h, size_t n, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift) {
  SCGPU_TRACE_CALL();
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  std::vector<scgpu_handle*> hs;
  if (h->is_group) hs = h->shards;
  else hs.push_back(h);
  return replay_fetch(hs, n, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

int scgpu_replay_batched(scgpu_handle* h, const void* pts, size_t n_scans, size_t pts_per_scan, size_t stride, int location,
                         int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift) {
  if (!loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  RET(scgpu_replay_async(h, pts, n_scans, pts_per_scan, stride, location));
  if (n_scans == 0) return SCGPU_OK;
  return scgpu_replay_results(h, n_scans, loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
}

int scgpu_query_batched(scgpu_handle* h, uint64_t first, size_t nq, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx,
                        int* nearest_shift) {
  if (!h || !loop_id || !yaw) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->is_group && h->cfg.shard_count != 1)
    return fail(SCGPU_E_INVALID, "scgpu_query_batched needs the whole database behind the handle; use the staged API for shards");
  if (nq == 0) return SCGPU_OK;
  if (first + nq > h->n_global || nq > 65535) return fail(SCGPU_E_INVALID, "query range outside the database");
  std::vector<uint64_t> ns(nq);
  const uint64_t excl = (uint64_t)h->cfg.exclude_recent;
  for (size_t i = 0; i < nq; ++i) {
    const uint64_t size = first + i + 1;
    ns[i] = size >= excl + 1 ? size - excl : 0;
  }
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) s->last_nq = 0;
    return group_query_range(h, first, nq, ns.data(), loop_id, yaw, nearest_dist, nearest_idx, nearest_shift);
  }
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
  RET(h->records.reserve(nq * h->L.rec_bytes));
  CK(cudaEventRecord(h->ev_t0, h->stream));
  CK(cudaEventRecord(h->ev_t1, h->stream));
  k_gather<<<(unsigned)nq, 128, 0, h->stream>>>(h->records.as<unsigned char>(), h->L, h->db, h->peers, first, nullptr);
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
  if (h->is_group) {  // the last detect ran on one shard: that shard holds the candidate table
    for (scgpu_handle* s : h->shards)
      if (s->last_nq) return scgpu_get_batch_candidates(s, q, cand_idx, cand_d2, cand_dist, cand_shift, n_search);
    return fail(SCGPU_E_INVALID, "no such query in the last call");
  }
  if (q >= h->last_nq) return fail(SCGPU_E_INVALID, "no such query in the last call");
  // the dump re-reads the query records and the candidate table of the last pipeline: anything that ran since then may have
  // overwritten them
  if (h->last_mutation != h->mutation) return fail(SCGPU_E_INVALID, "the database or the query workspace changed since the last detect / replay / query call");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventSynchronize(h->ev_qdone));
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
  if (h->is_group) return scgpu_get_entry(h->shards[i % h->shards.size()], i, sc, ring, sector);
  const int owner = (int)(i % (uint64_t)h->cfg.shard_count);
  const bool remote = owner != h->cfg.shard_rank;
  // an attached peer shard reads other shards' entries through its mapping of their memory (the caller makes sure the owner
  // has finished writing: scgpu_replay_results / a barrier)
  if (remote && !(h->peer && h->attached)) return fail(SCGPU_E_INVALID, "entry lives on another shard");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaStreamSynchronize(h->qstream));
  const uint64_t l = i / (uint64_t)h->cfg.shard_count;
  if (remote) {
    if (sc) CK(cudaMemcpy(sc, h->peers.sc[owner] + l * h->L.RS, sizeof(float) * h->L.RS, cudaMemcpyDeviceToHost));
    if (ring)
      CK(cudaMemcpy2D(ring, sizeof(float), h->db.ringT + ring_at(i, 0, h->L.R), 32 * sizeof(float), sizeof(float), h->L.R, cudaMemcpyDeviceToHost));
    if (sector) CK(cudaMemcpy(sector, h->peers.sector[owner] + l * h->L.S, sizeof(double) * h->L.S, cudaMemcpyDeviceToHost));
    return SCGPU_OK;
  }
  if (sc) CK(cudaMemcpy(sc, h->db.sc + l * h->L.RS, sizeof(float) * h->L.RS, cudaMemcpyDeviceToHost));
  if (ring)
    CK(cudaMemcpy2D(ring, sizeof(float), h->db.ringT + ring_at(ring_slot(h->db, i, l), 0, h->L.R), 32 * sizeof(float), sizeof(float), h->L.R,
                    cudaMemcpyDeviceToHost));
  if (sector) CK(cudaMemcpy(sector, h->db.sector + l * h->L.S, sizeof(double) * h->L.S, cudaMemcpyDeviceToHost));
  return SCGPU_OK;
}

int scgpu_truncate(scgpu_handle* h, uint64_t n) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) RET(scgpu_truncate(s, n));
    if (n < h->n_global) h->n_global = n;
    return SCGPU_OK;
  }
  h->mutation++;
  if (n < h->n_global) h->n_global = n;
  if (h->x_upto > local_count(h, h->n_global)) h->x_upto = local_count(h, h->n_global);
  // the snapshot may reference forgotten entries: the next detect takes a fresh one (counter % period == 0)
  h->counter = 0;
  h->n_tree = 0;
  return SCGPU_OK;
}

int scgpu_plan_n_search(scgpu_handle* h, uint64_t first_size, size_t n, uint64_t* out) {
  if (!h || (!out && n)) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) plan(s, first_size, n, out);
    return SCGPU_OK;
  }
  plan(h, first_size, n, out);
  return SCGPU_OK;
}

// ---- exhaustive search: shard-local core + reduction over the shards behind the handle -------------------------------
struct ExhWin {
  double dist = 10000000.0;
  int shift = 0;
  int64_t idx = 0;
  int flip = 0;
  bool found = false;
};
// the loop of SC.cpp:296-311 over every entry: strict-min in (index, forward-before-flipped) order
static void exh_take(ExhWin& w, double d, int shift, int64_t idx, int flip) {
  if (!(d < 10000000.0)) return;
  if (!w.found || d < w.dist || (d == w.dist && (idx < w.idx || (idx == w.idx && flip < w.flip)))) {
    w.dist = d;
    w.shift = shift;
    w.idx = idx;
    w.flip = flip;
    w.found = true;
  }
}

// Every local entry with global index < n_search scored by the exact FP64 pair kernel (no screening); query record on the device.
static int exhaustive_exact_local(scgpu_handle* h, const unsigned char* d_qrec, uint64_t n_search, int flipped, ExhWin* out) {
  cudaStream_t st = h->stream;
  const size_t n = (size_t)local_count(h, n_search);
  *out = ExhWin();
  if (n == 0) return SCGPU_OK;
  const uint64_t G = (uint64_t)h->cfg.shard_count, r = (uint64_t)h->cfg.shard_rank;
  DevBuf keys, pd, ps, ns;
  RET(keys.reserve(n * 8));
  RET(pd.reserve(n * 8 * 2));
  RET(ps.reserve(n * 4 * 2));
  RET(ns.reserve(8));
  std::vector<uint64_t> hk(n);
  for (size_t i = 0; i < n; ++i) hk[i] = i * G + r;
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
      p.qrecords = d_qrec;
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
      p.peers = PeerTab{};
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
  for (size_t i = 0; i < n; ++i)
    for (int f = 0; f <= (flipped ? 1 : 0); ++f) exh_take(*out, hd[f * n + i], hs[f * n + i], (int64_t)hk[i], f);
  return SCGPU_OK;
}

// phase 1 (asynchronous): gather the query records of stored entries q[0..nq) (from whichever shard holds them), screen +
// rescore this shard's entries; per-query Best -> h->x_best
static int exh_enqueue(scgpu_handle* h, scgpu_handle* group, const uint64_t* q, const uint64_t* n_search, size_t nq, int flipped) {
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
  cudaStream_t st = h->stream;
  if (group) RET(group_wait_appends(group, h, st));
  RET(h->records.reserve(nq * h->L.rec_bytes));
  RET(h->x_best.reserve(nq * sizeof(Best)));
  RET(h->q_idx.reserve(nq * sizeof(uint64_t)));
  CK(cudaMemcpyAsync(h->q_idx.p, q, nq * 8, cudaMemcpyHostToDevice, st));  // (pageable source: staged by the runtime before it returns)
  CK(cudaEventRecord(h->ev_t0, st));
  k_gather<<<(unsigned)nq, 128, 0, st>>>(h->records.as<unsigned char>(), h->L, h->db, h->peers, 0, h->q_idx.as<unsigned long long>());
  h->launches++;
  CK(cudaGetLastError());
  if (h->exh) {
    const size_t per = flipped ? EXH_MAX_BATCH / 2 : EXH_MAX_BATCH;
    for (size_t i0 = 0; i0 < nq; i0 += per) {
      const size_t m = nq - i0 < per ? nq - i0 : per;
      RET(launch_exhaustive_fast(h, h->records.as<unsigned char>() + i0 * h->L.rec_bytes, m, n_search + i0, h->x_best.as<Best>() + i0, st,
                                 nq == 1 ? h->ev_t0 : nullptr, nq == 1 ? h->ev_t1 : nullptr, flipped));
    }
  }
  if (nq != 1 || !h->exh) CK(cudaEventRecord(h->ev_t1, st));
  CK(cudaEventRecord(h->ev_t2, st));
  h->timing_valid = true;
  return SCGPU_OK;
}

// phase 2: wait, read the shard's winners; a query whose rescoring list overflowed (more near-ties than EXH_CAND_CAP, e.g. a
// database of duplicates -- the list is shared by the queries of a batch) is redone exactly over every local entry
static int exh_collect(scgpu_handle* h, const uint64_t* n_search, size_t nq, int flipped, std::vector<ExhWin>& out) {
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  out.assign(nq, ExhWin());
  std::vector<Best> b(nq);
  if (h->exh) CK(cudaMemcpyAsync(b.data(), h->x_best.p, nq * sizeof(Best), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->last_exh_rescored = 0;
  for (size_t i = 0; i < nq; ++i) {
    if (local_count(h, n_search[i]) == 0) continue;
    if (h->exh && (unsigned)b[i].rank <= EXH_CAND_CAP) {
      if (b[i].dist < 10000000.0) exh_take(out[i], b[i].dist, b[i].shift & 0x3fffffff, b[i].idx, (b[i].shift >> 30) & 1);
      h->last_exh_rescored = (unsigned)b[i].rank;
      continue;
    }
    RET(exhaustive_exact_local(h, h->records.as<unsigned char>() + i * h->L.rec_bytes, n_search[i], flipped, &out[i]));
    h->last_exh_rescored = (unsigned)local_count(h, n_search[i]);
  }
  return SCGPU_OK;
}

static int exhaustive_any(scgpu_handle* h, const uint64_t* q, const uint64_t* n_search, size_t nq, int flipped, std::vector<ExhWin>& win) {
  if (!h->is_group && h->cfg.shard_count != 1)
    return fail(SCGPU_E_INVALID, "scgpu_exhaustive* needs the whole database behind the handle; use scgpu_stage_exhaustive for shards");
  for (size_t i = 0; i < nq; ++i)
    if (q[i] >= h->n_global || n_search[i] > h->n_global) return fail(SCGPU_E_INVALID, "range outside the database");
  win.assign(nq, ExhWin());
  if (nq == 0) return SCGPU_OK;
  std::vector<scgpu_handle*> hs;
  if (h->is_group) hs = h->shards;
  else hs.push_back(h);
  for (scgpu_handle* s : hs) RET(exh_enqueue(s, h->is_group ? h : nullptr, q, n_search, nq, flipped));
  std::vector<ExhWin> part;
  for (scgpu_handle* s : hs) {
    RET(exh_collect(s, n_search, nq, flipped, part));
    for (size_t i = 0; i < nq; ++i)
      if (part[i].found) exh_take(win[i], part[i].dist, part[i].shift, part[i].idx, part[i].flip);
  }
  return SCGPU_OK;
}

int scgpu_exhaustive(scgpu_handle* h, uint64_t q, uint64_t n_search, int flipped, double* best_dist, int* best_shift, int64_t* best_idx,
                     int* best_flip) {
  if (!h || !best_dist || !best_shift || !best_idx) return fail(SCGPU_E_INVALID, "null argument");
  std::vector<ExhWin> w;
  RET(exhaustive_any(h, &q, &n_search, 1, flipped, w));
  *best_dist = w[0].dist;
  *best_shift = w[0].shift;
  *best_idx = w[0].idx;
  if (best_flip) *best_flip = w[0].flip;
  return SCGPU_OK;
}

int scgpu_exhaustive_batched(scgpu_handle* h, const uint64_t* q, const uint64_t* n_search, size_t nq, double* best_dist, int* best_shift,
                             int64_t* best_idx) {
  SCGPU_TRACE_CALL();
  if (!h || ((!q || !n_search || !best_dist || !best_shift || !best_idx) && nq)) return fail(SCGPU_E_INVALID, "null argument");
  std::vector<ExhWin> w;
  RET(exhaustive_any(h, q, n_search, nq, 0, w));
  for (size_t i = 0; i < nq; ++i) {
    best_dist[i] = w[i].dist;
    best_shift[i] = w[i].shift;
    best_idx[i] = w[i].idx;
  }
  return SCGPU_OK;
}

int scgpu_probe_screen(scgpu_handle* h, uint64_t q, uint64_t n_search, float* d32, uint32_t* shift) {
  if (!h || !d32) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group || h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "the screening probe runs on a single-device handle");
  if (!h->exh) return fail(SCGPU_E_INVALID, "no screening kernel is instantiated for this configuration");
  if (q >= h->n_global || n_search > h->n_global) return fail(SCGPU_E_INVALID, "range outside the database");
  if (n_search == 0) return SCGPU_OK;
  if (shift && h->exh_cfg != 3) return fail(SCGPU_E_INVALID, "per-entry shifts are reported by the full-shift configuration only");
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  h->mutation++;
  cudaStream_t st = h->stream;
  k_gather<<<1, 128, 0, st>>>(h->rec_single.as<unsigned char>(), h->L, h->db, h->peers, q, nullptr);
  h->launches++;
  RET(h->x_best.reserve(sizeof(Best)));
  h->tc_want_shifts = shift != nullptr;
  const int rc = launch_exhaustive_fast(h, h->rec_single.as<unsigned char>(), 1, &n_search, h->x_best.as<Best>(), st, nullptr, nullptr, 0);
  h->tc_want_shifts = false;
  RET(rc);
  CK(cudaMemcpyAsync(d32, h->x_d32.p, n_search * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (shift) CK(cudaMemcpyAsync(shift, h->tc_shift.p, n_search * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SCGPU_OK;
}

int scgpu_exhaustive_stats(scgpu_handle* h, uint64_t* rescored) {
  if (!h || !rescored) return fail(SCGPU_E_INVALID, "null argument");
  *rescored = h->last_exh_rescored;
  for (scgpu_handle* s : h->shards) *rescored += s->last_exh_rescored;
  return SCGPU_OK;
}

// Flat binary database file: header | float descriptors (column-major R*S each), entry 0 first.
struct SaveHeader {
  char magic[8];  // "SCGPUDB2"
  uint64_t R, S, n;
  uint64_t counter, n_tree;  // tree-snapshot state (SC.cpp:264-276): detects after a load continue the saved run's sequence
  double lidar_height, max_radius;  // the constants that shaped the stored descriptors: a handle with others must not load them
};

int scgpu_save(scgpu_handle* h, const char* path) {
  if (!h || !path) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->is_group && h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "save needs the whole database behind the handle");
  std::vector<scgpu_handle*> hs;
  if (h->is_group) hs = h->shards;
  else hs.push_back(h);
  const uint64_t n = h->n_global, G = hs.size();
  const size_t RS = (size_t)h->L.RS;
  try {
    std::vector<std::vector<float>> part(G);
    for (uint64_t r = 0; r < G; ++r) {
      scgpu_handle* s = hs[r];
      CK(cudaSetDevice(s->cfg.device));
      CK(cudaStreamSynchronize(s->stream));
      CK(cudaStreamSynchronize(s->qstream));
      part[r].resize((size_t)local_count(s, n) * RS);
      if (!part[r].empty()) CK(cudaMemcpy(part[r].data(), s->db.sc, part[r].size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SCGPU_E_IO, "cannot open %s for writing", path);
    SaveHeader hd{};
    memcpy(hd.magic, "SCGPUDB2", 8);
    hd.R = (uint64_t)h->L.R;
    hd.S = (uint64_t)h->L.S;
    hd.n = n;
    hd.counter = (uint64_t)hs[0]->counter;
    hd.n_tree = hs[0]->n_tree;
    hd.lidar_height = h->cfg.lidar_height;
    hd.max_radius = h->cfg.max_radius;
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1;
    if (G == 1) {
      ok = ok && (part[0].empty() || fwrite(part[0].data(), sizeof(float), part[0].size(), f) == part[0].size());
    } else {
      for (uint64_t i = 0; i < n && ok; ++i) ok = fwrite(part[i % G].data() + (size_t)(i / G) * RS, sizeof(float), RS, f) == RS;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? SCGPU_OK : fail(SCGPU_E_IO, "short write to %s", path);
  } catch (const std::exception& e) {
    return fail(SCGPU_E_IO, "save: %s", e.what());
  }
}

int scgpu_load(scgpu_handle* h, const char* path) {
  if (!h || !path) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->is_group && h->cfg.shard_count != 1) return fail(SCGPU_E_INVALID, "load needs the whole database behind the handle");
  if (h->n_global != 0) return fail(SCGPU_E_INVALID, "load replaces the database: the handle must be empty (scgpu_truncate(h, 0) first)");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(SCGPU_E_IO, "cannot open %s", path);
  SaveHeader hd{};
  if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "SCGPUDB2", 8) != 0) {
    fclose(f);
    return fail(SCGPU_E_IO, "%s is not a scgpu database", path);
  }
  if ((int)hd.R != h->L.R || (int)hd.S != h->L.S || hd.lidar_height != h->cfg.lidar_height || hd.max_radius != h->cfg.max_radius) {
    fclose(f);
    return fail(SCGPU_E_INVALID, "database was built with %llux%llu, lidar height %g, radius %g; the handle has %dx%d, %g, %g",
                (unsigned long long)hd.R, (unsigned long long)hd.S, hd.lidar_height, hd.max_radius, h->L.R, h->L.S, h->cfg.lidar_height,
                h->cfg.max_radius);
  }
  // the entry count comes from an untrusted file: check it against the file's size and the 32-bit index space before sizing anything
  const size_t RS = (size_t)h->L.RS;
  long here = ftell(f);
  fseek(f, 0, SEEK_END);
  const long end = ftell(f);
  fseek(f, here, SEEK_SET);
  if (hd.n > 0xffffffffull || here < 0 || end < here || (uint64_t)(end - here) != hd.n * RS * sizeof(float)) {
    fclose(f);
    return fail(SCGPU_E_IO, "%s: header says %llu entries, the file holds %lld bytes of descriptors", path, (unsigned long long)hd.n,
                (long long)(end - here));
  }
  int rc = SCGPU_OK;
  try {
    const size_t batch = 8192;
    std::vector<float> buf(std::min<size_t>(batch, (size_t)hd.n) * RS);
    for (uint64_t i0 = 0; i0 < hd.n && rc == SCGPU_OK; i0 += batch) {
      const size_t m = (size_t)std::min<uint64_t>(batch, hd.n - i0);
      if (fread(buf.data(), sizeof(float), m * RS, f) != m * RS) rc = fail(SCGPU_E_IO, "short read from %s", path);
      else rc = scgpu_append_descs(h, buf.data(), m);
    }
  } catch (const std::exception& e) {
    rc = fail(SCGPU_E_IO, "load: %s", e.what());
  }
  fclose(f);
  if (rc != SCGPU_OK) return rc;
  std::vector<scgpu_handle*> hs;
  if (h->is_group) hs = h->shards;
  else hs.push_back(h);
  for (scgpu_handle* s : hs) {
    s->counter = (long long)hd.counter;
    s->n_tree = hd.n_tree;
  }
  return SCGPU_OK;
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
  h = GROUP_FIRST(h);
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
  h = GROUP_FIRST(h);
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
  // one small device allocation per (host thread, device), made on the first call and kept: in / bin / height / theta cells
  struct Cells {
    int dev = -1;
    float* p = nullptr;
  };
  static thread_local Cells cells;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (cells.dev != dev) {
    cells.p = nullptr;  // (an allocation on another device is left to the driver's teardown)
    CK(cudaMalloc(&cells.p, 8 * sizeof(float)));
    cells.dev = dev;
  }
  float xyz[3] = {x, y, 0.f};
  CK(cudaMemcpy(cells.p, xyz, 12, cudaMemcpyHostToDevice));
  const BinConst bc = make_bin_const(20, 60, 2.0, 80.0, 0);
  k_probe_bins<<<1, 32>>>(cells.p, 1, bc, reinterpret_cast<int*>(cells.p + 4), cells.p + 5, cells.p + 6);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out_deg, cells.p + 6, 4, cudaMemcpyDeviceToHost);
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
  if (h->is_group) {
    for (scgpu_handle* s : h->shards) RET(scgpu_set_downsample_leaf(s, leaf));
    return SCGPU_OK;
  }
  if (!(leaf >= 0.f) || !(leaf < 1e30f)) return fail(SCGPU_E_INVALID, "leaf size must be >= 0 (0 = no downsampling) and finite");
  if (leaf > 0.f && vox_smem_bytes(h->L.RS) > 227 * 1024) return fail(SCGPU_E_INVALID, "descriptor too large for the voxel kernel's shared memory");
  h->voxel_leaf = leaf;
  return SCGPU_OK;
}

int scgpu_voxel_downsample(scgpu_handle* h, const void* pts, size_t n, size_t stride_bytes, float leaf, float* out_xyzn, uint32_t* out_idx,
                           size_t cap, size_t* out_n, int32_t* min_b, int32_t* div_b, int32_t* status) {
  if (!h || !out_n || (n && !pts)) return fail(SCGPU_E_INVALID, "null argument");
  if (stride_bytes < 12 || (stride_bytes & 3)) return fail(SCGPU_E_INVALID, "stride must be >= 12 and a multiple of 4");
  h = GROUP_FIRST(h);
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
  if (h->is_group) return fail(SCGPU_E_INVALID, "the staged API addresses one shard, not a device-list handle");
  // only slots that have been written may become searchable
  if (local_count(h, n_global) > h->n_written) return fail(SCGPU_E_INVALID, "size beyond this shard's stored entries");
  if (n_global < h->n_global) {  // shrinking = scgpu_truncate: screening copies above the cut and the snapshot state are void
    const uint64_t keep = local_count(h, n_global);
    if (h->x_upto > keep) h->x_upto = keep;
    h->counter = 0;
    h->n_tree = 0;
  }
  h->mutation++;
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
  if (h->exh && h->exh_cfg != 3 && !(h->cfg.flags & SCGPU_FLAG_NO_SCREENING)) RET(launch_score_screened(h, d_qrec, nq, d_keys, d_ns, st));
  else RET(launch_score(h, d_qrec, nq, d_keys, d_ns, h->K, h->pair_dist.as<double>(), h->pair_shift.as<int>(), 0, st));
  return launch_best(h, nq, d_keys, static_cast<Best*>(d_best_out), st);
}

int scgpu_stage_gather(scgpu_handle* h, uint64_t global_idx, void* d_record, void* stream) {
  if (!h || !d_record) return fail(SCGPU_E_INVALID, "null argument");
  if (global_idx >= h->n_global) return fail(SCGPU_E_INVALID, "entry out of range");
  if (!(h->peer && h->attached) && (int)(global_idx % (uint64_t)h->cfg.shard_count) != h->cfg.shard_rank)
    return fail(SCGPU_E_INVALID, "entry lives on another shard");
  CK(cudaSetDevice(h->cfg.device));
  k_gather<<<1, 128, 0, ST(stream)>>>(static_cast<unsigned char*>(d_record), h->L, h->db, h->peers, global_idx, nullptr);
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

int scgpu_stage_exhaustive2(scgpu_handle* h, const void* d_qrecords, size_t nq, const uint64_t* n_search, int flipped, void* d_best_out,
                            void* stream) {
  if (!h || !d_qrecords || !d_best_out || (!n_search && nq)) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) return fail(SCGPU_E_INVALID, "the staged API addresses one shard, not a device-list handle");
  if (!h->exh) return fail(SCGPU_E_INVALID, "screening kernels are instantiated for 20x60 (radius 3) and 40x120 (radius 6) only");
  CK(cudaSetDevice(h->cfg.device));
  const size_t per = flipped ? EXH_MAX_BATCH / 2 : EXH_MAX_BATCH;
  for (size_t i0 = 0; i0 < nq; i0 += per) {
    const size_t m = nq - i0 < per ? nq - i0 : per;
    RET(launch_exhaustive_fast(h, static_cast<const unsigned char*>(d_qrecords) + i0 * h->L.rec_bytes, m, n_search + i0,
                               static_cast<Best*>(d_best_out) + i0, ST(stream), nullptr, nullptr, flipped));
  }
  return SCGPU_OK;
}

int scgpu_stage_exhaustive(scgpu_handle* h, const void* d_qrecords, size_t nq, const uint64_t* n_search, void* d_best_out, void* stream) {
  return scgpu_stage_exhaustive2(h, d_qrecords, nq, n_search, 0, d_best_out, stream);
}

int scgpu_stage_finalize(scgpu_handle* h, const void* d_best_parts, int parts, size_t nq, const uint64_t* d_ns, int32_t* d_loop_id,
                         float* d_yaw, double* d_nearest_dist, int32_t* d_nearest_idx, int32_t* d_nearest_shift, void* stream) {
  if (!h || !d_best_parts || !d_ns || !d_loop_id || !d_yaw || parts < 1) return fail(SCGPU_E_INVALID, "bad argument");
  CK(cudaSetDevice(h->cfg.device));
  return launch_finalize(h, static_cast<const Best*>(d_best_parts), parts, nq, d_ns, d_loop_id, d_yaw, d_nearest_dist, d_nearest_idx,
                         d_nearest_shift, ST(stream));
}

// The exhaustive search of this shard for ONE query record, every local entry scored exactly (no screening): what a caller of
// scgpu_stage_exhaustive falls back to when a result reports more rescored candidates than the list holds (n_rescored >
// SCGPU_EXH_LIST_CAP).  Synchronises the handle's own stream; d_query_record must be complete when called.
int scgpu_stage_exhaustive_exact(scgpu_handle* h, const void* d_query_record, uint64_t n_search, void* d_best_out) {
  if (!h || !d_query_record || !d_best_out) return fail(SCGPU_E_INVALID, "null argument");
  if (h->is_group) return fail(SCGPU_E_INVALID, "the staged API addresses one shard, not a device-list handle");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaDeviceSynchronize());
  ExhWin w;
  RET(exhaustive_exact_local(h, static_cast<const unsigned char*>(d_query_record), n_search, 0, &w));
  Best b;
  b.dist = w.found ? w.dist : 10000000.0;
  b.rank = 0;
  b.shift = w.shift;
  b.idx = w.idx;
  CK(cudaMemcpy(d_best_out, &b, sizeof b, cudaMemcpyHostToDevice));
  return SCGPU_OK;
}

// ---- loop verification after the path ---------------------------------------------------------------
int scgpu_default_icp_params(scgpu_icp_params* p) {
  if (!p) return fail(SCGPU_E_INVALID, "null argument");
  memset(p, 0, sizeof *p);
  p->max_iterations = 100;               // icp.setMaximumIterations(100)            mapOpt.cpp:1055
  p->max_correspondence_distance = 100;  // icp.setMaxCorrespondenceDistance(100)    mapOpt.cpp:1054
  p->transformation_epsilon = 1e-6;      // icp.setTransformationEpsilon(1e-6)       mapOpt.cpp:1056
  p->euclidean_fitness_epsilon = 1e-6;   // icp.setEuclideanFitnessEpsilon(1e-6)     mapOpt.cpp:1057
  p->fitness_threshold = 1.5;            // historyKeyframeFitnessScore              utility.h:139
  p->seed_axis = -1;                     // identity initial guess, as the reference runs it (mapOpt.cpp:1066)
  p->seed_angle = 0.f;
  return SCGPU_OK;
}

// The ICP loop on device-resident clouds (mapOptmization.cpp:1053-1078); everything on the handle's stream.
static int icp_run_device(scgpu_handle* h, const unsigned char* d_src, size_t n_src, size_t stride_s, const unsigned char* d_tgt, size_t n_tgt,
                          size_t stride_t, const scgpu_icp_params* prm, double* T16, double* fitness, int* converged, int* iterations, int* accepted) {
  cudaStream_t st = h->stream;
  IcpState s0;
  memset(&s0, 0, sizeof s0);
  for (int k = 0; k < 4; ++k) s0.T[5 * k] = 1.0;
  if (prm->seed_axis >= 0 && prm->seed_axis <= 2) {  // seed with the yaw Scan Context reported, about the caller's vertical axis
    const double c = cos((double)prm->seed_angle), sn = sin((double)prm->seed_angle);
    const int a = (prm->seed_axis + 1) % 3, b = (prm->seed_axis + 2) % 3;
    s0.T[4 * a + a] = c;
    s0.T[4 * a + b] = -sn;
    s0.T[4 * b + a] = sn;
    s0.T[4 * b + b] = c;
  }
  s0.prev_mse = 1.79769313486231570e308;
  s0.fitness = 1.79769313486231570e308;
  IcpState out = s0;
  if (n_src && n_tgt) {
    const unsigned blocks = (unsigned)((n_src + ICP_BLOCK - 1) / ICP_BLOCK);
    RET(h->icp_state.reserve(sizeof(IcpState)));
    RET(h->icp_part.reserve((size_t)blocks * sizeof(IcpPartial)));
    CK(cudaMemcpyAsync(h->icp_state.p, &s0, sizeof s0, cudaMemcpyHostToDevice, st));
    IcpCriteria crit;
    crit.max_iterations = prm->max_iterations;
    crit.translation_threshold = prm->transformation_epsilon;
    crit.rotation_threshold = 0.99999;
    crit.mse_abs = prm->euclidean_fitness_epsilon;
    crit.mse_rel = 0.00001;
    const double md = prm->max_correspondence_distance;
    const float max_d2 = (float)(md * md);
    IcpState* d_st = h->icp_state.as<IcpState>();
    for (int it = 0; it < prm->max_iterations; ++it) {
      k_icp_nn<<<blocks, ICP_BLOCK, 0, st>>>(d_src, (unsigned)n_src, (unsigned)stride_s, d_tgt, (unsigned)n_tgt, (unsigned)stride_t, d_st, max_d2, 0,
                                             h->icp_part.as<IcpPartial>());
      k_icp_solve<<<1, 32, 0, st>>>(h->icp_part.as<IcpPartial>(), blocks, d_st, crit, 0);
      h->launches += 2;
      if ((it & 7) == 7) {  // peek at the flag now and then so that an early convergence does not cost the full launch list
        int done = 0;
        CK(cudaMemcpyAsync(&done, &d_st->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (done) break;
      }
    }
    k_icp_nn<<<blocks, ICP_BLOCK, 0, st>>>(d_src, (unsigned)n_src, (unsigned)stride_s, d_tgt, (unsigned)n_tgt, (unsigned)stride_t, d_st, max_d2, 1,
                                           h->icp_part.as<IcpPartial>());
    k_icp_solve<<<1, 32, 0, st>>>(h->icp_part.as<IcpPartial>(), blocks, d_st, crit, 1);
    h->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&out, d_st, sizeof out, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (T16) memcpy(T16, out.T, sizeof out.T);
  if (fitness) *fitness = out.fitness;
  if (converged) *converged = out.converged;
  if (iterations) *iterations = out.iterations;
  if (accepted) *accepted = out.converged && out.fitness <= prm->fitness_threshold;  // mapOpt.cpp:1068
  return SCGPU_OK;
}

int scgpu_verify_loop(scgpu_handle* h, const void* src, size_t n_src, const void* tgt, size_t n_tgt, size_t stride, const scgpu_icp_params* prm,
                      double* T16, double* fitness, int* converged, int* iterations, int* accepted) {
  if (!h || !prm || (!src && n_src) || (!tgt && n_tgt)) return fail(SCGPU_E_INVALID, "null argument");
  if (stride < 12 || (stride & 3)) return fail(SCGPU_E_INVALID, "stride must be >= 12 and a multiple of 4");
  if (n_src > 0x7fffffffull || n_tgt > 0x7fffffffull || prm->max_iterations < 1) return fail(SCGPU_E_INVALID, "bad cloud size / iteration cap");
  h = GROUP_FIRST(h);
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  if (n_src && n_tgt) {
    RET(h->icp_src.reserve(n_src * stride));
    RET(h->icp_tgt.reserve(n_tgt * stride));
    CK(cudaMemcpyAsync(h->icp_src.p, src, n_src * stride, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->icp_tgt.p, tgt, n_tgt * stride, cudaMemcpyHostToDevice, st));
  }
  return icp_run_device(h, h->icp_src.as<unsigned char>(), n_src, stride, h->icp_tgt.as<unsigned char>(), n_tgt, stride, prm, T16, fitness, converged,
                        iterations, accepted);
}

// Submap assembly (mapOptmization.cpp:928-949) on the device: the clouds -> staged input -> k_submap_transform[_filtered] -> `out`
// (float4 x, y, z, intensity; cloud after cloud).  *n_out = points written (known on the host unless `filter`).
static int submap_to_device(scgpu_handle* h, const void* const* clouds, const size_t* n_points, const scgpu_pose6* poses, size_t n_clouds,
                            bool one_pose, size_t stride, size_t intensity_off, bool filter, DevBuf& stage, DevBuf& out, size_t* n_out) {
  cudaStream_t st = h->stream;
  size_t total = 0;
  for (size_t i = 0; i < n_clouds; ++i) {
    if (n_points[i] && !clouds[i]) return fail(SCGPU_E_INVALID, "null cloud");
    total += n_points[i];
  }
  *n_out = 0;
  if (total == 0) return SCGPU_OK;
  if (total > 0x7fffffffull) return fail(SCGPU_E_INVALID, "submap too large");
  std::vector<SubmapCloud> desc(n_clouds);
  size_t off = 0, pts = 0;
  for (size_t i = 0; i < n_clouds; ++i) {
    const scgpu_pose6& ps = poses[one_pose ? 0 : i];
    SubmapCloud& c = desc[i];
    c.in_off = off;
    c.n = (unsigned)n_points[i];
    c.out_off = (unsigned)pts;
    c.cy = cosf(ps.yaw), c.sy = sinf(ps.yaw);      // cos(float) / sin(float) of mapOptmization.cpp:611-621 are the float overloads
    c.cr = cosf(ps.roll), c.sr = sinf(ps.roll);
    c.cp = cosf(ps.pitch), c.sp = sinf(ps.pitch);
    c.tx = ps.x, c.ty = ps.y, c.tz = ps.z;
    off += (n_points[i] * stride + 15) & ~(size_t)15;
    pts += n_points[i];
  }
  RET(stage.reserve(off + n_clouds * sizeof(SubmapCloud) + 64));
  RET(out.reserve(total * sizeof(float4) + 16));
  RET(h->x_small.reserve(256));
  unsigned char* d_in = stage.as<unsigned char>();
  for (size_t i = 0; i < n_clouds; ++i)
    if (n_points[i]) CK(cudaMemcpyAsync(d_in + desc[i].in_off, clouds[i], n_points[i] * stride, cudaMemcpyHostToDevice, st));
  SubmapCloud* d_desc = reinterpret_cast<SubmapCloud*>(d_in + ((off + 15) & ~(size_t)15));
  CK(cudaMemcpyAsync(d_desc, desc.data(), n_clouds * sizeof(SubmapCloud), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // `desc` (pageable) has been consumed
  if (!filter) {
    size_t big = 0;
    for (size_t i = 0; i < n_clouds; ++i) big = std::max(big, n_points[i]);
    const unsigned gx = (unsigned)std::min<size_t>((big + 255) / 256, 1024);
    k_submap_transform<<<dim3(gx ? gx : 1, (unsigned)n_clouds), 256, 0, st>>>(d_in, (unsigned)stride, (unsigned)intensity_off, d_desc, out.as<float4>());
    *n_out = total;
  } else {
    unsigned* d_n = reinterpret_cast<unsigned*>(h->x_small.as<unsigned char>() + 128);
    k_submap_transform_filtered<<<1, 1024, 0, st>>>(d_in, (unsigned)stride, (unsigned)intensity_off, d_desc, (unsigned)n_clouds, out.as<float4>(), d_n);
    unsigned n = 0;
    CK(cudaMemcpyAsync(&n, d_n, sizeof n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_out = n;
  }
  h->launches++;
  CK(cudaGetLastError());
  return SCGPU_OK;
}

// voxel grid over device-resident float4 points (the history side, mapOptmization.cpp:948-949); result in h->vox_pts, *n_out voxels
static int submap_voxel(scgpu_handle* h, const void* d_pts, size_t n, float leaf, size_t* n_out) {
  cudaStream_t st = h->stream;
  const unsigned ocap = (unsigned)(n ? n : 1);
  RET(h->vox_info.reserve(sizeof(VoxInfo)));
  RET(h->vox_pts.reserve((size_t)ocap * sizeof(float4)));
  RET(h->vox_idx.reserve((size_t)ocap * sizeof(unsigned)));
  CK(cudaMemsetAsync(h->vox_info.p, 0, sizeof(VoxInfo), st));
  VoxInfo vi;
  memset(&vi, 0, sizeof vi);
  if (n) {
    RET(launch_build_voxel(h, d_pts, 1, n, sizeof(float4), leaf, nullptr, h->vox_info.as<VoxInfo>(), h->vox_pts.as<float4>(), h->vox_idx.as<unsigned>(),
                           ocap, st));
    CK(cudaMemcpyAsync(&vi, h->vox_info.p, sizeof vi, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  *n_out = vi.n_out;
  return SCGPU_OK;
}

static int submap_check(scgpu_handle* h, size_t stride, size_t intensity_off) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (stride < 12 || (stride & 3)) return fail(SCGPU_E_INVALID, "stride must be >= 12 and a multiple of 4");
  if (intensity_off && (intensity_off < 12 || intensity_off + 4 > stride || (intensity_off & 3))) return fail(SCGPU_E_INVALID, "intensity offset outside the point record");
  return SCGPU_OK;
}

int scgpu_assemble_submap(scgpu_handle* h, const void* const* clouds, const size_t* n_points, const scgpu_pose6* poses, size_t n_clouds,
                          size_t stride, size_t intensity_off, int drop_negative_intensity, float leaf, float* out_xyzw, size_t cap, size_t* n_out) {
  RET(submap_check(h, stride, intensity_off));
  if (!n_out || (n_clouds && (!clouds || !n_points || !poses))) return fail(SCGPU_E_INVALID, "null argument");
  if (leaf < 0.f || !(leaf < 1e30f)) return fail(SCGPU_E_INVALID, "bad leaf size");
  h = GROUP_FIRST(h);
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  size_t n = 0;
  RET(submap_to_device(h, clouds, n_points, poses, n_clouds, false, stride, intensity_off, drop_negative_intensity != 0, h->vox_in, h->icp_tgt, &n));
  const void* d_res = h->icp_tgt.p;
  if (leaf > 0.f && n) {
    RET(submap_voxel(h, h->icp_tgt.p, n, leaf, &n));
    d_res = h->vox_pts.p;
  }
  *n_out = n;
  if (n > cap) return fail(SCGPU_E_INVALID, "output capacity %zu < %zu points", cap, n);
  if (n && out_xyzw) {
    CK(cudaMemcpyAsync(out_xyzw, d_res, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return SCGPU_OK;
}

int scgpu_verify_loop_keyframes(scgpu_handle* h, const void* const* src_clouds, const size_t* src_points, size_t n_src_clouds, const scgpu_pose6* src_pose,
                                const void* const* tgt_clouds, const size_t* tgt_points, const scgpu_pose6* tgt_poses, size_t n_tgt_clouds, size_t stride,
                                size_t intensity_off, float leaf, const scgpu_icp_params* prm, double* T16, double* fitness, int* converged,
                                int* iterations, int* accepted, size_t* n_src_used, size_t* n_tgt_used) {
  RET(submap_check(h, stride, intensity_off));
  if (!prm || prm->max_iterations < 1 || (n_src_clouds && (!src_clouds || !src_points || !src_pose)) ||
      (n_tgt_clouds && (!tgt_clouds || !tgt_points || !tgt_poses)))
    return fail(SCGPU_E_INVALID, "null argument");
  if (leaf < 0.f || !(leaf < 1e30f)) return fail(SCGPU_E_INVALID, "bad leaf size");
  h = GROUP_FIRST(h);
  CK(cudaSetDevice(h->cfg.device));
  RET(join_replay(h));
  size_t ns = 0, nt = 0;
  // query side: the latest keyframe's clouds in the frame of the matched keyframe, negative intensities dropped (924-939)
  RET(submap_to_device(h, src_clouds, src_points, src_pose, n_src_clouds, true, stride, intensity_off, intensity_off != 0, h->vox_in, h->icp_src, &ns));
  // history side: the keyframes around the match, each by its own pose, then the voxel grid (942-949)
  RET(submap_to_device(h, tgt_clouds, tgt_points, tgt_poses, n_tgt_clouds, false, stride, intensity_off, false, h->vox_in, h->icp_tgt, &nt));
  const unsigned char* d_tgt = h->icp_tgt.as<unsigned char>();
  if (leaf > 0.f && nt) {
    RET(submap_voxel(h, h->icp_tgt.p, nt, leaf, &nt));
    d_tgt = h->vox_pts.as<unsigned char>();
  }
  if (n_src_used) *n_src_used = ns;
  if (n_tgt_used) *n_tgt_used = nt;
  return icp_run_device(h, h->icp_src.as<unsigned char>(), ns, sizeof(float4), d_tgt, nt, sizeof(float4), prm, T16, fitness, converged, iterations,
                        accepted);
}

// ---- peer-sharded database, one process per GPU -------------------------------------------------------------------------

// Which scans of a batch a shard bins, stores and searches for: scan i of the batch becomes global entry first + i and belongs to
// shard (first + i) % G.  Host arithmetic only (no device needed): *first_index = the shard's first scan in the batch, *count = how
// many it has; its j-th scan is batch index first_index + j * G.  This is the partition scgpu_peer_replay_async applies.
int scgpu_peer_partition(uint64_t first, size_t n_total, int G, int rank, size_t* first_index, size_t* count) {
  if (!first_index || !count || G < 1 || rank < 0 || rank >= G) return fail(SCGPU_E_INVALID, "bad argument");
  const size_t off = (size_t)(((uint64_t)rank + (uint64_t)G - first % (uint64_t)G) % (uint64_t)G);
  *first_index = off;
  *count = n_total > off ? (n_total - 1 - off) / (size_t)G + 1 : 0;
  return SCGPU_OK;
}

int scgpu_peer_export(scgpu_handle* h, void* blob, size_t blob_bytes) {
  if (!h || !blob) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->peer || h->is_group) return fail(SCGPU_E_INVALID, "not a peer-sharded shard handle (SCGPU_FLAG_PEER)");
  if (blob_bytes < SCGPU_PEER_BLOB_BYTES) return fail(SCGPU_E_INVALID, "blob must hold %d bytes", SCGPU_PEER_BLOB_BYTES);
  CK(cudaSetDevice(h->cfg.device));
  memset(blob, 0, SCGPU_PEER_BLOB_BYTES);
  cudaIpcMemHandle_t mh;
  CK(cudaIpcGetMemHandle(&mh, h->slab));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  unsigned char* b = static_cast<unsigned char*>(blob);
  memcpy(b, &mh, 64);
  uint64_t meta[4] = {h->slab_bytes, h->db.cap, (uint64_t)h->cfg.shard_rank, (uint64_t)h->cfg.shard_count};
  memcpy(b + 64, meta, sizeof meta);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  memcpy(b + 96, &prop.uuid, 16);  // lets the peers see whether two shards share a physical GPU
  return SCGPU_OK;
}

int scgpu_peer_attach(scgpu_handle* h, const void* blobs, int n) {
  if (!h || !blobs) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->peer || h->is_group) return fail(SCGPU_E_INVALID, "not a peer-sharded shard handle (SCGPU_FLAG_PEER)");
  if (n != h->cfg.shard_count) return fail(SCGPU_E_INVALID, "%d blobs for %d shards", n, h->cfg.shard_count);
  if (h->attached) return fail(SCGPU_E_INVALID, "already attached");
  CK(cudaSetDevice(h->cfg.device));
  const unsigned char* b = static_cast<const unsigned char*>(blobs);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  for (int s = 0; s < n; ++s) {
    const unsigned char* bs = b + (size_t)s * SCGPU_PEER_BLOB_BYTES;
    uint64_t meta[4];
    memcpy(meta, bs + 64, sizeof meta);
    if (meta[0] != h->slab_bytes || meta[1] != h->db.cap || meta[2] != (uint64_t)s || meta[3] != (uint64_t)n)
      return fail(SCGPU_E_INVALID, "shard %d was created with a different configuration (slab %llu vs %llu bytes)", s,
                  (unsigned long long)meta[0], (unsigned long long)h->slab_bytes);
    if (s == h->cfg.shard_rank) {
      peer_fill(h, s, h->slab);
      continue;
    }
    if (memcmp(bs + 96, &prop.uuid, 16) == 0)
      return fail(SCGPU_E_INVALID, "shards %d and %d are on the same GPU: the cross-process barrier needs one GPU per process", s, h->cfg.shard_rank);
    cudaIpcMemHandle_t mh;
    memcpy(&mh, bs, 64);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_base[s] = base;
    peer_fill(h, s, base);
  }
  h->peers.G = n;
  h->peers.rank = h->cfg.shard_rank;
  h->peers.cap = h->db.cap;
  h->attached = true;
  return SCGPU_OK;
}

int scgpu_peer_replay_async(scgpu_handle* h, const void* pts, size_t n_total, size_t pts_per_scan, size_t stride, int location) {
  SCGPU_TRACE_CALL();
  if (!h || (!pts && n_total && pts_per_scan)) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->peer || h->is_group || !h->attached) return fail(SCGPU_E_INVALID, "needs an attached peer-sharded shard handle");
  if (n_total == 0) return SCGPU_OK;
  const uint64_t first = h->n_global;
  std::vector<ReplayShard> sh;
  sh.push_back({h, pts, 0});
  return replay_enqueue(sh, first, n_total, pts_per_scan, stride, location, true);
}

int scgpu_peer_barrier(scgpu_handle* h, void* stream) {
  if (!h) return fail(SCGPU_E_INVALID, "null argument");
  if (!h->peer || h->is_group || !h->attached) return fail(SCGPU_E_INVALID, "needs an attached peer-sharded shard handle");
  CK(cudaSetDevice(h->cfg.device));
  return launch_peer_barrier(h, 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
