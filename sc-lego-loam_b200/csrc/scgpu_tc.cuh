// scgpu_tc.cuh -- the column-shifted cosine distance as a tensor-core contraction (tcgen05 / TMEM / TMA, sm_100a) for the
// FULL-SHIFT search (SEARCH_RATIO = 1, SC.cpp:123-144 with every one of the S shifts in the search set) batched over queries.
//
// Why here and not in the reference's windowed search: with unit-normalised columns the distance at shift s is
//       1 - (1/n_s) * sum_r sum_j A^[r][j] * B^[r][(j - s) mod S]
// For the windowed search only W = 7 of the S = 60 wrapped diagonals are needed and the SIMT FFMA2 kernel
// (scgpu_exhaustive.cuh) computes exactly those; computing all S of them costs 8.6x the products.  In the full-shift mode all S
// ARE needed -- 72,000 multiply-adds per (query, entry) pair -- and that is GEMM-shaped work:
//
//       D[e][(q, s)] = sum_k  E[e][k] * Qs[(q, s)][k],      k = (r, column position) over R*S = 1200 values
//
// with E = the screening copy of the database exactly as the SIMT kernels store it (row-major per entry, unit-norm columns,
// pair-interleaved column positions) and Qs[(q, s)] = query q's normalised descriptor rotated by s columns, laid out in the same
// k order (a "circulant expansion" of the queries, built once per batch by k_tc_prep_queries: 60 x 1200 floats per query).
// The shift index lives in the GEMM's N dimension, so the accumulator tile IS the table of per-shift sums: no wrapped-diagonal
// reduction, no alignment stage; the epilogue normalises by n_s (valid-column masks, popcount) and takes the minimum over s.
//
// Precision: 3xTF32 -- every operand v is split into hi = v with the low 13 mantissa bits cleared (exactly a TF32 value) and
// lo = v - hi (exact in FP32); D accumulates hi*hi + hi*lo + lo*hi in FP32 in TMEM.  Each product then carries a relative
// error <= 3 * 2^-20 (the dropped lo*lo term and the truncation of lo), i.e. <= 2.9e-6 on a distance (the per-shift sum of
// |cos| terms is <= n_s), plus the FP32 accumulation of 3 * 1200 terms.  TC_EPS bounds both (observed: see DESIGN.md); the
// survivors are rescored by the exact FP64 pair kernel exactly as in the SIMT screening path.
//
// Kernel shape (one CTA per SM, persistent over (query group, entry tile) pairs):
//   warp 0   TMA producer: 2-D tensor maps (128-byte swizzle) over E_hi / E_lo [entries][1200] and Qs_hi / Qs_lo [rows][1200];
//            per K block of 32 floats one stage = E tile 128 x 32 (hi, lo) + Qs tile 240 x 32 (hi, lo) = 92 KB, 2 stages
//   warp 1   MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, M = 128 entries, N = 240 = 4 queries x 60
//            shifts, K = 8 per instruction -> 4 x 3 MMAs per stage; accumulator in TMEM (2 x 256 columns, double-buffered)
//   warps 2-9 epilogue (two per TMEM lane quarter, two queries each): tcgen05.ld 32x32b (thread = entry, columns = (query, shift)),
//            n_s from precomputed rotated query masks, min over s with the reference's smallest-shift tie rule, d32 + per-query
//            running minimum out
// Windowed mode (batches of the reference's windowed search): two more K blocks per tile carry the sector keys, a second small GEMM
// into TMEM columns [256, 496) gives the alignment correlations of all S shifts; the epilogue takes the argmax (ambiguous ones: lower
// bound over both candidates' windows) and the minimum over the 2*radius+1 shifts around it.  One tile in TMEM at a time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "scgpu_exhaustive.cuh"

namespace scgpu {

constexpr float TC_ALIGN_MARGIN = 6.4e-5f;  // relative (to |v1||v2|) gap below which the 3xTF32 alignment is called ambiguous
constexpr float TC_EPS = 1.0e-4f;  // |d32 - d| bound used for candidate selection on the tensor-core path (observed: < 2e-5)

constexpr int TC_S = 60, TC_R = 20, TC_K = TC_R * TC_S;  // instantiated for the reference's 20 x 60 descriptor
constexpr int TC_QG = 4;                                  // queries per accumulator tile
constexpr int TC_N = TC_QG * TC_S;                        // 240 accumulator columns
constexpr int TC_M = 128;                                 // entries per tile
constexpr int TC_BK = 32;                                 // floats per K block (128 bytes: one swizzle atom row)
constexpr int TC_KBLOCKS = (TC_K + TC_BK - 1) / TC_BK;    // 38 (the last one is half out of bounds: TMA zero-fills)
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_M * TC_BK * 4;              // 16 KB
constexpr int TC_B_BYTES = TC_N * TC_BK * 4;              // 30 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;
constexpr int TC_THREADS = 320;                            // TMA warp, MMA warp, eight epilogue warps
constexpr int TC_ACC_COLS = 256;                          // TMEM columns per accumulator stage
constexpr int TC_VK = 64;                                 // sector keys padded to two K blocks (windowed mode: alignment GEMM)
constexpr int TC_VKBLOCKS = TC_VK / TC_BK;

struct TcQueryAux {  // per query: rotated valid-column masks + flags (epilogue side data)
  unsigned long long qrot[TC_S];  // qrot[s] bit c set <=> query column (c + s) mod S is valid: popc(qrot[s] & vmask_e) = n_s
  unsigned flags;                 // bit 0: rescore everything (non-representable norms)
  float v1norm;                   // |sector key| of the query (alignment margin, windowed mode)
};

constexpr size_t tc_smem_bytes() { return 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + TC_QG * sizeof(TcQueryAux) + 64 * 4 + 256; }

// ---- operand preparation ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// database side: hi / lo split of the screening copy (rows [first, first + gridDim.x))
__global__ void __launch_bounds__(256) k_tc_split_db(const float* sc_hat, float* hi, float* lo, unsigned long long first) {
  const unsigned long long e = first + blockIdx.x;
  for (int i = threadIdx.x; i < TC_K; i += blockDim.x) {
    const float v = sc_hat[e * TC_K + i];
    const float h = tf32_hi(v);
    hi[e * TC_K + i] = h;
    lo[e * TC_K + i] = v - h;
  }
}

// database side, windowed mode: hi / lo split of the float sector keys (pair-interleaved, as stored), padded to TC_VK
__global__ void __launch_bounds__(64) k_tc_split_vk(const unsigned char* vk, float* hi, float* lo, unsigned long long first) {
  const unsigned long long e = first + blockIdx.x;
  const float* v = reinterpret_cast<const float*>(vk + e * sizeof(ExhVkRec<TC_S>));
  const int i = threadIdx.x;
  const float x = i < TC_S ? v[i] : 0.f;
  const float h = tf32_hi(x);
  hi[e * TC_VK + i] = h;
  lo[e * TC_VK + i] = x - h;
}

// query side: the circulant expansion.  Block (s, q): row q*S + s of Qs = query q rotated by s, in the database's k order.
__global__ void __launch_bounds__(128) k_tc_prep_queries(const ExhQuery* qs, float* qs_hi, float* qs_lo, TcQueryAux* aux, float* qv_hi, float* qv_lo) {
  const int s = blockIdx.x, q = blockIdx.y;
  const ExhQuery* Q = qs + q;
  const size_t row = ((size_t)q * TC_S + s) * TC_K;
  for (int i = threadIdx.x; i < TC_K; i += blockDim.x) {
    const int r = i / TC_S, pos = i - r * TC_S;
    const int c = (pos & 1) ? (pos >> 1) + TC_S / 2 : (pos >> 1);  // column stored at this position (inverse of pair_pos)
    int cq = c + s;
    if (cq >= TC_S) cq -= TC_S;
    const float v = Q->qhat[r * TC_S + cq];  // D[s] = sum_j A^[r][j] * B^[r][(j - s)]  =  sum_c B^[r][c] * A^[r][(c + s)]
    const float h = tf32_hi(v);
    qs_hi[row + i] = h;
    qs_lo[row + i] = v - h;
  }
  if (qv_hi && threadIdx.x < TC_VK) {  // windowed mode: the same expansion of the query's sector key (alignment GEMM, K = 60 -> 64)
    const int pos = threadIdx.x;
    float v = 0.f;
    if (pos < TC_S) {
      const int c = (pos & 1) ? (pos >> 1) + TC_S / 2 : (pos >> 1);
      int cq = c + s;
      if (cq >= TC_S) cq -= TC_S;
      v = Q->v1[cq];  // corr(s) = sum_c v2[c] * v1[(c + s) mod S]: the shift that minimises |v1 - circshift(v2, s)| (SC.cpp:93-113)
    }
    const float h = tf32_hi(v);
    const size_t o = ((size_t)q * TC_S + s) * TC_VK + pos;
    qv_hi[o] = h;
    qv_lo[o] = v - h;
  }
  if (threadIdx.x == 0) {
    const unsigned long long m = Q->qmask[0];
    const unsigned long long full = (1ull << TC_S) - 1;
    // bit c of qrot[s] = bit (c + s) mod S of the query mask
    aux[q].qrot[s] = s == 0 ? m : (((m >> s) | (m << (TC_S - s))) & full);
    if (s == 0) {
      aux[q].flags = Q->flags;
      aux[q].v1norm = Q->v1norm;
    }
  }
}

// ---- tcgen05 / TMA PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), LBO unused (1)
__device__ __forceinline__ uint64_t tc_smem_desc(const void* p) {
  return (uint64_t)((smem_u32(p) >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TcParams {
  const unsigned char* vk;            // [cap] ExhVkRec<60>: valid-column mask + flags of every entry
  const TcQueryAux* qaux;             // [n_groups * TC_QG]
  const unsigned long long* n_local;  // [nq] local entries to score per query
  unsigned nq;                        // queries of the batch (the last group may be partly empty)
  unsigned n_groups;                  // ceil(nq / TC_QG)
  unsigned n_tiles;                   // ceil(max n_local / TC_M)
  unsigned long long d32_pitch;
  float* d32;                         // [nq][d32_pitch]
  unsigned* min_bits;                 // [nq]
  unsigned* shift_out;                // optional [nq][d32_pitch]: argmin shift (tests)
  int windowed;                       // 1: the reference's windowed search -- a second small GEMM (sector keys, K = 64) gives the alignment,
  int radius;                         //    the distance is the minimum over the 2*radius+1 shifts around it (SC.cpp:121-144)
  // two-pass windowed mode (256-row tiles have no TMEM left for the alignment accumulator): k_tc_fullshift runs the alignment GEMM
  // ALONE and writes, per (query, entry), best shift | runner-up << 8 | ambiguous << 16 | flagged << 17 to align_out;
  // k_tc_fullshift2 then reads it (align_in) in its epilogue
  unsigned* align_out;
  const unsigned* align_in;
};

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_fullshift(const __grid_constant__ CUtensorMap map_e_hi, const __grid_constant__ CUtensorMap map_e_lo,
                                                                const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                                                                const __grid_constant__ CUtensorMap map_ev_hi, const __grid_constant__ CUtensorMap map_ev_lo,
                                                                const __grid_constant__ CUtensorMap map_qv_hi, const __grid_constant__ CUtensorMap map_qv_lo,
                                                                const TcParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  // carve: [stages][E_hi | E_lo | Q_hi | Q_lo] (each 1024-byte aligned: 16 KB, 16 KB, 30 KB, 30 KB) | query aux | rcp table | barriers
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  auto stage_ptr = [&](int s, int which) {  // which: 0 E_hi, 1 E_lo, 2 Q_hi, 3 Q_lo
    unsigned char* sp = base + (size_t)s * TC_STAGE_BYTES;
    return sp + (which == 0 ? 0 : which == 1 ? TC_A_BYTES : which == 2 ? 2 * TC_A_BYTES : 2 * TC_A_BYTES + TC_B_BYTES);
  };
  TcQueryAux* s_aux = reinterpret_cast<TcQueryAux*>(base + (size_t)TC_STAGES * TC_STAGE_BYTES);
  float* s_rcp = reinterpret_cast<float*>(s_aux + TC_QG);  // [64]: 1 / n
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rcp + 64);
  uint64_t* full = bars;                 // [TC_STAGES]
  uint64_t* empty = bars + TC_STAGES;    // [TC_STAGES]
  uint64_t* acc_full = bars + 2 * TC_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* aux_ready = acc_empty + 2;         // [1]  (unused slots keep the layout simple)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(aux_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 64) s_rcp[threadIdx.x] = threadIdx.x ? 1.0f / (float)threadIdx.x : 0.f;
  if (warp == 1) {  // TMEM: the whole 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // work items: (group g, entry tile m), g-major so that consecutive items of a CTA reuse the query operand in L2
  const unsigned long long n_items = (unsigned long long)p.n_groups * p.n_tiles;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      unsigned it = 0;
      for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x) {
        const unsigned g = (unsigned)(w / p.n_tiles), m = (unsigned)(w % p.n_tiles);
        const int nkb = TC_KBLOCKS + (p.windowed ? TC_VKBLOCKS : 0);
        for (int kb = p.align_out ? TC_KBLOCKS : 0; kb < nkb; ++kb, ++it) {
          const int s = it % TC_STAGES;
          if (it >= TC_STAGES) mbar_wait(&empty[s], ((it / TC_STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
          const bool al = kb >= TC_KBLOCKS;  // the last two stages of a windowed tile carry the sector keys
          const int k0 = (al ? kb - TC_KBLOCKS : kb) * TC_BK;
          tma_load_2d(stage_ptr(s, 0), al ? &map_ev_hi : &map_e_hi, k0, (int)(m * TC_M), &full[s]);
          tma_load_2d(stage_ptr(s, 1), al ? &map_ev_lo : &map_e_lo, k0, (int)(m * TC_M), &full[s]);
          tma_load_2d(stage_ptr(s, 2), al ? &map_qv_hi : &map_q_hi, k0, (int)(g * TC_N), &full[s]);
          tma_load_2d(stage_ptr(s, 3), al ? &map_qv_lo : &map_q_lo, k0, (int)(g * TC_N), &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = tc_idesc(TC_M, TC_N);
      unsigned it = 0, tile = 0;
      for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x, ++tile) {
        // full-shift: two accumulators (256 columns each), the epilogue of tile t overlaps the MMAs of tile t+1.
        // windowed: ONE tile in TMEM at a time -- columns [0, 240) the per-shift sums, [256, 496) the alignment correlations
        const int a = p.windowed ? 0 : (tile & 1);
        if (p.windowed) {
          if (tile >= 1) mbar_wait(&acc_empty[0], (tile - 1) & 1);
        } else if (tile >= 2) {
          mbar_wait(&acc_empty[a], ((tile >> 1) - 1) & 1);  // the epilogue has drained this accumulator
        }
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)a * TC_ACC_COLS;
        const int nkb = TC_KBLOCKS + (p.windowed ? TC_VKBLOCKS : 0);
        for (int kb = p.align_out ? TC_KBLOCKS : 0; kb < nkb; ++kb, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait(&full[s], (it / TC_STAGES) & 1);
          tc_fence_after();
          const uint64_t e_hi = tc_smem_desc(stage_ptr(s, 0)), e_lo = tc_smem_desc(stage_ptr(s, 1));
          const uint64_t q_hi = tc_smem_desc(stage_ptr(s, 2)), q_lo = tc_smem_desc(stage_ptr(s, 3));
          const bool al = kb >= TC_KBLOCKS;
          const uint32_t d = al ? d0 + TC_ACC_COLS : d0;
          const int kfirst = al ? TC_KBLOCKS : 0;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 8 * 4) >> 4;  // 32 bytes along K inside the swizzle atom
            tc_mma_tf32(d, e_hi + adv, q_hi + adv, idesc, ((kb - kfirst) | kk) != 0);
            tc_mma_tf32(d, e_hi + adv, q_lo + adv, idesc, 1);
            tc_mma_tf32(d, e_lo + adv, q_hi + adv, idesc, 1);
          }
          tc_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
        }
        tc_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: thread = entry (TMEM lane), columns = (query, shift) =====
    // eight warps: two per TMEM lane quarter (a warp may read lanes 32 * (warp % 4) .. +31), each taking two of the tile's four queries
    const int quarter = warp & 3;
    const int qhalf = (warp - 2) >> 2;         // 0: queries 0, 1 of the group; 1: queries 2, 3
    const int row = quarter * 32 + lane;       // entry within the tile
    unsigned tile = 0;
    unsigned my_min[TC_QG];
    unsigned cur_g = 0xffffffffu;
    for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x, ++tile) {
      const unsigned g = (unsigned)(w / p.n_tiles), m = (unsigned)(w % p.n_tiles);
      const int a = p.windowed ? 0 : (tile & 1);
      if (g != cur_g) {  // new query group: flush the running minima of the previous one, load the side data of this one
        if (cur_g != 0xffffffffu) {
#pragma unroll
          for (int q = 0; q < TC_QG; ++q) {
            const unsigned mn = __reduce_min_sync(FULL, my_min[q]);
            if (lane == 0 && mn != 0x7f800000u && cur_g * TC_QG + q < p.nq) atomicMin(p.min_bits + cur_g * TC_QG + q, mn);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done with the previous group's side data
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.qaux + (size_t)g * TC_QG);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_aux);
        for (int i = threadIdx.x - 64; i < (int)(TC_QG * sizeof(TcQueryAux) / 8); i += 256) dst[i] = src[i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_g = g;
#pragma unroll
        for (int q = 0; q < TC_QG; ++q) my_min[q] = 0x7f800000u;
      }
      const unsigned long long e = (unsigned long long)m * TC_M + row;
      const ExhAux* ax = reinterpret_cast<const ExhAux*>(p.vk + e * sizeof(ExhVkRec<TC_S>) + TC_S * sizeof(float));
      unsigned long long vmask = 0;
      unsigned eflags = 0;
      float vnorm = 0.f;
      unsigned long long nl_max = 0;
#pragma unroll
      for (int q = 0; q < TC_QG; ++q) {
        const unsigned qi = g * TC_QG + q;
        const unsigned long long nl = qi < p.nq ? p.n_local[qi] : 0;
        nl_max = nl > nl_max ? nl : nl_max;
      }
      if (e < nl_max) {
        vmask = ax->vmask[0];
        eflags = ax->flags;
        vnorm = ax->vnorm;
      }
      const int tpar = p.windowed ? (int)(tile & 1) : (int)((tile >> 1) & 1);
      mbar_wait(&acc_full[a], tpar);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * TC_ACC_COLS;
#pragma unroll 1
      for (int q = qhalf * (TC_QG / 2); q < (qhalf + 1) * (TC_QG / 2); ++q) {
        uint32_t v0[32], v1[32];
        const unsigned qi = g * TC_QG + q;
        const unsigned long long nl = qi < p.nq ? p.n_local[qi] : 0;
        // windowed: the alignment first -- argmax over the S correlations (smallest shift on ties), ambiguous when the runner-up is
        // within the error bound of the best (then the exact kernel decides, as in the SIMT screening)
        // An alignment whose runner-up is within the error bound is not decided here; instead the distance is taken over the UNION of
        // the two candidates' windows -- a lower bound of the true windowed distance whichever alignment the reference picks -- and
        // the entry is kept out of the per-query minimum (which must come from certain values).  A lower bound above the selection
        // threshold proves the entry cannot win; below it the exact kernel decides.  Three candidates within the bound: flagged.
        int a_cur = 0, a_2nd = 0;
        bool amb2 = false, amb3 = false;
        if (p.windowed) {
          tc_ld32(tbase + TC_ACC_COLS + q * TC_S, v0);
          tc_ld32(tbase + TC_ACC_COLS + q * TC_S + 28, v1);
          tc_ld_wait();
          float b1 = -3.4e38f, b2 = -3.4e38f, b3 = -3.4e38f;
#pragma unroll
          for (int s = 0; s < TC_S; ++s) {
            const float c = __uint_as_float(s < 32 ? v0[s] : v1[s - 28]);
            const bool g1 = c > b1;            // strict: the earlier shift keeps a tie
            const bool g2 = !g1 && c > b2;
            b3 = fmaxf(b3, g1 ? b2 : (g2 ? b2 : c));
            a_2nd = g1 ? a_cur : (g2 ? s : a_2nd);
            b2 = g1 ? b1 : (g2 ? c : b2);
            a_cur = g1 ? s : a_cur;
            b1 = g1 ? c : b1;
          }
          const float margin = TC_ALIGN_MARGIN * s_aux[q].v1norm * vnorm;
          amb2 = !((b1 - b2) > margin);
          amb3 = !((b1 - b3) > margin) || !(b1 == b1);
        }
        if (p.align_out) {  // alignment pass only
          if (e < nl) p.align_out[(size_t)qi * p.d32_pitch + e] = (unsigned)a_cur | ((unsigned)a_2nd << 8) | (amb2 ? 1u << 16 : 0u) | (amb3 ? 1u << 17 : 0u);
          continue;
        }
        tc_ld32(tbase + q * TC_S, v0);        // shifts 0..31
        tc_ld32(tbase + q * TC_S + 28, v1);   // shifts 28..59
        tc_ld_wait();
        float best = __int_as_float(0x7f800000);
        int best_s = 0;
        bool bad = false;
#pragma unroll
        for (int s = 0; s < TC_S; ++s) {
          const float sum = __uint_as_float(s < 32 ? v0[s] : v1[s - 28]);
          const int n = __popcll(s_aux[q].qrot[s] & vmask);
          float dist = n ? 1.0f - sum * s_rcp[n] : __int_as_float(0x7f800000);
          if (p.windowed) {  // only the shifts a-radius .. a+radius (mod S) are in the reference's search set (SC.cpp:123-130)
            int rel = s - a_cur, rel2 = s - a_2nd;
            rel += rel < 0 ? TC_S : 0;
            rel2 += rel2 < 0 ? TC_S : 0;
            const bool in = rel <= p.radius || rel >= TC_S - p.radius || (amb2 && (rel2 <= p.radius || rel2 >= TC_S - p.radius));
            dist = in ? dist : __int_as_float(0x7f800000);
          }
          bad |= !(dist == dist);
          if (dist < best) {  // ascending s, strict: the smallest shift keeps a tie (SC.cpp:136-143)
            best = dist;
            best_s = s;
          }
        }
        if (e < nl) {
          float out = best < 0.f ? 0.f : best;
          if (bad || amb3 || (eflags & 1u) || (s_aux[q].flags & 1u)) out = -1.0f;  // the exact kernel decides
          p.d32[(size_t)qi * p.d32_pitch + e] = out;
          if (p.shift_out) p.shift_out[(size_t)qi * p.d32_pitch + e] = (unsigned)best_s;
          if (out >= 0.f && !amb2) my_min[q] = min(my_min[q], __float_as_uint(out));  // lower bounds do not set the threshold
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
    if (cur_g != 0xffffffffu) {
#pragma unroll
      for (int q = 0; q < TC_QG; ++q) {
        const unsigned mn = __reduce_min_sync(FULL, my_min[q]);
        if (lane == 0 && mn != 0x7f800000u && cur_g * TC_QG + q < p.nq) atomicMin(p.min_bits + cur_g * TC_QG + q, mn);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- k_tc_fullshift2: the full-shift kernel with a 256 x 240 tile per CTA --------------------------------------------------
// k_tc_fullshift is bound by operand traffic, not by the tensor pipe (58 % active): every K block moves (128 + 240) rows x hi/lo
// through shared memory for 128 x 240 x 32 products, and both operands come from L2.  Here one CTA accumulates TWO 128-row halves
// against the same query tile -- (256 + 240) rows per 256 x 240 products: 1.48x fewer operand bytes per FLOP.  To keep three
// stages in 227 KB the K block is 16 floats (64-byte rows, 64-byte swizzle): stage = E 256 x 16 (hi, lo) + Qs 240 x 16 (hi, lo)
// = 62 KB.  The two accumulators take TMEM columns [0, 240) and [256, 496): one tile at a time (the epilogue -- eight warps, one
// entry x four queries per thread -- is not hidden; ~10 % of a tile's MMA time).  Full-shift mode only.
constexpr int TC2_M = 256, TC2_BK = 16, TC2_STAGES = 3;
constexpr int TC2_KBLOCKS = TC_K / TC2_BK;  // 75
constexpr int TC2_A_BYTES = TC2_M * TC2_BK * 4;  // 16 KB
constexpr int TC2_B_BYTES = TC_N * TC2_BK * 4;   // 15 KB
constexpr int TC2_STAGE_BYTES = 2 * TC2_A_BYTES + 2 * TC2_B_BYTES;
static_assert(TC_K % TC2_BK == 0, "whole K blocks");
constexpr size_t tc2_smem_bytes() { return 1024 + (size_t)TC2_STAGES * TC2_STAGE_BYTES + TC_QG * sizeof(TcQueryAux) + 64 * 4 + 256; }

// K-major operand tile, 64-byte swizzle: rows of 64 bytes, 8-row atoms of 512 bytes (SBO)
__device__ __forceinline__ uint64_t tc_smem_desc64(const void* p) {
  return (uint64_t)((smem_u32(p) >> 4) & 0x3fffu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_fullshift2(const __grid_constant__ CUtensorMap map_e_hi, const __grid_constant__ CUtensorMap map_e_lo,
                                                                 const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                                                                 const TcParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  auto stage_ptr = [&](int s, int which) {  // which: 0 E_hi, 1 E_lo, 2 Q_hi, 3 Q_lo
    unsigned char* sp = base + (size_t)s * TC2_STAGE_BYTES;
    return sp + (which == 0 ? 0 : which == 1 ? TC2_A_BYTES : which == 2 ? 2 * TC2_A_BYTES : 2 * TC2_A_BYTES + TC2_B_BYTES);
  };
  TcQueryAux* s_aux = reinterpret_cast<TcQueryAux*>(base + (size_t)TC2_STAGES * TC2_STAGE_BYTES);
  float* s_rcp = reinterpret_cast<float*>(s_aux + TC_QG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rcp + 64);
  uint64_t* full = bars;                          // [TC2_STAGES]
  uint64_t* empty = bars + TC2_STAGES;            // [TC2_STAGES]
  uint64_t* acc_full = bars + 2 * TC2_STAGES;     // [1]
  uint64_t* acc_empty = acc_full + 1;             // [1]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (threadIdx.x < 64) s_rcp[threadIdx.x] = threadIdx.x ? 1.0f / (float)threadIdx.x : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const unsigned long long n_items = (unsigned long long)p.n_groups * p.n_tiles;  // n_tiles counts 256-entry tiles here

  if (warp == 0) {
    if (lane == 0) {
      unsigned it = 0;
      for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x) {
        const unsigned g = (unsigned)(w / p.n_tiles), m = (unsigned)(w % p.n_tiles);
        for (int kb = 0; kb < TC2_KBLOCKS; ++kb, ++it) {
          const int s = it % TC2_STAGES;
          if (it >= TC2_STAGES) mbar_wait(&empty[s], ((it / TC2_STAGES) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], TC2_STAGE_BYTES);
          tma_load_2d(stage_ptr(s, 0), &map_e_hi, kb * TC2_BK, (int)(m * TC2_M), &full[s]);
          tma_load_2d(stage_ptr(s, 1), &map_e_lo, kb * TC2_BK, (int)(m * TC2_M), &full[s]);
          tma_load_2d(stage_ptr(s, 2), &map_q_hi, kb * TC2_BK, (int)(g * TC_N), &full[s]);
          tma_load_2d(stage_ptr(s, 3), &map_q_lo, kb * TC2_BK, (int)(g * TC_N), &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc_idesc(128, TC_N);
      unsigned it = 0, tile = 0;
      for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x, ++tile) {
        if (tile >= 1) mbar_wait(acc_empty, (tile - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < TC2_KBLOCKS; ++kb, ++it) {
          const int s = it % TC2_STAGES;
          mbar_wait(&full[s], (it / TC2_STAGES) & 1);
          tc_fence_after();
          const uint64_t q_hi = tc_smem_desc64(stage_ptr(s, 2)), q_lo = tc_smem_desc64(stage_ptr(s, 3));
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {  // the two 128-row halves of the entry tile share the query tile
            const uint64_t e_hi = tc_smem_desc64(stage_ptr(s, 0) + mh * 128 * TC2_BK * 4), e_lo = tc_smem_desc64(stage_ptr(s, 1) + mh * 128 * TC2_BK * 4);
            const uint32_t d = tmem_base + (uint32_t)mh * TC_ACC_COLS;
#pragma unroll
            for (int kk = 0; kk < TC2_BK / 8; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 8 * 4) >> 4;
              tc_mma_tf32(d, e_hi + adv, q_hi + adv, idesc, (kb | kk) != 0);
              tc_mma_tf32(d, e_hi + adv, q_lo + adv, idesc, 1);
              tc_mma_tf32(d, e_lo + adv, q_hi + adv, idesc, 1);
            }
          }
          tc_commit(&empty[s]);
        }
        tc_commit(acc_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int mhalf = (warp - 2) >> 2;
    const int row = mhalf * 128 + quarter * 32 + lane;
    unsigned tile = 0;
    unsigned my_min[TC_QG];
    unsigned cur_g = 0xffffffffu;
    for (unsigned long long w = blockIdx.x; w < n_items; w += gridDim.x, ++tile) {
      const unsigned g = (unsigned)(w / p.n_tiles), m = (unsigned)(w % p.n_tiles);
      if (g != cur_g) {
        if (cur_g != 0xffffffffu) {
#pragma unroll
          for (int q = 0; q < TC_QG; ++q) {
            const unsigned mn = __reduce_min_sync(FULL, my_min[q]);
            if (lane == 0 && mn != 0x7f800000u && cur_g * TC_QG + q < p.nq) atomicMin(p.min_bits + cur_g * TC_QG + q, mn);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.qaux + (size_t)g * TC_QG);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_aux);
        for (int i = threadIdx.x - 64; i < (int)(TC_QG * sizeof(TcQueryAux) / 8); i += 256) dst[i] = src[i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_g = g;
#pragma unroll
        for (int q = 0; q < TC_QG; ++q) my_min[q] = 0x7f800000u;
      }
      const unsigned long long e = (unsigned long long)m * TC2_M + row;
      const ExhAux* ax = reinterpret_cast<const ExhAux*>(p.vk + e * sizeof(ExhVkRec<TC_S>) + TC_S * sizeof(float));
      unsigned long long vmask = 0;
      unsigned eflags = 0;
      unsigned long long nl_max = 0;
#pragma unroll
      for (int q = 0; q < TC_QG; ++q) {
        const unsigned qi = g * TC_QG + q;
        const unsigned long long nl = qi < p.nq ? p.n_local[qi] : 0;
        nl_max = nl > nl_max ? nl : nl_max;
      }
      if (e < nl_max) {
        vmask = ax->vmask[0];
        eflags = ax->flags;
      }
      mbar_wait(acc_full, tile & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)mhalf * TC_ACC_COLS;
#pragma unroll 1
      for (int q = 0; q < TC_QG; ++q) {
        uint32_t v0[32], v1[32];
        tc_ld32(tbase + q * TC_S, v0);
        tc_ld32(tbase + q * TC_S + 28, v1);
        tc_ld_wait();
        const unsigned qi = g * TC_QG + q;
        const unsigned long long nl = qi < p.nq ? p.n_local[qi] : 0;
        // windowed search: the alignment of this (query, entry) pair comes from the alignment pass (k_tc_fullshift, align_out)
        int a_cur = 0, a_2nd = 0;
        bool amb2 = false, amb3 = false;
        if (p.align_in && e < nl) {
          const unsigned ai = p.align_in[(size_t)qi * p.d32_pitch + e];
          a_cur = (int)(ai & 0xffu);
          a_2nd = (int)((ai >> 8) & 0xffu);
          amb2 = (ai >> 16) & 1u;
          amb3 = (ai >> 17) & 1u;
        }
        float best = __int_as_float(0x7f800000);
        int best_s = 0;
        bool bad = false;
#pragma unroll
        for (int s = 0; s < TC_S; ++s) {
          const float sum = __uint_as_float(s < 32 ? v0[s] : v1[s - 28]);
          const int n = __popcll(s_aux[q].qrot[s] & vmask);
          float dist = n ? 1.0f - sum * s_rcp[n] : __int_as_float(0x7f800000);
          if (p.align_in) {
            int rel = s - a_cur, rel2 = s - a_2nd;
            rel += rel < 0 ? TC_S : 0;
            rel2 += rel2 < 0 ? TC_S : 0;
            const bool in = rel <= p.radius || rel >= TC_S - p.radius || (amb2 && (rel2 <= p.radius || rel2 >= TC_S - p.radius));
            dist = in ? dist : __int_as_float(0x7f800000);
          }
          bad |= !(dist == dist);
          if (dist < best) {
            best = dist;
            best_s = s;
          }
        }
        if (e < nl) {
          float out = best < 0.f ? 0.f : best;
          if (bad || amb3 || (eflags & 1u) || (s_aux[q].flags & 1u)) out = -1.0f;
          p.d32[(size_t)qi * p.d32_pitch + e] = out;
          if (p.shift_out) p.shift_out[(size_t)qi * p.d32_pitch + e] = (unsigned)best_s;
          if (out >= 0.f && !amb2) my_min[q] = min(my_min[q], __float_as_uint(out));  // lower bounds do not set the threshold
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (cur_g != 0xffffffffu) {
#pragma unroll
      for (int q = 0; q < TC_QG; ++q) {
        const unsigned mn = __reduce_min_sync(FULL, my_min[q]);
        if (lane == 0 && mn != 0x7f800000u && cur_g * TC_QG + q < p.nq) atomicMin(p.min_bits + cur_g * TC_QG + q, mn);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- SIMT counterpart for the A/B: the same full-shift screening on the FP32 pipes (FFMA2), one warp per entry ----------------
// Lane r < R owns descriptor row r of the entry (15 x 16-byte loads from the screening copy) and accumulates the S shifts in four
// passes of 15 against the query's pair table in shared memory (window_fma, as k_exh_screen); each pass ends in the
// transpose-reduce of its 15 sums.  72,000 FMA per pair, no alignment stage.
template <int R, int S>
__global__ void __launch_bounds__(256) k_fullshift_simt(const float* sc_hat, const unsigned char* vk, const ExhQuery* qs, const unsigned long long* n_local,
                                                        unsigned long long d32_pitch, float* d32, unsigned* min_bits, unsigned* shift_out) {
  constexpr int WP = 15, PASSES = S / WP;
  static_assert(S % WP == 0 && R <= 32, "shifts are covered by whole passes");
  constexpr int PITCH = 2 * qtab_pairs(S, WP);
  extern __shared__ __align__(16) unsigned char fs_smem[];
  float* qtable = reinterpret_cast<float*>(fs_smem);
  const unsigned qi = blockIdx.y;
  const ExhQuery* Q = qs + qi;
  qtab_fill<R, S, WP>(qtable, [Q](int r, int c) { return r < R ? Q->qhat[r * S + c] : 0.f; });
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long nl = n_local[qi];
  const unsigned long long qmask[2] = {Q->qmask[0], Q->qmask[1]};
  const bool q_flag = (Q->flags & 1u) != 0;
  const bool row_lane = lane < R;
  unsigned my_min = 0x7f800000u;
  const unsigned long long wstride = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  for (unsigned long long e = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + warp; e < nl; e += wstride) {
    const ExhAux ax = *reinterpret_cast<const ExhAux*>(vk + e * sizeof(ExhVkRec<S>) + S * sizeof(float));
    float best = __int_as_float(0x7f800000);
    int best_s = 0;
    bool bad = false;
    for (int pass = 0; pass < PASSES; ++pass) {
      float acc[WP];
#pragma unroll
      for (int d = 0; d < WP; ++d) acc[d] = 0.f;
      if (row_lane) window_fma<S, WP>(reinterpret_cast<const float4*>(sc_hat + e * (R * S) + lane * S), qtable + lane * PITCH, pass * WP, acc);
      int d_mine;
      const float total = transpose_reduce<WP>(acc, row_lane, lane, &d_mine);
      float dist = __int_as_float(0x7f800000);
      const int sft = pass * WP + d_mine;
      if (d_mine < WP) {
        const int n = valid_pairs<S>(qmask, ax.vmask, sft);
        if (n > 0) dist = 1.0f - __fdividef(total, (float)n);
      }
      const bool nan_here = !(dist == dist);
      bad |= __any_sync(FULL, nan_here);
      // minimum over this pass's shifts, smallest shift on ties: order-encoded distance in the high bits, shift in the low ones
      const unsigned long long key = ((unsigned long long)(unsigned)(enc_ord(nan_here ? __int_as_float(0x7f800000) : dist) ^ 0x80000000) << 32) | (unsigned)sft;
      unsigned long long k = key;
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(FULL, k, o);
        k = t < k ? t : k;
      }
      const float pd = dec_ord((int)((unsigned)(k >> 32) ^ 0x80000000));
      if (pd < best) {
        best = pd;
        best_s = (int)(k & 0xffffffffu);
      }
    }
    if (lane == 0) {
      float out = best < 0.f ? 0.f : best;
      if (bad || (ax.flags & 1u) || q_flag) out = -1.0f;
      d32[(size_t)qi * d32_pitch + e] = out;
      if (shift_out) shift_out[(size_t)qi * d32_pitch + e] = (unsigned)best_s;
      if (out >= 0.f) my_min = min(my_min, __float_as_uint(out));
    }
  }
  if (lane == 0 && my_min != 0x7f800000u) atomicMin(min_bits + qi, my_min);
}

}  // namespace scgpu
