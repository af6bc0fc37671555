// scgpu_exhaustive.cuh -- exhaustive loop search: EVERY database entry scored against the query
// (BASELINE config 4: 100k keyframes, 20x60) at HBM speed, with the reference's exact result.
//
// Two phases.
//   screen  (k_exh_screen, this file): an FP32 pass that streams the database once from HBM through a TMA-fed
//           shared-memory ring and produces, per entry, an approximation d32 of distanceBtnScanContext
//           (SC.cpp:116-148) with |d32 - d| <= EXH_EPS -- or a flag saying "cannot tell" (-1).
//   rescore (k_exh_compact / k_score / k_exh_final): the entries that can still be the argmin
//           (d32 <= min d32 + 2*EXH_EPS, plus every flagged entry) go through the bit-exact FP64 pair kernel; the
//           winner is the strict minimum in index order, exactly the loop of SC.cpp:296-311 over all entries.
//
// Screening math (the "query-column x candidate-column contraction" of SURVEY.md 8a): with unit-normalised
// columns A^ (query) and B^ (candidate; zero columns stay zero), distDirectSC at shift s is
//       1 - (1/n_s) * sum_j sum_r A^[r][j] * B^[r][(j - s) mod S],     n_s = #{ j : both columns non-zero }
// so a lane that owns ROW r loads B^[r][0..S) of the candidate into registers (the screening copy of the database is
// row-major: 15 conflict-free LDS.128), streams A^[r][*] of the query from a DOUBLED row table in shared memory
// (position p + shift never wraps, so the loads use immediate offsets -- no index arithmetic) through a
// (2*RAD+1)-deep register window and accumulates all 2*RAD+1 shifts of the reference's search window at once:
// one LDS per 2*RAD+1 FMAs.
// The sector-key alignment (SC.cpp:93-113) is argmax_s sum_p v2[p] * v1[(p + s) mod S]; it has exactly the same
// shape, so the lanes of the warp that own no row compute it -- for the NEXT entry, while the row lanes work on
// the current one (software pipeline across entries inside a warp; nothing but warp shuffles in between).
//
// Exactness: the alignment is accepted only when the best correlation beats the runner-up by more than the
// FP32 error bound (otherwise -> flag); d32 NaN -> flag.  Error budget in DESIGN.md ("exhaustive screening").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scgpu_kernels.cuh"

namespace scgpu {

constexpr float EXH_EPS = 1.0e-5f;         // |d32 - d| bound used for candidate selection (observed: < 2e-6)
constexpr float EXH_ALIGN_MARGIN = 1.6e-5f;  // relative (to |v1||v2|) gap below which the alignment is ambiguous

// ---- hand-written PTX wrappers: mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the warp sleeps in hardware instead of burning issue slots in a poll loop
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SCGPU_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra SCGPU_DONE;\n\t"
      "bra SCGPU_WAIT;\n\t"
      "SCGPU_DONE:\n\t"
      "}"
      :
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// per-entry auxiliary record kept next to the normalised descriptor (32 bytes: TMA copies need multiples of 16)
struct ExhAux {
  unsigned long long vmask[2];  // bit c: column c has a non-zero norm (up to 128 sectors)
  float vnorm;                  // |sector key| (float)
  unsigned flags;               // bit 0: a column norm is not representable / not finite in FP32 -> always rescore
  unsigned long long pad;
};

// float sector key + aux of one entry, adjacent in memory so that ONE bulk copy fetches both
template <int S>
struct ExhVkRec {
  float vkey[S];
  ExhAux aux;
};
__host__ __device__ __forceinline__ size_t exh_vk_bytes(int S) { return (size_t)S * sizeof(float) + sizeof(ExhAux); }

struct ExhDb {
  const float* sc_hat;       // [cap][R][S] unit-normalised columns, ROW-major per entry, pair-interleaved (window_fma)
  const unsigned char* vk;   // [cap] ExhVkRec<S>
};

// query pack built by k_exh_prep: normalised query, float sector key, valid-column mask, |v1|
struct ExhQuery {
  float qhat[40 * 128];  // row-major like the database copy; only R*S used
  float v1[128];
  unsigned long long qmask[2];
  float v1norm;
  unsigned flags;
};

// storage position of column c in the pair-interleaved screening copy (see window_fma)
__host__ __device__ __forceinline__ int pair_pos(int c, int S) { return c < S / 2 ? 2 * c : 2 * (c - S / 2) + 1; }

// Screening side data of one descriptor (stored entry or query record).  One block; every thread calls.
template <bool PAIRED>  // PAIRED: pair-interleaved column order (the database side of window_fma)
__device__ __forceinline__ void exh_normalise(const float* sc, const double* sector, const double* norm, const Layout& L, float* sc_hat,
                                              float* vkey32, unsigned long long* vmask, float* vnorm, unsigned* flags) {
  __shared__ unsigned long long s_mask[2];
  __shared__ unsigned s_flags;
  __shared__ float s_v2;
  __shared__ float s_inv[128];
  if (threadIdx.x == 0) {
    s_mask[0] = s_mask[1] = 0;
    s_flags = 0;
    s_v2 = 0.f;
  }
  __syncthreads();
  {
    // per column: 1 / norm, valid-column mask, |sector key|^2, "not a clean float" flag -- warp votes and a shuffle sum, one atomic
    // per warp (a same-address shared-memory atomic per column held the whole block at the barrier below)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    for (int c0 = warp * 32; c0 < L.S; c0 += nwarps * 32) {
      const int c = c0 + lane;
      const bool in = c < L.S;
      const double n = in ? norm[c] : 0.0;
      const float nf = (float)n, v = in ? (float)sector[c] : 0.f;
      if (in) {
        s_inv[c] = (n == 0.0) ? 0.f : (float)(1.0 / n);
        vkey32[PAIRED ? pair_pos(c, L.S) : c] = v;
      }
      const unsigned nz = __ballot_sync(FULL, in && n != 0.0);
      const bool bad = in && ((n != 0.0 && !(nf > 1e-30f && nf < 1e30f)) || !(fabsf(v) < 1e30f));  // subnormal-ish, inf or NaN
      const bool any_bad = __any_sync(FULL, bad);
      float v2 = v * v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v2 += __shfl_xor_sync(FULL, v2, o);
      if (lane == 0) {
        if (nz) atomicOr(&s_mask[c0 >> 6], (unsigned long long)nz << (c0 & 63));
        if (any_bad) atomicOr(&s_flags, 1u);
        atomicAdd(&s_v2, v2);
      }
    }
  }
  __syncthreads();
  // ROW-major output: element (r, c) at r*S + c.  Walk the INPUT (column-major, coalesced loads, four in flight per thread) and
  // scatter the stores -- the other way round every iteration waited for a strided load.
  for (int i0 = threadIdx.x; i0 < L.RS; i0 += blockDim.x * 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = i < L.RS ? sc[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i >= L.RS) break;
      const int c = i / L.R, r = i - c * L.R;
      sc_hat[r * L.S + (PAIRED ? pair_pos(c, L.S) : c)] = v[u] * s_inv[c];
    }
  }
  if (threadIdx.x == 0) {
    vmask[0] = s_mask[0];
    vmask[1] = s_mask[1];
    *vnorm = sqrtf(s_v2);
    *flags = s_flags;
  }
  __syncthreads();
}

// database side of the screening data: local entries [first_local, first_local + gridDim.x), derived from the stored
// descriptor / sector key / column norms (built lazily, right before the first search that needs them)
__global__ void __launch_bounds__(128) k_exh_append(Layout L, Db db, float* sc_hat, unsigned char* vk, unsigned long long first_local) {
  const unsigned long long l = first_local + blockIdx.x;
  float* vkey32 = reinterpret_cast<float*>(vk + l * exh_vk_bytes(L.S));
  ExhAux* aux = reinterpret_cast<ExhAux*>(vkey32 + L.S);
  exh_normalise<true>(db.sc + l * L.RS, db.sector + l * L.S, db.colnorm + l * L.S, L, sc_hat + l * L.RS, vkey32, aux->vmask, &aux->vnorm,
                      &aux->flags);
}

// query pack + reset of the per-query reduction cells (one launch instead of three)
__global__ void __launch_bounds__(256) k_exh_prep(const unsigned char* qrecs, Layout L, ExhQuery* qs, unsigned* min_bits, unsigned* count) {
  const unsigned q = blockIdx.x;  // one block per query of the batch
  if (threadIdx.x == 0) {
    min_bits[q] = 0x7f800000u;
    if (q == 0) *count = 0;
  }
  ExhQuery* dst = qs + q;
  const unsigned char* rec = qrecs + (size_t)q * L.rec_bytes;
  exh_normalise<false>(reinterpret_cast<const float*>(rec), reinterpret_cast<const double*>(rec + L.off_sector),
                reinterpret_cast<const double*>(rec + L.off_norm), L, dst->qhat, dst->v1, dst->qmask, &dst->v1norm, &dst->flags);
}

struct ExhScreenParams {     // grid (blocks, queries): blockIdx.y selects the query of the batch
  ExhDb db;
  const ExhQuery* q;                   // [nq]
  const unsigned long long* n_local;   // [nq] local entries to score per query: [0, n_local[q])
  unsigned long long d32_pitch;        // elements between the d32 rows of consecutive queries
  float* d32;                          // [rows][d32_pitch] out: approx distance; -1 = must be rescored; +inf = can never win
  unsigned* min_bits;                  // [nq] out: bit pattern of the smallest certain d32 (atomicMin; pre-set to +inf)
  int flip_mode;                       // 0: row = query, forward.  1: row = 2*query + f, f = 1 scores column-reversed candidates
};

// ---- pieces shared by k_exh_screen and k_cand_screen -------------------------------------------------------------
// Packed FP32 (sm_100: FFMA2 -- two independent FMAs per issued instruction on an aligned register pair).
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Pairing.  A lane's sum over the S columns is split into the halves p and p + S/2 (H = S/2):
//     acc[d] = sum_{p<H}  held[p] * q[base+p+d]  +  held[p+H] * q[base+p+H+d]
// so with the screening copy of the database stored PAIR-INTERLEAVED (position 2p = column p, 2p+1 = column p+H) and a
// query table of pairs  T[i] = (q[i mod S], q[(i+H) mod S])  every operand is an aligned 8-byte register pair / 8-byte
// shared-memory load for any base and shift: W FFMA2 + one LDS.64 per two columns (scalar form: 2W FFMA + two LDS.32).
// Rows 0..R-1 of the table are the unit-normalised query rows, row R the float sector key; NP pairs per row with NP odd,
// so the 8-byte loads of 16 consecutive rows fall into 16 distinct bank pairs.
__host__ __device__ constexpr int qtab_pairs(int S, int W) { return (3 * S / 2 + W) | 1; }
template <int R, int S, int W>
constexpr size_t qtab_bytes() {
  return (size_t)(R + 1) * qtab_pairs(S, W) * 2 * sizeof(float);
}
// fill by the whole block; value(r, c) = element c of table row r.  Source-driven: each of the (R+1)*S values is
// produced once (consecutive threads take consecutive r: the query record is column-major) and stored to the <= 4
// table positions that hold it.
template <int R, int S, int W, class F>
__device__ __forceinline__ void qtab_fill(float* qtable, F value) {
  constexpr int NP = qtab_pairs(S, W), H = S / 2;
  static_assert(NP <= 2 * S, "a value appears at most twice per half");
  // four values per thread per round: the loads behind value() (global memory: the query record) are all in flight before the
  // first store (k_cand_screen builds a table per query for ten candidates: one load at a time was a quarter of that kernel)
  constexpr int U = 4, N = (R + 1) * S;
  for (int t0 = threadIdx.x; t0 < N; t0 += blockDim.x * U) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * blockDim.x;
      const int c = t / (R + 1), r = t - c * (R + 1);
      v[u] = t < N ? value(r, c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * blockDim.x;
      if (t >= N) break;
      const int c = t / (R + 1), r = t - c * (R + 1);
      float* row = qtable + r * 2 * NP;
      const int i1 = c < H ? c + H : c - H;  // pairs whose second half is column c
      row[2 * c] = v[u];
      if (c + S < NP) row[2 * (c + S)] = v[u];
      row[2 * i1 + 1] = v[u];
      if (i1 + S < NP) row[2 * (i1 + S) + 1] = v[u];
    }
  }
}

// acc[d] += sum_p held[p] * q[base + p + d]   (held4: a pair-interleaved row of S floats, read with 16-byte loads;
// qrow: the lane's row of the pair table).  REV: the held sequence is read back to front (a candidate with its columns
// reversed -- the "flipped" search): the pair (held[H-1-p], held[S-1-p]) meets (q[base+p+H+d], q[base+p+d]) =
// T[base+H+p+d], i.e. the same table entered H later with the held pairs walked downwards.
template <int S, int W, bool REV = false>
__device__ __forceinline__ void window_fma(const float4* held4, const float* qrow, int base, float (&acc)[W]) {
  static_assert(S % 4 == 0, "rows are read with 16-byte loads");
  constexpr int H = S / 2;
  unsigned long long held[H];
#pragma unroll
  for (int i = 0; i < S / 4; ++i) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(held4)[i];
    held[2 * i] = v.x;
    held[2 * i + 1] = v.y;
  }
  if (REV) base = base + H >= S ? base + H - S : base + H;
  const unsigned long long* qp = reinterpret_cast<const unsigned long long*>(qrow) + base;
  unsigned long long acc2[W], win[W];
#pragma unroll
  for (int d = 0; d < W; ++d) acc2[d] = 0ull;
#pragma unroll
  for (int d = 0; d < W - 1; ++d) win[d] = qp[d];
#pragma unroll
  for (int pp = 0; pp < H; ++pp) {
    win[W - 1] = qp[pp + W - 1];
#pragma unroll
    for (int d = 0; d < W; ++d) acc2[d] = fma2(held[REV ? H - 1 - pp : pp], win[d], acc2[d]);
#pragma unroll
    for (int d = 0; d < W - 1; ++d) win[d] = win[d + 1];
  }
#pragma unroll
  for (int d = 0; d < W; ++d) acc[d] += sum2(acc2[d]);
}

// Sum acc[0..W) over the lanes flagged `contributes` with a transpose-reduce (NV -> NV/2 -> ... -> 1 values per lane
// while summing across xor 16, 8, ...); returns the total that belongs to this lane's shift index *d_mine.
template <int W>
__device__ __forceinline__ float transpose_reduce(const float (&acc)[W], bool contributes, int lane, int* d_mine) {
  static_assert(W <= 16, "at most 16 shifts per window");
  if (W <= 8) {
    float r8[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) r8[d] = (contributes && d < W) ? acc[d < W ? d : 0] : 0.f;
    float r4[4], r2[2], r1;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) r4[i] = (h16 ? r8[4 + i] : r8[i]) + __shfl_xor_sync(FULL, h16 ? r8[i] : r8[4 + i], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) r2[i] = (h8 ? r4[2 + i] : r4[i]) + __shfl_xor_sync(FULL, h8 ? r4[i] : r4[2 + i], 8);
    r1 = (h4 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, h4 ? r2[0] : r2[1], 4);
    r1 += __shfl_xor_sync(FULL, r1, 2);
    r1 += __shfl_xor_sync(FULL, r1, 1);
    *d_mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    return r1;
  } else {
    float r16[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) r16[d] = (contributes && d < W) ? acc[d < W ? d : 0] : 0.f;
    float r8[8], r4[4], r2[2], r1;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r8[i] = (h16 ? r16[8 + i] : r16[i]) + __shfl_xor_sync(FULL, h16 ? r16[i] : r16[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) r4[i] = (h8 ? r8[4 + i] : r8[i]) + __shfl_xor_sync(FULL, h8 ? r8[i] : r8[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) r2[i] = (h4 ? r4[2 + i] : r4[i]) + __shfl_xor_sync(FULL, h4 ? r4[i] : r4[2 + i], 4);
    r1 = (h2 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, h2 ? r2[0] : r2[1], 2);
    r1 += __shfl_xor_sync(FULL, r1, 1);
    *d_mine = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    return r1;
  }
}

// warp-wide max / min of order-encoded floats: one CREDUX instead of a five-round shuffle butterfly
__device__ __forceinline__ int enc_ord(float f) {  // order-preserving float -> int (enc_float without the branch)
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float dec_ord(int e) { return __int_as_float(e ^ ((e >> 31) & 0x7fffffff)); }

// valid-column mask of a candidate with its columns reversed: bit c <-> bit S-1-c
template <int S>
__device__ __forceinline__ void reverse_mask(unsigned long long (&m)[2]) {
  if (S <= 64) {
    m[0] = __brevll(m[0]) >> (64 - S);
  } else {
    typedef unsigned __int128 u128;
    const u128 r = (((u128)__brevll(m[0])) << 64) | (u128)__brevll(m[1]);  // 128-bit reversal
    const u128 v = r >> (128 - S);
    m[0] = (unsigned long long)v;
    m[1] = (unsigned long long)(v >> 64);
  }
}

// number of column pairs (query column j, candidate column (j - sft) mod S) with both columns non-zero
template <int S>
__device__ __forceinline__ int valid_pairs(const unsigned long long (&qmask)[2], const unsigned long long (&vmask)[2], int sft) {
  if (S <= 64) {
    const unsigned long long m = vmask[0];
    const unsigned long long rot = sft == 0 ? m : (((m << sft) | (m >> (S - sft))) & ((S == 64) ? ~0ull : ((1ull << (S & 63)) - 1)));
    return __popcll(qmask[0] & rot);
  } else {
    typedef unsigned __int128 u128;
    const u128 m = ((u128)vmask[1] << 64) | vmask[0];
    const u128 full = (S == 128) ? ~(u128)0 : ((((u128)1) << (S & 127)) - 1);
    const u128 rot = sft == 0 ? m : (((m << sft) | (m >> (S - sft))) & full);
    return __popcll(qmask[0] & (unsigned long long)rot) + __popcll(qmask[1] & (unsigned long long)(rot >> 64));
  }
}

// The screened distance of one entry from the per-shift totals: every lane evaluates "its" shift (acc index d belongs
// to shift a_cur - RAD + d), then a warp minimum.  Returns the value to store (-1 = rescore, +inf = cannot win).
template <int S, int RAD>
__device__ __forceinline__ float screened_distance(float total, int d_mine, int a_cur, const unsigned long long (&qmask)[2], const ExhAux& ax,
                                                   bool ambiguous, bool q_flag) {
  constexpr int W = 2 * RAD + 1;
  float dist = __int_as_float(0x7f800000);  // +inf: no valid column pair at this shift -> the reference yields NaN there
  bool nan_here = false;
  if (d_mine < W) {
    int sft = a_cur + d_mine - RAD;  // a_cur in [0, S), d_mine - RAD in [-RAD, RAD]: one conditional wrap, no division
    sft = sft < 0 ? sft + S : (sft >= S ? sft - S : sft);
    const int n = valid_pairs<S>(qmask, ax.vmask, sft);
    if (n > 0) {
      dist = 1.0f - __fdividef(total, (float)n);  // n <= 128: the approximate reciprocal costs < 2.4e-7 absolute
      nan_here = !(dist == dist);
    }
  }
  const bool any_nan = __any_sync(FULL, nan_here);
  const float best = dec_ord(__reduce_min_sync(FULL, enc_ord(nan_here ? __int_as_float(0x7f800000) : dist)));
  if (ambiguous || any_nan || (ax.flags & 1u) || q_flag) return -1.0f;
  return best < 0.f ? 0.f : best;  // tiny negative from rounding stays a valid "certain" value
}

// argmax over the alignment lanes' correlations acc[d] = corr(base + d): best shift (smallest on ties) and whether the
// runner-up is within the FP32 error bound (ambiguous -> the exact path must decide).  `base` must grow with the lane
// index (ties across lanes go to the lowest lane).  Branch-free top-2 per lane, then two CREDUX, a ballot and a shuffle.
template <int S, int W>
__device__ __forceinline__ void align_argmax(const float (&acc)[W], bool has, int base, float margin_scale, int* a_out, bool* amb_out) {
  constexpr int NONE = (int)0x807fffff;  // enc_ord(-inf)
  int e1 = NONE, e2 = NONE, s1 = 0x7fffffff;
#pragma unroll
  for (int d = 0; d < W; ++d) {
    const int s = base + d;
    const int e = (has && s < S) ? enc_ord(acc[d]) : NONE;
    const bool gt = e > e1;  // strict: the earlier (smaller) shift keeps a tie
    e2 = max(e2, gt ? e1 : e);
    s1 = gt ? s : s1;
    e1 = max(e1, e);
  }
  const int E1 = __reduce_max_sync(FULL, e1);
  const unsigned winners = __ballot_sync(FULL, e1 == E1);
  const int wl = __ffs(winners) - 1;
  const int S1 = __shfl_sync(FULL, s1, wl);
  const int E2 = __reduce_max_sync(FULL, (int)(threadIdx.x & 31) == wl ? e2 : e1);
  const float b1 = dec_ord(E1), b2 = dec_ord(E2);
  *a_out = (S1 == 0x7fffffff) ? 0 : S1;
  *amb_out = !((b1 - b2) > EXH_ALIGN_MARGIN * margin_scale) || !(b1 == b1);
}

// Shared memory: every consumer warp owns a private ring -- two descriptors, two (sector key + aux) records -- and the
// four mbarriers that guard it.  A warp feeds its own ring: lane 0 issues the TMA bulk copies for the entries the warp
// will need next as soon as the warp has finished reading a slot, so there is no producer warp, no "slot empty"
// barrier and no block-wide convoy (the first version staged groups of EW entries per slot: every warp waited for the
// whole 96 KB group and the next group could not be requested before the slowest warp had released the slot).
template <int R, int S>
struct ExhWarpRing {
  float sc_hat[2][R * S];
  ExhVkRec<S> vk[2];
};

template <int R, int S, int RAD, int EW, int WPE = 1>
constexpr size_t exh_smem_bytes() {
  return sizeof(ExhWarpRing<R, S>) * EW + (size_t)EW * 4 * sizeof(uint64_t) + qtab_bytes<R, S, 2 * RAD + 1>() +
         (WPE > 1 ? (size_t)EW * 2 * WPE * 16 * sizeof(float) : 0);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// R x S descriptor, search radius RAD, RPL descriptor rows per lane, EW entry streams (rings) per block, WPE warps per
// entry.  WPE = 2 (the 40 x 120 instantiation: a 19.2 KB descriptor leaves room for only four rings per SM): the two warps
// of a team share a ring, each takes half of the rows, both compute the (cheap) alignment redundantly, and they exchange
// their per-shift partial sums through shared memory around ONE named barrier per entry -- after which the slots are
// free, so the same barrier also orders the next TMA request.  Twice the warps per SM for the same shared memory.
template <int R, int S, int RAD, int RPL, int EW, int WPE = 1>
__global__ void __launch_bounds__(EW * WPE * 32, 1) k_exh_screen(const ExhScreenParams pp) {
  // this block's query
  struct {
    ExhDb db;
    const ExhQuery* q;
    unsigned long long n_local;
    float* d32;
    unsigned* min_bits;
  } p;
  const unsigned qi = pp.flip_mode ? blockIdx.y >> 1 : blockIdx.y;
  const bool rev = pp.flip_mode && (blockIdx.y & 1);
  p.db = pp.db;
  p.q = pp.q + qi;
  p.n_local = pp.n_local[qi];
  p.d32 = pp.d32 + blockIdx.y * pp.d32_pitch;
  p.min_bits = pp.min_bits + qi;
  constexpr int W = 2 * RAD + 1;
  constexpr int ROWS_PER_WARP = R / WPE;
  constexpr int ROW_LANES = ROWS_PER_WARP / RPL;
  constexpr int ALIGN_LANES = (S + W - 1) / W;
  constexpr int PITCH = 2 * qtab_pairs(S, W);  // floats per table row
  static_assert(R % WPE == 0 && ROWS_PER_WARP % RPL == 0 && ROW_LANES + ALIGN_LANES <= 32, "rows + alignment lanes must fit one warp");
  static_assert(S <= 128 && S % 4 == 0, "valid-column masks are 128 bits; rows are read with 16-byte loads");
  static_assert(sizeof(ExhWarpRing<R, S>) % 16 == 0, "TMA destinations are 16-byte aligned");
  static_assert(WPE == 1 || EW + 1 <= 15, "one named barrier per team");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / WPE, half = warp % WPE;  // team = entry stream; half = which share of the rows
  ExhWarpRing<R, S>& ring = reinterpret_cast<ExhWarpRing<R, S>*>(smem_raw)[team];
  uint64_t* full_sc = reinterpret_cast<uint64_t*>(smem_raw + sizeof(ExhWarpRing<R, S>) * EW) + team * 4;  // [2]
  uint64_t* full_vk = full_sc + 2;                                                                        // [2]
  float* qtable = reinterpret_cast<float*>(smem_raw + sizeof(ExhWarpRing<R, S>) * EW + (size_t)EW * 4 * sizeof(uint64_t));  // pair table
  float* xbuf = qtable + (R + 1) * PITCH + team * 2 * WPE * 16;  // [2][WPE][16] partial sums of the team (WPE > 1)

  // consecutive teams of the grid take consecutive entries: team gw scores gw, gw + TW, gw + 2 TW, ...
  const unsigned long long TW = (unsigned long long)gridDim.x * EW, gw = (unsigned long long)blockIdx.x * EW + team;
  const unsigned long long my_n = p.n_local > gw ? (p.n_local - 1 - gw) / TW + 1 : 0;
  auto issue_sc = [&](unsigned long long j) {  // descriptor of my j-th entry -> slot j & 1
    const int slot = (int)(j & 1);
    mbar_arrive_expect_tx(&full_sc[slot], R * S * 4u);
    tma_bulk_g2s(&ring.sc_hat[slot][0], p.db.sc_hat + (gw + j * TW) * (R * S), R * S * 4u, &full_sc[slot]);
  };
  auto issue_vk = [&](unsigned long long j) {  // sector key + aux of my j-th entry -> slot j & 1
    const int slot = (int)(j & 1);
    mbar_arrive_expect_tx(&full_vk[slot], (unsigned)sizeof(ExhVkRec<S>));
    tma_bulk_g2s(&ring.vk[slot], p.db.vk + (gw + j * TW) * sizeof(ExhVkRec<S>), (unsigned)sizeof(ExhVkRec<S>), &full_vk[slot]);
  };
  if (lane == 0 && half == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full_sc[s], 1);
    fence_barrier_init();
    if (my_n > 0) {  // the copies do not touch the query table: start them before it is built
      issue_vk(0);
      issue_sc(0);
    }
    if (my_n > 1) issue_vk(1);
  }
  {
    const ExhQuery* qq = p.q;
    qtab_fill<R, S, W>(qtable, [qq](int r, int c) { return r < R ? qq->qhat[r * S + c] : qq->v1[c]; });
  }
  __syncthreads();

  // lane roles: [0, ROW_LANES) own RPL descriptor rows each; [ROW_LANES, ROW_LANES+ALIGN_LANES) own W alignment shifts
  const bool row_lane = lane < ROW_LANES, align_lane = lane >= ROW_LANES && lane < ROW_LANES + ALIGN_LANES;
  const unsigned long long qmask[2] = {p.q->qmask[0], p.q->qmask[1]};
  const float v1norm = p.q->v1norm;
  const bool q_flag = (p.q->flags & 1u) != 0;

  int a_cur = 0;          // alignment of the entry whose window is scored in this iteration
  bool amb_cur = false;
  ExhAux ax_cur;          // ... and its aux record (copied out of the ring one iteration earlier)
  ax_cur.vmask[0] = ax_cur.vmask[1] = 0;
  ax_cur.vnorm = 0.f;
  ax_cur.flags = 0;
  unsigned my_min = 0x7f800000u;  // lane 0: smallest certain distance this warp has seen
  for (unsigned long long k = 0; k <= my_n; ++k) {
    // iteration k: window of my entry k-1 (row lanes) + alignment of my entry k (alignment lanes)
    const bool has_win = k >= 1, has_al = k < my_n;
    const int sc_w = (int)((k + 1) & 1), vk_a = (int)(k & 1);
    if (has_win) mbar_wait(&full_sc[sc_w], (uint32_t)(((k - 1) >> 1) & 1));
    if (has_al) mbar_wait(&full_vk[vk_a], (uint32_t)((k >> 1) & 1));
    float acc[W];
#pragma unroll
    for (int d = 0; d < W; ++d) acc[d] = 0.f;
    // One instruction stream for both lane roles (row lanes: window of entry k-1; alignment lanes: correlation of
    // entry k): each lane only differs in WHERE its held values and its query row come from.
    //   acc[d] = sum over my rows r, columns p of held_r[p] * q_r[(p + base + d) mod S]
    // Selects, not branches: with an if / else-if here the compiler lets the two lane groups run the window code one
    // after the other (every FFMA2 issued twice per entry).
    const bool rw = row_lane && has_win, al = align_lane && has_al;
    const int base = rw ? (a_cur >= RAD ? a_cur - RAD : a_cur - RAD + S) : (al ? (lane - ROW_LANES) * W : 0);
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      const bool on = rw || (i == 0 && al);
      const int r = rw ? half * ROWS_PER_WARP + lane + i * ROW_LANES : R;
      const float* held = rw ? &ring.sc_hat[sc_w][r * S] : &ring.vk[vk_a].vkey[0];
      const float* qrow = qtable + r * PITCH;
      __syncwarp();
      if (on) {
        if (rev) window_fma<S, W, true>(reinterpret_cast<const float4*>(held), qrow, base, acc);
        else window_fma<S, W, false>(reinterpret_cast<const float4*>(held), qrow, base, acc);
      }
    }
    if (WPE == 1) {
      // ---- window result of entry k-1 ------------------------------------------------------------------
      if (has_win) {
        int d_mine;
        const float total = transpose_reduce<W>(acc, row_lane, lane, &d_mine);
        ExhAux ax = ax_cur;
        if (rev) reverse_mask<S>(ax.vmask);
        const float out = screened_distance<S, RAD>(total, d_mine, a_cur, qmask, ax, amb_cur, q_flag);
        if (lane == 0) {
          p.d32[gw + (k - 1) * TW] = out;
          if (out >= 0.f) my_min = min(my_min, __float_as_uint(out));
        }
      }
      // ---- alignment result of entry k (becomes a_cur of the next iteration) -------------------------------
      if (has_al) {
        ax_cur = ring.vk[vk_a].aux;
        align_argmax<S, W>(acc, align_lane, base, v1norm * ax_cur.vnorm, &a_cur, &amb_cur);
      }
      // ---- both slots read in this iteration are free: request what they hold next ---------------------------
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();  // the warp's generic-proxy reads of the slots precede the async-proxy writes
        if (k + 1 < my_n) issue_sc(k + 1);  // -> slot sc_w (k = 0: the still unused second slot)
        if (k + 2 < my_n) issue_vk(k + 2);  // -> slot vk_a
      }
    } else {
      // ---- team of WPE warps: publish my partial sums, meet the others ONCE, then everything runs on registers ----
      ExhAux ax_new = ax_cur;
      if (has_al) ax_new = ring.vk[vk_a].aux;  // last read of the sector-key slot
      int d_mine = 0;
      float total = 0.f;
      float* xb = xbuf + (int)(k & 1) * WPE * 16;
      if (has_win) {
        total = transpose_reduce<W>(acc, row_lane, lane, &d_mine);
        xb[half * 16 + d_mine] = total;  // (lanes sharing a shift index write the same value)
      }
      __syncwarp();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(WPE * 32) : "memory");
      if (lane == 0 && half == 0) {  // every warp of the team has finished reading both slots
        fence_proxy_async();
        if (k + 1 < my_n) issue_sc(k + 1);
        if (k + 2 < my_n) issue_vk(k + 2);
      }
      if (has_win) {
        total = 0.f;
#pragma unroll
        for (int w = 0; w < WPE; ++w) total += xb[w * 16 + d_mine];  // same order in every warp of the team
        ExhAux ax = ax_cur;
        if (rev) reverse_mask<S>(ax.vmask);
        const float out = screened_distance<S, RAD>(total, d_mine, a_cur, qmask, ax, amb_cur, q_flag);
        if (lane == 0 && half == 0) {
          p.d32[gw + (k - 1) * TW] = out;
          if (out >= 0.f) my_min = min(my_min, __float_as_uint(out));
        }
      }
      ax_cur = ax_new;
      if (has_al) align_argmax<S, W>(acc, align_lane, base, v1norm * ax_cur.vnorm, &a_cur, &amb_cur);
    }
  }
  if (lane == 0 && my_min != 0x7f800000u) atomicMin(p.min_bits, my_min);
}

// ---- rescoring side -----------------------------------------------------------------------------------
// candidates = flagged entries + entries within 2*EXH_EPS of the query's smallest certain value -> one flat key list
// (flip << 63 | query index << 32 | global entry index); grid (blocks, rows)
__global__ void k_exh_compact(const float* d32, unsigned long long d32_pitch, const unsigned long long* n_local, const unsigned* min_bits,
                              int rank, int G, int flip_mode, unsigned long long* keys, unsigned* count, unsigned cap, float eps) {
  const unsigned row = blockIdx.y, q = flip_mode ? row >> 1 : row;
  const unsigned long long fbit = (flip_mode && (row & 1)) ? (1ull << 63) : 0ull;
  const float* rowp = d32 + row * d32_pitch;
  const unsigned long long n = n_local[q];
  const float thr = __uint_as_float(min_bits[q]) + 2.0f * eps;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const float v = rowp[i];
    if (v < 0.f || (v <= thr && v < __int_as_float(0x7f800000))) {
      const unsigned slot = atomicAdd(count, 1u);
      if (slot < cap) keys[slot] = fbit | ((unsigned long long)q << 32) | (i * (unsigned long long)G + rank);
    }
  }
}

// per query (one block each): strict minimum in index order over its rescored candidates (SC.cpp:296-311 over every
// entry); rank = total number of rescored candidates of the batch (overflow check on the host)
__global__ void __launch_bounds__(256) k_exh_final(const double* pair_dist, const int* pair_shift, const unsigned long long* keys,
                                                   const unsigned* count, unsigned cap, Best* out) {
  __shared__ Best s_best[256];
  const unsigned q = blockIdx.x;
  Best b;
  b.dist = 10000000.0;
  b.rank = 0;
  b.shift = 0;
  b.idx = -1;
  const unsigned n = min(*count, cap);
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long key = keys[i];
    if ((unsigned)((key >> 32) & 0x7fffffffull) != q) continue;
    const double d = pair_dist[i];
    // order: distance, then entry index, then forward before flipped -> sort key idx2 = idx * 2 + flip
    const long long idx2 = (long long)((key & 0xffffffffull) * 2 + (key >> 63));
    if (pair_shift[i] >= 0 && d < 10000000.0 && (d < b.dist || (d == b.dist && (b.idx < 0 || idx2 < b.idx)))) {
      b.dist = d;
      b.shift = pair_shift[i];
      b.idx = idx2;
    }
  }
  s_best[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const Best x = s_best[threadIdx.x + o];
      Best& y = s_best[threadIdx.x];
      if (x.idx >= 0 && (y.idx < 0 || x.dist < y.dist || (x.dist == y.dist && x.idx < y.idx))) y = x;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    Best r = s_best[0];
    r.rank = (int)min(*count, 0x7fffffffu);
    if (r.idx < 0) {
      r.idx = 0;
      r.shift = 0;
      r.dist = 10000000.0;
    } else {
      r.shift |= (int)(r.idx & 1) << 30;  // flipped winner flagged in bit 30 of the shift
      r.idx >>= 1;
    }
    out[q] = r;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Top-K path: the same FP32 screening applied to the K retrieved candidates of each query (SC.cpp:296-311).
//   k_cand_screen : one block per query (query table built from its record), one warp per candidate slot; the
//                   candidate's screening copy is read straight from global memory (L2-resident) into registers.
//                   The same block then selects (cand_select_warp): candidates whose screened distance is within 2*EXH_EPS
//                   of the smallest certain one (plus every flagged one) go on a list; the others cannot be the minimum
//                   and are marked skipped.
//   k_score_pairs : the exact FP64 pair kernel over that list (typically 1-2 of the K candidates per query).
// The strict-min in candidate order (k_best) then runs over exactly-scored candidates only, so the result is the
// reference's.  scgpu_get_candidates rescoring everything exactly on demand keeps the parity dumps complete.
// ------------------------------------------------------------------------------------------------------------
struct CandScreenParams {
  const unsigned char* qrecords;
  Layout L;
  Db db;                                // ownership (rank, G)
  ExhDb xdb;                            // screening copy
  const unsigned long long* keys;       // [nq][K]
  const unsigned long long* n_search;   // [nq]
  int K;
  float* d32;                           // [nq][K] out: approx distance; -1 = rescore; +inf = cannot win / not mine
  PeerTab peers;                        // peers.G > 0: every candidate is screened here; its screening copy is fetched from the
                                        // owner shard by the same TMA bulk copy, over NVLink peer memory
  // fused selection (k_cand_select's job, done by the block that screened the query): slots that need the exact kernel go on
  // `list` ((query << 32) | slot, counted in *count), the others are marked skipped in pair_dist / pair_shift
  unsigned long long* list;
  unsigned* count;
  double* pair_dist;
  int* pair_shift;
};

// which candidate slots of query q need the exact kernel: flagged ones and those within 2*EXH_EPS of the smallest certain value.
// One warp; d32 of the query's K slots must be visible.
__device__ __forceinline__ void cand_select_warp(const CandScreenParams& p, int q, int lane) {
  const float INF = __int_as_float(0x7f800000);
  float mn = INF;
  for (int k = lane; k < p.K; k += 32) {
    const float v = p.d32[(size_t)q * p.K + k];
    if (v >= 0.f) mn = fminf(mn, v);
  }
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));
  const float thr = mn + 2.0f * EXH_EPS;
  for (int k = lane; k < p.K; k += 32) {
    const size_t o = (size_t)q * p.K + k;
    const float v = p.d32[o];
    if (v < 0.f || (v <= thr && v < INF)) {
      p.list[atomicAdd(p.count, 1u)] = ((unsigned long long)q << 32) | (unsigned)k;
    } else {  // cannot be the minimum (or not this shard's): skipped by k_best
      p.pair_dist[o] = 10000000.0;
      p.pair_shift[o] = -1;
    }
  }
}

// one candidate staged in shared memory: screening copy of the descriptor + its sector key / aux record
template <int R, int S>
struct CandSlot {
  float sc_hat[R * S];
  ExhVkRec<S> vk;
};

template <int R, int S, int RAD, int CW, int NS>
constexpr size_t cand_smem_bytes() {
  return sizeof(CandSlot<R, S>) * CW * NS + (size_t)CW * NS * sizeof(uint64_t) + qtab_bytes<R, S, 2 * RAD + 1>();
}

// CW warps per block, NS staging slots per warp.  Warp w scores candidates w, w + CW, ... of the block's query.  Each
// warp fetches its candidates itself with TMA bulk copies (lane 0; up to NS in flight) -- the first fetch is issued
// before the block builds the query table, so the dependent chain "key -> entry -> sector key -> descriptor rows"
// (three L2 round trips when read with plain loads) overlaps the table build instead of following it.
template <int R, int S, int RAD, int RPL, int CW, int NS>
__global__ void __launch_bounds__(CW * 32) k_cand_screen(const CandScreenParams p) {
  constexpr int W = 2 * RAD + 1;
  constexpr int ROW_LANES = R / RPL;
  constexpr int PITCH = 2 * qtab_pairs(S, W);
  static_assert(R % RPL == 0 && ROW_LANES <= 32 && S <= 128 && S % 4 == 0, "see k_exh_screen");
  static_assert(sizeof(CandSlot<R, S>) % 16 == 0, "TMA destinations are 16-byte aligned");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_inv[S];
  __shared__ unsigned long long s_mask[2];
  __shared__ unsigned s_flags;
  __shared__ float s_v2;
  const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  CandSlot<R, S>* slots = reinterpret_cast<CandSlot<R, S>*>(smem_raw) + warp * NS;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(CandSlot<R, S>) * CW * NS) + warp * NS;
  float* qtable = reinterpret_cast<float*>(smem_raw + sizeof(CandSlot<R, S>) * CW * NS + (size_t)CW * NS * sizeof(uint64_t));
  const bool early = p.n_search[q] == 0;
  const unsigned char* qrec = p.qrecords + (size_t)q * p.L.rec_bytes;
  const float* qsc = reinterpret_cast<const float*>(qrec);
  const double* qsector = reinterpret_cast<const double*>(qrec + p.L.off_sector);
  const double* qnorm = reinterpret_cast<const double*>(qrec + p.L.off_norm);
  const unsigned long long* qkeys = p.keys + (size_t)q * p.K;

  // candidate k of this query: is it scored here, which shard holds it and which local entry (warp-uniform)
  const unsigned long long nshards = p.peers.G ? (unsigned long long)p.peers.G : (unsigned long long)p.db.G;
  auto owned = [&](int k, unsigned long long* l, int* owner) {
    const unsigned long long key = qkeys[k];
    const unsigned long long g = (key == KEY_NONE) ? 0ull : (key & 0xffffffffull);
    *l = g / nshards;
    *owner = (int)(g % nshards);
    return !early && (p.peers.G != 0 || *owner == p.db.rank);
  };
  int issued = 0, consumed = 0, k_issue = warp;  // fetches requested / used so far; next candidate to look at for a fetch
  auto fetch_more = [&]() {
    while (issued - consumed < NS && k_issue < p.K) {
      unsigned long long l;
      int owner;
      if (owned(k_issue, &l, &owner)) {
        if (lane == 0) {
          const int s = issued % NS;
          const float* src_hat = p.peers.G ? p.peers.sc_hat[owner] : p.xdb.sc_hat;
          const unsigned char* src_vk = p.peers.G ? p.peers.vk[owner] : p.xdb.vk;
          mbar_arrive_expect_tx(&full[s], (unsigned)sizeof(CandSlot<R, S>));
          tma_bulk_g2s(slots[s].sc_hat, src_hat + l * (R * S), R * S * 4u, &full[s]);
          tma_bulk_g2s(&slots[s].vk, src_vk + l * sizeof(ExhVkRec<S>), (unsigned)sizeof(ExhVkRec<S>), &full[s]);
        }
        ++issued;
      }
      k_issue += CW;
    }
  };
  if (lane == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  fetch_more();
  // sharded database: a rank that owns none of this query's candidates has nothing to score -- skip the table build
  if (!__syncthreads_or(issued > 0)) {
    for (int k = warp; k < p.K; k += CW)
      if (lane == 0) p.d32[(size_t)q * p.K + k] = __int_as_float(0x7f800000);
    if (p.list) {
      __syncthreads();
      if (warp == 0) cand_select_warp(p, q, lane);
    }
    return;
  }

  if (threadIdx.x == 0) {
    s_mask[0] = s_mask[1] = 0;
    s_flags = 0;
    s_v2 = 0.f;
  }
  __syncthreads();
  if (!early) {
    // per column: 1 / norm, valid-column mask, |sector key|^2, "not a clean float" flag -- warp votes and a shuffle sum, then one
    // atomic per warp (sixty same-address shared-memory atomics per block held every warp at the barrier below)
    for (int c0 = warp * 32; c0 < S; c0 += CW * 32) {
      const int c = c0 + lane;
      const bool in = c < S;
      const double n = in ? qnorm[c] : 0.0;
      const float nf = (float)n, v = in ? (float)qsector[c] : 0.f;
      if (in) s_inv[c] = (n == 0.0) ? 0.f : (float)(1.0 / n);
      const unsigned nz = __ballot_sync(FULL, in && n != 0.0);
      const bool bad = in && ((n != 0.0 && !(nf > 1e-30f && nf < 1e30f)) || !(fabsf(v) < 1e30f));
      const bool any_bad = __any_sync(FULL, bad);
      float v2 = v * v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v2 += __shfl_xor_sync(FULL, v2, o);
      if (lane == 0) {
        if (nz) atomicOr(&s_mask[c0 >> 6], (unsigned long long)nz << (c0 & 63));
        if (any_bad) atomicOr(&s_flags, 1u);
        atomicAdd(&s_v2, v2);
      }
    }
  }
  __syncthreads();
  if (!early) {
    const float* inv = s_inv;
    qtab_fill<R, S, W>(qtable, [qsc, qsector, inv](int r, int c) { return r < R ? qsc[c * R + r] * inv[c] : (float)qsector[c]; });
  }
  __syncthreads();
  const unsigned long long qmask[2] = {s_mask[0], s_mask[1]};
  const float v1norm = sqrtf(s_v2);
  const bool q_flag = s_flags != 0;
  const bool row_lane = lane < ROW_LANES;

  for (int k = warp; k < p.K; k += CW) {
    const size_t o = (size_t)q * p.K + k;
    unsigned long long l;
    int owner;
    if (!owned(k, &l, &owner)) {
      if (lane == 0) p.d32[o] = __int_as_float(0x7f800000);
      continue;
    }
    const CandSlot<R, S>& slot = slots[consumed % NS];
    mbar_wait(&full[consumed % NS], (uint32_t)((consumed / NS) & 1));
    const ExhAux ax = slot.vk.aux;
    // ---- alignment with ALL lanes (nothing else to do in this phase): lane l takes the WA shifts l*WA .. l*WA+WA-1
    constexpr int WA = (S + 31) / 32;
    int a_cur;
    bool amb;
    {
      float ca[WA];
#pragma unroll
      for (int d = 0; d < WA; ++d) ca[d] = 0.f;
      const int ab = lane * WA;
      if (ab < S) window_fma<S, WA>(reinterpret_cast<const float4*>(slot.vk.vkey), qtable + R * PITCH, ab, ca);
      align_argmax<S, WA>(ca, ab < S, ab, v1norm * ax.vnorm, &a_cur, &amb);
    }
    // ---- window on the row lanes: acc[d] belongs to shift a_cur - RAD + d
    float acc[W];
#pragma unroll
    for (int d = 0; d < W; ++d) acc[d] = 0.f;
    const int base = a_cur >= RAD ? a_cur - RAD : a_cur - RAD + S;
    if (row_lane) {
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int r = lane + i * ROW_LANES;
        window_fma<S, W>(reinterpret_cast<const float4*>(&slot.sc_hat[r * S]), qtable + r * PITCH, base, acc);
      }
    }
    // ---- the slot has been read: request the next candidate, then finish this one
    __syncwarp();
    ++consumed;
    if (lane == 0) fence_proxy_async();
    fetch_more();
    int d_mine;
    const float total = transpose_reduce<W>(acc, row_lane, lane, &d_mine);
    const float out = screened_distance<S, RAD>(total, d_mine, a_cur, qmask, ax, amb, q_flag);
    if (lane == 0) p.d32[o] = out;
  }
  if (p.list) {
    __syncthreads();  // every warp's d32 stores are visible to the block
    if (warp == 0) cand_select_warp(p, q, lane);
  }
}

// Stage-2 storage + screening copy in one launch (peer-sharded / batched replays: the screening rows must be current before the
// barrier that lets other shards fetch them).  The screening side data is derived from the RECORD (in hand), not re-read from
// the database.
__global__ void __launch_bounds__(128) k_append_hat(const unsigned char* records, Layout L, Db db, unsigned long long first_global,
                                                    unsigned long long step, PushList push, float* sc_hat, unsigned char* vk) {
  unsigned long long l;
  const unsigned char* rec = records + (size_t)blockIdx.x * L.rec_bytes;
  if (!append_entry(rec, L, db, first_global + blockIdx.x * step, push, &l)) return;
  float* vkey32 = reinterpret_cast<float*>(vk + l * exh_vk_bytes(L.S));
  ExhAux* aux = reinterpret_cast<ExhAux*>(vkey32 + L.S);
  exh_normalise<true>(reinterpret_cast<const float*>(rec), reinterpret_cast<const double*>(rec + L.off_sector),
                      reinterpret_cast<const double*>(rec + L.off_norm), L, sc_hat + l * L.RS, vkey32, aux->vmask, &aux->vnorm, &aux->flags);
}

// the exact pair kernel over a (query, slot) list; persistent grid
__global__ void __launch_bounds__(128) k_score_pairs(const ScoreParams p, const unsigned long long* list, const unsigned* count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned n = *count;
  for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
    const unsigned long long e = list[i];
    score_pair<false>(p, (int)(e & 0xffffffffull), (int)(e >> 32), smem_raw);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_build_tma: k_build with the point stream staged through shared memory by TMA bulk copies (UBLKCP).
// An elected thread keeps BUILD_STAGES chunks of BUILD_CHUNK points in flight per block, so HBM latency is covered by
// bytes in flight in the async proxy instead of by registers / resident warps; every thread then picks its points
// from shared memory with conflict-free 16-byte loads.  Binning, aggregation and the record epilogue are k_build's.
// ------------------------------------------------------------------------------------------------------------
constexpr int BUILD_STAGES = 3;
constexpr int BUILD_UNROLL = 4;
constexpr int BUILD_CHUNK = 256 * BUILD_UNROLL;  // points per chunk
constexpr int BUILD_QCAP = 128;                  // undecided points parked per block (overflow: decided on the spot)

template <int STRIDE, bool FAST, bool LH_FLOAT>  // STRIDE = 12, 16 or 32 (bytes per point, 16-byte aligned scan starts)
__global__ void __launch_bounds__(256) k_build_tma(const BuildParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                                             // [BUILD_STAGES][BUILD_CHUNK * STRIDE]
  int* s_bins = reinterpret_cast<int*>(smem_raw + (size_t)BUILD_STAGES * BUILD_CHUNK * STRIDE);
  __shared__ __align__(8) uint64_t full[BUILD_STAGES];
  __shared__ __align__(8) uint64_t empty[BUILD_STAGES];  // one arrival per warp once it has copied its points out
  __shared__ float s_queue[BUILD_QCAP * 3];               // points the front end could not decide: exact path, converged, at the end
  __shared__ unsigned s_qcount;
  __shared__ bool s_last;
  const int RS = p.L.RS;
  const unsigned scan = blockIdx.y;
  const unsigned start = blockIdx.x * p.pts_per_block;
  const unsigned end = min(start + p.pts_per_block, p.n_pts);
  const unsigned n_chunks = end > start ? (end - start + BUILD_CHUNK - 1) / BUILD_CHUNK : 0;
  const unsigned char* base = p.pts + (unsigned long long)scan * p.scan_pitch + (unsigned long long)start * STRIDE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < BUILD_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 256 / 32);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < RS; i += blockDim.x) s_bins[i] = SCGPU_ENC_NOPOINT;
  if (threadIdx.x == 0) s_qcount = 0;
  const uint32_t bins_addr = smem_u32(s_bins);
  __syncthreads();
  auto issue = [&](unsigned c) {
    const unsigned pts = min((unsigned)BUILD_CHUNK, end - start - c * BUILD_CHUNK);
    const int slot = (int)(c % BUILD_STAGES);
    if (c >= BUILD_STAGES) mbar_wait(&empty[slot], ((c / BUILD_STAGES) - 1) & 1);  // every warp has copied chunk c-STAGES out
    mbar_arrive_expect_tx(&full[slot], pts * STRIDE);
    tma_bulk_g2s(ring + (size_t)slot * BUILD_CHUNK * STRIDE, base + (unsigned long long)c * BUILD_CHUNK * STRIDE, pts * STRIDE, &full[slot]);
  };
  if (threadIdx.x == 0)
    for (unsigned c = 0; c < BUILD_STAGES - 1 && c < n_chunks; ++c) issue(c);

  // chunk loop, unrolled over the ring so that slot numbers (shared-memory offsets, barrier addresses) are constants
  for (unsigned c0 = 0; c0 < n_chunks; c0 += BUILD_STAGES) {
    const unsigned ring_parity = (c0 / BUILD_STAGES) & 1;
#pragma unroll
   for (int slot = 0; slot < BUILD_STAGES; ++slot) {
    const unsigned c = c0 + slot;
    if (c >= n_chunks) break;
    // refill the slot that chunk c-1 occupied (no block-wide barrier: warps only report "copied out" per slot)
    if (threadIdx.x == 0 && c + BUILD_STAGES - 1 < n_chunks) issue(c + BUILD_STAGES - 1);
    mbar_wait(&full[slot], ring_parity);
    const unsigned pts = min((unsigned)BUILD_CHUNK, end - start - c * BUILD_CHUNK);
    const unsigned char* sp = ring + (size_t)slot * BUILD_CHUNK * STRIDE + (size_t)threadIdx.x * STRIDE;
    float px[BUILD_UNROLL], py[BUILD_UNROLL], pz[BUILD_UNROLL];
    // STRIDE 16 / 32: one 16-byte load per point; STRIDE 12 (packed xyz): three 4-byte loads, lane stride 3 words -- both
    // conflict-free
    auto take = [&](int u) {
      if (STRIDE == 12) {
        const float* f = reinterpret_cast<const float*>(sp + (size_t)u * 256 * STRIDE);
        px[u] = f[0];
        py[u] = f[1];
        pz[u] = f[2];
      } else {
        const float4 v = *reinterpret_cast<const float4*>(sp + (size_t)u * 256 * STRIDE);
        px[u] = v.x;
        py[u] = v.y;
        pz[u] = v.z;
        if (STRIDE == 32 && p.val_off != 8)  // intensity descriptor: the value sits in the record's second half (already staged)
          pz[u] = *reinterpret_cast<const float*>(sp + (size_t)u * 256 * STRIDE + p.val_off);
      }
    };
    if (pts == BUILD_CHUNK) {  // full chunk: no bounds checks
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u) take(u);
    } else {
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u) {
        px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);  // NaN: dropped
        if (u * 256 + threadIdx.x < pts) take(u);
      }
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot]);  // this warp's points are in registers
    // front end for all points of the thread first; the (rare) undecided ones go through ONE divergent region
    int bin[BUILD_UNROLL];
    float hh[BUILD_UNROLL];
    bool undecided = false;
#pragma unroll
    for (int u = 0; u < BUILD_UNROLL; ++u) {
      bin[u] = bin_point_fast<LH_FLOAT>(p.bc, px[u], py[u], pz[u], hh[u]);
      undecided |= (bin[u] == BIN_UNDECIDED);
    }
    if (undecided) {  // rare (~3e-4 of the points): park the point; only if the queue is full decide it here
#pragma unroll
      for (int u = 0; u < BUILD_UNROLL; ++u)
        if (bin[u] == BIN_UNDECIDED) {
          const unsigned slot = atomicAdd(&s_qcount, 1u);
          if (slot < (unsigned)BUILD_QCAP) {
            s_queue[3 * slot] = px[u];
            s_queue[3 * slot + 1] = py[u];
            s_queue[3 * slot + 2] = pz[u];
            bin[u] = -1;
          } else {
            const unsigned long long r = bin_point_exact_noinline(p.bc.R, p.bc.S, p.bc.lidar_height, p.bc.max_radius, px[u], py[u], pz[u]);
            bin[u] = (int)(unsigned)(r >> 32);
            hh[u] = __uint_as_float((unsigned)r);
          }
        }
    }
    // max into the block's grid.  A bin's value only ever grows, so a plain read is a valid filter: a point that does
    // not beat the value read cannot beat the current one; after the first few points of a bin almost none does.
    // Branch-free: the four reads first, then one PREDICATED shared-memory reduction per point (as C++ this compiled
    // to a divergent branch per point that also re-derived the shared-memory window address every time).
    int cur[BUILD_UNROLL], enc[BUILD_UNROLL];
    uint32_t addr[BUILD_UNROLL];
#pragma unroll
    for (int u = 0; u < BUILD_UNROLL; ++u) {
      addr[u] = bins_addr + 4u * (uint32_t)max(bin[u], 0);
      enc[u] = enc_float(hh[u]);
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(cur[u]) : "r"(addr[u]));
    }
#pragma unroll
    for (int u = 0; u < BUILD_UNROLL; ++u)
      asm volatile(
          "{\n\t"
          ".reg .pred p, q;\n\t"
          "setp.ge.s32 q, %3, 0;\n\t"
          "setp.gt.and.s32 p, %1, %2, q;\n\t"
          "@p red.shared.max.s32 [%0], %1;\n\t"
          "}" ::"r"(addr[u]),
          "r"(enc[u]), "r"(cur[u]), "r"(bin[u])
          : "memory");
   }
  }
  __syncthreads();
  {  // the parked points: exact path, all lanes busy
    const unsigned nq = min(s_qcount, (unsigned)BUILD_QCAP);
    for (unsigned i = threadIdx.x; i < nq; i += blockDim.x) {
      float h;
      const int b = bin_point_exact(p.bc, s_queue[3 * i], s_queue[3 * i + 1], s_queue[3 * i + 2], h);
      if (b >= 0) atomicMax(&s_bins[b], enc_float(h));
    }
  }
  __syncthreads();

  if (gridDim.x > 1) {
    int* g = p.gbins + (unsigned long long)scan * RS;
    for (int i = threadIdx.x; i < RS; i += blockDim.x) {
      const int v = s_bins[i];
      if (v != SCGPU_ENC_NOPOINT) atomicMax(&g[i], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&p.tickets[scan], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < RS; i += blockDim.x) {
      s_bins[i] = __ldcg(&g[i]);
      g[i] = SCGPU_ENC_NOPOINT;
    }
    if (threadIdx.x == 0) p.tickets[scan] = 0;
    __syncthreads();
  }
  unsigned char* rec = p.records + (unsigned long long)scan * p.L.rec_bytes;
  float* s_sc = reinterpret_cast<float*>(s_bins);
  float* rec_sc = reinterpret_cast<float*>(rec);
  for (int i = threadIdx.x; i < RS; i += blockDim.x) {
    float f = dec_float(s_bins[i]);
    if (f == -1000.0f) f = 0.0f;
    s_sc[i] = f;
    rec_sc[i] = f;
  }
  __syncthreads();
  keys_from_sc<float>(s_sc, p.L.R, p.L.S, p.L.R, nullptr, reinterpret_cast<float*>(rec + p.L.off_ring),
                      reinterpret_cast<double*>(rec + p.L.off_sector), reinterpret_cast<double*>(rec + p.L.off_norm));
}

template <int STRIDE>
constexpr size_t build_tma_smem(int RS) {
  return (size_t)BUILD_STAGES * BUILD_CHUNK * STRIDE + (size_t)RS * sizeof(int);
}

}  // namespace scgpu
