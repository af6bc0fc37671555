// scgpu_voxel.cuh -- voxel-grid downsample fused in front of the descriptor build.
//
// The reference feeds the Scan Context path with the raw scan filtered by pcl::VoxelGrid, leaf 0.5 m:
// mapOptmization.cpp:264 (setLeafSize), 1235-1237 (filter), 1628-1630 (makeAndSaveScancontextAndKeys of the result).
// k_build_voxel does both steps for a batch of scans: points -> voxel centroids -> polar max-height bins -> record.
//
// One thread-block CLUSTER of 8 CTAs per scan, two CTAs resident per SM (one scan's barriers overlap another's work).
// The voxel table of the scan is spread over the cluster's shared memories (8 x 3,072 slots x 32 B = 768 KB): voxel index -> hash -> (owner CTA, slot).  Every CTA walks ALL points of the scan
// (the scan is L2-resident: phase 1 has just read it) and accumulates the points whose voxel it owns with LOCAL
// shared-memory atomics; distributed shared memory carries the min/max exchange, the overflow flag and the final merge
// of the eight partial polar grids.  (First version: every CTA read 1/8 of the points and updated the owner's table
// with remote DSMEM atomics -- 84 ms per 1,184 scans; remote atomics run at ~0.1 per clock per SM.)
//   phase 1  min / max of the finite points (block reduce, then across the cluster through DSMEM)
//            -> min_b, div_b exactly as PCL computes them (FP32 multiply by 1/leaf, floor, int conversion)
//   phase 2  per point: leaf index (bit-exact with PCL), insert / accumulate: count and the three coordinate sums.
//            The sums are INTEGER (coordinate * 2^24 rounded to int64: exact for |x| >= 1/64 m, error < 3e-8 m below),
//            so they do not depend on the order of the atomics: results are deterministic, and the centroid is the
//            exact sum times 1/(n 2^24) in FP64, rounded to FP32 (PCL sums in FP32 in std::sort order: its result differs
//            from the true mean by up to ~n ulp; tests compare within that bound).
//            If a table overflows, the whole scan is redone in 2, 4, ... key partitions (max-height binning is
//            idempotent, so bins emitted by an abandoned attempt are harmless).
//   phase 3  every CTA walks its own slots: centroid -> exact polar bin (bin_point_exact) -> atomicMax into its own copy
//            of the grid; optionally the downsampled cloud is written out (unordered, with leaf indices).
//   phase 4  CTA 0 takes the maximum of the eight grids (remote loads), derives the keys and writes the record.
// PCL's refusal ("leaf size too small", index overflow) passes the input through unfiltered; so does this kernel
// (the raw points are binned).  Intensity is not carried: the only consumer of the filtered cloud is the
// descriptor, which reads x, y, z (SC.cpp:166-183).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "scgpu_kernels.cuh"

namespace scgpu {
namespace cg = cooperative_groups;

constexpr int VOX_CLUSTER = 8;
constexpr int VOX_THREADS = 512;
constexpr int VOX_SLOTS = 3072;           // per CTA
constexpr int VOX_MAX_PROBE = 96;
constexpr int VOX_UNROLL = 8;             // independent point loads in flight per thread
constexpr unsigned VOX_EMPTY = 0xffffffffu;

struct VoxInfo {  // per scan, for tests / callers
  int min_b[3];
  int div_b[3];
  int passes;     // key partitions the scan needed (1 unless the table overflowed)
  int status;     // 0 ok, 1 = PCL's "leaf size too small": input passed through unfiltered
  unsigned n_out; // voxels
  unsigned pad;
};

struct VoxelBuildParams {
  const unsigned char* pts;
  unsigned long long scan_pitch;
  unsigned n_pts, stride;
  float inv_leaf;            // 1.0f / leaf (FP32 division, as Eigen::Array4f::Ones() / leaf_size)
  BinConst bc;
  Layout L;
  unsigned char* records;    // [n_scans] packed records (may be null: downsample only)
  VoxInfo* info;             // [n_scans] (may be null)
  float4* out_pts;           // [n_scans][out_cap] centroid x, y, z, w = number of points (may be null)
  unsigned* out_idx;         // [n_scans][out_cap] leaf index
  unsigned out_cap;
  unsigned* keys;            // [scans of this launch][n_pts] scratch, 8 bytes per point: (point, leaf index) pairs grouped by owner CTA
  int* passes_hint;          // one int per handle: key partitions recent scans needed (performance only)
};

__host__ __device__ inline size_t vox_smem_bytes(int RS) {
  return ((size_t)VOX_SLOTS * (4 + 4 + 3 * 8) + (size_t)RS * 4 + 256 + 15) / 16 * 16 + (size_t)(VOX_THREADS / 32) * 64 * 8;
}

__device__ __forceinline__ unsigned vox_hash(unsigned k) { return k * 0x9E3779B1u; }
__device__ __forceinline__ unsigned vox_hash2(unsigned k) { return (k * 0x85EBCA6Bu) >> 15; }

// 64-bit accumulation with NATIVE 32-bit shared-memory atomics: there is no 64-bit shared-memory add (atomicAdd on
// unsigned long long compiles to an ATOMS.CAST.SPIN compare-and-swap loop, which crawls when the lanes of a warp hit the
// same voxel).  sum mod 2^64 = (hi, lo) with lo = sum of the low halves mod 2^32 and hi = sum of the high halves plus the
// number of times lo wrapped; the add that wraps lo sees it in the value the atomic returns.
__device__ __forceinline__ void vox_add64(unsigned* lo, unsigned* hi, unsigned slot, long long v) {
  const unsigned vl = (unsigned)(unsigned long long)v, vh = (unsigned)((unsigned long long)v >> 32);
  const unsigned old = atomicAdd(&lo[slot], vl);
  const unsigned carry = (old + vl < old) ? 1u : 0u;
  if (vh + carry) atomicAdd(&hi[slot], vh + carry);
}
__device__ __forceinline__ long long vox_get64(const unsigned* lo, const unsigned* hi, int slot) {
  return (long long)(((unsigned long long)hi[slot] << 32) | (unsigned long long)lo[slot]);
}

// point loads: KEEP = plain load (the scan is read again by the next phase: keep it in L2), else streaming
template <int STRIDE, bool KEEP>
__device__ __forceinline__ void vox_load(const unsigned char* p, float& x, float& y, float& z) {
  if (!KEEP) {
    load_point<STRIDE>(p, x, y, z);
  } else if (STRIDE == 16 || STRIDE == 32) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x = v.x, y = v.y, z = v.z;
  } else {
    const float* f = reinterpret_cast<const float*>(p);
    x = f[0], y = f[1], z = f[2];
  }
}

template <int STRIDE>
__global__ void __cluster_dims__(VOX_CLUSTER, 1, 1) __launch_bounds__(VOX_THREADS, 2) k_build_voxel(const VoxelBuildParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* t_lo = reinterpret_cast<unsigned*>(smem_raw);                                    // [3][VOX_SLOTS] low halves of the int64 sums
  unsigned* t_hi = t_lo + 3 * VOX_SLOTS;                                                     // [3][VOX_SLOTS] high halves
  unsigned* t_key = reinterpret_cast<unsigned*>(smem_raw + (size_t)VOX_SLOTS * 24);         // [VOX_SLOTS]
  unsigned* t_cnt = t_key + VOX_SLOTS;                                                       // [VOX_SLOTS]
  int* s_bins = reinterpret_cast<int*>(t_cnt + VOX_SLOTS);                                   // [RS] (CTA 0's copy is THE grid)
  float* s_mm = reinterpret_cast<float*>(s_bins + p.L.RS);                                   // [6] this CTA's min xyz, max xyz
  unsigned* s_flag = reinterpret_cast<unsigned*>(s_mm + 8);                                  // [0] overflow (CTA 0's is THE flag)
  unsigned* s_cnt = s_flag + 2;            // [8] my points per owner CTA
  unsigned* s_cur = s_cnt + VOX_CLUSTER;   // [8] write cursors
  unsigned* s_off = s_cur + VOX_CLUSTER;   // [8] list offsets of (me, owner) + [2] my own segment (start, length)
  const int lane = threadIdx.x & 31;
  uint2* myq = reinterpret_cast<uint2*>(smem_raw + vox_smem_bytes(p.L.RS) - (size_t)(VOX_THREADS / 32) * 64 * sizeof(uint2)) +
               (threadIdx.x >> 5) * 64;                                                      // this warp's queue of owned points (index, leaf index)
  const int RS = p.L.RS;
  const unsigned scan = blockIdx.y;
  const unsigned char* base = p.pts + (unsigned long long)scan * p.scan_pitch;
  // this CTA's share of the points
  const unsigned per = (p.n_pts + VOX_CLUSTER - 1) / VOX_CLUSTER;
  const unsigned start = min(rank * per, p.n_pts), end = min(start + per, p.n_pts);

  // ---- phase 1: min / max over the finite points (getMinMax3D) ---------------------------------------------------
  const float FMAX = 3.402823466e+38f;
  float mn[3] = {FMAX, FMAX, FMAX}, mx[3] = {-FMAX, -FMAX, -FMAX};
  for (unsigned i0 = start + threadIdx.x; i0 < end; i0 += VOX_THREADS * VOX_UNROLL) {
    float px[VOX_UNROLL], py[VOX_UNROLL], pz[VOX_UNROLL];
#pragma unroll
    for (int u = 0; u < VOX_UNROLL; ++u) {
      const unsigned i = i0 + u * VOX_THREADS;
      px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);
      if (i < end) vox_load<STRIDE, true>(base + (unsigned long long)i * p.stride, px[u], py[u], pz[u]);
    }
#pragma unroll
    for (int u = 0; u < VOX_UNROLL; ++u) {
      const float x = px[u], y = py[u], z = pz[u];
      if (isfinite(x) && isfinite(y) && isfinite(z)) {
        mn[0] = fminf(mn[0], x), mn[1] = fminf(mn[1], y), mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x), mx[1] = fmaxf(mx[1], y), mx[2] = fmaxf(mx[2], z);
      }
    }
  }
  if (threadIdx.x < 3) {  // order-preserving int encoding: float min / max as integer atomics
    reinterpret_cast<int*>(s_mm)[threadIdx.x] = enc_float(FMAX);
    reinterpret_cast<int*>(s_mm)[3 + threadIdx.x] = enc_float(-FMAX);
  }
  if (threadIdx.x == 0) s_flag[0] = 0, s_flag[1] = 0;
  for (int i = threadIdx.x; i < RS; i += VOX_THREADS) s_bins[i] = SCGPU_ENC_NOPOINT;
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = mn[a], hi = mx[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(FULL, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(reinterpret_cast<int*>(&s_mm[a]), enc_float(lo));
      atomicMax(reinterpret_cast<int*>(&s_mm[3 + a]), enc_float(hi));
    }
  }
  if (threadIdx.x == 0) reinterpret_cast<int*>(s_mm)[6] = *reinterpret_cast<volatile int*>(p.passes_hint);  // CTA 0's copy is the one used
  cluster.sync();
  float gmn[3], gmx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int lo = 0x7fffffff, hi = (int)0x80000000;
    for (unsigned r = 0; r < VOX_CLUSTER; ++r) {
      const int* rm = reinterpret_cast<const int*>(cluster.map_shared_rank(s_mm, r));
      lo = min(lo, rm[a]);
      hi = max(hi, rm[3 + a]);
    }
    gmn[a] = dec_float(lo);
    gmx[a] = dec_float(hi);
  }
  const bool any_point = gmn[0] <= gmx[0];
  int min_b[3], div_b[3];
  bool refuse = false;
  if (any_point) {
    const float fx = __fmul_rn(__fsub_rn(gmx[0], gmn[0]), p.inv_leaf), fy = __fmul_rn(__fsub_rn(gmx[1], gmn[1]), p.inv_leaf),
                fz = __fmul_rn(__fsub_rn(gmx[2], gmn[2]), p.inv_leaf);
    if (!(fx < 2147483648.f && fy < 2147483648.f && fz < 2147483648.f)) {
      refuse = true;  // (the int64 conversions below would be out of range)
    } else {
      const long long dx = (long long)fx + 1, dy = (long long)fy + 1, dz = (long long)fz + 1;  // each <= 2^31: no int64 overflow
      refuse = dx * dy * dz > 2147483647ll;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      min_b[a] = __float2int_rz(floorf(__fmul_rn(gmn[a], p.inv_leaf)));
      div_b[a] = __float2int_rz(floorf(__fmul_rn(gmx[a], p.inv_leaf))) - min_b[a] + 1;
    }
  } else {
    min_b[0] = min_b[1] = min_b[2] = 0;
    div_b[0] = div_b[1] = div_b[2] = 0;
  }
  const float fmb0 = (float)min_b[0], fmb1 = (float)min_b[1], fmb2 = (float)min_b[2];
  const int mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  unsigned* flag0 = cluster.map_shared_rank(s_flag, 0);
  const bool want_out = p.out_pts != nullptr;
  float4* out_pts = want_out ? p.out_pts + (unsigned long long)scan * p.out_cap : nullptr;
  unsigned* out_idx = want_out ? p.out_idx + (unsigned long long)scan * p.out_cap : nullptr;
  unsigned* out_count = p.info ? &p.info[scan].n_out : nullptr;  // global counter (zeroed by the host)

  // key partitions to start with: what recent scans needed (all CTAs of the cluster must agree: CTA 0's reading)
  unsigned attempt = 0, gen = 0;  // gen = (attempt << 16 | partition + 1): the value the overflow flag takes
  int passes = max(1, min(64, reinterpret_cast<const int*>(cluster.map_shared_rank(s_mm, 0))[6]));  // 1, 2, 4, ...: only ever doubled
  if (refuse) {
    // ---- PCL: "Leaf size is too small for the input dataset" -> output = input: bin the raw points ----------------
    for (unsigned i = start + threadIdx.x; i < end; i += VOX_THREADS) {
      float x, y, z, h;
      load_point<STRIDE>(base + (unsigned long long)i * p.stride, x, y, z);
      const int b = bin_point_exact(p.bc, x, y, z, h);
      if (b >= 0) atomicMax(&s_bins[b], enc_float(h));
      if (want_out) {
        const unsigned o = atomicAdd(out_count, 1u);
        if (o < p.out_cap) out_pts[o] = make_float4(x, y, z, 1.f), out_idx[o] = i;
      }
    }
  } else if (any_point) {
    // ---- phase 1b: leaf index of every point, PARTITIONED BY OWNER (a counting sort across the cluster).  Each CTA
    // computes the indices of its 1/8 of the points once, keeps them in shared memory (the table is not in use yet) and
    // counts them per owner; the 8 x 8 counts are exchanged through distributed shared memory; then every CTA writes
    // its (point, leaf index) pairs straight to their place in the scan's list, owner by owner.  Phase 2 then reads only
    // the owner's own segment -- in the previous version every CTA scanned all leaf indices of the scan in every pass
    // (74 % of the kernel's instructions).  Scans too large for the stash keep that flat layout.
    const bool partitioned = per <= (unsigned)(VOX_SLOTS * 6);          // stash = the 24 bytes/slot sum area, 4 bytes per point
    unsigned* stash = t_lo;
    uint2* plist = reinterpret_cast<uint2*>(p.keys) + (unsigned long long)scan * p.n_pts;
    if (threadIdx.x < VOX_CLUSTER) s_cnt[threadIdx.x] = 0, s_cur[threadIdx.x] = 0;
    __syncthreads();
    for (unsigned i0 = start + threadIdx.x; i0 < end; i0 += VOX_THREADS * VOX_UNROLL) {
      float px[VOX_UNROLL], py[VOX_UNROLL], pz[VOX_UNROLL];
#pragma unroll
      for (int u = 0; u < VOX_UNROLL; ++u) {
        const unsigned i = i0 + u * VOX_THREADS;
        px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);
        if (i < end) vox_load<STRIDE, true>(base + (unsigned long long)i * p.stride, px[u], py[u], pz[u]);
      }
#pragma unroll
      for (int u = 0; u < VOX_UNROLL; ++u) {
        const unsigned i = i0 + u * VOX_THREADS;
        const float x = px[u], y = py[u], z = pz[u];
        const int i0v = __float2int_rz(__fsub_rn(floorf(__fmul_rn(x, p.inv_leaf)), fmb0));
        const int i1v = __float2int_rz(__fsub_rn(floorf(__fmul_rn(y, p.inv_leaf)), fmb1));
        const int i2v = __float2int_rz(__fsub_rn(floorf(__fmul_rn(z, p.inv_leaf)), fmb2));
        const bool fin = isfinite(x) && isfinite(y) && isfinite(z);
        const unsigned key = fin ? (unsigned)(i0v + i1v * mul1 + i2v * mul2) : VOX_EMPTY;
        if (i < end) {
          if (partitioned) {
            stash[i - start] = key;
            if (fin) atomicAdd(&s_cnt[vox_hash(key) >> 29], 1u);
          } else {
            __stcg(&plist[i], make_uint2(i, key));
          }
        }
      }
    }
    unsigned seg_start = 0, seg_len = p.n_pts;  // my part of the list (flat layout: all of it, filtered by owner below)
    if (partitioned) {
      cluster.sync();  // every CTA's counts are final
      if (threadIdx.x < VOX_CLUSTER) {
        // place of (source CTA = me, owner = threadIdx.x): after all lower owners, then after the lower source CTAs
        unsigned off = 0;
        for (unsigned o = 0; o < VOX_CLUSTER; ++o) {
          for (unsigned sr = 0; sr < VOX_CLUSTER; ++sr) {
            const unsigned c = cluster.map_shared_rank(s_cnt, sr)[o];
            if (o < threadIdx.x || (o == threadIdx.x && sr < rank)) off += c;
          }
        }
        s_off[threadIdx.x] = off;
      }
      if (threadIdx.x == 32) {  // my own segment as owner
        unsigned st0 = 0, len = 0;
        for (unsigned o = 0; o <= rank; ++o)
          for (unsigned sr = 0; sr < VOX_CLUSTER; ++sr) {
            const unsigned c = cluster.map_shared_rank(s_cnt, sr)[o];
            if (o < rank) st0 += c;
            else len += c;
          }
        s_off[VOX_CLUSTER] = st0;
        s_off[VOX_CLUSTER + 1] = len;
      }
      __syncthreads();
      for (unsigned j = threadIdx.x; j < end - start; j += VOX_THREADS) {
        const unsigned key = stash[j];
        if (key == VOX_EMPTY) continue;
        const unsigned o = vox_hash(key) >> 29;
        __stcg(&plist[s_off[o] + atomicAdd(&s_cur[o], 1u)], make_uint2(start + j, key));
      }
      seg_start = s_off[VOX_CLUSTER];
      seg_len = s_off[VOX_CLUSTER + 1];
    }
    cluster.sync();  // (release / acquire at cluster scope: the list is read below with ld.global.cg; the stash is free)
    for (;; passes *= 2) {  // retried with twice the key partitions when a table overflows
      if (want_out && rank == 0 && threadIdx.x == 0) atomicExch(out_count, 0u);
      bool overflow = false;
      ++attempt;
      for (int part = 0; part < passes && !overflow; ++part) {
        gen = (attempt << 16) | (unsigned)(part + 1);  // what the overflow flag is set to in this (attempt, partition)
        // ---- clear my table
        for (int i = threadIdx.x; i < VOX_SLOTS; i += VOX_THREADS) {
          t_key[i] = VOX_EMPTY;
          t_cnt[i] = 0;
#pragma unroll
          for (int a = 0; a < 3; ++a) t_lo[a * VOX_SLOTS + i] = 0u, t_hi[a * VOX_SLOTS + i] = 0u;
        }
        __syncthreads();
        // ---- phase 2: EVERY CTA walks ALL points of the scan (L2-resident after phase 1) and keeps the ones whose
        // voxel it owns: accumulation is local shared-memory atomics only (remote DSMEM atomics -- one table entry
        // owned by one CTA, fed by all eight -- measured ~0.1 atomic per clock per SM: 60x slower than this)
        // VOX_UNROLL independent loads in flight per thread (one load per iteration left the SM at ~20 GB/s).
        // Only 1/8 of the points are this CTA's: the owned ones are COMPACTED into a per-warp queue and inserted 32 at a
        // time by full warps (inserting straight away ran the probe / atomic code with ~4 active lanes per warp:
        // thread_inst_executed_per_inst 8.1 in the first ncu capture).
        int qn = 0;  // entries waiting in my warp's queue (warp-uniform)
        auto insert = [&](const uint2 e) {  // e = (point index, leaf index)
          float x, y, z;
          vox_load<STRIDE, true>(base + (unsigned long long)e.x * p.stride, x, y, z);
          const unsigned key = e.y;
          const unsigned hsh = vox_hash(key);
          unsigned slot = ((hsh >> 5) & 0xffffffu) % VOX_SLOTS;
          bool placed = false;
          for (int pr = 0; pr < VOX_MAX_PROBE; ++pr) {
            const unsigned old = atomicCAS(&t_key[slot], VOX_EMPTY, key);
            if (old == VOX_EMPTY || old == key) {
              placed = true;
              break;
            }
            slot = slot + 1 == VOX_SLOTS ? 0 : slot + 1;
          }
          if (!placed) {
            *flag0 = gen;  // (remote store; rare)
            return;
          }
          atomicAdd(&t_cnt[slot], 1u);
          vox_add64(t_lo, t_hi, slot, __double2ll_rn((double)x * 16777216.0));
          vox_add64(t_lo + VOX_SLOTS, t_hi + VOX_SLOTS, slot, __double2ll_rn((double)y * 16777216.0));
          vox_add64(t_lo + 2 * VOX_SLOTS, t_hi + 2 * VOX_SLOTS, slot, __double2ll_rn((double)z * 16777216.0));
        };
        for (unsigned i0 = threadIdx.x & ~31u; i0 < seg_len; i0 += VOX_THREADS * VOX_UNROLL) {  // warp-uniform trip count
          uint2 kk[VOX_UNROLL];
#pragma unroll
          for (int u = 0; u < VOX_UNROLL; ++u) {
            const unsigned i = i0 + lane + u * VOX_THREADS;
            kk[u] = i < seg_len ? __ldcg(&plist[seg_start + i]) : make_uint2(0u, VOX_EMPTY);
          }
#pragma unroll
          for (int u = 0; u < VOX_UNROLL; ++u) {
            const unsigned key = kk[u].y;
            bool own = key != VOX_EMPTY && (partitioned || (vox_hash(key) >> 29) == rank);  // VOX_CLUSTER == 8: owner = top three hash bits
            own = own && (int)(vox_hash2(key) & (unsigned)(passes - 1)) == part;  // passes is a power of two
            const unsigned m = __ballot_sync(FULL, own);
            if (own) myq[qn + __popc(m & ((1u << lane) - 1u))] = kk[u];
            qn += __popc(m);
            if (qn >= 32) {
              __syncwarp();
              qn -= 32;
              const uint2 e = myq[qn + lane];
              __syncwarp();
              insert(e);
            }
          }
        }
        __syncwarp();
        if (lane < qn) insert(myq[lane]);
        cluster.sync();
        // ONE barrier per partition: the flag is never cleared.  All writes of this (attempt, partition) precede the
        // barrier above; what a CTA that is already further along may have written since tells the same story: a later
        // partition of the same attempt means this one did not overflow, a later ATTEMPT exists only because it did.
        // So every CTA of the cluster takes the same decision whatever it reads.
        {
          const unsigned v = *reinterpret_cast<volatile unsigned*>(flag0);
          overflow = v == gen || (v >> 16) > attempt;
        }
        if (overflow) break;
        // ---- phase 3: my slots -> centroids -> my copy of the polar grid
        for (int i = threadIdx.x; i < VOX_SLOTS; i += VOX_THREADS) {
          const unsigned key = t_key[i];
          if (key == VOX_EMPTY) continue;
          const double rn = 1.0 / ((double)t_cnt[i] * 16777216.0);  // one FP64 division per voxel, three multiplications
          const float cx = __double2float_rn(__dmul_rn((double)vox_get64(t_lo, t_hi, i), rn));
          const float cy = __double2float_rn(__dmul_rn((double)vox_get64(t_lo + VOX_SLOTS, t_hi + VOX_SLOTS, i), rn));
          const float cz = __double2float_rn(__dmul_rn((double)vox_get64(t_lo + 2 * VOX_SLOTS, t_hi + 2 * VOX_SLOTS, i), rn));
          float h;
          const int b = bin_point<true, false>(p.bc, cx, cy, cz, h);  // FP32 front end + exact fallback: the reference's bin
          if (b >= 0) atomicMax(&s_bins[b], enc_float(h));
          if (want_out) {
            const unsigned o = atomicAdd(out_count, 1u);
            if (o < p.out_cap) out_pts[o] = make_float4(cx, cy, cz, (float)t_cnt[i]), out_idx[o] = key;
          }
        }
        __syncthreads();
      }
      if (!overflow) break;
    }
  }
  // ---- the eight partial grids -> CTA 0 (plain remote loads, coalesced)
  cluster.sync();
  if (rank == 0) {
    for (unsigned r = 1; r < VOX_CLUSTER; ++r) {
      const int* rb = cluster.map_shared_rank(s_bins, r);
      for (int i = threadIdx.x; i < RS; i += VOX_THREADS) s_bins[i] = max(s_bins[i], rb[i]);
    }
  }
  cluster.sync();  // nobody leaves while CTA 0 still reads its shared memory
  if (rank != 0) return;
  if (threadIdx.x == 0 && !refuse && any_point) atomicMax(p.passes_hint, passes);
  if (threadIdx.x == 0 && p.info) {
    VoxInfo* vi = &p.info[scan];
    for (int a = 0; a < 3; ++a) vi->min_b[a] = min_b[a], vi->div_b[a] = div_b[a];
    vi->passes = passes;
    vi->status = refuse ? 1 : 0;
  }
  if (!p.records) return;
  // ---- phase 4: record (as k_build)
  unsigned char* rec = p.records + (unsigned long long)scan * p.L.rec_bytes;
  float* s_sc = reinterpret_cast<float*>(s_bins);
  float* rec_sc = reinterpret_cast<float*>(rec);
  for (int i = threadIdx.x; i < RS; i += VOX_THREADS) {
    float f = dec_float(s_bins[i]);
    if (f == -1000.0f) f = 0.0f;
    s_sc[i] = f;
    rec_sc[i] = f;
  }
  __syncthreads();
  keys_from_sc<float>(s_sc, p.L.R, p.L.S, p.L.R, nullptr, reinterpret_cast<float*>(rec + p.L.off_ring),
                      reinterpret_cast<double*>(rec + p.L.off_sector), reinterpret_cast<double*>(rec + p.L.off_norm));
}

}  // namespace scgpu
