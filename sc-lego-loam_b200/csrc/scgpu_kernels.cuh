// scgpu_kernels.cuh -- hand-written sm_100a kernels of the Scan Context loop-closure hot path.
//
// One kernel per reference stage (citations: "SC.cpp" = SC-LeGO-LOAM/LeGO-LOAM/src/Scancontext.cpp,
// "SC.h" = include/Scancontext.h, "nf.hpp" = include/nanoflann.hpp):
//   k_build      stage 1+2  SC.cpp:151-195 (polar max-height binning) fused with SC.cpp:198-227 (keys)
//   k_append     records -> HBM-resident database (SC.cpp:237-240)
//   k_topk       stage 3    exact brute-force ring-key top-K replacing SC.cpp:283-289 / nf.hpp kd-tree
//   k_score      stage 4    SC.cpp:116-148 column-shifted cosine distance (FP64, reference order)
//   k_best/k_finalize       SC.cpp:296-336 candidate loop, threshold, yaw
//
// Bit-exactness rules used throughout: every arithmetic step of the reference is one IEEE operation here
// (__fmul_rn/__fadd_rn/__dmul_rn/... never contract into FMA; the file is also compiled with --fmad=false);
// FP64 reductions follow Eigen 3.3's SSE2 redux order (4 interleaved partial sums); the ring-key distance
// follows nanoflann's grouped-by-4 FP32 accumulation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scgpu {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned long long KEY_NONE = ~0ull;

// ------------------------------------------------------------------------------------------------
// Packed descriptor record (one per scan): float sc[R*S] (column-major) | float ring[R] | double sector[S]
// | double colnorm[S]; padded to 16 bytes.
// ------------------------------------------------------------------------------------------------
struct Layout {
  int R, S, RS;
  unsigned off_ring, off_sector, off_norm, rec_bytes;
};

__host__ __device__ inline Layout make_layout(int R, int S) {
  Layout L;
  L.R = R;
  L.S = S;
  L.RS = R * S;
  L.off_ring = 4u * (unsigned)L.RS;
  L.off_sector = (L.off_ring + 4u * (unsigned)R + 7u) & ~7u;
  L.off_norm = L.off_sector + 8u * (unsigned)S;
  L.rec_bytes = (L.off_norm + 8u * (unsigned)S + 15u) & ~15u;
  return L;
}

// ------------------------------------------------------------------------------------------------
// Exact scalar pieces
// ------------------------------------------------------------------------------------------------

// glibc 2.39 atanf == fdlibm s_atanf.c evaluated in plain binary32 (verified over every float against the
// host libm by tests/test_atanf.py on the oracle restatement, and against this device function on the GPU).
// The reference reaches it through SC.cpp:26-35 (atan -> atanf: <math.h> is in scope in the ROS build).
__device__ __forceinline__ float atanf_fdlibm(float x) {
  const unsigned hx = __float_as_uint(x), ix = hx & 0x7fffffffu;
  int id;
  if (ix >= 0x4c000000u) {  // |x| >= 2^25, inf, NaN
    if (ix > 0x7f800000u) return __fadd_rn(x, x);
    const float r = __fadd_rn(__uint_as_float(0x3fc90fdau), __uint_as_float(0x33a22168u));
    return (hx >> 31) ? -r : r;
  }
  float hi = 0.f, lo = 0.f;
  if (ix < 0x3ee00000u) {  // |x| < 7/16
    if (ix < 0x31000000u) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000u) {
      if (ix < 0x3f300000u) {
        id = 0;
        hi = __uint_as_float(0x3eed6338u);
        lo = __uint_as_float(0x31ac3769u);
        x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x));
      } else {
        id = 1;
        hi = __uint_as_float(0x3f490fdau);
        lo = __uint_as_float(0x33222168u);
        x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000u) {
        id = 2;
        hi = __uint_as_float(0x3f7b985eu);
        lo = __uint_as_float(0x33140fb4u);
        x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x)));
      } else {
        id = 3;
        hi = __uint_as_float(0x3fc90fdau);
        lo = __uint_as_float(0x33a22168u);
        x = __fdiv_rn(-1.0f, x);
      }
    }
  }
  const float aT0 = __uint_as_float(0x3eaaaaabu), aT1 = __uint_as_float(0xbe4ccccdu), aT2 = __uint_as_float(0x3e124925u),
              aT3 = __uint_as_float(0xbde38e38u), aT4 = __uint_as_float(0x3dba2e6eu), aT5 = __uint_as_float(0xbd9d8795u),
              aT6 = __uint_as_float(0x3d886b35u), aT7 = __uint_as_float(0xbd6ef16bu), aT8 = __uint_as_float(0x3d4bda59u),
              aT9 = __uint_as_float(0xbd15a221u), aT10 = __uint_as_float(0x3c8569d7u);
  const float z = __fmul_rn(x, x);
  const float w = __fmul_rn(z, z);
  float s1 = __fadd_rn(aT8, __fmul_rn(w, aT10));
  s1 = __fadd_rn(aT6, __fmul_rn(w, s1));
  s1 = __fadd_rn(aT4, __fmul_rn(w, s1));
  s1 = __fadd_rn(aT2, __fmul_rn(w, s1));
  s1 = __fadd_rn(aT0, __fmul_rn(w, s1));
  s1 = __fmul_rn(z, s1);
  float s2 = __fadd_rn(aT7, __fmul_rn(w, aT9));
  s2 = __fadd_rn(aT5, __fmul_rn(w, s2));
  s2 = __fadd_rn(aT3, __fmul_rn(w, s2));
  s2 = __fadd_rn(aT1, __fmul_rn(w, s2));
  s2 = __fmul_rn(w, s2);
  const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
  if (id < 0) return __fsub_rn(x, xs);
  const float r = __fsub_rn(hi, __fsub_rn(__fsub_rn(xs, lo), x));
  return (hx >> 31) ? -r : r;
}

// SC.cpp:23-36.  (180/M_PI) is a double constant; the float atanf result is widened, scaled and offset in
// double and narrowed by the float return.  NaN coordinates are rejected by the caller (reference: UB).
__device__ __forceinline__ float xy2theta_exact(float x, float y) {
  const double k = 180.0 / 3.14159265358979323846;
  if ((x >= 0.f) & (y >= 0.f)) return __double2float_rn(__dmul_rn(k, (double)atanf_fdlibm(__fdiv_rn(y, x))));
  if ((x < 0.f) & (y >= 0.f)) return __double2float_rn(__dsub_rn(180.0, __dmul_rn(k, (double)atanf_fdlibm(__fdiv_rn(y, -x)))));
  if ((x < 0.f) & (y < 0.f)) return __double2float_rn(__dadd_rn(180.0, __dmul_rn(k, (double)atanf_fdlibm(__fdiv_rn(y, x)))));
  return __double2float_rn(__dsub_rn(360.0, __dmul_rn(k, (double)atanf_fdlibm(__fdiv_rn(-y, x)))));
}

struct BinConst {
  int R, S;
  double lidar_height, max_radius;
  // fast-path gating (never the source of a result, only of the decision WHICH path computes it)
  float lh_f;         // (float)lidar_height
  int lh_is_float;    // lidar_height is exactly a float: float(double(z)+H) == fadd(z, H)  (53 >= 2*24+2)
  int fast;           // 0: every point takes the exact path
  float ring_scale;   // R / max_radius
  float sec_scale;    // S / (2 pi)
  float eps_r, eps_s; // distance to the nearest ring / sector boundary below which the exact path decides
};

__host__ inline BinConst make_bin_const(int R, int S, double lidar_height, double max_radius, int fast) {
  BinConst c;
  c.R = R;
  c.S = S;
  c.lidar_height = lidar_height;
  c.max_radius = max_radius;
  c.lh_f = (float)lidar_height;
  c.lh_is_float = ((double)c.lh_f == lidar_height);
  c.fast = fast;
  c.ring_scale = (float)((double)R / max_radius);
  c.sec_scale = (float)((double)S / 6.283185307179586476925);
  // error budget (DESIGN.md "binning fast path"): fast ring coordinate |err| <= 3.2e-5 at 64 rings, reference's own
  // deviation from the true value <= 4e-6; fast sector coordinate |err| <= 2.5e-7 S, reference's <= 1e-7 S.
  c.eps_r = 1.0e-4f * (R > 64 ? (float)R / 64.f : 1.f);
  c.eps_s = 3.4e-6f * (float)S;
  if (c.eps_s < 2.0e-4f) c.eps_s = 2.0e-4f;
  return c;
}

// SC.cpp:166-183 for one point.  Returns the 0-based bin (sector*R + ring) or -1 when the point does not
// contribute (outside the ROI, NaN coordinate, or a height that can never win the max).
__device__ __forceinline__ int bin_point_exact(const BinConst& c, float x, float y, float z, float& height) {
  height = __double2float_rn(__dadd_rn((double)z, c.lidar_height));                 // SC.cpp:168
  const float range = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));     // SC.cpp:171
  if (!(x == x) || !(y == y)) return -1;                                            // reference: undefined
  if ((double)range > c.max_radius) return -1;                                      // SC.cpp:175
  if (!(height == height)) return -1;                                               // NaN never passes SC.cpp:182
  const float theta = xy2theta_exact(x, y);                                         // SC.cpp:172
  const double qr = ceil(__dmul_rn(__ddiv_rn((double)range, c.max_radius), (double)c.R));  // SC.cpp:178
  const double qs = ceil(__dmul_rn(__ddiv_rn((double)theta, 360.0), (double)c.S));         // SC.cpp:179
  // int(): NaN -> INT_MIN on x86, 0 here; both clamp to 1.  Values are otherwise far inside int range.
  const int ring = max(min(c.R, __double2int_rz(qr)), 1);
  const int sector = max(min(c.S, __double2int_rz(qs)), 1);
  return (sector - 1) * c.R + (ring - 1);
}

// The same function with a cheap FP32 front end.  The approximate polar coordinates (one MUFU.RSQ, one MUFU.RCP,
// a degree-7 odd atan polynomial, |error| <= 1.3e-7 rad) decide the bin ONLY when the point is farther from every
// ring and sector boundary than the sum of the approximation error and the reference's own rounding error (see
// make_bin_const); everything else -- boundary neighbourhoods, points on an axis (where the reference's quadrant
// tests on -0.0 matter), the ROI edge, zero / subnormal / non-finite inputs (which the .ftz approximations turn into
// inf or NaN and thereby into "not safe") -- is decided by bin_point_exact.  Either way the result is the
// reference's.  Branch-free up to the (rare) hand-over.  tests: scgpu_probe_selfcheck compares the two paths over
// billions of device-generated points.
__device__ __forceinline__ float mufu_rsq(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// out-of-line copy of the exact path for the kernels that defer it (keeps their hot loop small).  Everything by
// value -- arguments and the (bin << 32 | height bits) result travel in registers: a version taking `const BinConst&`
// and `float&` forced the caller's BinConst copy and its height array into local memory (160-byte stack frame, every
// height store a local store: 1.2 GB of extra DRAM writes per 4,541-scan launch).
__device__ __noinline__ unsigned long long bin_point_exact_noinline(int R, int S, double lidar_height, double max_radius, float x, float y,
                                                                    float z) {
  BinConst c;
  c.R = R;
  c.S = S;
  c.lidar_height = lidar_height;
  c.max_radius = max_radius;
  float h;
  const int b = bin_point_exact(c, x, y, z, h);
  return ((unsigned long long)(unsigned)b << 32) | (unsigned long long)__float_as_uint(h);
}

constexpr int BIN_UNDECIDED = -2;  // bin_point_fast: the exact path has to decide this point

// The front end alone: returns the bin, -1 (outside the ROI by more than the error), or BIN_UNDECIDED.  Branch-free.
template <bool LH_FLOAT>
__device__ __forceinline__ int bin_point_fast(const BinConst& c, float x, float y, float z, float& height);

template <bool FAST, bool LH_FLOAT>
__device__ __forceinline__ int bin_point(const BinConst& c, float x, float y, float z, float& height, bool* fell_back = nullptr) {
  if (!FAST) return bin_point_exact(c, x, y, z, height);
  const int b = bin_point_fast<LH_FLOAT>(c, x, y, z, height);
  if (b != BIN_UNDECIDED) return b;
  if (fell_back) *fell_back = true;
  return bin_point_exact(c, x, y, z, height);
}

template <bool LH_FLOAT>
__device__ __forceinline__ int bin_point_fast(const BinConst& c, float x, float y, float z, float& height) {
  const float fR = (float)c.R, fS = (float)c.S;
  const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const float qr = __fmul_rn(__fmul_rn(r2, mufu_rsq(r2)), c.ring_scale);  // ~ range * R / max_radius
  const float h = LH_FLOAT ? __fadd_rn(z, c.lh_f) : __double2float_rn(__dadd_rn((double)z, c.lidar_height));
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = __fmul_rn(mn, mufu_rcp(mx));
  const float s = __fmul_rn(a, a);
  float p = __uint_as_float(0xbb84dc15u);
  p = __fmaf_rn(p, s, __uint_as_float(0x3cb319dcu));
  p = __fmaf_rn(p, s, __uint_as_float(0xbd650442u));
  p = __fmaf_rn(p, s, __uint_as_float(0x3dc578dcu));
  p = __fmaf_rn(p, s, __uint_as_float(0xbe0e6ca2u));
  p = __fmaf_rn(p, s, __uint_as_float(0x3e4c40b9u));
  p = __fmaf_rn(p, s, __uint_as_float(0xbeaaa61du));
  p = __fmaf_rn(p, s, __uint_as_float(0x3f7ffff5u));
  float t = __fmul_rn(__fmul_rn(a, p), c.sec_scale);  // octant angle in sector units, [0, S/8]
  t = (ay > ax) ? __fsub_rn(0.25f * fS, t) : t;
  t = (x < 0.f) ? __fsub_rn(0.5f * fS, t) : t;
  t = (y < 0.f) ? __fsub_rn(fS, t) : t;
  // nearest integers and the signed distances to them on the FP32 add pipe: v + 2^23 rounds v (< 2^22) to an
  // integer held in the low mantissa bits
  const float MAGIC = 8388608.0f;
  const float qr_m = __fadd_rn(qr, MAGIC), t_m = __fadd_rn(t, MAGIC);
  const float dr = __fsub_rn(qr, __fsub_rn(qr_m, MAGIC)), ds = __fsub_rn(t, __fsub_rn(t_m, MAGIC));
  const int kr = (__float_as_int(qr_m) & 0x7fffff) - (dr < 0.f);  // floor(qr)
  const int ks = (__float_as_int(t_m) & 0x7fffff) - (ds < 0.f);   // floor(t)
  // beyond the ROI by more than the error (finite qr only: inf / NaN come from 0, subnormal or overflowing r2)
  const bool outside = (qr > fR + c.eps_r) & (qr <= 3.0e38f);
  // |dr| > eps also excludes qr in [R, R+eps] and NaN; |ds| > eps excludes t near 0 and S and NaN / inf
  const bool safe = (fabsf(dr) > c.eps_r) & (fabsf(ds) > c.eps_s) & (mn > 0.f) & (h == h);
  height = h;
  return outside ? -1 : (safe ? ks * c.R + kr : BIN_UNDECIDED);
}

// order-preserving float <-> int map so that atomicMax on ints is max on floats (SC.cpp:182-183)
__device__ __forceinline__ int enc_float(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float dec_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
#define SCGPU_ENC_NOPOINT (-1148846081) /* enc_float(-1000.0f): bits 0xc47a0000 ^ 0x7fffffff = 0xbb85ffff */

// Eigen 3.3 Core/Redux.h (LinearVectorizedTraversal, NoUnrolling, Packet2d, alignedStart 0): the order in which
// the reference's sum()/mean()/norm()/dot() add their terms.  coeff(i) returns the i-th (transformed) scalar.
template <class Coeff>
__device__ __forceinline__ double redux_eigen(int size, Coeff coeff) {
  const int alignedSize2 = (size / 4) * 4;
  const int alignedSize = (size / 2) * 2;
  if (size == 0) return 0.0;
  double res;
  if (alignedSize) {
    double p00 = coeff(0), p01 = coeff(1);
    if (alignedSize > 2) {
      double p10 = coeff(2), p11 = coeff(3);
      for (int i = 4; i < alignedSize2; i += 4) {
        p00 = __dadd_rn(p00, coeff(i));
        p01 = __dadd_rn(p01, coeff(i + 1));
        p10 = __dadd_rn(p10, coeff(i + 2));
        p11 = __dadd_rn(p11, coeff(i + 3));
      }
      p00 = __dadd_rn(p00, p10);
      p01 = __dadd_rn(p01, p11);
      if (alignedSize > alignedSize2) {
        p00 = __dadd_rn(p00, coeff(alignedSize2));
        p01 = __dadd_rn(p01, coeff(alignedSize2 + 1));
      }
    }
    res = __dadd_rn(p00, p01);
    for (int i = alignedSize; i < size; ++i) res = __dadd_rn(res, coeff(i));
  } else {
    res = coeff(0);
    for (int i = 1; i < size; ++i) res = __dadd_rn(res, coeff(i));
  }
  return res;
}

template <class T>
struct CoeffStrided {
  const T* p;
  int stride;
  __device__ __forceinline__ double operator()(int i) const { return (double)p[i * stride]; }
};
template <class T>
struct CoeffSquare {
  const T* p;
  __device__ __forceinline__ double operator()(int i) const {
    const double v = (double)p[i];
    return __dmul_rn(v, v);
  }
};
template <class T>
struct CoeffProduct {
  const T* a;
  const T* b;
  __device__ __forceinline__ double operator()(int i) const { return __dmul_rn((double)a[i], (double)b[i]); }
};
// (vkey1[j] - vkey2[(j - s) mod S])^2  -- SC.cpp:99-103 with circshift SC.cpp:39-59
struct CoeffShiftDiffSq {
  const double* v1;
  const double* v2;
  int s, S;
  __device__ __forceinline__ double operator()(int j) const {
    int jj = j - s;
    if (jj < 0) jj += S;
    const double d = __dsub_rn(v1[j], v2[jj]);
    return __dmul_rn(d, d);
  }
};

// Keys of one descriptor held in shared memory (T = float from binning, or double from the public API):
// ring key (SC.cpp:198-211), sector key (SC.cpp:214-227), column norms (SC.cpp:78).  Whole block cooperates.
template <class T>
__device__ __forceinline__ void keys_from_sc(const T* s_sc, int R, int S, int ld, double* ring_d, float* ring_f, double* sector,
                                             double* colnorm) {  // ld: elements between consecutive columns (>= R)
  for (int t = threadIdx.x; t < R + S; t += blockDim.x) {
    if (t < R) {
      const double m = __ddiv_rn(redux_eigen(S, CoeffStrided<T>{s_sc + t, ld}), (double)S);
      if (ring_d) ring_d[t] = m;
      if (ring_f) ring_f[t] = __double2float_rn(m);  // eig2stdvec, SC.cpp:62-66
    } else {
      const int c = t - R;
      const T* col = s_sc + c * ld;
      if (sector) sector[c] = __ddiv_rn(redux_eigen(R, CoeffStrided<T>{col, 1}), (double)R);
      if (colnorm) colnorm[c] = __dsqrt_rn(redux_eigen(R, CoeffSquare<T>{col}));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 1+2: k_build
//   grid = (tiles per scan, scans), block = 256.  Each block bins a tile of consecutive points into a
//   shared-memory R*S grid of order-encoded ints (warp-aggregated atomicMax), merges its non-empty bins into
//   the scan's global grid, and the last block of a scan (ticket counter) decodes the grid, derives the keys
//   and writes the packed record.  The global grid and the ticket are left reset for the next launch.
// ------------------------------------------------------------------------------------------------
struct BuildParams {
  const unsigned char* pts;  // scans contiguous: scan s at pts + s * scan_pitch
  unsigned long long scan_pitch;
  unsigned n_pts;     // points per scan
  unsigned stride;    // bytes between points
  unsigned val_off;   // byte offset of the value that is binned: 8 = z (the reference), 16 = intensity of a pcl::PointXYZI
                      // (the descriptor variant of Scancontext.h:41; lidar_height is then 0)
  unsigned pts_per_block;
  BinConst bc;
  Layout L;
  int* gbins;         // [scans][RS], pre-set to SCGPU_ENC_NOPOINT
  unsigned* tickets;  // [scans], pre-set to 0
  unsigned char* records;
};

template <int STRIDE>  // 16 / 32: one aligned 16-byte load per point; 0: three 4-byte loads (any 4-aligned stride)
__device__ __forceinline__ void load_point(const unsigned char* p, float& x, float& y, float& z, unsigned val_off = 8) {
  if (STRIDE == 16 || STRIDE == 32) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    x = v.x;
    y = v.y;
    z = v.z;
  } else {
    const float* f = reinterpret_cast<const float*>(p);
    x = __ldcs(f);
    y = __ldcs(f + 1);
    z = __ldcs(f + 2);
  }
  if (val_off != 8) z = __ldcs(reinterpret_cast<const float*>(p + val_off));  // (uniform branch) intensity instead of height
}

// One atomicMax per distinct bin per warp: consecutive points of a scan share a beam and neighbouring azimuths,
// so a warp touches a handful of bins.  match.any groups the lanes by bin, redux.sync takes each group's maximum,
// the lowest lane of each group issues the shared-memory atomic.
__device__ __forceinline__ void warp_bin_max(int* s_bins, int bin, int enc) {
  const unsigned peers = __match_any_sync(FULL, bin);
  const int m = __reduce_max_sync(peers, enc);
  if (bin >= 0 && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicMax(&s_bins[bin], m);
}

template <int STRIDE, bool FAST, bool LH_FLOAT>
__global__ void __launch_bounds__(256) k_build(const BuildParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* s_bins = reinterpret_cast<int*>(smem_raw);
  __shared__ bool s_last;
  const int RS = p.L.RS;
  const unsigned scan = blockIdx.y;
  for (int i = threadIdx.x; i < RS; i += blockDim.x) s_bins[i] = SCGPU_ENC_NOPOINT;
  __syncthreads();

  const unsigned start = blockIdx.x * p.pts_per_block;
  const unsigned end = min(start + p.pts_per_block, p.n_pts);
  const unsigned char* base = p.pts + (unsigned long long)scan * p.scan_pitch;
  // whole warps iterate together (the aggregation uses full-mask warp primitives); UNROLL independent loads are
  // issued before the first point is processed so that enough bytes are in flight per SM to cover HBM latency
  constexpr int UNROLL = 4;
  const unsigned lane_base = start + threadIdx.x;
  for (unsigned i0 = start + (threadIdx.x & ~31u); i0 < end; i0 += UNROLL * blockDim.x) {
    float px[UNROLL], py[UNROLL], pz[UNROLL];
    const unsigned i = lane_base + (i0 - (start + (threadIdx.x & ~31u)));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned iu = i + u * blockDim.x;
      px[u] = py[u] = pz[u] = __int_as_float(0x7fc00000);  // NaN: dropped
      if (iu < end) load_point<STRIDE>(base + (unsigned long long)iu * p.stride, px[u], py[u], pz[u], p.val_off);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (i0 + u * blockDim.x >= end) break;  // warp-uniform
      float h;
      const int bin = bin_point<FAST, LH_FLOAT>(p.bc, px[u], py[u], pz[u], h);
      warp_bin_max(s_bins, bin, bin >= 0 ? enc_float(h) : INT_MIN);
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

  // decode (SC.cpp:187-190: cells still at NO_POINT become 0) and write the record
  unsigned char* rec = p.records + (unsigned long long)scan * p.L.rec_bytes;
  float* s_sc = reinterpret_cast<float*>(smem_raw);
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

__global__ void k_fill_int(int* p, size_t n, int v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Records from ready-made float descriptors (database load / pre-fill): keys as SC.cpp:233-235.
__global__ void __launch_bounds__(128) k_records_from_sc(const float* sc, Layout L, unsigned char* records) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_sc = reinterpret_cast<float*>(smem_raw);
  const float* src = sc + (size_t)blockIdx.x * L.RS;
  unsigned char* rec = records + (size_t)blockIdx.x * L.rec_bytes;
  for (int i = threadIdx.x; i < L.RS; i += blockDim.x) {
    const float f = src[i];
    s_sc[i] = f;
    reinterpret_cast<float*>(rec)[i] = f;
  }
  __syncthreads();
  keys_from_sc<float>(s_sc, L.R, L.S, L.R, nullptr, reinterpret_cast<float*>(rec + L.off_ring),
                      reinterpret_cast<double*>(rec + L.off_sector), reinterpret_cast<double*>(rec + L.off_norm));
}

// ------------------------------------------------------------------------------------------------
// Database (one shard): sc[cap][RS] float | ringT (ring keys, tiled dimension-major: ring_at) | sector[cap][S] double |
// colnorm[cap][S] double.  Global entry g lives on shard g % G at g / G.
// ------------------------------------------------------------------------------------------------
struct Db {
  float* sc;
  float* ringT;
  double* sector;
  double* colnorm;
  unsigned long long cap;  // local capacity
  int rank, G;
  // ring-key matrix addressing.  Sharded-by-all_gather / single shard: ringT is [R][cap] over LOCAL slots (ring_global = 0,
  // ring_cap = cap).  Peer-sharded (PeerTab below): ringT is this shard's REPLICA of every shard's ring keys, [R][ring_cap]
  // over GLOBAL indices (ring_global = 1), so that retrieval for a query runs on one device without a merge.
  unsigned long long ring_cap;
  int ring_global;
};

__host__ __device__ __forceinline__ unsigned long long ring_slot(const Db& db, unsigned long long g, unsigned long long l) {
  return db.ring_global ? g : l;
}
// Ring-key matrix layout: tiles of 32 consecutive slots, dimension-major inside a tile -- element (slot, d) at
// (slot / 32) * (R * 32) + d * 32 + slot % 32.  A warp that takes 32 consecutive slots reads one aligned 128-byte line per
// dimension at a CONSTANT offset (d * 128 bytes) from its tile base: no per-dimension address arithmetic in k_topk (the plain
// dimension-major [R][cap] layout of round 1 cost ~35 integer instructions per key), and the layout does not depend on the capacity.
__host__ __device__ __forceinline__ size_t ring_at(unsigned long long slot, int d, int R) {
  return (size_t)(slot >> 5) * (size_t)(R * 32) + (size_t)d * 32 + (size_t)(slot & 31);
}

// Peer-sharded database: every shard's arrays as seen from THIS device -- its own allocations and the other shards'
// through NVLink peer mappings (cudaIpcOpenMemHandle between processes, cudaDeviceEnablePeerAccess inside one).  Global entry g
// lives on shard g % G, local slot g / G.  G == 0: table unused.
constexpr int MAX_SHARDS = 8;
struct PeerTab {
  int G, rank;
  unsigned long long cap;  // local capacity, identical on every shard
  const float* sc[MAX_SHARDS];
  const double* sector[MAX_SHARDS];
  const double* colnorm[MAX_SHARDS];
  const float* sc_hat[MAX_SHARDS];          // screening copies (scgpu_exhaustive.cuh)
  const unsigned char* vk[MAX_SHARDS];
};

// where a kernel that produces per-entry / per-query values also delivers them on other devices (peer stores)
struct PushList {
  int n;
  void* dst[MAX_SHARDS];
};

// records [0,n) -> entries first_global + i*step (kept when owned)
// push: the other shards' ring-key replicas (peer stores over NVLink); the ring key of a record is written there too
// returns true (block-uniform) when the entry is stored on this shard; *l_out = its local slot
__device__ __forceinline__ bool append_entry(const unsigned char* rec, const Layout& L, const Db& db, unsigned long long g, const PushList& push,
                                             unsigned long long* l_out) {
  const unsigned long long l = g / (unsigned long long)db.G;
  *l_out = l;
  const bool mine = (int)(g % (unsigned long long)db.G) == db.rank;
  if (db.ring_global) {  // every shard keeps every ring key
    const float* ring = reinterpret_cast<const float*>(rec + L.off_ring);
    for (int i = threadIdx.x; i < L.R; i += blockDim.x) {
      const float v = ring[i];
      const size_t o = ring_at(g, i, L.R);
      db.ringT[o] = v;
      for (int s = 0; s < push.n; ++s) static_cast<float*>(push.dst[s])[o] = v;
    }
  }
  if (!mine) return false;
  const float4* src4 = reinterpret_cast<const float4*>(rec);
  float4* dst4 = reinterpret_cast<float4*>(db.sc + l * L.RS);
  if ((L.RS & 3) == 0) {
    for (int i = threadIdx.x; i < L.RS / 4; i += blockDim.x) dst4[i] = src4[i];
  } else {
    for (int i = threadIdx.x; i < L.RS; i += blockDim.x) db.sc[l * L.RS + i] = reinterpret_cast<const float*>(rec)[i];
  }
  const float* ring = reinterpret_cast<const float*>(rec + L.off_ring);
  const double* sector = reinterpret_cast<const double*>(rec + L.off_sector);
  const double* norm = reinterpret_cast<const double*>(rec + L.off_norm);
  if (!db.ring_global)
    for (int i = threadIdx.x; i < L.R; i += blockDim.x) db.ringT[ring_at(l, i, L.R)] = ring[i];
  for (int i = threadIdx.x; i < L.S; i += blockDim.x) {
    db.sector[l * L.S + i] = sector[i];
    db.colnorm[l * L.S + i] = norm[i];
  }
  return true;
}

__global__ void __launch_bounds__(128) k_append(const unsigned char* records, Layout L, Db db, unsigned long long first_global,
                                                unsigned long long step, PushList push) {
  unsigned long long l;
  append_entry(records + (size_t)blockIdx.x * L.rec_bytes, L, db, first_global + blockIdx.x * step, push, &l);
}

// stored entries -> records (query records for stored entries).  Block b takes global entry idx[b] (idx != null) or
// first_global + b.  Single shard: the entry is local.  Peer-sharded (peers.G > 0): the entry is read from its owner
// shard through the peer table, whichever shard that is.
__global__ void __launch_bounds__(128) k_gather(unsigned char* records, Layout L, Db db, PeerTab peers, unsigned long long first_global,
                                                const unsigned long long* idx) {
  const unsigned long long g = idx ? idx[blockIdx.x] : first_global + blockIdx.x;
  const unsigned long long G = peers.G ? (unsigned long long)peers.G : (unsigned long long)db.G;
  const unsigned long long l = g / G;
  const int o = (int)(g % G);
  const float* sc = peers.G ? peers.sc[o] : db.sc;
  const double* sector = peers.G ? peers.sector[o] : db.sector;
  const double* colnorm = peers.G ? peers.colnorm[o] : db.colnorm;
  unsigned char* rec = records + (size_t)blockIdx.x * L.rec_bytes;
  for (int i = threadIdx.x; i < L.RS; i += blockDim.x) reinterpret_cast<float*>(rec)[i] = sc[l * L.RS + i];
  for (int i = threadIdx.x; i < L.R; i += blockDim.x)
    reinterpret_cast<float*>(rec + L.off_ring)[i] = db.ringT[ring_at(ring_slot(db, g, l), i, L.R)];
  for (int i = threadIdx.x; i < L.S; i += blockDim.x) {
    reinterpret_cast<double*>(rec + L.off_sector)[i] = sector[l * L.S + i];
    reinterpret_cast<double*>(rec + L.off_norm)[i] = colnorm[l * L.S + i];
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 3: exact brute-force ring-key top-K (replaces the nanoflann KD-tree and its rebuild).
//   Keys are (dist2 bits << 32 | global index): squared distances are non-negative floats, so unsigned
//   64-bit order == (dist2, index) order -- the deterministic tie-break.
//   Each warp keeps a sorted list of 32*SLOTS keys spread over its lanes (slot s lives in lane s%32,
//   register s/32); a candidate is inserted only if it beats the current K-th key.
// ------------------------------------------------------------------------------------------------
template <int SLOTS>
struct WarpList {
  unsigned long long v[SLOTS];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) v[s] = KEY_NONE;
  }
  __device__ __forceinline__ unsigned long long kth(int K) const {  // value of slot K-1 (warp-uniform)
    unsigned long long r = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const unsigned long long t = __shfl_sync(FULL, v[s], (K - 1) & 31);
      if (((K - 1) >> 5) == s) r = t;
    }
    return r;
  }
  __device__ __forceinline__ void insert(unsigned long long c) {  // c is warp-uniform
    const int lane = threadIdx.x & 31;
    int pos = 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) pos += __popc(__ballot_sync(FULL, v[s] < c));
    unsigned long long carry = 0;  // last element of the previous register row
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      unsigned long long up = __shfl_up_sync(FULL, v[s], 1);
      const unsigned long long last = __shfl_sync(FULL, v[s], 31);
      if (lane == 0) up = carry;
      carry = last;
      const int slot = s * 32 + lane;
      if (slot > pos) v[s] = up;
      else if (slot == pos) v[s] = c;
    }
  }
  // offer one key per lane (KEY_NONE = nothing)
  __device__ __forceinline__ void offer(unsigned long long key, int K) {
    unsigned long long w = kth(K);
    unsigned cand = __ballot_sync(FULL, key < w);
    while (cand) {
      const int src = __ffs(cand) - 1;
      cand &= cand - 1;
      const unsigned long long c = __shfl_sync(FULL, key, src);
      if (c < w) {
        insert(c);
        w = kth(K);
      }
    }
  }
  // offer `count` keys from memory
  __device__ __forceinline__ void offer_from(const unsigned long long* src, int count, int K) {
    const int lane = threadIdx.x & 31;
    for (int b = 0; b < count; b += 32) {
      const int i = b + lane;
      offer(i < count ? src[i] : KEY_NONE, K);
    }
  }
  __device__ __forceinline__ void store(unsigned long long* dst, int K) const {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if (s * 32 + lane < K) dst[s * 32 + lane] = v[s];
  }
};

// nf.hpp:383-408: four squared differences added left to right, then added to the running sum; FP32, no FMA.
// RC > 0: compile-time dimension -- the loop unrolls and all RC loads are in flight together (with a run-time bound
// every group of four loads was a separate round trip to L2); RC = 0: any dimension.
template <int RC>
__device__ __forceinline__ float ringkey_dist2(const float* q, const float* ringT, unsigned long long l, int R_rt) {
  const int R = RC > 0 ? RC : R_rt;
  const float* base = ringT + (size_t)(l >> 5) * (size_t)(R * 32) + (size_t)(l & 31);  // element d at base[d * 32] (ring_at)
  float result = 0.f;
  if (RC > 0) {
    float kv[RC > 0 ? RC : 1];
#pragma unroll
    for (int d = 0; d < RC; ++d) kv[d] = __ldg(base + d * 32);
    int d = 0;
#pragma unroll
    for (; d + 3 < RC; d += 4) {
      const float d0 = __fsub_rn(q[d], kv[d]), d1 = __fsub_rn(q[d + 1], kv[d + 1]);
      const float d2 = __fsub_rn(q[d + 2], kv[d + 2]), d3 = __fsub_rn(q[d + 3], kv[d + 3]);
      const float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
      result = __fadd_rn(result, t);
    }
#pragma unroll
    for (; d < RC; ++d) {
      const float d0 = __fsub_rn(q[d], kv[d]);
      result = __fadd_rn(result, __fmul_rn(d0, d0));
    }
    return result;
  }
  int d = 0;
  for (; d + 3 < R; d += 4) {
    const float d0 = __fsub_rn(q[d], __ldg(base + d * 32));
    const float d1 = __fsub_rn(q[d + 1], __ldg(base + (d + 1) * 32));
    const float d2 = __fsub_rn(q[d + 2], __ldg(base + (d + 2) * 32));
    const float d3 = __fsub_rn(q[d + 3], __ldg(base + (d + 3) * 32));
    const float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
    result = __fadd_rn(result, t);
  }
  for (; d < R; ++d) {
    const float d0 = __fsub_rn(q[d], __ldg(base + d * 32));
    result = __fadd_rn(result, __fmul_rn(d0, d0));
  }
  return result;
}

struct TopkParams {
  const unsigned char* qrecords;     // query records
  Layout L;
  Db db;
  const unsigned long long* n_search;  // per query: global entries [0, n_search) are searchable
  unsigned long long n_local;          // local entries stored
  unsigned chunk;                      // local entries per block
  int K;
  unsigned long long* partial;         // [nq][chunks][K]
  unsigned* tickets;                   // [nq], pre-set to 0
  unsigned long long* keys_out;        // [nq][K]
};

template <int RC>
__device__ __forceinline__ float ringkey_dist2_regs(const float* q, const float (&kv)[RC]) {
  float result = 0.f;
  int d = 0;
#pragma unroll
  for (; d + 3 < RC; d += 4) {
    const float d0 = __fsub_rn(q[d], kv[d]), d1 = __fsub_rn(q[d + 1], kv[d + 1]);
    const float d2 = __fsub_rn(q[d + 2], kv[d + 2]), d3 = __fsub_rn(q[d + 3], kv[d + 3]);
    const float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
    result = __fadd_rn(result, t);
  }
#pragma unroll
  for (; d < RC; ++d) {
    const float d0 = __fsub_rn(q[d], kv[d]);
    result = __fadd_rn(result, __fmul_rn(d0, d0));
  }
  return result;
}

// the streaming loop of k_topk for a compile-time key dimension: keys [first, end) in steps of `step`, two per lane
template <int SLOTS, int RC>
__device__ __forceinline__ void topk_stream(WarpList<SLOTS>& list, const float* q, const float* ringT, unsigned long long first, unsigned long long end,
                                            unsigned step, unsigned long long G, unsigned long long rank, int K) {
  const int lane = threadIdx.x & 31;
  float kn[2][RC];
  auto load = [&](unsigned long long base) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned long long l = base + 32 * u + lane;
      const float* b = ringT + (size_t)(l >> 5) * (size_t)(RC * 32) + (size_t)(l & 31);
#pragma unroll
      for (int d = 0; d < RC; ++d) kn[u][d] = l < end ? __ldg(b + d * 32) : 0.f;
    }
  };
  if (first < end) load(first);
  for (unsigned long long base = first; base < end; base += step) {
    float kv[2][RC];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int d = 0; d < RC; ++d) kv[u][d] = kn[u][d];
    if (base + step < end) load(base + step);  // in flight during the work below
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned long long l = base + 32 * u + lane;
      const float d2 = ringkey_dist2_regs<RC>(q, kv[u]);
      const unsigned long long key = l < end ? (((unsigned long long)__float_as_uint(d2) << 32) | (l * G + rank)) : KEY_NONE;
      list.offer(key, K);
    }
  }
}

// k_topk_qs -- one WARP per query, QPB consecutive queries per block, the ring keys staged ONCE per block in shared memory
// (cp.async, two stages of 20 KB: the tiled layout makes a run of keys one contiguous span).  Against k_topk for batches of
// hundreds to thousands of queries over a few thousand keys (the replay of a run): every key leaves L2 once per QPB queries
// instead of once per query (4,541 queries x ~2,270 visible keys x 80 B = 825 MB of L2 reads per step: the kernel was
// L2-bound), the distance reads are shared-memory loads (29 cycles instead of an L2 round trip: few resident warps suffice),
// and a query has ONE list fed by its whole stream (k_topk: two to eight lists per query, each paying its warm-up
// insertions, plus a merge).  Consecutive queries see almost the same prefix of the database (n_search = index - 50), so a
// block streams to the largest bound of its queries and every warp stops at its own.  K <= 32.
template <int RC, int QPB>
__global__ void __launch_bounds__(QPB * 32) k_topk_qs(const TopkParams p, unsigned nq) {
  constexpr int TILE = RC == 20 ? 256 : 128;                 // keys per stage
  constexpr int STAGE_F4 = TILE * RC / 4;                    // float4 per stage
  __shared__ __align__(16) float s_keys[2][TILE * RC];
  __shared__ float s_q[QPB][RC];
  __shared__ unsigned long long s_vis[QPB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, K = p.K;
  const unsigned q = blockIdx.x * QPB + warp;
  const unsigned long long G = (unsigned long long)p.db.G, rk = (unsigned long long)p.db.rank;
  unsigned long long vis = 0;
  if (q < nq) {
    const unsigned long long ns = p.n_search[q];
    if (ns > rk) vis = (ns - 1 - rk) / G + 1;
    if (vis > p.n_local) vis = p.n_local;
    const float* qring = reinterpret_cast<const float*>(p.qrecords + (size_t)q * p.L.rec_bytes + p.L.off_ring);
    if (lane < RC) s_q[warp][lane] = qring[lane];
    if (RC > 32 && lane + 32 < RC) s_q[warp][lane + 32] = qring[lane + 32];
  }
  if (lane == 0) s_vis[warp] = vis;
  __syncthreads();
  unsigned long long vmax = 0;
#pragma unroll
  for (int w = 0; w < QPB; ++w) vmax = s_vis[w] > vmax ? s_vis[w] : vmax;
  float qr[RC];
#pragma unroll
  for (int d = 0; d < RC; ++d) qr[d] = s_q[warp][d];
  const unsigned n_tiles = (unsigned)((vmax + TILE - 1) / TILE);
  const unsigned long long alloc_f4 = p.db.cap * (unsigned long long)RC / 4;  // float4 of the key array (capacities are multiples of 64 keys)
  const float4* g4 = reinterpret_cast<const float4*>(p.db.ringT);
  auto stage = [&](unsigned t) {
    float4* dst = reinterpret_cast<float4*>(s_keys[t & 1]);
    const unsigned long long base = (unsigned long long)t * STAGE_F4;
    for (int i = threadIdx.x; i < STAGE_F4; i += QPB * 32)
      if (base + i < alloc_f4) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g4 + base + i) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  WarpList<1> list;
  list.init();
  if (n_tiles) stage(0);
  for (unsigned t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) {
      stage(t + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // stage t is complete for every thread's copies
    const unsigned long long base = (unsigned long long)t * TILE;
    const float* sk = s_keys[t & 1];
    if (base < vis) {  // (warp-uniform)
#pragma unroll 2
      for (int g = 0; g < TILE / 32; ++g) {
        const unsigned long long l = base + g * 32 + lane;
        if (base + g * 32 >= vis) break;
        float kv[RC];
#pragma unroll
        for (int d = 0; d < RC; ++d) kv[d] = sk[(g * RC + d) * 32 + lane];
        const float d2 = ringkey_dist2_regs<RC>(qr, kv);
        list.offer(l < vis ? (((unsigned long long)__float_as_uint(d2) << 32) | (l * G + rk)) : KEY_NONE, K);
      }
    }
    __syncthreads();  // nobody still reads the stage that the next iteration's copy overwrites
  }
  if (q < nq) list.store(p.keys_out + (size_t)q * K, K);
}

// TOPK_WARPS warps share one (query, chunk).  Every warp's list has a warm-up phase of ~K(1 + ln(n_warp / K)) insertions,
// so many queries per launch use few warps per query (long streams per warp), few queries use many (latency).
// RC: compile-time key dimension (20, 40; 0 = any).  PIPE: software-pipelined key loads (more registers: chosen for small grids,
// where too few warps are resident to hide the L2 latency of the loads).
template <int SLOTS, int TOPK_WARPS, int RC, bool PIPE>
__global__ void __launch_bounds__(TOPK_WARPS * 32) k_topk(const TopkParams p) {
  constexpr int TOPK_THREADS = TOPK_WARPS * 32;
  __shared__ float s_q[64];
  __shared__ unsigned long long s_lists[TOPK_WARPS * 32 * SLOTS];
  __shared__ bool s_last;
  const int q = blockIdx.y, K = p.K, R = p.L.R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qring = reinterpret_cast<const float*>(p.qrecords + (size_t)q * p.L.rec_bytes + p.L.off_ring);
  if (threadIdx.x < R) s_q[threadIdx.x] = qring[threadIdx.x];
  __syncthreads();
  const unsigned long long ns = p.n_search[q];
  // local entries l with l*G + rank < ns
  unsigned long long vis = 0;
  if (ns > (unsigned long long)p.db.rank) vis = (ns - 1 - p.db.rank) / p.db.G + 1;
  if (vis > p.n_local) vis = p.n_local;
  const unsigned long long start = (unsigned long long)blockIdx.x * p.chunk;
  unsigned long long end = start + p.chunk;
  if (end > vis) end = vis;

  WarpList<SLOTS> list;
  list.init();
  // Two keys per lane per iteration, software-pipelined: the 2 x R loads of the NEXT iteration are issued before the current keys are
  // scored and offered, so the L2 latency of the key loads (the kernel's top stall: ~15 warps per SM cannot hide it) overlaps the
  // arithmetic and the list maintenance of the current ones.
  const unsigned long long G = (unsigned long long)p.db.G, rk = (unsigned long long)p.db.rank;
  if (PIPE && RC > 0) {
    topk_stream<SLOTS, (RC > 0 ? RC : 4)>(list, s_q, p.db.ringT, start + (unsigned long long)warp * 64, end, 2 * TOPK_THREADS, G, rk, K);
  } else {
    for (unsigned long long base = start + (unsigned long long)warp * 64; base < end; base += 2 * TOPK_THREADS) {
      unsigned long long k2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const unsigned long long l = base + 32 * u + lane;
        k2[u] = KEY_NONE;
        if (l < end) k2[u] = ((unsigned long long)__float_as_uint(ringkey_dist2<RC>(s_q, p.db.ringT, l, R)) << 32) | (l * G + rk);
      }
      list.offer(k2[0], K);
      list.offer(k2[1], K);
    }
  }
  // block merge: warps > 0 publish, warp 0 absorbs
  if (warp > 0) list.store(s_lists + warp * 32 * SLOTS, K);
  __syncthreads();
  unsigned long long* part = p.partial + ((size_t)q * gridDim.x + blockIdx.x) * K;
  if (warp == 0) {
    for (int w = 1; w < TOPK_WARPS; ++w) list.offer_from(s_lists + w * 32 * SLOTS, K, K);
    if (gridDim.x == 1) {
      list.store(p.keys_out + (size_t)q * K, K);
      return;
    }
    list.store(part, K);
    __threadfence();
  }
  if (gridDim.x == 1) return;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.tickets[q], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  // last block of this query: merge every chunk's list (volatile-ish reads through L2)
  list.init();
  const unsigned long long* all = p.partial + (size_t)q * gridDim.x * K;
  const int total = (int)gridDim.x * K;
  for (int b = 0; b < total; b += 32) {
    const int i = b + lane;
    list.offer(i < total ? __ldcg(all + i) : KEY_NONE, K);
  }
  list.store(p.keys_out + (size_t)q * K, K);
  if (lane == 0) p.tickets[q] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// k_topk_tile: the form for LARGE databases searched by LARGE query batches.  A block takes one chunk of the database
// and a GROUP of QT queries: every ring key is loaded from L2 once per group instead of once per query (k_topk at 4,096
// queries x 100k keys moves 33 GB through L2 in 3.5 ms: L2-bandwidth-bound), the loads of the next 32 keys are issued
// before the current ones are used, and the K-th key of every list is cached in a register so that the common case (no
// insertion) costs one vote.  Same arithmetic, same (dist2, index) order, same partial-list / last-block-done merge as
// k_topk; one ticket per group.  Only used when every warp's stream is long (>= 8k keys): each (warp, query) list pays
// ~K(1 + ln(n/K)) warm-up insertions, which on short streams costs more than the saved L2 traffic (measured: the
// 4,541-keyframe run's query stage 0.40 -> 0.51 ms when this kernel was used unconditionally).
// ------------------------------------------------------------------------------------------------------------
constexpr int TOPK_QT = 8;        // queries per group
constexpr int TOPK_TILE_WARPS = 4;

template <int SLOTS, int RC>
__global__ void __launch_bounds__(TOPK_TILE_WARPS * 32) k_topk_tile(const TopkParams p, unsigned nq) {
  constexpr int QT = TOPK_QT, TW = TOPK_TILE_WARPS;
  __shared__ __align__(16) float s_q[QT][RC];
  __shared__ unsigned long long s_vis[QT];
  __shared__ unsigned long long s_lists[TW][QT][32 * SLOTS];
  __shared__ bool s_last;
  const int K = p.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned q0 = blockIdx.y * QT;
  const int nqg = (int)min((unsigned)QT, nq - q0);  // queries in this group
  for (int i = threadIdx.x; i < QT * RC; i += blockDim.x) {
    const int q = i / RC, d = i - q * RC;
    s_q[q][d] = q < nqg ? reinterpret_cast<const float*>(p.qrecords + (size_t)(q0 + q) * p.L.rec_bytes + p.L.off_ring)[d] : 0.f;
  }
  if (threadIdx.x < QT) {
    unsigned long long vis = 0;
    if ((int)threadIdx.x < nqg) {
      const unsigned long long ns = p.n_search[q0 + threadIdx.x];
      if (ns > (unsigned long long)p.db.rank) vis = (ns - 1 - p.db.rank) / p.db.G + 1;  // local entries l with l*G + rank < ns
      if (vis > p.n_local) vis = p.n_local;
    }
    s_vis[threadIdx.x] = vis;
  }
  __syncthreads();
  unsigned long long vmax = 0;
#pragma unroll
  for (int q = 0; q < QT; ++q) vmax = s_vis[q] > vmax ? s_vis[q] : vmax;
  const unsigned long long start = (unsigned long long)blockIdx.x * p.chunk;
  unsigned long long end = start + p.chunk;
  if (end > vmax) end = vmax;

  WarpList<SLOTS> list[QT];
  unsigned long long wk[QT];  // K-th key of every list (warp-uniform)
#pragma unroll
  for (int q = 0; q < QT; ++q) {
    list[q].init();
    wk[q] = KEY_NONE;
  }
  float kv[RC], kn[RC];
  const float* col = p.db.ringT;
  {
    const unsigned long long l = start + (unsigned long long)warp * 32 + lane;
#pragma unroll
    for (int d = 0; d < RC; ++d) kn[d] = l < end ? __ldg(col + ring_at(l, d, RC)) : 0.f;
  }
  for (unsigned long long base = start + (unsigned long long)warp * 32; base < end; base += TW * 32) {
#pragma unroll
    for (int d = 0; d < RC; ++d) kv[d] = kn[d];
    const unsigned long long lcur = base + lane, lnext = lcur + TW * 32;
#pragma unroll
    for (int d = 0; d < RC; ++d) kn[d] = lnext < end ? __ldg(col + ring_at(lnext, d, RC)) : 0.f;  // in flight during the work below
    const unsigned long long idx = lcur * p.db.G + p.db.rank;
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const float d2 = ringkey_dist2_regs<RC>(s_q[q], kv);
      const unsigned long long key = lcur < s_vis[q] ? (((unsigned long long)__float_as_uint(d2) << 32) | idx) : KEY_NONE;
      if (__any_sync(FULL, key < wk[q])) {
        list[q].offer(key, K);
        wk[q] = list[q].kth(K);
      }
    }
  }
  // block merge: every warp publishes its QT lists, warp w then owns queries w, w + TW, ...
#pragma unroll
  for (int q = 0; q < QT; ++q) list[q].store(&s_lists[warp][q][0], K);
  __syncthreads();
  const unsigned chunks = gridDim.x;
  for (int q = warp; q < nqg; q += TW) {
    WarpList<SLOTS> m;
    m.init();
    for (int w = 0; w < TW; ++w) m.offer_from(&s_lists[w][q][0], K, K);
    if (chunks == 1) m.store(p.keys_out + (size_t)(q0 + q) * K, K);
    else m.store(p.partial + ((size_t)(q0 + q) * chunks + blockIdx.x) * K, K);
  }
  if (chunks == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&p.tickets[blockIdx.y], 1u) == chunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block of this group: merge every chunk's list of each of its queries (reads through L2)
  for (int q = warp; q < nqg; q += TW) {
    WarpList<SLOTS> m;
    m.init();
    const unsigned long long* all = p.partial + (size_t)(q0 + q) * chunks * K;
    const int total = (int)chunks * K;
    for (int b = 0; b < total; b += 32) {
      const int i = b + lane;
      m.offer(i < total ? __ldcg(all + i) : KEY_NONE, K);
    }
    m.store(p.keys_out + (size_t)(q0 + q) * K, K);
  }
  if (threadIdx.x == 0) p.tickets[blockIdx.y] = 0;
}

// merge `parts` lists per query: in [parts][nq][K] -> out [nq][K]; one warp per query
template <int SLOTS>
__global__ void __launch_bounds__(128) k_merge(const unsigned long long* in, int parts, unsigned nq, int K, unsigned long long* out) {
  const unsigned q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  WarpList<SLOTS> list;
  list.init();
  for (int p = 0; p < parts; ++p) list.offer_from(in + ((size_t)p * nq + q) * K, K, K);
  list.store(out + (size_t)q * K, K);
}

// ------------------------------------------------------------------------------------------------
// Stage 4: SC distance of one (query, candidate) pair -- SC.cpp:116-148 in FP64, reference order.
//   T = float for database descriptors (bins are exact floats), double for the public pairwise API.
// ------------------------------------------------------------------------------------------------
struct PairSmem {
  double* vk1;   // [S] sector key of sc1
  double* vk2;   // [S]
  double* n1;    // [S] column norms of sc1
  double* n2;    // [S]
  double* work;  // [max(S, WB*S)] per-shift norms, then per (shift, column) similarities
  int* shifts;   // [W]
  double* dist;  // [W]
};

constexpr int SCORE_WB = 16;  // shifts evaluated per pass (bounds the shared-memory footprint)

// descriptors sit in shared memory with an ODD column pitch: threads that work on consecutive columns hit distinct banks
__host__ __device__ inline int pair_pitch(int R) { return R | 1; }

__host__ __device__ inline size_t pair_smem_bytes(int R, int S, int W, size_t elem) {
  size_t b = 2 * (size_t)pair_pitch(R) * S * elem;  // the two descriptors
  b = (b + 7) & ~(size_t)7;
  b += 4 * (size_t)S * 8;                      // vk1 vk2 n1 n2
  const int wb = W < SCORE_WB ? W : SCORE_WB;
  b += (size_t)(wb > 1 ? wb : 1) * S * 8;      // work
  b += (size_t)W * 8;                          // dist
  b += (size_t)W * 4;                          // shifts
  return (b + 15) & ~(size_t)15;
}

template <class T>
__device__ __forceinline__ PairSmem carve_pair_smem(unsigned char* raw, int R, int S, int W, T*& a, T*& b) {
  a = reinterpret_cast<T*>(raw);
  b = a + (size_t)pair_pitch(R) * S;
  size_t off = 2 * (size_t)pair_pitch(R) * S * sizeof(T);
  off = (off + 7) & ~(size_t)7;
  PairSmem m;
  m.vk1 = reinterpret_cast<double*>(raw + off);
  m.vk2 = m.vk1 + S;
  m.n1 = m.vk2 + S;
  m.n2 = m.n1 + S;
  m.work = m.n2 + S;
  const int wb = W < SCORE_WB ? W : SCORE_WB;
  m.dist = m.work + (size_t)(wb > 1 ? wb : 1) * S;
  m.shifts = reinterpret_cast<int*>(m.dist + W);
  return m;
}

// SC.cpp:93-113.  Needs m.vk1/m.vk2; returns the argmin shift to every thread.  Strict '<' in ascending shift
// order == lexicographic min of (norm, shift) over norms < 1e7.
__device__ __forceinline__ int fast_align_block(const PairSmem& m, int S) {
  __shared__ int s_arg;
  for (int s = threadIdx.x; s < S; s += blockDim.x)
    m.work[s] = __dsqrt_rn(redux_eigen(S, CoeffShiftDiffSq{m.vk1, m.vk2, s, S}));
  __syncthreads();
  if (threadIdx.x < 32) {
    double best = 10000000.0;
    int arg = 0x7fffffff;
    for (int s = threadIdx.x; s < S; s += 32) {
      const double v = m.work[s];
      if (v < best) {  // ascending s within a lane: strict '<' keeps the smallest s
        best = v;
        arg = s;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oa = __shfl_xor_sync(FULL, arg, o);
      if (ob < best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (threadIdx.x == 0) s_arg = (arg == 0x7fffffff) ? 0 : arg;
  }
  __syncthreads();
  const int r = s_arg;
  __syncthreads();
  return r;
}

// SC.cpp:69-90 for the shifts m.shifts[0..W): m.dist[w] = distDirectSC(sc1, circshift(sc2, shift_w)).
// Needs a, b (descriptors) and m.n1/m.n2 (column norms).
template <class T>
__device__ __forceinline__ void dist_direct_block(const PairSmem& m, const T* a, const T* b, int R, int S, int W) {
  for (int w0 = 0; w0 < W; w0 += SCORE_WB) {
    const int wn = min(SCORE_WB, W - w0);
    for (int t = threadIdx.x; t < wn * S; t += blockDim.x) {
      const int w = t / S, j = t - w * S;
      int jb = j - m.shifts[w0 + w];
      if (jb < 0) jb += S;
      const double na = m.n1[j], nb = m.n2[jb];
      double sim = 0.0;
      if (!((na == 0.0) | (nb == 0.0))) {
        const double dot = redux_eigen(R, CoeffProduct<T>{a + (size_t)j * pair_pitch(R), b + (size_t)jb * pair_pitch(R)});
        sim = __ddiv_rn(dot, __dmul_rn(na, nb));
      }
      m.work[t] = sim;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < wn; w += blockDim.x) {
      const int s = m.shifts[w0 + w];
      double sum = 0.0;
      int num = 0;
      for (int j = 0; j < S; ++j) {
        int jb = j - s;
        if (jb < 0) jb += S;
        if ((m.n1[j] == 0.0) | (m.n2[jb] == 0.0)) continue;
        sum = __dadd_rn(sum, m.work[w * S + j]);
        ++num;
      }
      m.dist[w0 + w] = __dsub_rn(1.0, __ddiv_rn(sum, (double)num));  // 0/0 -> NaN like the reference
    }
    __syncthreads();
  }
}

// SC.cpp:123-144: the shift search space around `a` and the strict-min over it in ascending shift order
// (== lexicographic min of (dist, shift) over dist < 1e7; default (1e7, 0)).
__device__ __forceinline__ void fill_shifts(const PairSmem& m, int a, int radius, int S) {
  if (threadIdx.x == 0) {
    m.shifts[0] = a;
    for (int ii = 1; ii <= radius; ++ii) {
      m.shifts[2 * ii - 1] = (a + ii + S) % S;
      m.shifts[2 * ii] = ((a - ii + S) % S + S) % S;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void pick_min(const PairSmem& m, int W, double& dist, int& shift) {
  dist = 10000000.0;
  shift = 0;
  bool any = false;
  for (int w = 0; w < W; ++w) {
    const double d = m.dist[w];
    const int s = m.shifts[w];
    if (d < 10000000.0 && (!any || d < dist || (d == dist && s < shift))) {
      dist = d;
      shift = s;
      any = true;
    }
  }
}

struct ScoreParams {
  const unsigned char* qrecords;
  Layout L;
  Db db;
  const unsigned long long* keys;      // [nq][K] candidate keys (KEY_NONE = unfilled slot -> entry 0, SC.cpp:283-284)
  const unsigned long long* n_search;  // [nq]; 0 = query took the early return (SC.cpp:257-261)
  int K, radius;
  double* pair_dist;  // [nq][K]
  int* pair_shift;    // [nq][K]; -1 = candidate not owned by this shard
  int flip;           // score the candidate with its columns reversed (composed "reverse loop" search)
  const unsigned* active;  // optional: only slots k < *active are scored (fixed-size grid over a device-side count)
  PeerTab peers;      // peers.G > 0: every candidate is scored here, its data read from the owner shard (NVLink peer loads)
};

template <bool LIST>  // LIST: slot k of one flat list whose keys are (query index << 32 | global entry index)
__device__ __forceinline__ void score_pair(const ScoreParams& p, const int k, int q, unsigned char* smem_raw) {
  const int R = p.L.R, S = p.L.S, W = 2 * p.radius + 1;
  const size_t o = LIST ? (size_t)k : (size_t)q * p.K + k;
  if (LIST) q = (int)((p.keys[o] >> 32) & 0x7fffffffull);
  if (!LIST && p.n_search[q] == 0) {
    if (threadIdx.x == 0) {
      p.pair_dist[o] = 10000000.0;
      p.pair_shift[o] = -1;
    }
    return;
  }
  const unsigned long long key = p.keys[o];
  const unsigned long long g = (!LIST && key == KEY_NONE) ? 0ull : (key & 0xffffffffull);
  const unsigned long long G = p.peers.G ? (unsigned long long)p.peers.G : (unsigned long long)p.db.G;
  const int owner = (int)(g % G);
  if (!p.peers.G && owner != p.db.rank) {
    if (threadIdx.x == 0) {
      p.pair_dist[o] = 10000000.0;
      p.pair_shift[o] = -1;
    }
    return;
  }
  const unsigned long long l = g / G;
  float *a, *b;
  PairSmem m = carve_pair_smem<float>(smem_raw, R, S, W, a, b);
  const unsigned char* qrec = p.qrecords + (size_t)q * p.L.rec_bytes;
  const float* qsc = reinterpret_cast<const float*>(qrec);
  const float* csc = (p.peers.G ? p.peers.sc[owner] : p.db.sc) + l * p.L.RS;
  const double* csector = (p.peers.G ? p.peers.sector[owner] : p.db.sector) + l * S;
  const double* cnorm = (p.peers.G ? p.peers.colnorm[owner] : p.db.colnorm) + l * S;
  const int RP = pair_pitch(R);
  const bool flip = LIST ? (key >> 63) != 0 : p.flip != 0;
  for (int i = threadIdx.x; i < p.L.RS; i += blockDim.x) {
    const int c = i / R, r = i - c * R;
    a[c * RP + r] = qsc[i];
    b[c * RP + r] = csc[flip ? (S - 1 - c) * R + r : i];
  }
  const double* qv = reinterpret_cast<const double*>(qrec + p.L.off_sector);
  const double* qn = reinterpret_cast<const double*>(qrec + p.L.off_norm);
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const int ci = flip ? (S - 1 - i) : i;
    m.vk1[i] = qv[i];
    m.n1[i] = qn[i];
    m.vk2[i] = csector[ci];
    m.n2[i] = cnorm[ci];
  }
  __syncthreads();
  const int align = fast_align_block(m, S);
  fill_shifts(m, align, p.radius, S);
  dist_direct_block<float>(m, a, b, R, S, W);
  if (threadIdx.x == 0) {
    double d;
    int s;
    pick_min(m, W, d, s);
    p.pair_dist[o] = d;
    p.pair_shift[o] = s;
  }
}

// grid (K, nq): block (k, q) scores candidate slot k of query q.
__global__ void __launch_bounds__(128) k_score(const ScoreParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  score_pair<false>(p, blockIdx.x, blockIdx.y, smem_raw);
}

// Exhaustive rescoring: a flat candidate list (keys = query << 32 | entry) with a device-side count (p.active); a
// persistent grid strides over it.
__global__ void __launch_bounds__(128) k_score_list(const ScoreParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned n = min(*p.active, (unsigned)p.K);
  for (unsigned k = blockIdx.x; k < n; k += gridDim.x) {
    score_pair<true>(p, (int)k, 0, smem_raw);
    __syncthreads();
  }
}

// The public pairwise functions on caller-provided double matrices (SC.h:64-69).
//   mode 0: distanceBtnScanContext -> out_d[0] = dist, out_i[0] = shift
//   mode 1: distDirectSC           -> out_d[0]
//   mode 2: fastAlignUsingVkey on keys passed in sc1/sc2 (S doubles each) -> out_i[0]
//   mode 3: ring key + sector key of sc1 -> out_d[0..R) ring, out_d[R..R+S) sector
__global__ void __launch_bounds__(128) k_pair_api(const double* sc1, const double* sc2, int R, int S, int radius, int mode,
                                                  double* out_d, int* out_i) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int W = (mode == 0) ? 2 * radius + 1 : 1;
  double *a, *b;
  PairSmem m = carve_pair_smem<double>(smem_raw, R, S, W, a, b);
  if (mode == 2) {
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      m.vk1[i] = sc1[i];
      m.vk2[i] = sc2[i];
    }
    __syncthreads();
    const int al = fast_align_block(m, S);
    if (threadIdx.x == 0) out_i[0] = al;
    return;
  }
  const int RP = pair_pitch(R);
  for (int i = threadIdx.x; i < R * S; i += blockDim.x) {
    const int c = i / R, r = i - c * R;
    a[c * RP + r] = sc1[i];
    if (mode != 3) b[c * RP + r] = sc2[i];
  }
  __syncthreads();
  if (mode == 3) {
    keys_from_sc<double>(a, R, S, RP, out_d, nullptr, out_d + R, nullptr);
    return;
  }
  keys_from_sc<double>(a, R, S, RP, nullptr, nullptr, m.vk1, m.n1);
  keys_from_sc<double>(b, R, S, RP, nullptr, nullptr, m.vk2, m.n2);
  __syncthreads();
  if (mode == 1) {
    if (threadIdx.x == 0) m.shifts[0] = 0;
    __syncthreads();
    dist_direct_block<double>(m, a, b, R, S, 1);
    if (threadIdx.x == 0) out_d[0] = m.dist[0];
    return;
  }
  const int align = fast_align_block(m, S);
  fill_shifts(m, align, radius, S);
  dist_direct_block<double>(m, a, b, R, S, W);
  if (threadIdx.x == 0) {
    double d;
    int s;
    pick_min(m, W, d, s);
    out_d[0] = d;
    out_i[0] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// SC.cpp:296-311: per query, strict-min over this shard's candidates in retrieval order.
// ------------------------------------------------------------------------------------------------
struct Best {
  double dist;
  int rank;   // position in the candidate list; K = none
  int shift;
  long long idx;
};

// count (optional): the rescoring list's counter, re-armed for the next batch
__global__ void k_best(const double* pair_dist, const int* pair_shift, const unsigned long long* keys, unsigned nq, int K, Best* out,
                       unsigned* count) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (count && q == 0) *count = 0;
  if (q >= nq) return;
  Best b;
  b.dist = 10000000.0;
  b.rank = K;
  b.shift = 0;
  b.idx = 0;
  for (int k = 0; k < K; ++k) {
    const size_t o = (size_t)q * K + k;
    if (pair_shift[o] < 0) continue;
    const double d = pair_dist[o];
    if (d < b.dist) {
      b.dist = d;
      b.rank = k;
      b.shift = pair_shift[o];
      const unsigned long long key = keys[o];
      b.idx = (key == KEY_NONE) ? 0 : (long long)(key & 0xffffffffull);
    }
  }
  out[q] = b;
}

// SC.cpp:317-336 over `parts` shard results: threshold and yaw.  yaw = deg2rad(float(shift * 360/S)) with
// deg2rad(d) = float(double(d) * M_PI / 180.0) (SC.cpp:17-20, 333).
// Result placement (peer-sharded replay: the queries of one shard are every G-th entry of the batch, and every shard gets every
// result): query q goes to slot out_off + q * out_step of the local arrays and of the same arrays on the devices in `push`,
// whose dst[s] is the base of a result block laid out like ResultBlock below.
struct ResultBlock {  // capacity cap_q queries: double dist[cap_q] | int loop[cap_q] | float yaw[cap_q] | int idx[cap_q] | int shift[cap_q]
  unsigned long long cap_q;
  __host__ __device__ double* dist(void* base) const { return static_cast<double*>(base); }
  __host__ __device__ int* loop(void* base) const { return reinterpret_cast<int*>(static_cast<unsigned char*>(base) + cap_q * 8); }
  __host__ __device__ float* yaw(void* base) const { return reinterpret_cast<float*>(static_cast<unsigned char*>(base) + cap_q * 12); }
  __host__ __device__ int* idx(void* base) const { return reinterpret_cast<int*>(static_cast<unsigned char*>(base) + cap_q * 16); }
  __host__ __device__ int* shift(void* base) const { return reinterpret_cast<int*>(static_cast<unsigned char*>(base) + cap_q * 20); }
  __host__ __device__ size_t bytes() const { return (size_t)cap_q * 24; }
};

__global__ void k_finalize(const Best* parts_in, int parts, unsigned nq, const unsigned long long* n_search, int K, int S,
                           double thres, int* loop_id, float* yaw, double* nearest_dist, int* nearest_idx, int* nearest_shift,
                           unsigned out_off, unsigned out_step, ResultBlock rb, PushList push) {
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double dist = 10000000.0;
  int rank = K, shift = 0;
  long long idx = 0;
  if (n_search[q] != 0) {
    for (int p = 0; p < parts; ++p) {
      const Best b = parts_in[(size_t)p * nq + q];
      if (b.rank >= K) continue;
      if (b.dist < dist || (b.dist == dist && b.rank < rank)) {
        dist = b.dist;
        rank = b.rank;
        shift = b.shift;
        idx = b.idx;
      }
    }
  }
  int lid = -1;
  float y = 0.0f;
  if (n_search[q] != 0) {
    if (dist < thres) lid = (int)idx;
    const double unit = __ddiv_rn(360.0, (double)S);                               // SC.h:82
    const float deg = __double2float_rn(__dmul_rn((double)shift, unit));           // SC.cpp:333 argument narrowing
    y = __double2float_rn(__ddiv_rn(__dmul_rn((double)deg, 3.14159265358979323846), 180.0));  // SC.cpp:17-20
  }
  const unsigned o = out_off + q * out_step;
  loop_id[o] = lid;
  yaw[o] = y;
  if (nearest_dist) nearest_dist[o] = dist;
  if (nearest_idx) nearest_idx[o] = (int)idx;
  if (nearest_shift) nearest_shift[o] = shift;
  for (int s = 0; s < push.n; ++s) {
    void* b = push.dst[s];
    rb.loop(b)[o] = lid;
    rb.yaw(b)[o] = y;
    rb.dist(b)[o] = dist;
    rb.idx(b)[o] = (int)idx;
    rb.shift(b)[o] = shift;
  }
}

// ------------------------------------------------------------------------------------------------
// Cross-device barrier of a peer-sharded database whose shards are driven by DIFFERENT processes (one per GPU): thread t
// announces "this shard reached generation `epoch`" in shard t's cell block (a peer store), then waits until shard t's
// announcement has arrived in its own block.  cells: [channel][MAX_SHARDS] generations + cell 2*MAX_SHARDS = timeout flag.
// Everything a shard stored into peer memory before the barrier (ring keys by k_append, results by k_finalize) is
// ordered before its announcement by the system-scope fence.  Generations only grow; comparison is wrap-safe.
// The wait is bounded (timeout_ns of globaltimer, 20 s by default): a missing peer turns into an error code, not a hung GPU.
// Only for shards on DISTINCT devices: kernels that wait on one another must not share a GPU.
// ------------------------------------------------------------------------------------------------
struct BarrierCells {
  unsigned* cells[MAX_SHARDS];
};
__global__ void k_peer_barrier(BarrierCells peers, int G, int rank, int channel, unsigned epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t < G) {
    __threadfence_system();
    volatile unsigned* theirs = peers.cells[t] + channel * MAX_SHARDS + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned* mine = peers.cells[rank] + channel * MAX_SHARDS + t;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        peers.cells[rank][2 * MAX_SHARDS] = 1u;
        break;
      }
      __nanosleep(200);
    }
    __threadfence_system();
  }
}

// k_best + k_finalize in one launch for the paths where ONE device holds every candidate's score (single shard, peer-sharded):
// one warp per query -- lanes take the candidate slots, the strict-min in retrieval order is a warp reduction by (dist, slot).
// Also re-arms the rescoring list's counter for the next batch (count, optional).
__global__ void __launch_bounds__(128) k_best_finalize(const double* pair_dist, const int* pair_shift, const unsigned long long* keys, unsigned nq,
                                                       int K, const unsigned long long* n_search, int S, double thres, int* loop_id, float* yaw,
                                                       double* nearest_dist, int* nearest_idx, int* nearest_shift, unsigned out_off,
                                                       unsigned out_step, ResultBlock rb, PushList push, unsigned* count) {
  const unsigned q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (count && blockIdx.x == 0 && threadIdx.x == 0) *count = 0;
  if (q >= nq) return;
  double dist = 10000000.0;
  int rank = K;
  for (int k = lane; k < K; k += 32) {  // ascending k within a lane: strict '<' keeps the earliest slot
    const size_t o = (size_t)q * K + k;
    if (pair_shift[o] < 0) continue;
    const double d = pair_dist[o];
    if (d < dist) {
      dist = d;
      rank = k;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(FULL, dist, o);
    const int orank = __shfl_xor_sync(FULL, rank, o);
    if (orank < K && (rank >= K || od < dist || (od == dist && orank < rank))) {
      dist = od;
      rank = orank;
    }
  }
  if (lane != 0) return;
  int shift = 0;
  long long idx = 0;
  const bool live = n_search[q] != 0;
  if (rank < K && live) {
    const size_t o = (size_t)q * K + rank;
    shift = pair_shift[o];
    const unsigned long long key = keys[o];
    idx = (key == KEY_NONE) ? 0 : (long long)(key & 0xffffffffull);
  } else {
    dist = 10000000.0;
  }
  int lid = -1;
  float y = 0.0f;
  if (live) {
    if (dist < thres) lid = (int)idx;
    const double unit = __ddiv_rn(360.0, (double)S);                               // SC.h:82
    const float deg = __double2float_rn(__dmul_rn((double)shift, unit));           // SC.cpp:333 argument narrowing
    y = __double2float_rn(__ddiv_rn(__dmul_rn((double)deg, 3.14159265358979323846), 180.0));  // SC.cpp:17-20
  }
  const unsigned o = out_off + q * out_step;
  loop_id[o] = lid;
  yaw[o] = y;
  if (nearest_dist) nearest_dist[o] = dist;
  if (nearest_idx) nearest_idx[o] = (int)idx;
  if (nearest_shift) nearest_shift[o] = shift;
  for (int s = 0; s < push.n; ++s) {
    void* b = push.dst[s];
    rb.loop(b)[o] = lid;
    rb.yaw(b)[o] = y;
    rb.dist(b)[o] = dist;
    rb.idx(b)[o] = (int)idx;
    rb.shift(b)[o] = shift;
  }
}

// device-side probes used by the parity tests (tests/test_gpu_parity.py): atanf / xy2theta / bin of many points
__global__ void k_probe_atanf(const float* x, float* out, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = atanf_fdlibm(x[i]);
}
__global__ void k_probe_bins(const float* xyz, size_t n, BinConst bc, int* bin, float* height, float* theta) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float h;
  const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const int b = bc.fast ? (bc.lh_is_float ? bin_point<true, true>(bc, x, y, z, h) : bin_point<true, false>(bc, x, y, z, h))
                        : bin_point_exact(bc, x, y, z, h);   // the function k_build uses
  bin[i] = b;
  if (b < 0) h = __double2float_rn(__dadd_rn((double)z, bc.lidar_height));
  height[i] = h;
  theta[i] = xy2theta_exact(x, y);
}

// On-device self check of the binning front end: n pseudo-random points per launch, bin_point vs bin_point_exact.
//   mode 0: uniform in the square [-1.15, 1.15] * max_radius;
//   mode 1: adversarial -- on a sector boundary angle or a ring boundary radius, displaced by 2^-e (e = 6..45) of
//           a sector / ring width to either side, so that the hand-over between the two paths is swept densely.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void k_selfcheck(unsigned long long n, unsigned long long seed, int mode, BinConst bc, unsigned long long* mismatches,
                            unsigned long long* fallbacks, float* first_bad) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long h0 = mix64(seed ^ (i * 0x100000001b3ull)), h1 = mix64(h0), h2 = mix64(h1), h3 = mix64(h2);
  const double u0 = (double)(h0 >> 11) * (1.0 / 9007199254740992.0), u1 = (double)(h1 >> 11) * (1.0 / 9007199254740992.0);
  float x, y;
  const float z = (float)(16.0 * ((double)(h2 >> 11) * (1.0 / 9007199254740992.0)) - 4.0);
  if (mode == 0) {
    x = (float)((2.0 * u0 - 1.0) * 1.15 * bc.max_radius);
    y = (float)((2.0 * u1 - 1.0) * 1.15 * bc.max_radius);
  } else {
    const int e = 6 + (int)(h3 % 40);
    const double off = ((h3 >> 8) & 1 ? 1.0 : -1.0) * exp2(-(double)e) * (((h3 >> 9) & 1) ? 1.0 : 0.0);
    double ang, rad;
    if ((h3 >> 10) & 1) {  // sector boundary
      const int k = (int)(h0 % (unsigned long long)(bc.S + 1));
      ang = ((double)k + off) * (6.283185307179586476925 / (double)bc.S);
      rad = u1 * 1.05 * bc.max_radius;
    } else {               // ring boundary
      const int k = 1 + (int)(h0 % (unsigned long long)bc.R);
      rad = ((double)k + off) * (bc.max_radius / (double)bc.R);
      ang = u1 * 6.283185307179586476925;
    }
    x = (float)(rad * cos(ang));
    y = (float)(rad * sin(ang));
    if (((h3 >> 11) & 7) == 0) x = __uint_as_float(__float_as_uint(x) + (unsigned)((h3 >> 14) & 3) - 1u);  // +-1 ulp nudges
    if (((h3 >> 16) & 7) == 0) y = __uint_as_float(__float_as_uint(y) + (unsigned)((h3 >> 19) & 3) - 1u);
  }
  float ha, hb;
  bool fb = false;
  const int a = bc.lh_is_float ? bin_point<true, true>(bc, x, y, z, ha, &fb) : bin_point<true, false>(bc, x, y, z, ha, &fb);
  const int b = bin_point_exact(bc, x, y, z, hb);
  const bool bad = (a != b) || (a >= 0 && __float_as_uint(ha) != __float_as_uint(hb));
  if (fb) atomicAdd(fallbacks, 1ull);
  if (bad) {
    if (atomicAdd(mismatches, 1ull) == 0) {
      first_bad[0] = x;
      first_bad[1] = y;
      first_bad[2] = z;
      first_bad[3] = (float)a;
      first_bad[4] = (float)b;
    }
  }
}

}  // namespace scgpu
