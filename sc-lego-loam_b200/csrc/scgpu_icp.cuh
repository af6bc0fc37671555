// scgpu_icp.cuh -- loop verification after the path: point-to-point ICP of the latest keyframe cloud
// against the history submap, as mapOptmization.cpp:1053-1078 does with pcl::IterativeClosestPoint (max 100 iterations,
// correspondence distance 100, transformation epsilon 1e-6, euclidean fitness epsilon 1e-6, no RANSAC) before it accepts a Scan
// Context loop (fitness <= historyKeyframeFitnessScore = 1.5, utility.h:139).
//
// PARITY UNPINNED: PCL is a third-party dependency that is neither under /root/reference nor in this image.  The algorithm
// below restates PCL 1.8's ICP as published (registration/impl/icp.hpp, default_convergence_criteria.hpp,
// transformation_estimation_svd.hpp): nearest-neighbour correspondences of the transformed source in the target, rejected beyond
// the maximum correspondence distance; rigid transform from the correspondences' cross-covariance (Umeyama / Horn, no scaling);
// convergence when the iteration cap is reached, the incremental transform is below the rotation (cos >= 0.99999) and translation
// (squared <= epsilon) thresholds, or the mean squared correspondence distance changes by less than the absolute (epsilon) or
// relative (1e-5) threshold; fitness = mean squared nearest-neighbour distance of the final alignment.  The oracle
// (oracle/icp_oracle.cpp) is the same restatement on the CPU.
//
// Device side: the nearest-neighbour search is exact brute force -- a keyframe cloud (a few thousand points) against a submap
// (tens of thousands) is 1e8 distance evaluations per iteration, ~30 us on a B200, which a spatial index would not beat at this
// size.  One block handles 256 source points against target tiles staged in shared memory; per-block partial sums (FP64) go to
// global memory and a one-thread kernel reduces them in a fixed order (deterministic), solves for the rotation with Horn's
// quaternion method (Jacobi eigen-decomposition of the 4x4 symmetric matrix, FP64) and applies the convergence tests, so the
// whole loop runs without a host round trip: later iterations see the `done` flag and return immediately.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace scgpu {

constexpr int ICP_BLOCK = 256;
constexpr int ICP_TILE = 1024;

struct IcpState {
  double T[16];        // current source -> target transform, row-major 4x4
  double prev_mse;     // mean squared correspondence distance of the previous iteration
  double mse;          // ... of the last evaluated iteration
  double fitness;      // mean squared NN distance of the final alignment (all points, no distance gate)
  int iterations;      // completed iterations
  int done;            // 1: a convergence criterion fired
  int converged;       // PCL's hasConverged()
  int reason;          // 1 iterations, 2 transform, 3 absolute mse, 4 relative mse, 5 no correspondences
  unsigned long long n_corr;
};

struct IcpPartial {   // per-block sums over accepted correspondences (p = transformed source, q = target)
  double sp[3], sq[3], spq[9], sd2;
  unsigned long long n;
};

// One NN pass.  fitness_pass: no gate, only sd2 / n are meaningful.
__global__ void __launch_bounds__(ICP_BLOCK) k_icp_nn(const unsigned char* src, unsigned n_src, unsigned stride_s, const unsigned char* tgt,
                                                      unsigned n_tgt, unsigned stride_t, const IcpState* st, float max_d2, int fitness_pass,
                                                      IcpPartial* partials) {
  __shared__ float4 s_t[ICP_TILE];
  __shared__ double s_red[ICP_BLOCK / 32][24];
  if (st->done && !fitness_pass) return;
  const unsigned i = blockIdx.x * ICP_BLOCK + threadIdx.x;
  float px = 0.f, py = 0.f, pz = 0.f;
  const bool live = i < n_src;
  if (live) {
    const float* f = reinterpret_cast<const float*>(src + (size_t)i * stride_s);
    const double x = f[0], y = f[1], z = f[2];
    const double* T = st->T;
    px = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
    py = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
    pz = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
  }
  float best = 3.4e38f;
  unsigned arg = 0;
  for (unsigned t0 = 0; t0 < n_tgt; t0 += ICP_TILE) {
    const unsigned nt = min((unsigned)ICP_TILE, n_tgt - t0);
    __syncthreads();
    for (unsigned j = threadIdx.x; j < nt; j += ICP_BLOCK) {
      const float* f = reinterpret_cast<const float*>(tgt + (size_t)(t0 + j) * stride_t);
      s_t[j] = make_float4(f[0], f[1], f[2], 0.f);
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (unsigned j = 0; j < nt; ++j) {
        const float4 q = s_t[j];
        const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < best) {  // strict: the first (lowest-index) target wins a tie
          best = d2;
          arg = t0 + j;
        }
      }
    }
  }
  // accepted correspondence -> 16 sums (FP64), reduced over the block in a fixed order
  double v[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) v[k] = 0.0;
  const bool ok = live && n_tgt > 0 && (fitness_pass || best <= max_d2) && best == best;
  if (ok) {
    const float* f = reinterpret_cast<const float*>(tgt + (size_t)arg * stride_t);
    const double p[3] = {(double)px, (double)py, (double)pz}, q[3] = {(double)f[0], (double)f[1], (double)f[2]};
    for (int a = 0; a < 3; ++a) {
      v[a] = p[a];
      v[3 + a] = q[a];
      for (int b = 0; b < 3; ++b) v[6 + 3 * a + b] = p[a] * q[b];
    }
    v[15] = (double)best;
    v[16] = 1.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 17; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
    for (int k = 0; k < 17; ++k) s_red[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    IcpPartial out;
    double tot[17];
    for (int k = 0; k < 17; ++k) {
      tot[k] = 0.0;
      for (int w = 0; w < ICP_BLOCK / 32; ++w) tot[k] += s_red[w][k];
    }
    for (int a = 0; a < 3; ++a) {
      out.sp[a] = tot[a];
      out.sq[a] = tot[3 + a];
    }
    for (int k = 0; k < 9; ++k) out.spq[k] = tot[6 + k];
    out.sd2 = tot[15];
    out.n = (unsigned long long)(tot[16] + 0.5);
    partials[blockIdx.x] = out;
  }
}

// largest eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations (FP64; converges in a handful of sweeps)
__device__ inline void jacobi4_largest(double A[4][4], double q[4]) {
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0.0;
    for (int a = 0; a < 4; ++a)
      for (int b = a + 1; b < 4; ++b) off += A[a][b] * A[a][b];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (fabs(A[p][r]) < 1e-300) continue;
        const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = A[k][p], akr = A[k][r];
          A[k][p] = c * akp - s * akr;
          A[k][r] = s * akp + c * akr;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = A[p][k], ark = A[r][k];
          A[p][k] = c * apk - s * ark;
          A[r][k] = s * apk + c * ark;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkr = V[k][r];
          V[k][p] = c * vkp - s * vkr;
          V[k][r] = s * vkp + c * vkr;
        }
      }
  }
  int best = 0;
  for (int a = 1; a < 4; ++a)
    if (A[a][a] > A[best][best]) best = a;
  for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

// rigid transform p -> q from the correspondence sums (Horn 1987: quaternion of the rotation = dominant eigenvector of N(H))
__host__ __device__ inline void icp_compose(const double D[16], const double T[16], double out[16]) {  // out = D * T
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += D[4 * r + k] * T[4 * k + c];
      out[4 * r + c] = s;
    }
}

struct IcpCriteria {
  int max_iterations;
  double translation_threshold;  // transformation epsilon (squared translation)
  double rotation_threshold;     // cos of the smallest rotation that counts (PCL: 0.99999)
  double mse_abs, mse_rel;       // euclidean fitness epsilon; 1e-5
};

__global__ void k_icp_solve(const IcpPartial* partials, unsigned n_blocks, IcpState* st, IcpCriteria crit, int fitness_pass) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (st->done && !fitness_pass) return;
  double sp[3] = {0, 0, 0}, sq[3] = {0, 0, 0}, spq[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sd2 = 0;
  unsigned long long n = 0;
  for (unsigned b = 0; b < n_blocks; ++b) {  // fixed order: deterministic
    const IcpPartial& P = partials[b];
    for (int a = 0; a < 3; ++a) {
      sp[a] += P.sp[a];
      sq[a] += P.sq[a];
    }
    for (int k = 0; k < 9; ++k) spq[k] += P.spq[k];
    sd2 += P.sd2;
    n += P.n;
  }
  if (fitness_pass) {
    st->fitness = n ? sd2 / (double)n : 1.79769313486231570e308;
    return;
  }
  st->n_corr = n;
  if (n < 3) {  // PCL: "Not enough correspondences found" -> not converged
    st->done = 1;
    st->converged = 0;
    st->reason = 5;
    return;
  }
  const double inv = 1.0 / (double)n;
  double mp[3], mq[3], H[3][3];
  for (int a = 0; a < 3; ++a) {
    mp[a] = sp[a] * inv;
    mq[a] = sq[a] * inv;
  }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) H[a][b] = spq[3 * a + b] * inv - mp[a] * mq[b];  // cross-covariance of the centred sets
  double N[4][4];
  N[0][0] = H[0][0] + H[1][1] + H[2][2];
  N[0][1] = N[1][0] = H[1][2] - H[2][1];
  N[0][2] = N[2][0] = H[2][0] - H[0][2];
  N[0][3] = N[3][0] = H[0][1] - H[1][0];
  N[1][1] = H[0][0] - H[1][1] - H[2][2];
  N[1][2] = N[2][1] = H[0][1] + H[1][0];
  N[1][3] = N[3][1] = H[2][0] + H[0][2];
  N[2][2] = -H[0][0] + H[1][1] - H[2][2];
  N[2][3] = N[3][2] = H[1][2] + H[2][1];
  N[3][3] = -H[0][0] - H[1][1] + H[2][2];
  double q[4];
  jacobi4_largest(N, q);
  const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] / qn, x = q[1] / qn, y = q[2] / qn, z = q[3] / qn;
  double D[16] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z),     2 * (x * z + w * y),     0,
                  2 * (x * y + w * z),     1 - 2 * (x * x + z * z), 2 * (y * z - w * x),     0,
                  2 * (x * z - w * y),     2 * (y * z + w * x),     1 - 2 * (x * x + y * y), 0,
                  0,                       0,                       0,                       1};
  for (int a = 0; a < 3; ++a) D[4 * a + 3] = mq[a] - (D[4 * a] * mp[0] + D[4 * a + 1] * mp[1] + D[4 * a + 2] * mp[2]);
  double Tn[16];
  icp_compose(D, st->T, Tn);
  for (int k = 0; k < 16; ++k) st->T[k] = Tn[k];
  st->iterations += 1;
  // ---- PCL DefaultConvergenceCriteria ----
  const double mse = sd2 * inv;
  st->mse = mse;
  const double cos_angle = 0.5 * (D[0] + D[5] + D[10] - 1.0);
  const double tr2 = D[3] * D[3] + D[7] * D[7] + D[11] * D[11];
  int reason = 0;
  if (st->iterations >= crit.max_iterations) reason = 1;
  else if (cos_angle >= crit.rotation_threshold && tr2 <= crit.translation_threshold) reason = 2;
  else if (fabs(mse - st->prev_mse) < crit.mse_abs) reason = 3;
  else if (fabs(mse - st->prev_mse) / st->prev_mse < crit.mse_rel) reason = 4;
  st->prev_mse = mse;
  if (reason) {
    st->done = 1;
    st->converged = 1;
    st->reason = reason;
  }
}


// ---- submap assembly in front of the ICP (mapOptmization.cpp:928-949) ------------------------------------------------------
// The reference builds the ICP's two inputs from stored keyframe clouds: every cloud is moved by a 6-DoF key pose with
// LeGO-LOAM's own transformPointCloud (mapOptmization.cpp:598-627: yaw about z, then roll about x, then pitch about y, then the
// translation -- all in float), the clouds are concatenated, the query side drops points with (int)intensity < 0 (932-939) and
// the history side goes through a voxel grid (leaf 0.3 m, 264-268 / 948-949).  k_submap_transform restates the point arithmetic
// operation by operation in FP32 (the library is built with --fmad=false); the six sines and cosines of a pose are taken on the
// HOST with the C library's cosf / sinf -- the functions the reference's cos(float) / sin(float) resolve to -- so the
// transformed coordinates are bit-identical to the reference's.
struct SubmapCloud {
  unsigned long long in_off;  // byte offset of the cloud in the staged input
  unsigned n;                 // points
  unsigned out_off;           // first output point (unfiltered launches)
  float cy, sy, cr, sr, cp, sp, tx, ty, tz;
};

__device__ __forceinline__ float4 submap_point(const SubmapCloud& c, float x, float y, float z, float intensity) {
  const float x1 = __fsub_rn(__fmul_rn(c.cy, x), __fmul_rn(c.sy, y));  // 611-613
  const float y1 = __fadd_rn(__fmul_rn(c.sy, x), __fmul_rn(c.cy, y));
  const float z1 = z;
  const float x2 = x1;                                                   // 615-617
  const float y2 = __fsub_rn(__fmul_rn(c.cr, y1), __fmul_rn(c.sr, z1));
  const float z2 = __fadd_rn(__fmul_rn(c.sr, y1), __fmul_rn(c.cr, z1));
  float4 o;                                                              // 619-622
  o.x = __fadd_rn(__fadd_rn(__fmul_rn(c.cp, x2), __fmul_rn(c.sp, z2)), c.tx);
  o.y = __fadd_rn(y2, c.ty);
  o.z = __fadd_rn(__fadd_rn(__fmul_rn(-c.sp, x2), __fmul_rn(c.cp, z2)), c.tz);
  o.w = intensity;
  return o;
}

// grid (chunks, clouds): every point of every cloud, order kept (cloud after cloud: the reference's operator+=)
__global__ void __launch_bounds__(256) k_submap_transform(const unsigned char* in, unsigned stride, unsigned intensity_off, const SubmapCloud* clouds,
                                                          float4* out) {
  const SubmapCloud c = clouds[blockIdx.y];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const float* f = reinterpret_cast<const float*>(in + c.in_off + (unsigned long long)i * stride);
    const float w = intensity_off ? *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(f) + intensity_off) : 0.f;
    out[c.out_off + i] = submap_point(c, f[0], f[1], f[2], w);
  }
}

// one block: the clouds one after the other, points with (int)intensity < 0 dropped, order kept (mapOptmization.cpp:932-939)
__global__ void __launch_bounds__(1024) k_submap_transform_filtered(const unsigned char* in, unsigned stride, unsigned intensity_off,
                                                                    const SubmapCloud* clouds, unsigned n_clouds, float4* out, unsigned* n_out) {
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (unsigned ci = 0; ci < n_clouds; ++ci) {
    const SubmapCloud c = clouds[ci];
    for (unsigned i0 = 0; i0 < c.n; i0 += blockDim.x) {
      const unsigned i = i0 + threadIdx.x;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      bool keep = false;
      if (i < c.n) {
        const float* f = reinterpret_cast<const float*>(in + c.in_off + (unsigned long long)i * stride);
        const float w = intensity_off ? *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(f) + intensity_off) : 0.f;
        o = submap_point(c, f[0], f[1], f[2], w);
        // (int)w >= 0  <=>  -1 < w < 2^31: the cast truncates toward zero; out of range (and NaN) it yields INT_MIN on x86-64
        keep = !intensity_off || (w > -1.f && w < 2147483648.f);
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      unsigned before = s_base;
      for (int w2 = 0; w2 < warp; ++w2) before += s_warp[w2];
      if (keep) out[before + __popc(m & ((1u << lane) - 1u))] = o;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned t = s_base;
        for (unsigned w2 = 0; w2 < (blockDim.x >> 5); ++w2) t += s_warp[w2];
        s_base = t;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) *n_out = s_base;
}

}  // namespace scgpu
