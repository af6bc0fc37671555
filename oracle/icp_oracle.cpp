// oracle/icp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
// CPU restatement of pcl::IterativeClosestPoint as mapOptmization.cpp:1053-1078 configures it (point-to-point, SVD/Horn
// transform estimation, DefaultConvergenceCriteria, getFitnessScore).  PARITY UNPINNED: PCL is a third-party dependency that
// is absent from /root/reference and from this image; this follows PCL 1.8's published sources
// (registration/impl/icp.hpp:115-225, default_convergence_criteria.hpp:46-120, transformation_estimation_svd.hpp).
// The rotation is obtained with Horn's closed form (dominant eigenvector of the 4x4 matrix N of the cross-covariance; power
// iteration on a shifted matrix here, an independent method from the product's Jacobi sweeps) -- equal to PCL's SVD (Umeyama)
// solution whenever that one is a proper rotation.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

namespace {
void compose(const double D[16], const double T[16], double out[16]) {
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += D[4 * r + k] * T[4 * k + c];
      out[4 * r + c] = s;
    }
}
}  // namespace

extern "C" int icp_oracle(const float* src, size_t n_src, const float* tgt, size_t n_tgt, size_t stride_floats, int max_iter, double max_corr,
                          double trans_eps, double fit_eps, const double* T0, double* T_out, double* fitness, int* converged, int* iterations) {
  double T[16];
  std::memcpy(T, T0, sizeof T);
  double prev_mse = DBL_MAX;
  int it = 0, conv = 0;
  const float max_d2 = (float)(max_corr * max_corr);
  std::vector<float> p(3 * n_src);
  auto transform = [&]() {
    for (size_t i = 0; i < n_src; ++i) {
      const double x = src[i * stride_floats], y = src[i * stride_floats + 1], z = src[i * stride_floats + 2];
      p[3 * i] = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
      p[3 * i + 1] = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
      p[3 * i + 2] = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
    }
  };
  auto nearest = [&](size_t i, size_t* arg) {
    float best = 3.4e38f;
    size_t a = 0;
    for (size_t j = 0; j < n_tgt; ++j) {
      const float dx = p[3 * i] - tgt[j * stride_floats], dy = p[3 * i + 1] - tgt[j * stride_floats + 1], dz = p[3 * i + 2] - tgt[j * stride_floats + 2];
      const float d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best) {
        best = d2;
        a = j;
      }
    }
    *arg = a;
    return best;
  };
  while (!conv && it < max_iter) {
    transform();
    double sp[3] = {0, 0, 0}, sq[3] = {0, 0, 0}, spq[9] = {0}, sd2 = 0;
    size_t n = 0;
    for (size_t i = 0; i < n_src; ++i) {
      size_t a;
      const float d2 = nearest(i, &a);
      if (!(d2 <= max_d2)) continue;
      const double pp[3] = {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
      const double qq[3] = {tgt[a * stride_floats], tgt[a * stride_floats + 1], tgt[a * stride_floats + 2]};
      for (int u = 0; u < 3; ++u) {
        sp[u] += pp[u];
        sq[u] += qq[u];
        for (int v = 0; v < 3; ++v) spq[3 * u + v] += pp[u] * qq[v];
      }
      sd2 += d2;
      ++n;
    }
    if (n < 3) break;
    const double inv = 1.0 / (double)n;
    double mp[3], mq[3], H[3][3];
    for (int u = 0; u < 3; ++u) {
      mp[u] = sp[u] * inv;
      mq[u] = sq[u] * inv;
    }
    for (int u = 0; u < 3; ++u)
      for (int v = 0; v < 3; ++v) H[u][v] = spq[3 * u + v] * inv - mp[u] * mq[v];
    double N[4][4] = {{H[0][0] + H[1][1] + H[2][2], H[1][2] - H[2][1], H[2][0] - H[0][2], H[0][1] - H[1][0]},
                      {H[1][2] - H[2][1], H[0][0] - H[1][1] - H[2][2], H[0][1] + H[1][0], H[2][0] + H[0][2]},
                      {H[2][0] - H[0][2], H[0][1] + H[1][0], -H[0][0] + H[1][1] - H[2][2], H[1][2] + H[2][1]},
                      {H[0][1] - H[1][0], H[2][0] + H[0][2], H[1][2] + H[2][1], -H[0][0] - H[1][1] + H[2][2]}};
    // dominant eigenvector: power iteration on N + shift*I (shift = Frobenius norm makes the matrix positive semi-definite)
    double fro = 0;
    for (auto& row : N)
      for (double v : row) fro += v * v;
    const double shift = std::sqrt(fro) + 1e-30;
    double q[4] = {1, 0.01, 0.02, 0.03};
    for (int k = 0; k < 20000; ++k) {
      double r[4];
      for (int u = 0; u < 4; ++u) {
        r[u] = shift * q[u];
        for (int v = 0; v < 4; ++v) r[u] += N[u][v] * q[v];
      }
      const double nr = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
      double diff = 0;
      for (int u = 0; u < 4; ++u) {
        r[u] /= nr;
        diff += std::fabs(r[u] - q[u]);
        q[u] = r[u];
      }
      if (diff < 1e-15) break;
    }
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    double D[16] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z),     2 * (x * z + w * y),     0,
                    2 * (x * y + w * z),     1 - 2 * (x * x + z * z), 2 * (y * z - w * x),     0,
                    2 * (x * z - w * y),     2 * (y * z + w * x),     1 - 2 * (x * x + y * y), 0,
                    0,                       0,                       0,                       1};
    for (int u = 0; u < 3; ++u) D[4 * u + 3] = mq[u] - (D[4 * u] * mp[0] + D[4 * u + 1] * mp[1] + D[4 * u + 2] * mp[2]);
    double Tn[16];
    compose(D, T, Tn);
    std::memcpy(T, Tn, sizeof T);
    ++it;
    const double mse = sd2 * inv;
    const double cos_angle = 0.5 * (D[0] + D[5] + D[10] - 1.0), tr2 = D[3] * D[3] + D[7] * D[7] + D[11] * D[11];
    if (it >= max_iter) conv = 1;
    else if (cos_angle >= 0.99999 && tr2 <= trans_eps) conv = 1;
    else if (std::fabs(mse - prev_mse) < fit_eps) conv = 1;
    else if (std::fabs(mse - prev_mse) / prev_mse < 0.00001) conv = 1;
    prev_mse = mse;
  }
  transform();
  double sd2 = 0;
  for (size_t i = 0; i < n_src; ++i) {
    size_t a;
    sd2 += nearest(i, &a);
  }
  std::memcpy(T_out, T, sizeof T);
  *fitness = n_src && n_tgt ? sd2 / (double)n_src : DBL_MAX;
  *converged = conv;
  *iterations = it;
  return 0;
}
