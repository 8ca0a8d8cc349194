// Host-side quadrature helpers shared by the dealiasing (nsb_conv.cu) and pressure-mesh (nsb_ns.cu) set-up:
// Gauss-Legendre nodes and weights, Lagrange interpolation and derivative matrices on arbitrary nodes.
#pragma once
#include <cmath>
#include <vector>

namespace nsb {

// ---- host: Gauss-Legendre nodes, interpolation and derivative matrices --------------------------
inline void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5));
    double pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
  if (n % 2) x[n / 2] = 0.0;
}

inline std::vector<double> bary_weights(const std::vector<double> &z) {
  const int n = (int)z.size();
  std::vector<double> w(n, 1.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j) w[i] /= (z[i] - z[j]);
  return w;
}

// J[I * nf + i] = l_i(zto_I)
inline std::vector<double> interp_matrix(const std::vector<double> &zfrom, const std::vector<double> &zto) {
  const int nf = (int)zfrom.size(), nt = (int)zto.size();
  const std::vector<double> bw = bary_weights(zfrom);
  std::vector<double> J((size_t)nt * nf, 0.0);
  for (int I = 0; I < nt; ++I) {
    int hit = -1;
    for (int i = 0; i < nf; ++i)
      if (std::fabs(zto[I] - zfrom[i]) < 1e-15) hit = i;
    if (hit >= 0) {
      J[(size_t)I * nf + hit] = 1.0;
      continue;
    }
    double s = 0.0;
    for (int i = 0; i < nf; ++i) {
      J[(size_t)I * nf + i] = bw[i] / (zto[I] - zfrom[i]);
      s += J[(size_t)I * nf + i];
    }
    for (int i = 0; i < nf; ++i) J[(size_t)I * nf + i] /= s;
  }
  return J;
}

// D[i * n + j] = l_j'(z_i)
inline std::vector<double> deriv_matrix(const std::vector<double> &z) {
  const int n = (int)z.size();
  const std::vector<double> bw = bary_weights(z);
  std::vector<double> D((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      D[(size_t)i * n + j] = (bw[j] / bw[i]) / (z[i] - z[j]);
      s += D[(size_t)i * n + j];
    }
    D[(size_t)i * n + i] = -s;
  }
  return D;
}

}  // namespace nsb
