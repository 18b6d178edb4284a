// TEST INFRASTRUCTURE — stand-in for TooN/SymEigen.h: cyclic Jacobi eigen-decomposition of a small
// symmetric matrix; eigenvalues ascending, eigenvectors as ROWS of get_evectors() (TooN's layout).
// Not on either hot path (MapMaker::CalcPlaneAligner, initialisation only).
#pragma once
#include <algorithm>
#include <numeric>
#include "TooN.h"
namespace TooN {
inline void jacobi_eigen(std::vector<double>& a, int n, std::vector<double>& vals, std::vector<double>& vecs_rows) {
  std::vector<double> v((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) v[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq; a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk; a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q];
          v[(size_t)k * n + p] = c * vkp - s * vkq; v[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int i, int j) { return a[(size_t)i * n + i] < a[(size_t)j * n + j]; });
  vals.resize(n); vecs_rows.assign((size_t)n * n, 0.0);
  for (int r = 0; r < n; r++) {
    vals[r] = a[(size_t)order[r] * n + order[r]];
    for (int k = 0; k < n; k++) vecs_rows[(size_t)r * n + k] = v[(size_t)k * n + order[r]];
  }
}
template <int N = Dynamic, class P = double> class SymEigen {
 public:
  template <class M, TOON_IF(is_mat<M>::value)> SymEigen(const M& m) : evecs(MakeMat<N, N>::make(m.num_rows(), m.num_rows())), evals(m.num_rows()) {
    const int n = m.num_rows();
    std::vector<double> a((size_t)n * n), vals, vecs;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[(size_t)i * n + j] = 0.5 * (m(i, j) + m(j, i));
    jacobi_eigen(a, n, vals, vecs);
    for (int i = 0; i < n; i++) { evals[i] = vals[i]; for (int j = 0; j < n; j++) evecs(i, j) = vecs[(size_t)i * n + j]; }
  }
  Matrix<N, N>& get_evectors() { return evecs; }
  Vector<N>& get_evalues() { return evals; }
 private:
  Matrix<N, N> evecs;
  Vector<N> evals;
};
}  // namespace TooN
