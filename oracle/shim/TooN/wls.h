// TEST INFRASTRUCTURE — stand-in for TooN/wls.h: weighted least squares accumulator.
// add_mJ(m, J, w): C_inv += (J w) J^T, vector += m (J w); compute(): LDL^T solve.
#pragma once
#include "Cholesky.h"
namespace TooN {
template <int N = Dynamic, class P = double> class WLS {
 public:
  WLS() { clear(); }
  void clear() { C = Zeros; v = Zeros; }
  void add_prior(double val) { for (int i = 0; i < N; i++) C(i, i) += val; }
  template <class V, TOON_IF(is_vec<V>::value)> void add_mJ(double m, const V& J, double weight = 1) {
    Vector<N> Jw = J * weight;
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) C(i, j) += Jw[i] * J[j];
    v += Jw * m;
  }
  void compute() { mu = Cholesky<N>(C).backsub(v); }
  Vector<N>& get_mu() { return mu; }
  Matrix<N, N>& get_C_inv() { return C; }
  Vector<N>& get_vector() { return v; }
 private:
  Matrix<N, N> C;
  Vector<N> v, mu;
};
}  // namespace TooN
