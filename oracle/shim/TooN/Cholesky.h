// TEST INFRASTRUCTURE — stand-in for TooN/Cholesky.h: square-root-free LDL^T, no pivoting, lower
// triangle read, back-substitution L y = b, D z = y, L^T x = z — the same loops as
// orc::ldlt_factor / ldlt_backsub (oracle_math.h), which restate TooN's.
#pragma once
#include "TooN.h"
#include "../../oracle_math.h"
namespace TooN {
template <int N = Dynamic, class P = double> class Cholesky {
 public:
  template <class M, TOON_IF(is_mat<M>::value)> Cholesky(const M& m) : L(m) { orc::ldlt_factor(L.get_data_ptr(), L.num_rows(), L.num_cols()); }
  template <class V, TOON_IF(is_vec<V>::value)> Vector<N> backsub(const V& v) const {
    const int n = L.num_rows();
    Vector<N> b(v), x(n);
    orc::ldlt_backsub(const_cast<Matrix<N, N>&>(L).get_data_ptr(), n, n, b.get_data_ptr(), x.get_data_ptr());
    return x;
  }
  Matrix<N, N> get_inverse() const {
    const int n = L.num_rows();
    Matrix<N, N> inv = MakeMat<N, N>::make(n, n);
    std::vector<double> e(n), x(n);
    for (int c = 0; c < n; c++) {
      for (int i = 0; i < n; i++) e[i] = i == c ? 1.0 : 0.0;
      orc::ldlt_backsub(const_cast<Matrix<N, N>&>(L).get_data_ptr(), n, n, e.data(), x.data());
      for (int i = 0; i < n; i++) inv(i, c) = x[i];
    }
    return inv;
  }
  double determinant() const { double d = 1; for (int i = 0; i < L.num_rows(); i++) d *= L(i, i); return d; }
 private:
  Matrix<N, N> L;
};
}  // namespace TooN
