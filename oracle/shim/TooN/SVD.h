// TEST INFRASTRUCTURE — stand-in for TooN/SVD.h (LAPACK dgesvd in the real library): singular values
// descending, right singular vectors as rows of get_VT(), computed from the Jacobi eigen-decomposition
// of A^T A (adequate for the 4-column systems of MapMaker::Triangulate; the null vector it is used
// for is defined up to sign, which project() removes).
#pragma once
#include "SymEigen.h"
namespace TooN {
template <int R = Dynamic, int C = R, class P = double> class SVD {
 public:
  template <class M, TOON_IF(is_mat<M>::value)> SVD(const M& m) : U(MakeMat<R, C>::make(m.num_rows(), m.num_cols())), VT(MakeMat<C, C>::make(m.num_cols(), m.num_cols())), d(m.num_cols()) {
    const int r = m.num_rows(), c = m.num_cols();
    std::vector<double> a((size_t)c * c, 0.0), vals, vecs;
    for (int i = 0; i < c; i++) for (int j = 0; j < c; j++) { double s = 0; for (int k = 0; k < r; k++) s += m(k, i) * m(k, j); a[(size_t)i * c + j] = s; }
    jacobi_eigen(a, c, vals, vecs);
    for (int i = 0; i < c; i++) {  // ascending eigenvalues -> descending singular values
      const int src = c - 1 - i;
      d[i] = std::sqrt(vals[src] > 0 ? vals[src] : 0.0);
      for (int j = 0; j < c; j++) VT(i, j) = vecs[(size_t)src * c + j];
      for (int k = 0; k < r; k++) { double s = 0; for (int j = 0; j < c; j++) s += m(k, j) * VT(i, j); U(k, i) = d[i] > 0 ? s / d[i] : 0.0; }
    }
  }
  Matrix<R, C>& get_U() { return U; }
  Matrix<C, C>& get_VT() { return VT; }
  Vector<C>& get_diagonal() { return d; }
 private:
  Matrix<R, C> U;
  Matrix<C, C> VT;
  Vector<C> d;
};
}  // namespace TooN
