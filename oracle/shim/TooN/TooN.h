// TEST INFRASTRUCTURE — minimal stand-in for the TooN 2.2 headers, written from TooN's documented
// interface so that the reference's OWN sources (src/Bundle.cc, src/ATANCamera.cc, src/PatchFinder.cc,
// src/KeyFrame.cc, src/ImageProcess.cc, compiled where they lie under /root/reference by
// oracle/Makefile.ref) build without the real library, which is absent from this image.
// Everything here is eager fixed-size arithmetic with the plain left-to-right summation order of
// TooN's dot products; it is not a copy of TooN and implements only what those files use.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <type_traits>
#include <vector>

namespace TooN {

static const int Dynamic = -1;
struct ZerosT {};
static const ZerosT Zeros = ZerosT();
struct IdentityT { double s; IdentityT() : s(1) {} explicit IdentityT(double s_) : s(s_) {} };
static const IdentityT Identity = IdentityT();
inline IdentityT operator*(double k, const IdentityT& i) { return IdentityT(k * i.s); }
inline IdentityT operator*(const IdentityT& i, double k) { return IdentityT(i.s * k); }

struct vec_tag {};
struct mat_tag {};
template <class T> using is_vec = std::is_base_of<vec_tag, typename std::decay<T>::type>;
template <class T> using is_mat = std::is_base_of<mat_tag, typename std::decay<T>::type>;
#define TOON_IF(...) typename std::enable_if<(__VA_ARGS__), int>::type = 0

template <int N = Dynamic, class P = double> struct Vector;
template <int R = Dynamic, int C = R, class P = double> struct Matrix;

// storage helpers ---------------------------------------------------------------------------
template <int N> struct Store {
  double d[N > 0 ? N : 1];
  Store() {}
  explicit Store(int) {}
  int n() const { return N; }
  double* p() { return d; }
  const double* p() const { return d; }
};
template <> struct Store<Dynamic> {
  std::vector<double> d;
  Store() {}
  explicit Store(int n_) : d(n_) {}
  int n() const { return (int)d.size(); }
  double* p() { return d.data(); }
  const double* p() const { return d.data(); }
  void resize(int n_) { d.resize(n_); }
};

// strided vector view -------------------------------------------------------------------------
template <int N> struct VecView : vec_tag {
  static const int Size = N;
  double* ptr;
  int len, stride;
  VecView(double* p_, int len_, int stride_) : ptr(p_), len(len_), stride(stride_) {}
  int size() const { return len; }
  double& operator[](int i) const { return ptr[i * stride]; }
  template <class V, TOON_IF(is_vec<V>::value)> const VecView& operator=(const V& v) const {
    assert(v.size() == len);
    double tmp[64]; std::vector<double> big; double* t = tmp;
    if (len > 64) { big.resize(len); t = big.data(); }
    for (int i = 0; i < len; i++) t[i] = v[i];  // the source may alias the destination
    for (int i = 0; i < len; i++) ptr[i * stride] = t[i];
    return *this;
  }
  const VecView& operator=(const VecView& v) const { return this->template operator=<VecView>(v); }
  const VecView& operator=(const ZerosT&) const { for (int i = 0; i < len; i++) ptr[i * stride] = 0; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> const VecView& operator+=(const V& v) const { for (int i = 0; i < len; i++) ptr[i * stride] += v[i]; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> const VecView& operator-=(const V& v) const { for (int i = 0; i < len; i++) ptr[i * stride] -= v[i]; return *this; }
  const VecView& operator*=(double s) const { for (int i = 0; i < len; i++) ptr[i * stride] *= s; return *this; }
  const VecView& operator/=(double s) const { for (int i = 0; i < len; i++) ptr[i * stride] /= s; return *this; }
  template <int S, int L> VecView<L> slice() const { return VecView<L>(ptr + S * stride, L, stride); }
  VecView<Dynamic> slice(int s, int l) const { return VecView<Dynamic>(ptr + s * stride, l, stride); }
};

template <int R, int C> struct MatView;
template <int N, class P> struct Vector : vec_tag {
  static const int Size = N;
  MatView<N, 1> as_col() const;
  MatView<1, N> as_row() const;
  Store<N> s;
  Vector() {}
  explicit Vector(int n_) : s(n_) {}
  Vector(const ZerosT&) { for (int i = 0; i < size(); i++) s.p()[i] = 0; }
  template <class V, TOON_IF(is_vec<V>::value)> Vector(const V& v) : s(v.size()) { assert(N == Dynamic || v.size() == N); for (int i = 0; i < size(); i++) s.p()[i] = v[i]; }
  int size() const { return s.n(); }
  double& operator[](int i) { return s.p()[i]; }
  const double& operator[](int i) const { return s.p()[i]; }
  double* get_data_ptr() { return s.p(); }
  const double* get_data_ptr() const { return s.p(); }
  template <class V, TOON_IF(is_vec<V>::value)> Vector& operator=(const V& v) {
    if (N == Dynamic && size() != v.size()) resize_(v.size());
    assert(v.size() == size());
    Vector<N> t(size());
    for (int i = 0; i < size(); i++) t.s.p()[i] = v[i];
    for (int i = 0; i < size(); i++) s.p()[i] = t.s.p()[i];
    return *this;
  }
  Vector& operator=(const ZerosT&) { for (int i = 0; i < size(); i++) s.p()[i] = 0; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> Vector& operator+=(const V& v) { for (int i = 0; i < size(); i++) s.p()[i] += v[i]; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> Vector& operator-=(const V& v) { for (int i = 0; i < size(); i++) s.p()[i] -= v[i]; return *this; }
  Vector& operator*=(double k) { for (int i = 0; i < size(); i++) s.p()[i] *= k; return *this; }
  Vector& operator/=(double k) { for (int i = 0; i < size(); i++) s.p()[i] /= k; return *this; }
  template <int S, int L> VecView<L> slice() { return VecView<L>(s.p() + S, L, 1); }
  template <int S, int L> VecView<L> slice() const { return VecView<L>(const_cast<double*>(s.p()) + S, L, 1); }
  VecView<Dynamic> slice(int st, int l) { return VecView<Dynamic>(s.p() + st, l, 1); }
  VecView<Dynamic> slice(int st, int l) const { return VecView<Dynamic>(const_cast<double*>(s.p()) + st, l, 1); }
  VecView<N> as_view() const { return VecView<N>(const_cast<double*>(s.p()), size(), 1); }
 private:
  template <int M = N> typename std::enable_if<M == Dynamic>::type resize_(int n_) { s.resize(n_); }
  template <int M = N> typename std::enable_if<M != Dynamic>::type resize_(int) {}
};

// strided matrix view -------------------------------------------------------------------------
template <int R, int C> struct MatView : mat_tag {
  static const int Rows = R, Cols = C;
  double* ptr;
  int nr, nc, rs, cs;
  MatView(double* p_, int nr_, int nc_, int rs_, int cs_) : ptr(p_), nr(nr_), nc(nc_), rs(rs_), cs(cs_) {}
  int num_rows() const { return nr; }
  int num_cols() const { return nc; }
  double& operator()(int r, int c) const { return ptr[r * rs + c * cs]; }
  VecView<C> operator[](int r) const { return VecView<C>(ptr + r * rs, nc, cs); }
  MatView<C, R> T() const { return MatView<C, R>(ptr, nc, nr, cs, rs); }
  template <class M, TOON_IF(is_mat<M>::value)> const MatView& operator=(const M& m) const {
    assert(m.num_rows() == nr && m.num_cols() == nc);
    std::vector<double> t((size_t)nr * nc);
    for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) t[(size_t)r * nc + c] = m(r, c);
    for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) = t[(size_t)r * nc + c];
    return *this;
  }
  const MatView& operator=(const MatView& m) const { return this->template operator=<MatView>(m); }
  const MatView& operator=(const ZerosT&) const { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) = 0; return *this; }
  template <class M, TOON_IF(is_mat<M>::value)> const MatView& operator+=(const M& m) const { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) += m(r, c); return *this; }
  template <class M, TOON_IF(is_mat<M>::value)> const MatView& operator-=(const M& m) const { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) -= m(r, c); return *this; }
  template <int R0, int C0, int NR, int NC> MatView<NR, NC> slice() const { return MatView<NR, NC>(ptr + R0 * rs + C0 * cs, NR, NC, rs, cs); }
  MatView<Dynamic, Dynamic> slice(int r0, int c0, int nr_, int nc_) const { return MatView<Dynamic, Dynamic>(ptr + r0 * rs + c0 * cs, nr_, nc_, rs, cs); }
};

template <int R, int C, class P> struct Matrix : mat_tag {
  static const int Rows = R, Cols = C;
  Store<(R == Dynamic || C == Dynamic) ? Dynamic : R * C> s;
  int nr, nc;
  Matrix() : nr(R == Dynamic ? 0 : R), nc(C == Dynamic ? 0 : C) {}
  Matrix(int r, int c) : s(r * c), nr(r), nc(c) {}
  Matrix(const ZerosT&) : nr(R), nc(C) { static_assert(R != Dynamic && C != Dynamic, "sized"); for (int i = 0; i < nr * nc; i++) s.p()[i] = 0; }
  Matrix(const IdentityT& id) : nr(R), nc(C) { static_assert(R != Dynamic && C != Dynamic, "sized"); for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) s.p()[r * nc + c] = r == c ? id.s : 0.0; }
  template <class M, TOON_IF(is_mat<M>::value)> Matrix(const M& m) : s(m.num_rows() * m.num_cols()), nr(m.num_rows()), nc(m.num_cols()) {
    for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) s.p()[r * nc + c] = m(r, c);
  }
  int num_rows() const { return nr; }
  int num_cols() const { return nc; }
  double& operator()(int r, int c) { return s.p()[r * nc + c]; }
  const double& operator()(int r, int c) const { return s.p()[r * nc + c]; }
  VecView<C> operator[](int r) { return VecView<C>(s.p() + r * nc, nc, 1); }
  VecView<C> operator[](int r) const { return VecView<C>(const_cast<double*>(s.p()) + r * nc, nc, 1); }
  MatView<C, R> T() { return MatView<C, R>(s.p(), nc, nr, 1, nc); }
  MatView<C, R> T() const { return MatView<C, R>(const_cast<double*>(s.p()), nc, nr, 1, nc); }
  MatView<R, C> as_view() const { return MatView<R, C>(const_cast<double*>(s.p()), nr, nc, nc, 1); }
  double* get_data_ptr() { return s.p(); }
  template <class M, TOON_IF(is_mat<M>::value)> Matrix& operator=(const M& m) {
    Matrix t(m);
    assert((nr == t.nr && nc == t.nc) || R == Dynamic || C == Dynamic);
    s = t.s; nr = t.nr; nc = t.nc;
    return *this;
  }
  Matrix& operator=(const ZerosT&) { for (int i = 0; i < nr * nc; i++) s.p()[i] = 0; return *this; }
  Matrix& operator=(const IdentityT& id) { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) s.p()[r * nc + c] = r == c ? id.s : 0.0; return *this; }
  template <class M, TOON_IF(is_mat<M>::value)> Matrix& operator+=(const M& m) { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) += m(r, c); return *this; }
  template <class M, TOON_IF(is_mat<M>::value)> Matrix& operator-=(const M& m) { for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) (*this)(r, c) -= m(r, c); return *this; }
  Matrix& operator*=(double k) { for (int i = 0; i < nr * nc; i++) s.p()[i] *= k; return *this; }
  Matrix& operator/=(double k) { for (int i = 0; i < nr * nc; i++) s.p()[i] /= k; return *this; }
  template <int R0, int C0, int NR, int NC> MatView<NR, NC> slice() { return MatView<NR, NC>(s.p() + R0 * nc + C0, NR, NC, nc, 1); }
  template <int R0, int C0, int NR, int NC> MatView<NR, NC> slice() const { return MatView<NR, NC>(const_cast<double*>(s.p()) + R0 * nc + C0, NR, NC, nc, 1); }
  MatView<Dynamic, Dynamic> slice(int r0, int c0, int nr_, int nc_) { return MatView<Dynamic, Dynamic>(s.p() + r0 * nc + c0, nr_, nc_, nc, 1); }
  MatView<Dynamic, Dynamic> slice(int r0, int c0, int nr_, int nc_) const { return MatView<Dynamic, Dynamic>(const_cast<double*>(s.p()) + r0 * nc + c0, nr_, nc_, nc, 1); }
};

template <int N, class P> inline MatView<N, 1> Vector<N, P>::as_col() const { return MatView<N, 1>(const_cast<double*>(s.p()), size(), 1, 1, 1); }
template <int N, class P> inline MatView<1, N> Vector<N, P>::as_row() const { return MatView<1, N>(const_cast<double*>(s.p()), 1, size(), size(), 1); }

// helpers to build result objects --------------------------------------------------------------
template <int N> inline Vector<N> make_vec(int n) { return Vector<N>(n); }
template <> inline Vector<Dynamic> make_vec<Dynamic>(int n) { return Vector<Dynamic>(n); }
template <int R, int C> struct MakeMat { static Matrix<R, C> make(int, int) { return Matrix<R, C>(); } };
template <int C> struct MakeMat<Dynamic, C> { static Matrix<Dynamic, C> make(int r, int c) { return Matrix<Dynamic, C>(r, c); } };
template <int R> struct MakeMat<R, Dynamic> { static Matrix<R, Dynamic> make(int r, int c) { return Matrix<R, Dynamic>(r, c); } };
template <> struct MakeMat<Dynamic, Dynamic> { static Matrix<Dynamic, Dynamic> make(int r, int c) { return Matrix<Dynamic, Dynamic>(r, c); } };
template <int A, int B> struct SizeOf2 { static const int value = A != Dynamic ? A : B; };

// arithmetic -----------------------------------------------------------------------------------
template <class A, class B, TOON_IF(is_vec<A>::value && is_vec<B>::value)> inline double operator*(const A& a, const B& b) {
  assert(a.size() == b.size());
  double sum = 0;
  for (int i = 0; i < a.size(); i++) sum += a[i] * b[i];
  return sum;
}
template <class A, class B, TOON_IF(is_vec<A>::value && is_vec<B>::value)> inline Vector<SizeOf2<A::Size, B::Size>::value> operator+(const A& a, const B& b) {
  Vector<SizeOf2<A::Size, B::Size>::value> r(a.size());
  for (int i = 0; i < a.size(); i++) r[i] = a[i] + b[i];
  return r;
}
template <class A, class B, TOON_IF(is_vec<A>::value && is_vec<B>::value)> inline Vector<SizeOf2<A::Size, B::Size>::value> operator-(const A& a, const B& b) {
  Vector<SizeOf2<A::Size, B::Size>::value> r(a.size());
  for (int i = 0; i < a.size(); i++) r[i] = a[i] - b[i];
  return r;
}
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<A::Size> operator-(const A& a) { Vector<A::Size> r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = -a[i]; return r; }
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<A::Size> operator*(const A& a, double k) { Vector<A::Size> r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] * k; return r; }
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<A::Size> operator*(double k, const A& a) { Vector<A::Size> r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = k * a[i]; return r; }
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<A::Size> operator/(const A& a, double k) { Vector<A::Size> r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] / k; return r; }

template <class A, class B, TOON_IF(is_mat<A>::value && is_mat<B>::value)> inline Matrix<A::Rows, B::Cols> operator*(const A& a, const B& b) {
  assert(a.num_cols() == b.num_rows());
  Matrix<A::Rows, B::Cols> r = MakeMat<A::Rows, B::Cols>::make(a.num_rows(), b.num_cols());
  for (int i = 0; i < a.num_rows(); i++)
    for (int j = 0; j < b.num_cols(); j++) {
      double sum = 0;
      for (int k = 0; k < a.num_cols(); k++) sum += a(i, k) * b(k, j);
      r(i, j) = sum;
    }
  return r;
}
template <class A, class B, TOON_IF(is_mat<A>::value && is_vec<B>::value)> inline Vector<A::Rows> operator*(const A& a, const B& b) {
  assert(a.num_cols() == b.size());
  Vector<A::Rows> r(a.num_rows());
  for (int i = 0; i < a.num_rows(); i++) {
    double sum = 0;
    for (int k = 0; k < a.num_cols(); k++) sum += a(i, k) * b[k];
    r[i] = sum;
  }
  return r;
}
template <class A, class B, TOON_IF(is_vec<A>::value && is_mat<B>::value)> inline Vector<B::Cols> operator*(const A& a, const B& b) {
  assert(a.size() == b.num_rows());
  Vector<B::Cols> r(b.num_cols());
  for (int j = 0; j < b.num_cols(); j++) {
    double sum = 0;
    for (int k = 0; k < b.num_rows(); k++) sum += a[k] * b(k, j);
    r[j] = sum;
  }
  return r;
}
template <class A, class B, TOON_IF(is_mat<A>::value && is_mat<B>::value)> inline Matrix<SizeOf2<A::Rows, B::Rows>::value, SizeOf2<A::Cols, B::Cols>::value> operator+(const A& a, const B& b) {
  auto r = MakeMat<SizeOf2<A::Rows, B::Rows>::value, SizeOf2<A::Cols, B::Cols>::value>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = a(i, j) + b(i, j);
  return r;
}
template <class A, class B, TOON_IF(is_mat<A>::value && is_mat<B>::value)> inline Matrix<SizeOf2<A::Rows, B::Rows>::value, SizeOf2<A::Cols, B::Cols>::value> operator-(const A& a, const B& b) {
  auto r = MakeMat<SizeOf2<A::Rows, B::Rows>::value, SizeOf2<A::Cols, B::Cols>::value>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = a(i, j) - b(i, j);
  return r;
}
template <class A, TOON_IF(is_mat<A>::value)> inline Matrix<A::Rows, A::Cols> operator*(const A& a, double k) {
  auto r = MakeMat<A::Rows, A::Cols>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = a(i, j) * k;
  return r;
}
template <class A, TOON_IF(is_mat<A>::value)> inline Matrix<A::Rows, A::Cols> operator*(double k, const A& a) {
  auto r = MakeMat<A::Rows, A::Cols>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = k * a(i, j);
  return r;
}
template <class A, TOON_IF(is_mat<A>::value)> inline Matrix<A::Rows, A::Cols> operator/(const A& a, double k) {
  auto r = MakeMat<A::Rows, A::Cols>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = a(i, j) / k;
  return r;
}
template <class A, TOON_IF(is_mat<A>::value)> inline Matrix<A::Rows, A::Cols> operator-(const A& a) {
  auto r = MakeMat<A::Rows, A::Cols>::make(a.num_rows(), a.num_cols());
  for (int i = 0; i < a.num_rows(); i++) for (int j = 0; j < a.num_cols(); j++) r(i, j) = -a(i, j);
  return r;
}
// cross product: TooN spells it v1 ^ v2
template <class A, class B, TOON_IF(is_vec<A>::value && is_vec<B>::value)> inline Vector<3> operator^(const A& a, const B& b) {
  Vector<3> r;
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}

// numbers are formatted with snprintf: inside a dlopen()ed test library the stream's locale facets
// are not always initialised, and these operators only serve the reference's diagnostics
inline void put_number(std::ostream& os, double x) { char b[40]; std::snprintf(b, sizeof b, "%g ", x); os << b; }
template <class A, TOON_IF(is_vec<A>::value)> inline std::ostream& operator<<(std::ostream& os, const A& v) { for (int i = 0; i < v.size(); i++) put_number(os, v[i]); return os; }
template <class A, TOON_IF(is_mat<A>::value)> inline std::ostream& operator<<(std::ostream& os, const A& m) {
  for (int i = 0; i < m.num_rows(); i++) { for (int j = 0; j < m.num_cols(); j++) put_number(os, m(i, j)); os << "\n"; }
  return os;
}
template <int N> inline std::istream& operator>>(std::istream& is, Vector<N>& v) { for (int i = 0; i < v.size(); i++) is >> v[i]; return is; }

// makeVector(a, b, ...)
template <class... T> inline Vector<(int)sizeof...(T)> makeVector(T... a) {
  Vector<(int)sizeof...(T)> r;
  const double tmp[] = {(double)a...};
  for (int i = 0; i < (int)sizeof...(T); i++) r[i] = tmp[i];
  return r;
}

}  // namespace TooN
#include "helpers.h"
