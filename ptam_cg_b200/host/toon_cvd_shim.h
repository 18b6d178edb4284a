// Minimal stand-ins for the TooN 2.2 / libCVD 20150407 value types that appear at the reference's
// class boundary (Bundle.h:110-118, KeyFrame.h:55-141, Tracker.h:158-163, ATANCamera.h:66-90).
// Neither library is vendored by the reference nor installed here, so the host mirror carries these
// look-alikes: same names, same accessors, only what the boundary needs.  A tree that has the real
// libraries defines PTAM_B200_HAVE_TOON_CVD and this header just includes them; the host classes
// touch the types only through the accessors both provide (operator[], get_rotation(),
// get_translation(), get_matrix(), size(), data(), row_stride()).
#pragma once

#ifdef PTAM_B200_HAVE_TOON_CVD
#include <TooN/TooN.h>
#include <TooN/se3.h>
#include <cvd/image.h>
#include <cvd/byte.h>
#else

#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

namespace TooN {

template <int N = 3, class P = double>
struct Vector {
  P v[N];
  Vector() { for (int i = 0; i < N; i++) v[i] = P(); }
  P& operator[](int i) { return v[i]; }
  const P& operator[](int i) const { return v[i]; }
  static int size() { return N; }
  Vector operator+(const Vector& o) const { Vector r; for (int i = 0; i < N; i++) r[i] = v[i] + o[i]; return r; }
  Vector operator-(const Vector& o) const { Vector r; for (int i = 0; i < N; i++) r[i] = v[i] - o[i]; return r; }
  Vector operator*(P s) const { Vector r; for (int i = 0; i < N; i++) r[i] = v[i] * s; return r; }
  P operator*(const Vector& o) const { P s = P(); for (int i = 0; i < N; i++) s += v[i] * o[i]; return s; }
};

template <class P = double> inline Vector<2, P> makeVector(P a, P b) { Vector<2, P> r; r[0] = a; r[1] = b; return r; }
template <class P = double> inline Vector<3, P> makeVector(P a, P b, P c) { Vector<3, P> r; r[0] = a; r[1] = b; r[2] = c; return r; }
inline Vector<5> makeVector(double a, double b, double c, double d, double e) { Vector<5> r; r[0] = a; r[1] = b; r[2] = c; r[3] = d; r[4] = e; return r; }
inline Vector<6> makeVector(double a, double b, double c, double d, double e, double f) { Vector<6> r; r[0] = a; r[1] = b; r[2] = c; r[3] = d; r[4] = e; r[5] = f; return r; }

template <int R = 3, int C = R, class P = double>
struct Matrix {
  P m[R][C];
  Matrix() { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) m[i][j] = P(); }
  P* operator[](int r) { return m[r]; }
  const P* operator[](int r) const { return m[r]; }
  P& operator()(int r, int c) { return m[r][c]; }
  const P& operator()(int r, int c) const { return m[r][c]; }
  Vector<R, P> operator*(const Vector<C, P>& x) const {
    Vector<R, P> y;
    for (int i = 0; i < R; i++) { P s = P(); for (int j = 0; j < C; j++) s += m[i][j] * x[j]; y[i] = s; }
    return y;
  }
  template <int K> Matrix<R, K, P> operator*(const Matrix<C, K, P>& o) const {
    Matrix<R, K, P> r;
    for (int i = 0; i < R; i++) for (int k = 0; k < K; k++) { P s = P(); for (int j = 0; j < C; j++) s += m[i][j] * o.m[j][k]; r.m[i][k] = s; }
    return r;
  }
  Matrix<C, R, P> T() const { Matrix<C, R, P> r; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r.m[j][i] = m[i][j]; return r; }
};

inline Vector<3> operator^(const Vector<3>& a, const Vector<3>& b) {
  return makeVector(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

// SO3 / SE3 with the exp/ln conventions of TooN so3.h / se3.h (translation first in the 6-vector;
// Taylor branches at theta^2 < 1e-8 and < 1e-6).
template <class P = double>
class SO3 {
 public:
  SO3() { for (int i = 0; i < 3; i++) mat[i][i] = 1; }
  const Matrix<3, 3, P>& get_matrix() const { return mat; }
  Matrix<3, 3, P>& get_matrix_mut() { return mat; }
  Vector<3, P> operator*(const Vector<3, P>& x) const { return mat * x; }
  SO3 operator*(const SO3& o) const { SO3 r; r.mat = mat * o.mat; return r; }
  SO3 inverse() const { SO3 r; r.mat = mat.T(); return r; }
  static void rodrigues(const Vector<3, P>& w, P A, P B, Matrix<3, 3, P>& R) {
    const P wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
    R[0][0] = 1 - B * (wy2 + wz2); R[1][1] = 1 - B * (wx2 + wz2); R[2][2] = 1 - B * (wx2 + wy2);
    P a = A * w[2], b = B * (w[0] * w[1]);
    R[0][1] = b - a; R[1][0] = b + a;
    a = A * w[1]; b = B * (w[0] * w[2]);
    R[0][2] = b + a; R[2][0] = b - a;
    a = A * w[0]; b = B * (w[1] * w[2]);
    R[1][2] = b - a; R[2][1] = b + a;
  }
  static SO3 exp(const Vector<3, P>& w) {
    const P th2 = w * w, th = std::sqrt(th2);
    P A, B;
    if (th2 < 1e-8) { A = 1 - th2 / 6; B = 0.5; }
    else if (th2 < 1e-6) { B = 0.5 - 0.25 * th2 / 6; A = 1 - th2 * (1.0 / 6) * (1 - th2 / 20); }
    else { A = std::sin(th) / th; B = (1 - std::cos(th)) / th2; }
    SO3 r;
    rodrigues(w, A, B, r.mat);
    return r;
  }
  Vector<3, P> ln() const {
    const P kSqrtHalf = 0.70710678118654752440, kPi = 3.14159265358979323846;
    const Matrix<3, 3, P>& R = mat;
    const P ca = (R[0][0] + R[1][1] + R[2][2] - 1) * 0.5;
    Vector<3, P> o = makeVector((R[2][1] - R[1][2]) / 2, (R[0][2] - R[2][0]) / 2, (R[1][0] - R[0][1]) / 2);
    const P sa = std::sqrt(o * o);
    if (ca > kSqrtHalf) { if (sa > 0) o = o * (std::asin(sa) / sa); }
    else if (ca > -kSqrtHalf) { o = o * (std::acos(ca) / sa); }
    else {
      const P angle = kPi - std::asin(sa);
      const P d0 = R[0][0] - ca, d1 = R[1][1] - ca, d2 = R[2][2] - ca;
      Vector<3, P> r2;
      if (d0 * d0 > d1 * d1 && d0 * d0 > d2 * d2) r2 = makeVector(d0, (R[1][0] + R[0][1]) / 2, (R[0][2] + R[2][0]) / 2);
      else if (d1 * d1 > d2 * d2) r2 = makeVector((R[1][0] + R[0][1]) / 2, d1, (R[2][1] + R[1][2]) / 2);
      else r2 = makeVector((R[0][2] + R[2][0]) / 2, (R[2][1] + R[1][2]) / 2, d2);
      if (r2 * o < 0) r2 = r2 * P(-1);
      o = r2 * (angle / std::sqrt(r2 * r2));
    }
    return o;
  }
 private:
  Matrix<3, 3, P> mat;
};

template <class P = double>
class SE3 {
 public:
  SE3() {}
  SE3(const SO3<P>& R, const Vector<3, P>& t) : rot(R), trans(t) {}
  const SO3<P>& get_rotation() const { return rot; }
  SO3<P>& get_rotation() { return rot; }
  const Vector<3, P>& get_translation() const { return trans; }
  Vector<3, P>& get_translation() { return trans; }
  Vector<3, P> operator*(const Vector<3, P>& x) const { return rot * x + trans; }
  SE3 operator*(const SE3& o) const { return SE3(rot * o.rot, trans + rot * o.trans); }
  SE3 inverse() const { SO3<P> ri = rot.inverse(); return SE3(ri, (ri * trans) * P(-1)); }
  static SE3 exp(const Vector<6, P>& mu) {
    const Vector<3, P> u = makeVector(mu[0], mu[1], mu[2]), w = makeVector(mu[3], mu[4], mu[5]);
    const P th2 = w * w, th = std::sqrt(th2);
    const Vector<3, P> cr = w ^ u;
    SE3 r;
    P A, B;
    if (th2 < 1e-8) { A = 1 - th2 / 6; B = 0.5; r.trans = u + cr * P(0.5); }
    else {
      P Cc;
      if (th2 < 1e-6) { Cc = (1.0 / 6) * (1 - th2 / 20); A = 1 - th2 * Cc; B = 0.5 - 0.25 * th2 / 6; }
      else { const P it = 1 / th; A = std::sin(th) * it; B = (1 - std::cos(th)) * (it * it); Cc = (1 - A) * (it * it); }
      r.trans = u + cr * B + (w ^ cr) * Cc;
    }
    SO3<P>::rodrigues(w, A, B, r.rot.get_matrix_mut());
    return r;
  }
  Vector<6, P> ln() const {
    const Vector<3, P> w = rot.ln();
    const P th = std::sqrt(w * w);
    P shtot = 0.5;
    if (th > 0.00001) shtot = std::sin(th / 2) / th;
    const SO3<P> half = SO3<P>::exp(w * P(-0.5));
    Vector<3, P> rt = half * trans;
    if (th > 0.001) rt = rt - w * ((trans * w) * (1 - 2 * shtot) / (w * w));
    else rt = rt - w * ((trans * w) / 24);
    rt = rt * (1 / (2 * shtot));
    Vector<6, P> o;
    for (int i = 0; i < 3; i++) { o[i] = rt[i]; o[3 + i] = w[i]; }
    return o;
  }
 private:
  SO3<P> rot;
  Vector<3, P> trans;
};

}  // namespace TooN

namespace CVD {

typedef unsigned char byte;

struct ImageRef {
  int x, y;
  ImageRef() : x(0), y(0) {}
  ImageRef(int xx, int yy) : x(xx), y(yy) {}
  bool operator==(const ImageRef& o) const { return x == o.x && y == o.y; }
  bool operator!=(const ImageRef& o) const { return !(*this == o); }
  ImageRef operator/(int d) const { return ImageRef(x / d, y / d); }
  ImageRef operator*(int d) const { return ImageRef(x * d, y * d); }
  ImageRef operator+(const ImageRef& o) const { return ImageRef(x + o.x, y + o.y); }
  ImageRef operator-(const ImageRef& o) const { return ImageRef(x - o.x, y - o.y); }
  unsigned mag_squared() const { return (unsigned)(x * x + y * y); }
};

// Non-owning view (CVD::BasicImage) and owning image (CVD::Image), row-major with a row stride.
template <class T>
class BasicImage {
 public:
  BasicImage(T* d, const ImageRef& sz, int stride = -1) : my_data(d), my_size(sz), my_stride(stride < 0 ? sz.x : stride) {}
  virtual ~BasicImage() {}
  const ImageRef& size() const { return my_size; }
  int row_stride() const { return my_stride; }
  T* data() { return my_data; }
  const T* data() const { return my_data; }
  T* operator[](int row) { return my_data + (size_t)row * my_stride; }
  const T* operator[](int row) const { return my_data + (size_t)row * my_stride; }
  T& operator[](const ImageRef& p) { return my_data[(size_t)p.y * my_stride + p.x]; }
  const T& operator[](const ImageRef& p) const { return my_data[(size_t)p.y * my_stride + p.x]; }
  bool in_image(const ImageRef& p) const { return p.x >= 0 && p.y >= 0 && p.x < my_size.x && p.y < my_size.y; }
  bool in_image_with_border(const ImageRef& p, int b) const { return p.x >= b && p.y >= b && p.x < my_size.x - b && p.y < my_size.y - b; }
 protected:
  T* my_data;
  ImageRef my_size;
  int my_stride;
};

template <class T>
class Image : public BasicImage<T> {
 public:
  Image() : BasicImage<T>(nullptr, ImageRef(0, 0)) {}
  explicit Image(const ImageRef& sz) : BasicImage<T>(nullptr, ImageRef(0, 0)) { resize(sz); }
  Image(const Image& o) : BasicImage<T>(nullptr, ImageRef(0, 0)) { *this = o; }
  Image& operator=(const Image& o) {  // deep copy (what Level::operator= enforces, KeyFrame.h:66-75)
    if (this != &o) { resize(o.size()); for (int y = 0; y < o.size().y; y++) std::memcpy((*this)[y], o[y], sizeof(T) * o.size().x); }
    return *this;
  }
  void resize(const ImageRef& sz) {
    store.assign((size_t)sz.x * sz.y, T());
    this->my_data = store.data(); this->my_size = sz; this->my_stride = sz.x;
  }
 private:
  std::vector<T> store;
};

}  // namespace CVD
#endif  // PTAM_B200_HAVE_TOON_CVD

namespace ptam_b200 {
// SE3 <-> the C-ABI's 12-double layout (rotation row-major, then translation)
inline void se3_to_array(const TooN::SE3<>& s, double* a) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[3 * i + j] = s.get_rotation().get_matrix()[i][j];
  for (int i = 0; i < 3; i++) a[9 + i] = s.get_translation()[i];
}
inline TooN::SE3<> se3_from_array(const double* a) {
  TooN::Vector<3> t;
  for (int i = 0; i < 3; i++) t[i] = a[9 + i];
#ifdef PTAM_B200_HAVE_TOON_CVD
  TooN::Matrix<3> R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = a[3 * i + j];
  return TooN::SE3<>(TooN::SO3<>(R), t);  // note: TooN re-orthonormalises (coerce) on construction
#else
  TooN::SO3<> R;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.get_matrix_mut()[i][j] = a[3 * i + j];
  return TooN::SE3<>(R, t);
#endif
}
}  // namespace ptam_b200
