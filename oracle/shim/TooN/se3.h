// TEST INFRASTRUCTURE — stand-in for TooN/se3.h (see so3.h).
#pragma once
#include "so3.h"
namespace TooN {
template <class P = double> class SE3 {
 public:
  SE3() : t(Zeros) {}
  template <class V, TOON_IF(is_vec<V>::value)> SE3(const SO3<P>& R_, const V& t_) : R(R_), t(t_) {}
  template <class V, TOON_IF(is_vec<V>::value)> SE3(const V& mu) { *this = exp(mu); }
  SO3<P>& get_rotation() { return R; }
  const SO3<P>& get_rotation() const { return R; }
  Vector<3>& get_translation() { return t; }
  const Vector<3>& get_translation() const { return t; }
  template <class V, TOON_IF(is_vec<V>::value)> static SE3 exp(const V& mu) {
    double m[6];
    for (int i = 0; i < 6; i++) m[i] = mu[i];
    const orc::SE3 e = orc::se3_exp(m);
    return from_orc(e);
  }
  static Vector<6> ln(const SE3& s) { return s.ln(); }
  Vector<6> ln() const { Vector<6> r; orc::se3_ln(to_orc(), r.get_data_ptr()); return r; }
  SE3 inverse() const { const SO3<P> Ri = R.inverse(); return SE3(Ri, -(Ri * t)); }
  SE3 operator*(const SE3& rhs) const { return SE3(R * rhs.R, t + R * rhs.t); }
  SE3& operator*=(const SE3& rhs) { t = t + R * rhs.t; R = R * rhs.R; return *this; }
  SE3& left_multiply_by(const SE3& l) { t = l.t + l.R * t; R = l.R * R; return *this; }
  template <class V, TOON_IF(is_vec<V>::value && (V::Size == 3))> Vector<3> operator*(const V& v) const { return R * v + t; }
  template <class V, TOON_IF(is_vec<V>::value && (V::Size == 4))> Vector<4> operator*(const V& v) const {
    Vector<4> r;
    r.template slice<0, 3>() = R * v.template slice<0, 3>() + t * v[3];
    r[3] = v[3];
    return r;
  }
  template <class V, TOON_IF(is_vec<V>::value)> static Vector<4> generator_field(int i, const V& pos) {
    Vector<4> r(Zeros);
    if (i < 3) { r[i] = pos[3]; return r; }
    r[(i + 1) % 3] = -pos[(i + 2) % 3];
    r[(i + 2) % 3] = pos[(i + 1) % 3];
    return r;
  }
  orc::SE3 to_orc() const {
    orc::SE3 s;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) s.R[3 * i + j] = R.get_matrix()(i, j); s.t[i] = t[i]; }
    return s;
  }
  static SE3 from_orc(const orc::SE3& e) {
    SE3 r;
    Matrix<3> m;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) m(i, j) = e.R[3 * i + j]; r.t[i] = e.t[i]; }
    r.R = raw(m);
    return r;
  }
  // builds an SO3 from a matrix without the re-orthonormalisation of the converting constructor
  static SO3<P> raw(const Matrix<3>& m) { SO3<P> s; std::memcpy((void*)&s.get_matrix(), &m, sizeof(m)); return s; }
 private:
  SO3<P> R;
  Vector<3> t;
};
template <class P> inline std::ostream& operator<<(std::ostream& os, const SE3<P>& s) { return os << s.get_rotation() << s.get_translation() << "\n"; }
}  // namespace TooN
