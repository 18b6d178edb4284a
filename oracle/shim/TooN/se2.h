// TEST INFRASTRUCTURE — stand-in for TooN/se2.h (see TooN.h).
#pragma once
#include "so2.h"
namespace TooN {
template <class P = double> class SE2 {
 public:
  SE2() : t(Zeros) {}
  template <class V, TOON_IF(is_vec<V>::value)> SE2(const SO2<P>& R_, const V& t_) : R(R_), t(t_) {}
  SO2<P>& get_rotation() { return R; }
  const SO2<P>& get_rotation() const { return R; }
  Vector<2>& get_translation() { return t; }
  const Vector<2>& get_translation() const { return t; }
  SE2 inverse() const { const SO2<P> Ri = R.inverse(); return SE2(Ri, -(Ri * t)); }
  SE2 operator*(const SE2& o) const { return SE2(R * o.R, t + R * o.t); }
  SE2& operator*=(const SE2& o) { *this = *this * o; return *this; }
  template <class V, TOON_IF(is_vec<V>::value && (V::Size == 2))> Vector<2> operator*(const V& v) const { return R * v + t; }
 private:
  SO2<P> R;
  Vector<2> t;
};
}  // namespace TooN
