// TEST INFRASTRUCTURE — stand-in for TooN/so2.h (see TooN.h).
#pragma once
#include "TooN.h"
namespace TooN {
template <class P = double> class SO2 {
 public:
  SO2() : m(Identity) {}
  explicit SO2(double a) { *this = exp(a); }
  static SO2 exp(double a) { SO2 r; r.m(0, 0) = r.m(1, 1) = std::cos(a); r.m(1, 0) = std::sin(a); r.m(0, 1) = -r.m(1, 0); return r; }
  double ln() const { return std::atan2(m(1, 0), m(0, 0)); }
  SO2 inverse() const { SO2 r; r.m = m.T(); return r; }
  const Matrix<2>& get_matrix() const { return m; }
  SO2 operator*(const SO2& o) const { SO2 r; r.m = m * o.m; return r; }
  SO2& operator*=(const SO2& o) { m = m * o.m; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> Vector<2> operator*(const V& v) const { return m * v; }
 private:
  Matrix<2> m;
};
}  // namespace TooN
