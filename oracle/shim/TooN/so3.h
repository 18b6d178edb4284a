// TEST INFRASTRUCTURE — stand-in for TooN/so3.h; exp / ln evaluate the same formulas as the oracle
// (oracle_math.h restates them from TooN's so3.h), so that differences between the compiled
// reference sources and the oracle come from the reference's own code only.
#pragma once
#include "TooN.h"
#include "../../oracle_math.h"
namespace TooN {
template <class P = double> class SO3 {
 public:
  SO3() : m(Identity) {}
  template <class V, TOON_IF(is_vec<V>::value)> SO3(const V& w) { *this = exp(w); }
  template <class M, TOON_IF(is_mat<M>::value)> SO3(const M& rhs) : m(rhs) { coerce(); }
  template <class M, TOON_IF(is_mat<M>::value)> SO3& operator=(const M& rhs) { m = rhs; coerce(); return *this; }
  // Gram-Schmidt on the rows, as TooN::SO3::coerce
  void coerce() {
    normalize(m[0]);
    m[1] -= m[0] * (m[0] * m[1]);
    normalize(m[1]);
    m[2] -= m[0] * (m[0] * m[2]);
    m[2] -= m[1] * (m[1] * m[2]);
    normalize(m[2]);
  }
  template <class V, TOON_IF(is_vec<V>::value)> static SO3 exp(const V& w) {
    const double ww[3] = {w[0], w[1], w[2]};
    SO3 r;
    orc::so3_exp(ww, r.m.get_data_ptr());
    return r;
  }
  Vector<3> ln() const { Vector<3> r; orc::so3_ln(const_cast<Matrix<3>&>(m).get_data_ptr(), r.get_data_ptr()); return r; }
  SO3 inverse() const { SO3 r; r.m = m.T(); return r; }
  const Matrix<3>& get_matrix() const { return m; }
  SO3 operator*(const SO3& rhs) const { SO3 r; r.m = m * rhs.m; return r; }
  SO3& operator*=(const SO3& rhs) { m = m * rhs.m; return *this; }
  template <class V, TOON_IF(is_vec<V>::value)> Vector<3> operator*(const V& v) const { return m * v; }
  static Matrix<3> generator(int i) {
    Matrix<3> r(Zeros);
    r((i + 1) % 3, (i + 2) % 3) = -1;
    r((i + 2) % 3, (i + 1) % 3) = 1;
    return r;
  }
  template <class V, TOON_IF(is_vec<V>::value)> static Vector<3> generator_field(int i, const V& pos) {
    Vector<3> r;
    r[i] = 0;
    r[(i + 1) % 3] = -pos[(i + 2) % 3];
    r[(i + 2) % 3] = pos[(i + 1) % 3];
    return r;
  }
 private:
  Matrix<3> m;
  template <class A, TOON_IF(is_vec<A>::value)> static void normalize(const A& v) { const double n = std::sqrt(v * v); for (int i = 0; i < v.size(); i++) v[i] = v[i] / n; }
};
template <class P> inline std::ostream& operator<<(std::ostream& os, const SO3<P>& r) { return os << r.get_matrix(); }
}  // namespace TooN
