// TEST INFRASTRUCTURE — stand-in for TooN/helpers.h (see TooN.h in this directory).
#pragma once
#include "TooN.h"
namespace TooN {
template <class A, TOON_IF(is_vec<A>::value)> inline double norm_sq(const A& v) { return v * v; }
template <class A, TOON_IF(is_vec<A>::value)> inline double norm(const A& v) { return std::sqrt(v * v); }
template <class A, TOON_IF(is_vec<A>::value)> inline void normalize(A& v) { const double n = std::sqrt(v * v); for (int i = 0; i < v.size(); i++) v[i] = v[i] / n; }
template <class A, TOON_IF(is_vec<A>::value)> inline void normalize(const A& v) { const double n = std::sqrt(v * v); for (int i = 0; i < v.size(); i++) v[i] = v[i] / n; }
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<A::Size> unit(const A& v) { return v * (1 / std::sqrt(v * v)); }
// project: divide by the last element and drop it; unproject: append 1
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<(A::Size == Dynamic ? Dynamic : A::Size - 1)> project(const A& v) {
  Vector<(A::Size == Dynamic ? Dynamic : A::Size - 1)> r(v.size() - 1);
  const double last = v[v.size() - 1];
  for (int i = 0; i < v.size() - 1; i++) r[i] = v[i] / last;
  return r;
}
template <class A, TOON_IF(is_vec<A>::value)> inline Vector<(A::Size == Dynamic ? Dynamic : A::Size + 1)> unproject(const A& v) {
  Vector<(A::Size == Dynamic ? Dynamic : A::Size + 1)> r(v.size() + 1);
  for (int i = 0; i < v.size(); i++) r[i] = v[i];
  r[v.size()] = 1;
  return r;
}
template <class A, TOON_IF(is_mat<A>::value)> inline double trace(const A& m) { double t = 0; for (int i = 0; i < m.num_rows(); i++) t += m(i, i); return t; }
template <class A, TOON_IF(is_vec<A>::value)> inline void Fill(A& v, double x) { for (int i = 0; i < v.size(); i++) v[i] = x; }
}  // namespace TooN
