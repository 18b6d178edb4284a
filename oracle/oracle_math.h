// TEST INFRASTRUCTURE — CPU oracle for the PTAM hot paths. NOT part of the product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may build, load or call anything under oracle/.
//
// LIBRARY ARITHMETIC RESTATED, NOT PINNED: the reference (cggos/ptam_cg) ships no tests or golden vectors
// and part of its arithmetic lives in three un-vendored libraries (TooN 2.2, libCVD 20150407, GVars3
// 3.0, pinned only by URL in install_deps.sh:39-60) that are absent here.  This file restates, in plain
// single-threaded C++, the published algorithms of the TooN pieces the hot paths call (SE3/SO3 exp+ln,
// square-root-free LDL^T "Cholesky", WLS) and the reference's own ATANCamera model (the latter IS
// pinned: oracle/_ref compiles src/ATANCamera.cc itself).  Each function cites the reference call site
// it serves.  oracle/shim/TooN/{so3,se3,Cholesky}.h — the header stand-ins the reference's own sources
// are compiled against for oracle/_ref — call these same functions, so that every difference between
// oracle/_ref and the oracle comes from the reference's own code.
//
// Numeric contract shared (by specification, not by code) with the CUDA product:
//   * IEEE-754 binary64, round-to-nearest, NO fused multiply-add (-ffp-contract=off here,
//     -fmad=false there), so that integer-valued outputs (template bytes, patch offsets, search
//     levels) are bit-identical on CPU and GPU.
//   * atan() is evaluated by the fdlibm s_atan.c algorithm (< 1 ulp), spelled out below, instead of
//     the platform libm: glibc and CUDA libdevice differ from each other by an ulp, and the
//     bilinear->byte truncation in CVD::sample turns such ulps into different template bytes on
//     flat image regions.  Build with -DORACLE_LIBM_ATAN to use libm (tests quantify the effect).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

// ---------------------------------------------------------------------------------------------
// atan: fdlibm s_atan.c (Sun Microsystems, 1993) algorithm.  Used by ATANCamera::rtrans_factor
// (include/ATANCamera.h:143-149).
// ---------------------------------------------------------------------------------------------
inline double spec_atan(double x) {
#ifdef ORACLE_LIBM_ATAN
  return std::atan(x);
#else
  static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                                   9.82793723247329054082e-01, 1.57079632679489655800e+00};
  static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                                   1.39033110312309984516e-17, 6.12323399573676603587e-17};
  static const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01,
                                1.42857142725034663711e-01,  -1.11111104054623557880e-01,
                                9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                6.66107313738753120669e-02,  -5.83357013379057348645e-02,
                                4.97687799461593236017e-02,  -3.65315727442169155270e-02,
                                1.62858201153657823623e-02};
  if (x != x) return x;
  const bool neg = std::signbit(x);
  double ax = std::fabs(x);
  int id;
  if (ax >= 1.8446744073709552e19) {  // |x| >= 2^64
    double r = atanhi[3] + atanlo[3];
    return neg ? -r : r;
  }
  if (ax < 0.4375) {
    if (ax < 3.7252902984619141e-09) return x;  // |x| < 2^-28
    id = -1;
  } else if (ax < 1.1875) {
    if (ax < 0.6875) { id = 0; ax = (2.0 * ax - 1.0) / (2.0 + ax); }
    else             { id = 1; ax = (ax - 1.0) / (ax + 1.0); }
  } else {
    if (ax < 2.4375) { id = 2; ax = (ax - 1.5) / (1.0 + 1.5 * ax); }
    else             { id = 3; ax = -1.0 / ax; }
  }
  const double z = ax * ax;
  const double w = z * z;
  const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) {
    double r = ax - ax * (s1 + s2);
    return neg ? -r : r;
  }
  double r = atanhi[id] - ((ax * (s1 + s2) - atanlo[id]) - ax);
  return neg ? -r : r;
#endif
}

// ---------------------------------------------------------------------------------------------
// SE3 / SO3 (TooN se3.h / so3.h as used at Tracker.cc:567,641,1028,1037 and Bundle.cc:297,501).
// Storage: R row-major 3x3, then t.  12 doubles at the C boundary in the same order.
// ---------------------------------------------------------------------------------------------
struct SE3 {
  double R[9];
  double t[3];
  SE3() { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0; t[0] = t[1] = t[2] = 0; }
  static SE3 from12(const double* p) { SE3 s; std::memcpy(s.R, p, 72); std::memcpy(s.t, p + 9, 24); return s; }
  void to12(double* p) const { std::memcpy(p, R, 72); std::memcpy(p + 9, t, 24); }
  void apply(const double* x, double* y) const {
    for (int r = 0; r < 3; r++) y[r] = (R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2]) + t[r];
  }
  void rotate(const double* x, double* y) const {
    for (int r = 0; r < 3; r++) y[r] = R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2];
  }
};

inline void rodrigues(const double* w, double A, double B, double* R) {
  const double wx2 = w[0] * w[0], wy2 = w[1] * w[1], wz2 = w[2] * w[2];
  R[0] = 1.0 - B * (wy2 + wz2);
  R[4] = 1.0 - B * (wx2 + wz2);
  R[8] = 1.0 - B * (wx2 + wy2);
  double a = A * w[2], b = B * (w[0] * w[1]);
  R[1] = b - a; R[3] = b + a;
  a = A * w[1]; b = B * (w[0] * w[2]);
  R[2] = b + a; R[6] = b - a;
  a = A * w[0]; b = B * (w[1] * w[2]);
  R[5] = b - a; R[7] = b + a;
}

inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// SE3<>::exp: mu[0:3] translation part, mu[3:6] rotation vector.
inline SE3 se3_exp(const double* mu) {
  SE3 r;
  const double* w = mu + 3;
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = std::sqrt(theta_sq);
  double A, B;
  double cr[3];
  cross3(w, mu, cr);
  if (theta_sq < 1e-8) {
    A = 1.0 - (1.0 / 6.0) * theta_sq;
    B = 0.5;
    for (int i = 0; i < 3; i++) r.t[i] = mu[i] + 0.5 * cr[i];
  } else {
    double C;
    if (theta_sq < 1e-6) {
      C = (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq);
      A = 1.0 - theta_sq * C;
      B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq;
    } else {
      const double inv_theta = 1.0 / theta;
      A = std::sin(theta) * inv_theta;
      B = (1 - std::cos(theta)) * (inv_theta * inv_theta);
      C = (1 - A) * (inv_theta * inv_theta);
    }
    double wcr[3];
    cross3(w, cr, wcr);
    for (int i = 0; i < 3; i++) r.t[i] = mu[i] + B * cr[i] + C * wcr[i];
  }
  rodrigues(w, A, B, r.R);
  return r;
}

inline void so3_exp(const double* w, double* R) {
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double theta = std::sqrt(theta_sq);
  double A, B;
  if (theta_sq < 1e-8) {
    A = 1.0 - (1.0 / 6.0) * theta_sq;
    B = 0.5;
  } else if (theta_sq < 1e-6) {
    B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq;
    A = 1.0 - theta_sq * (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq);
  } else {
    const double inv_theta = 1.0 / theta;
    A = std::sin(theta) * inv_theta;
    B = (1 - std::cos(theta)) * (inv_theta * inv_theta);
  }
  rodrigues(w, A, B, R);
}

inline void so3_ln(const double* R, double* out) {
  const double cos_angle = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  out[0] = (R[7] - R[5]) / 2;
  out[1] = (R[2] - R[6]) / 2;
  out[2] = (R[3] - R[1]) / 2;
  double sin_angle_abs = std::sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
  if (cos_angle > M_SQRT1_2) {
    if (sin_angle_abs > 0) {
      const double f = std::asin(sin_angle_abs) / sin_angle_abs;
      for (int i = 0; i < 3; i++) out[i] *= f;
    }
  } else if (cos_angle > -M_SQRT1_2) {
    const double f = std::acos(cos_angle) / sin_angle_abs;
    for (int i = 0; i < 3; i++) out[i] *= f;
  } else {
    // Near pi: use the symmetric part (TooN so3.h).
    const double angle = M_PI - std::asin(sin_angle_abs);
    const double d0 = R[0] - cos_angle, d1 = R[4] - cos_angle, d2 = R[8] - cos_angle;
    double r2[3];
    if (d0 * d0 > d1 * d1 && d0 * d0 > d2 * d2) {
      r2[0] = d0; r2[1] = (R[3] + R[1]) / 2; r2[2] = (R[2] + R[6]) / 2;
    } else if (d1 * d1 > d2 * d2) {
      r2[0] = (R[3] + R[1]) / 2; r2[1] = d1; r2[2] = (R[7] + R[5]) / 2;
    } else {
      r2[0] = (R[2] + R[6]) / 2; r2[1] = (R[7] + R[5]) / 2; r2[2] = d2;
    }
    if (r2[0] * out[0] + r2[1] * out[1] + r2[2] * out[2] < 0) { r2[0] = -r2[0]; r2[1] = -r2[1]; r2[2] = -r2[2]; }
    const double n = std::sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    for (int i = 0; i < 3; i++) out[i] = angle * (r2[i] / n);
  }
}

inline void se3_ln(const SE3& s, double* out6) {
  double rot[3];
  so3_ln(s.R, rot);
  const double rr = rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2];
  const double theta = std::sqrt(rr);
  double shtot = 0.5;
  if (theta > 0.00001) shtot = std::sin(theta / 2) / theta;
  double half[3] = {rot[0] * -0.5, rot[1] * -0.5, rot[2] * -0.5};
  double H[9];
  so3_exp(half, H);
  double rt[3];
  for (int r = 0; r < 3; r++) rt[r] = H[3 * r] * s.t[0] + H[3 * r + 1] * s.t[1] + H[3 * r + 2] * s.t[2];
  const double tr = s.t[0] * rot[0] + s.t[1] * rot[1] + s.t[2] * rot[2];
  if (theta > 0.001) {
    const double f = (tr * (1 - 2 * shtot)) / rr;
    for (int i = 0; i < 3; i++) rt[i] -= rot[i] * f;
  } else {
    const double f = tr / 24;
    for (int i = 0; i < 3; i++) rt[i] -= rot[i] * f;
  }
  for (int i = 0; i < 3; i++) out6[i] = rt[i] / (2 * shtot);
  for (int i = 0; i < 3; i++) out6[3 + i] = rot[i];
}

inline SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
  for (int i = 0; i < 3; i++)
    r.t[i] = a.t[i] + (a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2]);
  return r;
}

inline SE3 se3_inverse(const SE3& a) {
  SE3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * j + i];
  for (int i = 0; i < 3; i++) r.t[i] = -(r.R[3 * i] * a.t[0] + r.R[3 * i + 1] * a.t[1] + r.R[3 * i + 2] * a.t[2]);
  return r;
}

// ---------------------------------------------------------------------------------------------
// TooN::Cholesky — square-root-free LDL^T, no pivoting, no failure path (NaN/Inf propagate).
// A is n x n row-major with leading dimension ld; only the lower triangle is read; overwritten
// with L (unit, strictly lower), D (diagonal) and L*D cached in the strict upper triangle.
// Serves PatchFinder.cc:235-236, Bundle.cc:356-357, Bundle.cc:458 and WLS<6>::compute.
// ---------------------------------------------------------------------------------------------
inline void ldlt_factor(double* A, int n, int ld) {
  for (int col = 0; col < n; col++) {
    double inv_diag = 1;
    for (int row = col; row < n; row++) {
      double val = A[row * ld + col];
      for (int c2 = 0; c2 < col; c2++) val -= A[c2 * ld + col] * A[row * ld + c2];
      if (row == col) {
        A[row * ld + col] = val;
        if (val == 0) return;
        inv_diag = 1 / val;
      } else {
        A[col * ld + row] = val;
        A[row * ld + col] = val * inv_diag;
      }
    }
  }
}

inline void ldlt_backsub(const double* A, int n, int ld, const double* b, double* x) {
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) {  // L y = b
    double v = b[i];
    for (int j = 0; j < i; j++) v -= A[i * ld + j] * y[j];
    y[i] = v;
  }
  for (int i = 0; i < n; i++) y[i] /= A[i * ld + i];  // D
  for (int i = n - 1; i >= 0; i--) {  // L^T x = y
    double v = y[i];
    for (int j = i + 1; j < n; j++) v -= A[j * ld + i] * x[j];
    x[i] = v;
  }
}

// get_inverse(): back-substitute the identity, column by column.
inline void ldlt_inverse(double* A, int n, double* inv) {
  ldlt_factor(A, n, n);
  std::vector<double> e(n), x(n);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) e[i] = (i == c) ? 1.0 : 0.0;
    ldlt_backsub(A, n, n, e.data(), x.data());
    for (int i = 0; i < n; i++) inv[i * n + c] = x[i];
  }
}

// ---------------------------------------------------------------------------------------------
// ATANCamera as a pure function (src/ATANCamera.cc:27-140,179-209; include/ATANCamera.h:143-157).
// The reference object is stateful (GetProjectionDerivs uses values cached by the last Project,
// ATANCamera.h:12-16); here Project returns that cached state explicitly.
// ---------------------------------------------------------------------------------------------
struct Camera {
  double p[5];
  double size[2];
  double focal[2], center[2], inv_focal[2];
  double w, winv, tan2, one_over_tan2, dist_enabled;
  double largest_radius, max_r, one_pixel_dist;

  double invrtrans(double r) const { return w == 0.0 ? r : std::tan(r * w) * one_over_tan2; }
  double rtrans_factor(double r) const {
    if (r < 0.001 || w == 0.0) return 1.0;
    return winv * spec_atan(r * tan2) / r;
  }
  void init(const double* params, double W, double H) {  // RefreshParams, ATANCamera.cc:27-70
    for (int i = 0; i < 5; i++) p[i] = params[i];
    size[0] = W; size[1] = H;
    focal[0] = W * p[0]; focal[1] = H * p[1];
    center[0] = W * p[2] - 0.5; center[1] = H * p[3] - 0.5;
    inv_focal[0] = 1.0 / focal[0]; inv_focal[1] = 1.0 / focal[1];
    w = p[4];
    if (w != 0.0) {
      tan2 = 2.0 * std::tan(w / 2.0);
      one_over_tan2 = 1.0 / tan2;
      winv = 1.0 / w;
      dist_enabled = 1.0;
    } else {
      winv = 0.0; tan2 = 0.0; one_over_tan2 = 0.0; dist_enabled = 0.0;
    }
    double v0 = std::max(p[2], 1.0 - p[2]) / p[0];
    double v1 = std::max(p[3], 1.0 - p[3]) / p[1];
    largest_radius = invrtrans(std::sqrt(v0 * v0 + v1 * v1));
    max_r = 1.5 * largest_radius;
    {  // mdOnePixelDist (ATANCamera.cc:59-64)
      const double c[2] = {W / 2, H / 2}, a[2] = {W / 2 + 1.0, H / 2 + 1.0};
      double uc[2], ua[2];
      unproject(c, uc); unproject(a, ua);
      const double d0 = uc[0] - ua[0], d1 = uc[1] - ua[1];
      one_pixel_dist = std::sqrt(d0 * d0 + d1 * d1) / std::sqrt(2.0);
    }
  }
  struct Proj { double im[2]; double cam[2]; double r; double factor; bool invalid; };
  Proj project(const double* cam) const {  // ATANCamera.cc:109-121
    Proj q;
    q.cam[0] = cam[0]; q.cam[1] = cam[1];
    q.r = std::sqrt(cam[0] * cam[0] + cam[1] * cam[1]);
    q.invalid = q.r > max_r;
    q.factor = rtrans_factor(q.r);
    q.im[0] = center[0] + focal[0] * (q.factor * cam[0]);
    q.im[1] = center[1] + focal[1] * (q.factor * cam[1]);
    return q;
  }
  void unproject(const double* im, double* cam) const {  // ATANCamera.cc:125-140
    double d0 = (im[0] - center[0]) * inv_focal[0];
    double d1 = (im[1] - center[1]) * inv_focal[1];
    double dr = std::sqrt(d0 * d0 + d1 * d1);
    double r = invrtrans(dr);
    double f = dr > 0.01 ? r / dr : 1.0;
    cam[0] = f * d0; cam[1] = f * d1;
  }
  void derivs(const Proj& q, double* m /*2x2 row-major*/) const {  // ATANCamera.cc:179-209
    double fx, fy;
    const double k = tan2, x = q.cam[0], y = q.cam[1];
    const double r = q.r * dist_enabled;
    if (r < 0.01) {
      fx = 0.0; fy = 0.0;
    } else {
      fx = winv * (k * x) / (r * r * (1 + k * k * r * r)) - x * q.factor / (r * r);
      fy = winv * (k * y) / (r * r * (1 + k * k * r * r)) - y * q.factor / (r * r);
    }
    m[0] = focal[0] * (fx * x + q.factor);
    m[2] = focal[1] * (fx * y);
    m[1] = focal[0] * (fy * x);
    m[3] = focal[1] * (fy * y + q.factor);
  }
};

// ---------------------------------------------------------------------------------------------
// M-estimators (include/Tools.h:128-254).  est: 0 Tukey, 1 Cauchy, 2 Huber.
// ---------------------------------------------------------------------------------------------
inline double mest_find_sigma_squared(std::vector<double>& v, int est) {
  std::sort(v.begin(), v.end());
  double med = v[v.size() / 2];
  double sigma = 1.4826 * (1 + 5.0 / (v.size() * 2 - 6)) * std::sqrt(med);
  sigma = (est == 2 ? 1.345 : 4.6851) * sigma;
  return sigma * sigma;
}
inline double mest_sqrt_weight(double e2, double s2, int est) {
  if (est == 0) return e2 > s2 ? 0.0 : 1.0 - (e2 / s2);
  if (est == 1) return std::sqrt(1.0 / (1.0 + e2 / s2));
  return std::sqrt(e2 < s2 ? 1.0 : std::sqrt(s2 / e2));
}
inline double mest_weight(double e2, double s2, int est) {
  if (est == 0) { double d = mest_sqrt_weight(e2, s2, 0); return d * d; }
  if (est == 1) return 1.0 / (1.0 + e2 / s2);
  return e2 < s2 ? 1.0 : std::sqrt(s2 / e2);
}
inline double mest_objective(double e2, double s2, int est) {
  if (est == 0) {
    if (e2 > s2) return 1.0;
    double d = 1.0 - e2 / s2;
    return 1.0 - d * d * d;
  }
  if (est == 1) return std::log(1.0 + e2 / s2);
  if (e2 < s2) return 0.5 * e2;
  double s = std::sqrt(s2), e = std::sqrt(e2);
  return s * (e - 0.5 * s);
}

}  // namespace orc
