// Shared device/host helpers for the B200 PTAM hot paths (sm_100a only).
// Small fixed-size f64 math written for this library: SE3/SO3 exp+ln, square-root-free LDL^T,
// the ATAN camera model as a pure function, M-estimators, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#define PTAM_DEV __device__ __forceinline__
#define PTAM_HD __host__ __device__ __forceinline__

namespace ptam {

constexpr int kLevels = 4;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// atan with a fixed, portable evaluation order (argument reduction at 7/16, 11/16, 19/16, 39/16
// and an odd degree-23 polynomial, the classic fdlibm scheme; < 1 ulp).  The tracker's integer
// outputs depend on bit-identical camera projections on every platform, so the library carries
// its own atan instead of libdevice's (2 ulp, implementation-defined).  Used by
// ATANCamera::rtrans_factor (reference include/ATANCamera.h:143-149).  Compiled with -fmad=false.
// ------------------------------------------------------------------------------------------
PTAM_HD double atan_portable(double x) {
  const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                        9.82793723247329054082e-01, 1.57079632679489655800e+00};
  const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                        1.39033110312309984516e-17, 6.12323399573676603587e-17};
  const double c0 = 3.33333333333329318027e-01, c1 = -1.99999999998764832476e-01,
               c2 = 1.42857142725034663711e-01, c3 = -1.11111104054623557880e-01,
               c4 = 9.09088713343650656196e-02, c5 = -7.69187620504482999495e-02,
               c6 = 6.66107313738753120669e-02, c7 = -5.83357013379057348645e-02,
               c8 = 4.97687799461593236017e-02, c9 = -3.65315727442169155270e-02,
               c10 = 1.62858201153657823623e-02;
  if (x != x) return x;
  const bool neg = x < 0.0 || (x == 0.0 && 1.0 / x < 0.0);
  double a = neg ? -x : x;
  int id;
  if (a >= 1.8446744073709552e19) {
    const double r = hi[3] + lo[3];
    return neg ? -r : r;
  }
  if (a < 0.4375) {
    if (a < 3.7252902984619141e-09) return x;
    id = -1;
  } else if (a < 1.1875) {
    if (a < 0.6875) { id = 0; a = (2.0 * a - 1.0) / (2.0 + a); }
    else            { id = 1; a = (a - 1.0) / (a + 1.0); }
  } else {
    if (a < 2.4375) { id = 2; a = (a - 1.5) / (1.0 + 1.5 * a); }
    else            { id = 3; a = -1.0 / a; }
  }
  const double z = a * a;
  const double w = z * z;
  const double s1 = z * (c0 + w * (c2 + w * (c4 + w * (c6 + w * (c8 + w * c10)))));
  const double s2 = w * (c1 + w * (c3 + w * (c5 + w * (c7 + w * c9))));
  double r;
  if (id < 0) r = a - a * (s1 + s2);
  else r = hi[id] - ((a * (s1 + s2) - lo[id]) - a);
  return neg ? -r : r;
}

// ------------------------------------------------------------------------------------------
// ATAN (FOV) camera — reference src/ATANCamera.cc:27-140,179-209.  Parameters are derived once on
// the host (tan() of the distortion angle included) and passed to kernels by value.
// ------------------------------------------------------------------------------------------
struct CamModel {
  double focal[2], center[2], inv_focal[2];
  double w, winv, tan2, one_over_tan2, dist_enabled;
  double largest_radius, max_r;
  double img_w, img_h;
};

struct CamProj {
  double im[2];
  double cam[2];
  double r, factor;
  bool invalid;
};

PTAM_HD CamProj cam_project(const CamModel& c, double x, double y) {  // ATANCamera.cc:109-121
  CamProj q;
  q.cam[0] = x; q.cam[1] = y;
  q.r = sqrt(x * x + y * y);
  q.invalid = q.r > c.max_r;
  q.factor = (q.r < 0.001 || c.w == 0.0) ? 1.0 : c.winv * atan_portable(q.r * c.tan2) / q.r;
  q.im[0] = c.center[0] + c.focal[0] * (q.factor * x);
  q.im[1] = c.center[1] + c.focal[1] * (q.factor * y);
  return q;
}

PTAM_HD void cam_derivs(const CamModel& c, const CamProj& q, double* m) {  // ATANCamera.cc:179-209
  double fx, fy;
  const double k = c.tan2, x = q.cam[0], y = q.cam[1];
  const double r = q.r * c.dist_enabled;
  if (r < 0.01) {
    fx = 0.0; fy = 0.0;
  } else {
    fx = c.winv * (k * x) / (r * r * (1 + k * k * r * r)) - x * q.factor / (r * r);
    fy = c.winv * (k * y) / (r * r * (1 + k * k * r * r)) - y * q.factor / (r * r);
  }
  m[0] = c.focal[0] * (fx * x + q.factor);
  m[2] = c.focal[1] * (fx * y);
  m[1] = c.focal[0] * (fy * x);
  m[3] = c.focal[1] * (fy * y + q.factor);
}

// ------------------------------------------------------------------------------------------
// SE3: p[0..8] rotation row-major, p[9..11] translation ("camera from world").
// exp/ln follow TooN se3.h / so3.h (Taylor branches at theta^2 < 1e-8 and < 1e-6).
// ------------------------------------------------------------------------------------------
PTAM_HD void se3_apply(const double* p, const double* x, double* y) {
  y[0] = (p[0] * x[0] + p[1] * x[1] + p[2] * x[2]) + p[9];
  y[1] = (p[3] * x[0] + p[4] * x[1] + p[5] * x[2]) + p[10];
  y[2] = (p[6] * x[0] + p[7] * x[1] + p[8] * x[2]) + p[11];
}
PTAM_HD void so3_rotate(const double* p, const double* x, double* y) {
  y[0] = p[0] * x[0] + p[1] * x[1] + p[2] * x[2];
  y[1] = p[3] * x[0] + p[4] * x[1] + p[5] * x[2];
  y[2] = p[6] * x[0] + p[7] * x[1] + p[8] * x[2];
}
PTAM_HD void rodrigues(const double* w, double A, double B, double* R) {
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
PTAM_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
PTAM_HD void se3_exp(const double* mu, double* out) {
  const double* w = mu + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double A, B, cr[3];
  cross3(w, mu, cr);
  if (th2 < 1e-8) {
    A = 1.0 - (1.0 / 6.0) * th2;
    B = 0.5;
    for (int i = 0; i < 3; i++) out[9 + i] = mu[i] + 0.5 * cr[i];
  } else {
    double Cc;
    if (th2 < 1e-6) {
      Cc = (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * th2);
      A = 1.0 - th2 * Cc;
      B = 0.5 - 0.25 * (1.0 / 6.0) * th2;
    } else {
      const double it = 1.0 / th;
      A = sin(th) * it;
      B = (1 - cos(th)) * (it * it);
      Cc = (1 - A) * (it * it);
    }
    double wcr[3];
    cross3(w, cr, wcr);
    for (int i = 0; i < 3; i++) out[9 + i] = mu[i] + B * cr[i] + Cc * wcr[i];
  }
  rodrigues(w, A, B, out);
}
PTAM_HD void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(th2);
  double A, B;
  if (th2 < 1e-8) { A = 1.0 - (1.0 / 6.0) * th2; B = 0.5; }
  else if (th2 < 1e-6) { B = 0.5 - 0.25 * (1.0 / 6.0) * th2; A = 1.0 - th2 * (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * th2); }
  else { const double it = 1.0 / th; A = sin(th) * it; B = (1 - cos(th)) * (it * it); }
  rodrigues(w, A, B, R);
}
// c = a * b (compose), all 12-double SE3
PTAM_HD void se3_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  for (int i = 0; i < 3; i++)
    c[9 + i] = a[9 + i] + (a[3 * i] * b[9] + a[3 * i + 1] * b[10] + a[3 * i + 2] * b[11]);
}
PTAM_HD void se3_inverse(const double* a, double* r) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * j + i];
  for (int i = 0; i < 3; i++) r[9 + i] = -(r[3 * i] * a[9] + r[3 * i + 1] * a[10] + r[3 * i + 2] * a[11]);
}
PTAM_HD void so3_ln(const double* R, double* out) {
  const double kSqrtHalf = 0.70710678118654752440;
  const double kPi = 3.14159265358979323846;
  const double ca = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  out[0] = (R[7] - R[5]) / 2;
  out[1] = (R[2] - R[6]) / 2;
  out[2] = (R[3] - R[1]) / 2;
  const double sa = sqrt(out[0] * out[0] + out[1] * out[1] + out[2] * out[2]);
  if (ca > kSqrtHalf) {
    if (sa > 0) { const double f = asin(sa) / sa; out[0] *= f; out[1] *= f; out[2] *= f; }
  } else if (ca > -kSqrtHalf) {
    const double f = acos(ca) / sa; out[0] *= f; out[1] *= f; out[2] *= f;
  } else {
    const double angle = kPi - asin(sa);
    const double d0 = R[0] - ca, d1 = R[4] - ca, d2 = R[8] - ca;
    double r2[3];
    if (d0 * d0 > d1 * d1 && d0 * d0 > d2 * d2) { r2[0] = d0; r2[1] = (R[3] + R[1]) / 2; r2[2] = (R[2] + R[6]) / 2; }
    else if (d1 * d1 > d2 * d2) { r2[0] = (R[3] + R[1]) / 2; r2[1] = d1; r2[2] = (R[7] + R[5]) / 2; }
    else { r2[0] = (R[2] + R[6]) / 2; r2[1] = (R[7] + R[5]) / 2; r2[2] = d2; }
    if (r2[0] * out[0] + r2[1] * out[1] + r2[2] * out[2] < 0) { r2[0] = -r2[0]; r2[1] = -r2[1]; r2[2] = -r2[2]; }
    const double n = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    for (int i = 0; i < 3; i++) out[i] = angle * (r2[i] / n);
  }
}
PTAM_HD void se3_ln(const double* s, double* out6) {
  double rot[3];
  so3_ln(s, rot);
  const double rr = rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2];
  const double th = sqrt(rr);
  double shtot = 0.5;
  if (th > 0.00001) shtot = sin(th / 2) / th;
  const double half[3] = {rot[0] * -0.5, rot[1] * -0.5, rot[2] * -0.5};
  double H[9], rt[3];
  so3_exp(half, H);
  so3_rotate(H, s + 9, rt);
  const double tr = s[9] * rot[0] + s[10] * rot[1] + s[11] * rot[2];
  const double f = th > 0.001 ? (tr * (1 - 2 * shtot)) / rr : tr / 24;
  for (int i = 0; i < 3; i++) rt[i] -= rot[i] * f;
  for (int i = 0; i < 3; i++) out6[i] = rt[i] / (2 * shtot);
  for (int i = 0; i < 3; i++) out6[3 + i] = rot[i];
}

// ------------------------------------------------------------------------------------------
// Square-root-free LDL^T (TooN::Cholesky semantics: no pivoting, no failure path) for small
// fixed N, fully unrolled.  A row-major N x N, lower triangle read.
// ------------------------------------------------------------------------------------------
template <int N>
PTAM_HD void ldlt_factor(double* A) {
#pragma unroll
  for (int col = 0; col < N; col++) {
    double inv_diag = 1;
#pragma unroll
    for (int row = col; row < N; row++) {
      double val = A[row * N + col];
#pragma unroll
      for (int c2 = 0; c2 < col; c2++) val -= A[c2 * N + col] * A[row * N + c2];
      if (row == col) {
        A[row * N + col] = val;
        inv_diag = 1 / val;  // val == 0 gives inf/NaN downstream, as TooN does not signal failure
      } else {
        A[col * N + row] = val;
        A[row * N + col] = val * inv_diag;
      }
    }
  }
}
template <int N>
PTAM_HD void ldlt_backsub(const double* A, const double* b, double* x) {
  double y[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double v = b[i];
#pragma unroll
    for (int j = 0; j < i; j++) v -= A[i * N + j] * y[j];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < N; i++) y[i] /= A[i * N + i];
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double v = y[i];
#pragma unroll
    for (int j = i + 1; j < N; j++) v -= A[j * N + i] * x[j];
    x[i] = v;
  }
}
template <int N>
PTAM_HD void ldlt_inverse(double* A, double* inv) {
  ldlt_factor<N>(A);
#pragma unroll
  for (int c = 0; c < N; c++) {
    double e[N], x[N];
#pragma unroll
    for (int i = 0; i < N; i++) e[i] = (i == c) ? 1.0 : 0.0;
    ldlt_backsub<N>(A, e, x);
#pragma unroll
    for (int i = 0; i < N; i++) inv[i * N + c] = x[i];
  }
}

// ------------------------------------------------------------------------------------------
// M-estimators — reference include/Tools.h:128-254.  est: 0 Tukey, 1 Cauchy, 2 Huber.
// ------------------------------------------------------------------------------------------
PTAM_HD double mest_sigma_from_median(double med, long long n, int est) {
  double sigma = 1.4826 * (1 + 5.0 / (double)(unsigned long long)(n * 2 - 6)) * sqrt(med);
  sigma = (est == 2 ? 1.345 : 4.6851) * sigma;
  return sigma * sigma;
}
PTAM_HD double mest_sqrt_weight(double e2, double s2, int est) {
  if (est == 0) return e2 > s2 ? 0.0 : 1.0 - (e2 / s2);
  if (est == 1) return sqrt(1.0 / (1.0 + e2 / s2));
  return sqrt(e2 < s2 ? 1.0 : sqrt(s2 / e2));
}
PTAM_HD double mest_weight(double e2, double s2, int est) {
  if (est == 0) { const double d = e2 > s2 ? 0.0 : 1.0 - (e2 / s2); return d * d; }
  if (est == 1) return 1.0 / (1.0 + e2 / s2);
  return e2 < s2 ? 1.0 : sqrt(s2 / e2);
}
PTAM_HD double mest_objective(double e2, double s2, int est) {
  if (est == 0) {
    if (e2 > s2) return 1.0;
    const double d = 1.0 - e2 / s2;
    return 1.0 - d * d * d;
  }
  if (est == 1) return log(1.0 + e2 / s2);
  if (e2 < s2) return 0.5 * e2;
  const double s = sqrt(s2), e = sqrt(e2);
  return s * (e - 0.5 * s);
}

// ------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------
PTAM_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
PTAM_DEV int warp_sum_int(int v) { return __reduce_add_sync(kFull, v); }

}  // namespace ptam

// host-side error plumbing shared by tracker.cu / bundle.cu
#define PTAM_CUDA_TRY(self, expr)                                                              \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (self)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return PTAM_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)
