// sm_100a kernels of path B (Bundle::Compute, reference src/Bundle.cc:209-551).
//
// per LM step      k_ba_project      ProjectAndFindSquaredError per measurement (Bundle.cc:219-225)
//                  k_ba_select       exact order statistic of the squared errors (sigma^2, :230-237)
//                  k_ba_jacobian     weights, A (2x6), B (2x3), W = A^T B, warp-aggregated f64 atomics
//                                    into U / epsA (per camera) and V / epsB (per point) (:251-332)
// per lambda trial k_ba_vinv         V*_i^-1 by 3x3 LDL^T (:341-359)
//                  k_ba_init_s       S diagonal blocks <- U*, vE <- epsA (:374-392)
//                  k_ba_schur        warp per point: S_jk -= W_ij V*^-1 W_ik^T, vE_j -= W_ij V*^-1 epsB
//                                    (:396-446) with atomics into the dense lower triangle
//                  k_ldlt_*          blocked square-root-free LDL^T of S (DMMA trailing update) + solve (:457-458)
//                  k_ba_point_update delta_b_i (:461-483), k_ba_cam_update exp(delta_a) (:496-504)
//                  k_ba_new_error    FindNewError (:188-207)
#pragma once
#include "common.cuh"
#include "../../include/ptam_b200.h"

namespace ptam {

enum : int { M_ALIVE = 0, M_BAD = 1, M_ERASED = 2 };

struct BundleDev {
  CamModel cam;
  int n_cams, n_pts, n_meas, n;  // n = 6 * non-fixed cameras; n_meas = measurements held by THIS shard
  int est;
  int p_lo, p_hi;        // points owned by this shard [p_lo, p_hi); everything for world == 1
  int add_cam_update;    // 1 on the rank that contributes the (replicated) camera part of |delta|^2
  // cameras
  double* cam_se3;      // [C][12]
  double* cam_se3_new;  // [C][12]
  const int* cam_fixed; // [C]
  const int* cam_row;   // [C] start row or -1
  double* U;            // [C][21] lower triangle, row-major packed
  double* epsA;         // [C][6]
  // points
  double* pt_pos;       // [P][3]
  double* pt_pos_new;   // [P][3]
  double* V;            // [P][6] lower triangle packed (00,10,11,20,21,22)
  double* epsB;         // [P][3]
  double* Vinv;         // [P][9]
  double* Ve;           // [P][3]  V*^-1 epsB
  const int* pt_off;    // [P+1] CSR by point (measurement ids sorted by camera id)
  const int* pt_meas;   // [M]
  // measurements (insertion order)
  const int* m_cam; const int* m_pt;
  const double* m_found;  // [M][2]
  const double* m_sin;    // [M] dSqrtInvNoise
  int* m_state;           // [M]
  double* m_v3cam;        // [M][3]
  double* m_derivs;       // [M][4]
  double* m_eps;          // [M][2]
  double* m_e2;           // [M]
  double* m_W;            // [M][18]
  double* e2_compact;     // [M] squared errors of the non-bad measurements
  // reduced system
  double* S;   // [n][n] (lower triangle valid; mirrored on request)
  double* vE;  // [n]
  double* upd; // [n] camera update
  // scalars: 0 n_valid (as double), 1 sigma^2, 2 current error, 3 new error, 4 sum sq update,
  //          5 abort votes, 6 lambda, 7 median   (2..5 are the slots summed across shards)
  double* scal;
  int* hist16;                   // [2048] digit histogram of the distributed radix select
  unsigned long long* sel_state; // [0] key prefix found so far, [1] rank still to find inside it
  int* counters;  // 0 n_valid, 1 n_outliers_total, 2 n_bad_this_step
  int* outliers;  // [M][2] (point, camera) in erase order
  int* m_erase_step;  // [M] LM step (1-based) at which the measurement was erased, 0 = still in the graph
};

PTAM_DEV double block_sum(double v, double* sh /*32*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_project(BundleDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.n_meas) return;
  if (d.m_state[m] == M_ERASED) return;
  const int c = d.m_cam[m], p = d.m_pt[m];
  double v3[3];
  se3_apply(d.cam_se3 + 12 * c, d.pt_pos + 3 * p, v3);
  d.m_v3cam[3 * m] = v3[0]; d.m_v3cam[3 * m + 1] = v3[1]; d.m_v3cam[3 * m + 2] = v3[2];
  if (v3[2] <= 0) { d.m_state[m] = M_BAD; return; }
  d.m_state[m] = M_ALIVE;
  const CamProj q = cam_project(d.cam, v3[0] / v3[2], v3[1] / v3[2]);
  double dv[4];
  cam_derivs(d.cam, q, dv);
  for (int k = 0; k < 4; k++) d.m_derivs[4 * m + k] = dv[k];
  const double s = d.m_sin[m];
  const double e0 = s * (d.m_found[2 * m] - q.im[0]), e1 = s * (d.m_found[2 * m + 1] - q.im[1]);
  d.m_eps[2 * m] = e0; d.m_eps[2 * m + 1] = e1;
  d.m_e2[m] = e0 * e0 + e1 * e1;
}

// compaction of the valid squared errors (order irrelevant for an order statistic)
__global__ void __launch_bounds__(256) k_ba_gather_e2(BundleDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = m < d.n_meas && d.m_state[m] == M_ALIVE;
  const unsigned b = __ballot_sync(kFull, ok);
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0 && b) base = atomicAdd(&d.counters[0], __popc(b));
  base = __shfl_sync(kFull, base, 0);
  if (ok) d.e2_compact[base + __popc(b & ((1u << lane) - 1))] = d.m_e2[m];
}

// sigma^2 = MEstimator::FindSigmaSquared (Tools.h:152-162), clamped to MinTukeySigma^2 (Bundle.cc:234-237).
// One CTA: MSB-first radix select (8 bits per pass) of element n/2 on the IEEE bit patterns.
__global__ void __launch_bounds__(1024) k_ba_select(BundleDev d, double min_sigma_sq) {
  __shared__ int hist[256];
  __shared__ unsigned long long prefix;
  __shared__ int kk;
  const int n = d.counters[0];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { prefix = 0ull; kk = n / 2; }
  for (int pass = 0; pass < 8 && n > 0; pass++) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long pf = prefix;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(d.e2_compact[i]);
      if (pass == 0 || (key >> (shift + 8)) == (pf >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (warp == 0) {
      int c[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { c[q] = hist[8 * lane + q]; tot += c[q]; }
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      const int k0 = kk;
      __syncwarp();  // every lane has read kk before the one owner rewrites it
      const int excl = inc - tot;
      if (k0 >= excl && k0 < inc) {
        int r = k0 - excl, q = 0;
        while (q < 7 && r >= c[q]) { r -= c[q]; q++; }
        kk = r;
        prefix = pf | ((unsigned long long)(8 * lane + q) << shift);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double med = __longlong_as_double((long long)prefix);
    double s2 = mest_sigma_from_median(med, n, d.est);
    if (s2 < min_sigma_sq) s2 = min_sigma_sq;
    d.scal[7] = med;
    d.scal[1] = s2;
    d.scal[0] = (double)n;
  }
}

// ---------------------------------------------------------------------------------------------
// Exact order statistic for large or sharded problems: MSB-first radix select over the IEEE bit
// patterns, six passes with digit widths 11,11,11,11,11,9.  Every pass: k_ba_hist (all CTAs: digits
// aggregated per warp with match.any, per-CTA histogram in shared memory, non-zero bins flushed with
// global atomics) -> [all-reduce of the 2048 bins across shards] -> k_ba_pick (one CTA finds the bin
// holding the wanted rank).  After the last pass the prefix IS the bit pattern of the floor(n/2)-th
// smallest squared error (Tools.h:152-162 sorts and takes element n/2), identical on every shard.
// ---------------------------------------------------------------------------------------------
constexpr int kSelPasses = 6;
constexpr int kSelBins = 2048;
PTAM_HD int sel_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
PTAM_HD int sel_width(int pass) { return pass < 5 ? 11 : 9; }

__global__ void __launch_bounds__(256) k_ba_hist(BundleDev d, int pass) {
  __shared__ int h[kSelBins];
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int shift = sel_shift(pass), width = sel_width(pass);
  const unsigned long long prefix = d.sel_state[0];
  const int lane = threadIdx.x & 31;
  for (int m0 = blockIdx.x * blockDim.x; m0 < d.n_meas; m0 += gridDim.x * blockDim.x) {
    const int m = m0 + threadIdx.x;
    bool take = false;
    unsigned digit = 0;
    if (m < d.n_meas && d.m_state[m] == M_ALIVE) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(d.m_e2[m]);
      take = pass == 0 || (key >> (shift + width)) == (prefix >> (shift + width));
      digit = (unsigned)(key >> shift) & ((1u << width) - 1u);
    }
    const unsigned active = __ballot_sync(kFull, take);
    if (take) {
      const unsigned same = __match_any_sync(active, digit);
      if (lane == __ffs(same) - 1) atomicAdd(&h[digit], __popc(same));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
    if (h[i]) atomicAdd(&d.hist16[i], h[i]);
}

__global__ void __launch_bounds__(1024) k_ba_pick(BundleDev d, int pass, double min_sigma_sq) {
  __shared__ int wsum[32];
  __shared__ int total_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int shift = sel_shift(pass);
  const int c0 = d.hist16[2 * t], c1 = d.hist16[2 * t + 1];
  const int tot = c0 + c1;
  int inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int ws = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, ws, o);
      if (lane >= o) ws += v;
    }
    wsum[lane] = ws;
    if (lane == 31) total_s = ws;
  }
  __syncthreads();
  const int n_all = total_s;  // pass 0: number of valid measurements over all shards
  const long long kk = pass == 0 ? (long long)(n_all / 2) : (long long)d.sel_state[1];
  const int excl = (warp ? wsum[warp - 1] : 0) + inc - tot;
  if (n_all > 0 && kk >= excl && kk < excl + tot) {  // exactly one thread
    int r = (int)(kk - excl), q = 0;
    if (r >= c0) { r -= c0; q = 1; }
    const unsigned long long prefix = (pass == 0 ? 0ull : d.sel_state[0]) | ((unsigned long long)(2 * t + q) << shift);
    d.sel_state[0] = prefix;
    d.sel_state[1] = (unsigned long long)r;
    if (pass == kSelPasses - 1) {
      const double med = __longlong_as_double((long long)prefix);
      const long long n = (long long)d.scal[0];
      double s2 = mest_sigma_from_median(med, n, d.est);
      if (s2 < min_sigma_sq) s2 = min_sigma_sq;
      d.scal[7] = med;
      d.scal[1] = s2;
    }
  }
  if (pass == 0 && t == 0) {
    d.scal[0] = (double)n_all;
    d.counters[0] = n_all;
    if (n_all == 0) {  // no valid measurement anywhere: same result as the single-CTA select on n = 0
      double s2 = mest_sigma_from_median(0.0, 0, d.est);
      if (s2 < min_sigma_sq) s2 = min_sigma_sq;
      d.scal[7] = 0.0; d.scal[1] = s2;
      d.sel_state[0] = 0ull; d.sel_state[1] = 0ull;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_ba_jacobian — one thread per observation.  Camera accumulators: measurements usually arrive
// camera-major (MapMaker.cc:871-882), so a whole warp mostly shares one camera: warp-reduce the 27
// values and issue one atomic per value; otherwise per-lane atomics.  Point accumulators: per-lane
// atomics (a point's few observations are scattered over warps).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ba_jacobian(BundleDev d) {
  __shared__ double sh[32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const double sigma2 = d.scal[1];
  double err = 0.0;
  bool active = false;
  int c = -1;
  double A[12], eps[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < 12; i++) A[i] = 0;
  bool cam_free = false;
  if (m < d.n_meas && d.m_state[m] != M_ERASED) {
    if (d.m_state[m] == M_BAD) err = 1.0;
    else {
      const double e2 = d.m_e2[m];
      const double w = mest_sqrt_weight(e2, sigma2, d.est);
      eps[0] = w * d.m_eps[2 * m]; eps[1] = w * d.m_eps[2 * m + 1];
      d.m_eps[2 * m] = eps[0]; d.m_eps[2 * m + 1] = eps[1];
      if (w == 0) { d.m_state[m] = M_BAD; err = 1.0; }
      else {
        active = true;
        err = mest_objective(e2, sigma2, d.est);
        c = d.m_cam[m];
        const int p = d.m_pt[m];
        const double s = d.m_sin[m];
        const double d0 = s * (w * d.m_derivs[4 * m]), d1 = s * (w * d.m_derivs[4 * m + 1]);
        const double d2 = s * (w * d.m_derivs[4 * m + 2]), d3 = s * (w * d.m_derivs[4 * m + 3]);
        const double X = d.m_v3cam[3 * m], Y = d.m_v3cam[3 * m + 1], Z = d.m_v3cam[3 * m + 2];
        const double ooz = 1.0 / Z;
        cam_free = !d.cam_fixed[c];
        if (cam_free) {
          const double gx[6] = {1, 0, 0, 0, Z, -Y}, gy[6] = {0, 1, 0, -Z, 0, X}, gz[6] = {0, 0, 1, Y, -X, 0};
#pragma unroll
          for (int q = 0; q < 6; q++) {
            const double a0 = (gx[q] - X * gz[q] * ooz) * ooz, a1 = (gy[q] - Y * gz[q] * ooz) * ooz;
            A[q] = d0 * a0 + d1 * a1;
            A[6 + q] = d2 * a0 + d3 * a1;
          }
        }
        double B[6];
        const double* R = d.cam_se3 + 12 * c;
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const double a0 = (R[q] - X * R[6 + q] * ooz) * ooz, a1 = (R[3 + q] - Y * R[6 + q] * ooz) * ooz;
          B[q] = d0 * a0 + d1 * a1;
          B[3 + q] = d2 * a0 + d3 * a1;
        }
        // V (lower, packed) and epsB: per-lane atomics
        double* Vp = d.V + 6 * p;
        int o = 0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int cc = 0; cc <= r; cc++) atomicAdd(&Vp[o++], B[r] * B[cc] + B[3 + r] * B[3 + cc]);
#pragma unroll
        for (int r = 0; r < 3; r++) atomicAdd(&d.epsB[3 * p + r], B[r] * eps[0] + B[3 + r] * eps[1]);
        // W = A^T B (6x3), zero for a fixed camera
        double* Wm = d.m_W + 18 * m;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) Wm[3 * r + cc] = A[r] * B[cc] + A[6 + r] * B[3 + cc];
      }
    }
  }
  // camera accumulators U (lower, packed 21) and epsA (6)
  // Lanes without a contribution (erased / outlier measurements, fixed cameras: A and eps are zero there)
  // ride along in the warp sums: the warp is uniform when all CONTRIBUTING lanes share one camera, which
  // leaves only the ~3 % of warps that straddle a camera boundary on the per-lane path.
  const int c_acc = (active && cam_free) ? c : -1;
  if (c_acc < 0) { eps[0] = 0.0; eps[1] = 0.0; }  // 0 x inf of a wild outlier must not reach the sums
  const unsigned contrib = __ballot_sync(kFull, c_acc >= 0);
  const int c0 = contrib ? __shfl_sync(kFull, c_acc, __ffs(contrib) - 1) : -1;
  const bool uniform = __all_sync(kFull, c_acc < 0 || c_acc == c0);
  if (uniform) {
    if (c0 >= 0) {
      int o = 0;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = 0; cc <= r; cc++) {
          const double v = warp_sum(A[r] * A[cc] + A[6 + r] * A[6 + cc]);
          if (lane == 0) atomicAdd(&d.U[21 * c0 + o], v);
          o++;
        }
#pragma unroll
      for (int r = 0; r < 6; r++) {
        const double v = warp_sum(A[r] * eps[0] + A[6 + r] * eps[1]);
        if (lane == 0) atomicAdd(&d.epsA[6 * c0 + r], v);
      }
    }
  } else if (c_acc >= 0) {
    int o = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc <= r; cc++) atomicAdd(&d.U[21 * c_acc + o++], A[r] * A[cc] + A[6 + r] * A[6 + cc]);
#pragma unroll
    for (int r = 0; r < 6; r++) atomicAdd(&d.epsA[6 * c_acc + r], A[r] * eps[0] + A[6 + r] * eps[1]);
  }
  const double t = block_sum(err, sh);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(&d.scal[2], t);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_vinv(BundleDev d) {
  const int i = d.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.p_hi) return;
  const double lambda = d.scal[6];
  const double* v = d.V + 6 * i;
  double Vs[9] = {v[0], v[1], v[3], v[1], v[2], v[4], v[3], v[4], v[5]};
  double inv[9];
  if (Vs[0] * Vs[4] * Vs[8] == 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) inv[k] = 0;
  } else {
    Vs[0] *= (1.0 + lambda); Vs[4] *= (1.0 + lambda); Vs[8] *= (1.0 + lambda);
    ldlt_inverse<3>(Vs, inv);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) d.Vinv[9 * i + k] = inv[k];
  const double* e = d.epsB + 3 * i;
#pragma unroll
  for (int r = 0; r < 3; r++) d.Ve[3 * i + r] = inv[3 * r] * e[0] + inv[3 * r + 1] * e[1] + inv[3 * r + 2] * e[2];
}

// S <- 0 except diagonal blocks U*_j (lambda-damped, both triangles); vE <- epsA
__global__ void __launch_bounds__(256) k_ba_init_s(BundleDev d) {
  const size_t tot = (size_t)d.n * d.n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) d.S[i] = 0.0;
}
__global__ void __launch_bounds__(64) k_ba_init_diag(BundleDev d) {
  const int c = blockIdx.x;
  const int row = d.cam_row[c];
  if (row < 0) return;
  const double lambda = d.scal[6];
  const int t = threadIdx.x;
  if (t < 36) {
    const int r = t / 6, cc = t % 6;
    const int a = r >= cc ? r : cc, b = r >= cc ? cc : r;
    double v = d.U[21 * c + a * (a + 1) / 2 + b];
    if (r == cc) v *= (1.0 + lambda);
    d.S[(size_t)(row + r) * d.n + row + cc] = v;
  } else if (t < 42) d.vE[row + t - 36] = d.epsA[6 * c + t - 36];
}

// Schur complement (Bundle.cc:365-453), two kernels.
// k_ba_schur_diag — the diagonal blocks and vE: one thread per observation in list order,
//   S_jj -= W_ij V*_i^-1 W_ij^T (lower triangle, 21 values),  vE_j -= W_ij V*_i^-1 epsB_i.
// The list is usually camera-major, so a warp mostly shares one camera: the 27 values are summed over the
// warp and added with one atomic each (as k_ba_jacobian does for U / epsA) instead of 1 200 atomics per
// address and camera at C4; warps that straddle a camera boundary use per-lane atomics.
__global__ void __launch_bounds__(128) k_ba_schur_diag(BundleDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int jrow = -1;
  double D[21], e[6];
#pragma unroll
  for (int q = 0; q < 21; q++) D[q] = 0.0;
#pragma unroll
  for (int q = 0; q < 6; q++) e[q] = 0.0;
  if (m < d.n_meas && d.m_state[m] == M_ALIVE) {
    jrow = d.cam_row[d.m_cam[m]];
    if (jrow >= 0) {
      const int i = d.m_pt[m];
      const double* Vi = d.Vinv + 9 * (size_t)i;
      const double* ve = d.Ve + 3 * (size_t)i;
      const double* W = d.m_W + 18 * (size_t)m;
      double Wr[18], WV[18];
#pragma unroll
      for (int q = 0; q < 18; q++) Wr[q] = W[q];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) WV[3 * r + c] = Wr[3 * r] * Vi[c] + Wr[3 * r + 1] * Vi[3 + c] + Wr[3 * r + 2] * Vi[6 + c];
      int o = 0;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) D[o++] = -(WV[3 * r] * Wr[3 * c] + WV[3 * r + 1] * Wr[3 * c + 1] + WV[3 * r + 2] * Wr[3 * c + 2]);
#pragma unroll
      for (int r = 0; r < 6; r++) e[r] = -(Wr[3 * r] * ve[0] + Wr[3 * r + 1] * ve[1] + Wr[3 * r + 2] * ve[2]);
    }
  }
  const unsigned contrib = __ballot_sync(kFull, jrow >= 0);
  if (!contrib) return;
  const int j0 = __shfl_sync(kFull, jrow, __ffs(contrib) - 1);
  const bool uniform = __all_sync(kFull, jrow < 0 || jrow == j0);
  if (uniform) {
    int o = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c <= r; c++) {
        const double v = warp_sum(D[o++]);
        if (lane == 0) atomicAdd(&d.S[(size_t)(j0 + r) * d.n + j0 + c], v);
      }
#pragma unroll
    for (int r = 0; r < 6; r++) {
      const double v = warp_sum(e[r]);
      if (lane == 0) atomicAdd(&d.vE[j0 + r], v);
    }
  } else if (jrow >= 0) {
    int o = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c <= r; c++) atomicAdd(&d.S[(size_t)(jrow + r) * d.n + jrow + c], D[o++]);
#pragma unroll
    for (int r = 0; r < 6; r++) atomicAdd(&d.vE[jrow + r], e[r]);
  }
}

// k_ba_schur — the off-diagonal blocks: warp per point, lane per camera pair (j > k) observing it, both
// cameras free, both measurements good:  S_jk -= W_ij V*_i^-1 W_ik^T  (36 f64 atomics into the lower S).
__global__ void __launch_bounds__(256) k_ba_schur(BundleDev d) {
  const int i = d.p_lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= d.p_hi) return;
  const int lane = threadIdx.x & 31;
  const int o0 = d.pt_off[i], k = d.pt_off[i + 1] - o0;
  if (k < 2) return;
  double Vi[9];
#pragma unroll
  for (int q = 0; q < 9; q++) Vi[q] = d.Vinv[9 * i + q];
  const int npairs = k * (k - 1) / 2;
  for (int pr = lane; pr < npairs; pr += 32) {
    // pr -> (a, b), b < a:  pr = a (a - 1) / 2 + b
    int a = (int)((sqrt(8.0 * pr + 1.0) + 1.0) * 0.5);
    while (a * (a - 1) / 2 > pr) a--;
    while ((a + 1) * a / 2 <= pr) a++;
    const int b = pr - a * (a - 1) / 2;
    const int mj = d.pt_meas[o0 + a], mk = d.pt_meas[o0 + b];
    if (d.m_state[mj] != M_ALIVE || d.m_state[mk] != M_ALIVE) continue;
    const int jrow = d.cam_row[d.m_cam[mj]], krow = d.cam_row[d.m_cam[mk]];
    if (jrow < 0 || krow < 0) continue;
    const double* Wj = d.m_W + 18 * mj;
    const double* Wk = d.m_W + 18 * mk;
    double WV[18];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) WV[3 * r + c] = Wj[3 * r] * Vi[c] + Wj[3 * r + 1] * Vi[3 + c] + Wj[3 * r + 2] * Vi[6 + c];
    double* Sb = d.S + (size_t)jrow * d.n + krow;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double v = WV[3 * r] * Wk[3 * c] + WV[3 * r + 1] * Wk[3 * c + 1] + WV[3 * r + 2] * Wk[3 * c + 2];
        atomicAdd(&Sb[(size_t)r * d.n + c], -v);
      }
  }
}

// mirror lower -> upper (Bundle.cc:451-453); only needed when S is exported
__global__ void __launch_bounds__(256) k_ba_mirror(double* S, int n) {
  const size_t tot = (size_t)n * n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t % n);
    if (j > i) S[t] = S[(size_t)j * n + i];
  }
}

// ---------------------------------------------------------------------------------------------
// Dense solve  delta_a = S^-1 vE  (Bundle.cc:457-458, TooN Cholesky<>: square-root-free LDL^T, no
// pivoting, no failure path — a non-positive-definite S yields inf/NaN exactly as in the reference).
// Right-looking blocked factorisation, panel width 64, with the forward substitution folded in:
//   k_ldlt_panel   every CTA applies the previous panel's pending update to the panel's 64 columns (its
//                  diagonal block and its own 64 rows), factors the 64x64 diagonal block in shared memory
//                  (redundantly: it saves a launch and a dependency), then solves its 64 rows of the panel,
//                  four threads per row:  w = a L11^-T  (= L21 D1),  L21 = w D1^-1, and applies the panel
//                  to the right-hand side:  y2 -= L21 y1.  Operands arrive by cp.async.bulk + mbarrier.
//   k_ldlt_update  trailing update  A22 -= (L21 D1) L21^T  on the lower triangle, 128x64 tiles,
//                  K = 64: the one genuine dense contraction of either hot path.  FP64 has no
//                  tcgen05 form, so it runs on the f64 tensor pipe (DMMA, mma.sync m8n8k4);
//                  operand tiles are staged in shared memory by the TMA engine (one 512-byte
//                  cp.async.bulk per row, completion on an mbarrier), rows padded to 68 doubles so
//                  that fragment loads are bank-conflict free.
//   k_ldlt_step    panel k and the tail of panel k-1's trailing update in one grid (the late, latency-bound
//                  part of the factorisation as back-to-back launches on one stream).
//   k_ldlt_back    z = D^-1 y (k_ldlt_scale),  L^T x = z: all panels in one launch by an 8-CTA cluster.
// ---------------------------------------------------------------------------------------------
constexpr int kNB = 64;     // panel width
constexpr int kUTM = 128;   // trailing-update tile rows
constexpr int kUTN = 64;    // trailing-update tile columns
constexpr int kLds = kNB + 4;  // padded shared-memory row, doubles
constexpr int kUpdateSmem = (kUTM + kUTN) * kLds * (int)sizeof(double) + 16;

// LDL^T of a 64x64 block by 256 threads, eight sub-panels of eight columns.  The trailing matrix lives in
// registers (thread (ty = tid / 16, tx = tid % 16) owns rows ty + 16 i x columns tx + 16 j, i, j < 4); the
// current 64 x 8 sub-panel is handled by threads 0..63, one row each.  Every row thread factors the 8x8
// diagonal block of the sub-panel REDUNDANTLY in its own registers (36 broadcast loads), so that pivots,
// reciprocals and the L D values of the pivot rows need no exchange: the serial chain per pivot is
// reciprocal -> multiply -> FMA, with the thread's own row riding along.  Then all warps apply the rank-8
// update to their register tiles from shared memory (L in `a`, L D in `us`) and the owners of the next
// eight columns hand them over: two block barriers per sub-panel, 16 per block instead of 64.
// The right-hand side rides along with the row threads (y_r -= l_r y_col: the forward substitution L y' = y).
// L = value * (1 / d) as TooN's Cholesky does, subtractions in ascending column order as in its
// left-looking loop.  Result: `a` holds L (strict lower) and D (diagonal), `ysh` the forward-substituted
// right-hand side, `dinv` the reciprocals of D.  Entries above the diagonal are scratch.
constexpr int kPanelThreads = 256;
constexpr int kFuseTailTiles = 576;  // tails of at most this many 128x64 tiles ride in the next panel's launch (k_ldlt_step)
constexpr int kPanelRows = 64;  // rows of the panel solved per CTA (four threads per row; more CTAs beat fuller CTAs here)
#ifdef PTAM_PANEL_DEBUG
__device__ long long g_dbg[8];
#define DBG_T(k) if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now = clock64(); atomicAdd((unsigned long long*)&g_dbg[k], (unsigned long long)(now - t_prev)); t_prev = now; }
#else
#define DBG_T(k)
#endif
constexpr int kLda = kNB + 2;  // even row pitch: (row, even column) pairs are 16-byte aligned

PTAM_DEV void block_ldlt64(double (*a)[kLda], double (*us)[8], double* dinv, double* ysh, int nb) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  double ar[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) ar[i][j] = a[ty + 16 * i][tx + 16 * j];
  double yr = tid < kNB ? ysh[tid] : 0.0;  // threads 0..63 own one row of the current sub-panel each
  if (tid < kNB) dinv[tid] = 1.0;
  __syncthreads();
#pragma unroll 1
  for (int c0 = 0; c0 < kNB; c0 += 8) {
    if (c0 >= nb) break;  // the identity padding of a short last block needs no work
    if (tid < kNB) {
      const int r = tid;
      // the 8x8 diagonal block of the sub-panel and its right-hand side, redundantly in every row thread
      // (broadcast loads): pivots, reciprocals and the L D values then need no exchange at all
      double dg[8][8], yv[8], pv[8], uv[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) dg[i][j] = a[c0 + i][c0 + j];
        yv[i] = ysh[c0 + i];
      }
      {
        const double2* row = reinterpret_cast<const double2*>(&a[r][c0]);
#pragma unroll
        for (int j = 0; j < 8; j += 2) { const double2 v = row[j >> 1]; pv[j] = v.x; pv[j + 1] = v.y; }
      }
      // the two row warps have read the diagonal block before its owners overwrite it below
      asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double rcp = 1.0 / dg[j][j];  // (a MUFU seed + two Newton steps was measured slower: 15.6 -> 16.6 us per block)
        if (r == c0 + j) dinv[c0 + j] = rcp;
        // own row (rows of the finished part and the pivot row itself stay as they are)
        const bool below = r > c0 + j;
        const double v = pv[j];
        const double l = below ? v * rcp : 0.0;
#pragma unroll
        for (int q = j + 1; q < 8; q++) pv[q] -= l * dg[q][j];  // dg[q][j] still holds L D of row c0 + q
        yr -= l * yv[j];
        uv[j] = below ? v : 0.0;
        if (below) pv[j] = l;
        // the diagonal block itself
#pragma unroll
        for (int i = j + 1; i < 8; i++) {
          const double li = dg[i][j] * rcp;
#pragma unroll
          for (int q = j + 1; q <= i; q++) dg[i][q] -= li * dg[q][j];
          yv[i] -= li * yv[j];
        }
      }
      if (r >= c0) {
        double2* row = reinterpret_cast<double2*>(&a[r][c0]);
#pragma unroll
        for (int j = 0; j < 8; j += 2) row[j >> 1] = make_double2(pv[j], pv[j + 1]);
      }
      double2* urow = reinterpret_cast<double2*>(&us[r][0]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) urow[j >> 1] = make_double2(uv[j], uv[j + 1]);
      ysh[r] = yr;
    }
    __syncthreads();
    const int t0 = c0 + 8;  // first row / column of the trailing matrix
    if (t0 < kNB) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (16 * i + 15 < t0) continue;
        double l[8];
        {
          const double2* row = reinterpret_cast<const double2*>(&a[ty + 16 * i][c0]);
#pragma unroll
          for (int q = 0; q < 8; q += 2) { const double2 v = row[q >> 1]; l[q] = v.x; l[q + 1] = v.y; }
        }
        if (ty + 16 * i < t0) {  // rows of the finished sub-panels: their entries here are scratch, keep them finite
#pragma unroll
          for (int q = 0; q < 8; q++) l[q] = 0.0;
        }
#pragma unroll
        for (int j = 0; j <= i; j++) {
          if (16 * j + 15 < t0) continue;
          const double2* urow = reinterpret_cast<const double2*>(&us[tx + 16 * j][0]);
          double acc = ar[i][j];
#pragma unroll
          for (int q = 0; q < 8; q += 2) { const double2 v = urow[q >> 1]; acc -= l[q] * v.x; acc -= l[q + 1] * v.y; }
          ar[i][j] = acc;
        }
      }
      // the owners of the next eight columns hand them to warp 0 (rows >= t0)
      const int jn = t0 >> 4;
      if ((tx >> 3) == ((t0 >> 3) & 1)) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int r = ty + 16 * i;
          if (r >= t0) {
            // ar[i][jn] with a run-time jn: select without dynamic register indexing
            const double v = jn == 0 ? ar[i][0] : jn == 1 ? ar[i][1] : jn == 2 ? ar[i][2] : ar[i][3];
            a[r][tx + 16 * jn] = v;
          }
        }
      }
    }
    __syncthreads();
  }
}

// Shared memory of k_ldlt_panel (dynamic): the diagonal block, the rank-8 operand, the right-hand side and
// the reciprocals, plus three 64x64 operands of the PENDING update (see below); the third is reused for
// the updated rows of this CTA.
constexpr int kPanelSmem = (4 * kNB * kLda + kNB * 8 + 2 * kNB) * (int)sizeof(double) + 16;  // + the mbarrier of the bulk loads

// Panel k, fused with the head of panel k-1's trailing update.  The columns of panel k still miss the
// contribution of panel k-1 (the tail kernel of panel k-1 only covers the column blocks from k+1 on), so
// every CTA first applies it itself:  A[rows][cols k] += Wp_prev[rows] L_head^T  for the diagonal block
// (all CTAs, redundantly, like the factorisation) and for its own 64 rows, 4x4 register tiles over K = 64.
// That makes the chain one kernel per panel instead of panel -> head update -> panel.
PTAM_DEV void ldlt_panel_body(double* A, double* Wp, const double* Wprev, double* y, int n, int k0, int block) {
  extern __shared__ __align__(16) unsigned char panel_smem[];
  double (*a)[kLda] = reinterpret_cast<double (*)[kLda]>(panel_smem);
  double (*lh)[kLda] = a + kNB;    // L of the diagonal block's rows in panel k-1's columns
  double (*wd)[kLda] = lh + kNB;   // Wp_prev rows of the diagonal block
  double (*wo)[kLda] = wd + kNB;   // Wp_prev rows of this CTA, then the updated rows themselves
  double (*us)[8] = reinterpret_cast<double (*)[8]>(wo + kNB);
  double* y1 = reinterpret_cast<double*>(us + kNB);
  double* dinv = y1 + kNB;
  const int nb = min(kNB, n - k0);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int row0 = k0 + nb + block * kPanelRows;
  const int rows_own = min(kPanelRows, n - row0);  // <= 0: the last panel's single CTA has no rows below the block
  const bool pend = Wprev != nullptr;
#ifdef PTAM_PANEL_DEBUG
  long long t_prev = clock64();
#endif
  if (nb == kNB) {
    // full panel: the four 64x64 operands arrive as 512-byte rows through the TMA engine (one cp.async.bulk
    // per row and thread, completion on an mbarrier) while the threads fetch their register tiles below.
    // `a` then also holds S's (finite, never used) values above the diagonal: block_ldlt64 treats them as scratch.
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(dinv + kNB);
    const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(mbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int rows_w = max(0, rows_own);
    if (tid == 0) {
      const unsigned bytes = (unsigned)(kNB + (pend ? 2 * kNB + rows_w : 0)) * kNB * (unsigned)sizeof(double);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
    }
    {
      const int which = tid >> 6, r = tid & 63;  // 0: a, 1: lh, 2: wd, 3: wo
      const double* src = nullptr;
      double* dst = which == 0 ? a[r] : which == 1 ? lh[r] : which == 2 ? wd[r] : wo[r];
      if (which == 0) src = A + (size_t)(k0 + r) * n + k0;
      else if (pend) {
        if (which == 1) src = A + (size_t)(k0 + r) * n + (k0 - kNB);
        else if (which == 2) src = Wprev + (size_t)(k0 + r) * kNB;
        else if (r < rows_w) src = Wprev + (size_t)(row0 + r) * kNB;
      }
      if (src) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"((unsigned)(kNB * sizeof(double))), "r"(mbar_a) : "memory");
      } else if (pend && which == 3) {
        for (int c = 0; c < kNB; c++) dst[c] = 0.0;  // rows past the matrix
      }
    }
    if (tid < kNB) y1[tid] = y[k0 + tid];
  } else {
    for (int i = tid; i < kNB * kNB; i += blockDim.x) {  // the short last panel (identity padding, no rows below it)
      const int r = i / kNB, c = i % kNB;
      a[r][c] = (r < nb && c <= r) ? A[(size_t)(k0 + r) * n + k0 + c] : (r == c ? 1.0 : 0.0);
      if (pend) {
        lh[r][c] = r < nb ? A[(size_t)(k0 + r) * n + (k0 - kNB) + c] : 0.0;
        wd[r][c] = r < nb ? Wprev[(size_t)(k0 + r) * kNB + c] : 0.0;
        wo[r][c] = r < rows_own ? Wprev[(size_t)(row0 + r) * kNB + c] : 0.0;
      }
    }
    if (tid < kNB) y1[tid] = tid < nb ? y[k0 + tid] : 0.0;
  }
  // this CTA's rows of the panel, as 4x4 register tiles (rows ty + 16 i, columns tx + 16 j)
  double co[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * i;
      co[i][j] = r < rows_own ? A[(size_t)(row0 + r) * n + k0 + tx + 16 * j] : 0.0;
    }
  if (nb == kNB) {
    const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(dinv + kNB);
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(mbar_a) : "memory");
  }
  __syncthreads();  // also covers the plain stores
  DBG_T(0)
  if (pend) {
    double cd[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) cd[i][j] = a[ty + 16 * i][tx + 16 * j];
#pragma unroll 2
    for (int q = 0; q < kNB; q += 2) {
      double2 vd[4], vo[4], vl[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        vd[i] = *reinterpret_cast<const double2*>(&wd[ty + 16 * i][q]);
        vo[i] = *reinterpret_cast<const double2*>(&wo[ty + 16 * i][q]);
        vl[i] = *reinterpret_cast<const double2*>(&lh[tx + 16 * i][q]);
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (j <= i) { cd[i][j] += vd[i].x * vl[j].x; cd[i][j] += vd[i].y * vl[j].y; }  // tiles above the diagonal are never stored
          co[i][j] += vo[i].x * vl[j].x; co[i][j] += vo[i].y * vl[j].y;
        }
    }
    __syncthreads();  // every thread is done with wd / wo / lh
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        if (r < nb && c <= r) a[r][c] = cd[i][j];  // the identity padding of a short last block stays
      }
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) wo[ty + 16 * i][tx + 16 * j] = co[i][j];
  __syncthreads();
  DBG_T(1)
  block_ldlt64(a, us, dinv, y1, nb);
  DBG_T(2)
  if (block == 0) {
    for (int i = tid; i < nb * nb; i += blockDim.x) {
      const int r = i / nb, c = i % nb;
      if (c <= r) A[(size_t)(k0 + r) * n + k0 + c] = a[r][c];
    }
    if (tid < nb) y[k0 + tid] = y1[tid];
  }
  DBG_T(3)
  // ---- rows below the block, w L11^T = a: FOUR threads per row (r = tid / 4), thread q = tid % 4 owns the 16
  // columns c = q (mod 4).  Right-looking, two columns per step (same expressions, same order per element as
  // a thread-per-row loop): the two pivots of the step travel by shuffle from their owners, then every thread
  // updates its own columns to the right with one 16-byte broadcast load of (L[c2][c], L[c2][c+1]) per
  // column.  All 256 threads work and a thread holds 16 values instead of 64 (6.0 -> 4.95 us per panel; a
  // four-columns-per-step variant with the 4x4 triangle solved redundantly was slower, 5.85 us).
  const int r = tid >> 2, q = tid & 3, lane = tid & 31;
  const int row = row0 + r;
  double x[kNB / 4];
#pragma unroll
  for (int j = 0; j < kNB / 4; j++) x[j] = wo[r][4 * j + q];
  DBG_T(4)
#pragma unroll
  for (int c = 0; c < kNB; c += 2) {
    const int jc = c >> 2;  // the owners of columns c and c + 1 hold them in x[jc]
    const double xc0 = __shfl_sync(kFull, x[jc], (lane & ~3) | (c & 3));
    if (q == ((c + 1) & 3)) x[jc] -= xc0 * a[c + 1][c];
    const double xc1 = __shfl_sync(kFull, x[jc], (lane & ~3) | ((c + 1) & 3));
#pragma unroll
    for (int j = jc; j < kNB / 4; j++) {
      const int c2 = 4 * j + q;
      if (j > jc || c2 > c + 1) {
        const double2 l = *reinterpret_cast<const double2*>(&a[c2][c]);
        x[j] -= xc0 * l.x + xc1 * l.y;
      }
    }
  }
  DBG_T(5)
  double dot = 0.0;
  if (row < n) {
    double* Ar = A + (size_t)row * n + k0;
    double* Wr = Wp + (size_t)row * kNB;  // holds -(L21 D1): the update kernel accumulates C += Wp L21^T
#pragma unroll
    for (int j = 0; j < kNB / 4; j++) {
      const int c = 4 * j + q;
      Wr[c] = -x[j];
      const double l = x[j] * dinv[c];  // value * (1 / d), as TooN does
      Ar[c] = l;
      dot += l * y1[c];
    }
  }
  dot += __shfl_xor_sync(kFull, dot, 1);
  dot += __shfl_xor_sync(kFull, dot, 2);
  if (q == 0 && row < n) y[row] -= dot;
  DBG_T(6)
}

__global__ void __launch_bounds__(kPanelThreads) k_ldlt_panel(double* A, double* Wp, const double* Wprev, double* y, int n, int k0) {
  ldlt_panel_body(A, Wp, Wprev, y, n, k0, (int)blockIdx.x);
}

PTAM_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// part 0: all tiles; part 1: only the first column block (the next panel's 64 columns: look-ahead
// head); part 2: everything else (look-ahead tail, runs on the second stream).
PTAM_DEV void ldlt_update_body(double* A, const double* Wp, int n, int k0, int part, int block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sW = reinterpret_cast<double*>(smem_raw);            // [kUTM][kLds]  -(L21 D1) rows of the i-tile
  double* sL = sW + kUTM * kLds;                                // [kUTN][kLds]  L21 rows of the j-tile
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sL + kUTN * kLds);
  const int r0 = k0 + kNB;
  int bi, bj;
  if (part == 1) { bi = block; bj = 0; }
  else if (part == 2) {
    // row block bi has column blocks bj = 1 .. 2 bi + 1: 2 bi + 1 tiles, bi^2 before it
    bi = (int)sqrt((double)block);
    while (bi * bi > block) bi--;
    while ((bi + 1) * (bi + 1) <= block) bi++;
    bj = block - bi * bi + 1;
  } else {
    // row block bi (128 rows) has column blocks bj = 0 .. 2 bi + 1 (64 columns): bi (bi + 1) tiles before it
    bi = (int)((sqrt(4.0 * block + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) > block) bi--;
    while ((bi + 1) * (bi + 2) <= block) bi++;
    bj = block - bi * (bi + 1);
  }
  const int i0 = r0 + bi * kUTM, j0 = r0 + bj * kUTN;
  if (j0 >= n) return;  // the last row block may be short of its second diagonal column block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned mbar_a = smem_u32(mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- TMA-engine staging: one 512-byte bulk copy per tile row (192 rows, threads 0..191)
  {
    const int rows_i = min(kUTM, n - i0), rows_j = min(kUTN, n - j0);
    if (tid == 0) {
      const unsigned bytes = (unsigned)(rows_i + rows_j) * kNB * (unsigned)sizeof(double);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
    }
    if (tid < kUTM + kUTN) {
      const double* src = nullptr;
      double* dst;
      if (tid < kUTM) { dst = sW + tid * kLds; if (tid < rows_i) src = Wp + (size_t)(i0 + tid) * kNB; }
      else { const int r = tid - kUTM; dst = sL + r * kLds; if (r < rows_j) src = A + (size_t)(j0 + r) * n + k0; }
      if (src) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"((unsigned)(kNB * sizeof(double))), "r"(mbar_a) : "memory");
      } else {
        for (int c = 0; c < kNB; c++) dst[c] = 0.0;  // rows past the matrix: defined operands, results never stored
      }
    }
  }
  // ---- accumulators start as the C tile (prefetched while the operand tiles are in flight)
  // 8 warps = 4 (32-row slabs) x 2 (32-column slabs); 4 x 4 m8n8k4 tiles per warp
  const int wm = warp & 3, wn = warp >> 2;
  const bool active = !(j0 + wn * 32 > i0 + wm * 32 + 31);  // warp tile not entirely above the diagonal
  double acc[4][4][2];
  if (active) {
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
      const int i = i0 + wm * 32 + mi * 8 + (lane >> 2);
      const double* Ci = A + (size_t)min(i, n - 1) * n;
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        const int j = j0 + wn * 32 + ni * 8 + 2 * (lane & 3);
        double2 v = make_double2(0.0, 0.0);
        if (i < n && j + 1 <= i) v = *reinterpret_cast<const double2*>(Ci + j);
        else if (i < n && j <= i) v.x = Ci[j];
        acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
      }
    }
  }
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(mbar_a) : "memory");
  }
  __syncthreads();  // also covers the plain zero-fill stores
  if (!active) return;
  const double* pa = sW + (wm * 32 + (lane >> 2)) * kLds + (lane & 3);
  const double* pb = sL + (wn * 32 + (lane >> 2)) * kLds + (lane & 3);
#pragma unroll 4
  for (int kk = 0; kk < kNB; kk += 4) {
    double fa[4], fb[4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) fa[mi] = pa[mi * 8 * kLds + kk];
#pragma unroll
    for (int ni = 0; ni < 4; ni++) fb[ni] = pb[ni * 8 * kLds + kk];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
      for (int ni = 0; ni < 4; ni++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1]) : "d"(fa[mi]), "d"(fb[ni]));
  }
#pragma unroll
  for (int mi = 0; mi < 4; mi++) {
    const int i = i0 + wm * 32 + mi * 8 + (lane >> 2);
    if (i >= n) continue;
    double* Ci = A + (size_t)i * n;
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      const int j = j0 + wn * 32 + ni * 8 + 2 * (lane & 3);
      if (j + 1 <= i) *reinterpret_cast<double2*>(Ci + j) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
      else if (j <= i) Ci[j] = acc[mi][ni][0];
    }
  }
}

__global__ void __launch_bounds__(256, 2) k_ldlt_update(double* A, const double* Wp, int n, int k0, int part) {
  ldlt_update_body(A, Wp, n, k0, part, (int)blockIdx.x);
}

// One launch per step of the late factorisation: the CTAs of panel k (blocks 0 .. n_panel_ctas-1, dispatched
// first: they are the chain) and, behind them, the tail of panel k-1's trailing update (column blocks from
// k+1 on), which only depends on panel k-1 and touches nothing panel k reads or writes.  With both in one
// grid the chain is a single stream of back-to-back kernels: no second stream, no event record / wait
// between the panels (those cost ~9 us per panel).  Used once the tail is small enough to hide behind the
// panel at one CTA per SM; the early, large tails keep their own two-CTAs-per-SM launches on the second stream.
__global__ void __launch_bounds__(256) k_ldlt_step(double* A, double* Wp_cur, double* Wp_prev, double* y, int n, int k0, int n_panel_ctas) {
  if ((int)blockIdx.x < n_panel_ctas) ldlt_panel_body(A, Wp_cur, Wp_prev, y, n, k0, (int)blockIdx.x);
  else ldlt_update_body(A, Wp_prev, n, k0 - kNB, 2, (int)blockIdx.x - n_panel_ctas);
}

// Backward substitution  L^T x = D^-1 y  in ONE launch.  The panels are walked from the bottom up by a
// thread-block cluster of kBackCtas CTAs; the steps are separated by the hardware cluster barrier
// (arrive.release / wait.acquire, which also orders the z updates in global memory between the CTAs)
// instead of 47 kernel boundaries (C4: 47 x 11 us before).  Per 64-row panel every CTA first solves the
// panel's 64x64 block itself (x_p = L11^-T z_p; z_p is complete by then), CTA 0 stores it in x, then each
// CTA applies the panel to its slice of the rows above:  z[i] -= sum_c L[k0 + c][i] x[k0 + c], i < k0
// (coalesced along i).  The next panel's diagonal block is fetched into registers while the current one
// is being solved.  `z` must hold D^-1 y (k_ldlt_scale).
__global__ void __launch_bounds__(256) k_ldlt_scale(const double* A, const double* y, double* z, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = y[i] / A[(size_t)i * n + i];
}

constexpr int kBackCtas = 8;       // portable cluster size
constexpr int kBackThreads = 512;  // 4096 threads: one row of z per thread up to n = 4160

__global__ void __cluster_dims__(kBackCtas, 1, 1) __launch_bounds__(kBackThreads, 1) k_ldlt_back(const double* A, double* z, double* x, int n) {
  __shared__ double a[kNB][kNB + 1];
  __shared__ double xs[kNB];
  const int tid = threadIdx.x;
  unsigned rank;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  constexpr int kPer = kNB * kNB / kBackThreads;
  double nxt[kPer];
  auto fetch = [&](int k0) {  // strict lower triangle of the diagonal block at k0, zero elsewhere
    const int nb = min(kNB, n - k0);
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      const int i = tid + q * kBackThreads, r = i / kNB, c = i % kNB;
      nxt[q] = (r < nb && c < r) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0;
    }
  };
  const int np = (n + kNB - 1) / kNB;
  fetch((np - 1) * kNB);
  for (int p = np - 1; p >= 0; p--) {
    const int k0 = p * kNB, nb = min(kNB, n - k0);
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      const int i = tid + q * kBackThreads;
      a[i / kNB][i % kNB] = nxt[q];
    }
    __syncthreads();
    if (p > 0) fetch(k0 - kNB);
    if (tid < 32) {  // L11^T x = z inside the block: lane r holds rows r and r + 32, pivots travel by shuffle
      double x0 = tid < nb ? __ldcg(&z[k0 + tid]) : 0.0, x1 = tid + 32 < nb ? __ldcg(&z[k0 + tid + 32]) : 0.0;
      for (int c = kNB - 1; c >= 32; c--) {
        const double xc = __shfl_sync(kFull, x1, c - 32);
        x0 -= a[c][tid] * xc;
        if (tid + 32 < c) x1 -= a[c][tid + 32] * xc;
      }
      for (int c = 31; c >= 0; c--) {
        const double xc = __shfl_sync(kFull, x0, c);
        if (tid < c) x0 -= a[c][tid] * xc;
      }
      xs[tid] = x0; xs[tid + 32] = x1;
    }
    __syncthreads();
    if (rank == 0 && tid < nb) x[k0 + tid] = xs[tid];
    for (int i = (int)rank * kBackThreads + tid; i < k0; i += kBackCtas * kBackThreads) {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
      const double* Ac = A + (size_t)k0 * n + i;
#pragma unroll 4
      for (int c = 0; c < kNB; c += 4) {  // rows past nb are never touched: xs is zero there, but stay in bounds
        if (c + 3 < nb) {
          v0 += Ac[(size_t)c * n] * xs[c]; v1 += Ac[(size_t)(c + 1) * n] * xs[c + 1];
          v2 += Ac[(size_t)(c + 2) * n] * xs[c + 2]; v3 += Ac[(size_t)(c + 3) * n] * xs[c + 3];
        } else {
          for (int q = c; q < nb; q++) v0 += Ac[(size_t)q * n] * xs[q];
        }
      }
      __stcg(&z[i], __ldcg(&z[i]) - ((v0 + v1) + (v2 + v3)));
    }
    // every CTA of the cluster is done with this panel (and with a / xs) before the next one starts
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}


// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_point_update(BundleDev d) {
  __shared__ double sh[32];
  const int i = d.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  double ss = 0.0;
  if (i < d.p_hi) {
    double sum[3] = {0, 0, 0};
    for (int o = d.pt_off[i]; o < d.pt_off[i + 1]; o++) {
      const int m = d.pt_meas[o];
      if (d.m_state[m] != M_ALIVE) continue;
      const int row = d.cam_row[d.m_cam[m]];
      if (row < 0) continue;
      const double* W = d.m_W + 18 * m;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double a = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) a += W[3 * q + r] * d.upd[row + q];
        sum[r] += a;
      }
    }
    const double v0 = d.epsB[3 * i] - sum[0], v1 = d.epsB[3 * i + 1] - sum[1], v2 = d.epsB[3 * i + 2] - sum[2];
    const double* Vi = d.Vinv + 9 * i;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double u = Vi[3 * r] * v0 + Vi[3 * r + 1] * v1 + Vi[3 * r + 2] * v2;
      d.pt_pos_new[3 * i + r] = d.pt_pos[3 * i + r] + u;
      ss += u * u;
    }
  }
  const double t = block_sum(ss, sh);
  if (threadIdx.x == 0) atomicAdd(&d.scal[4], t);
}

__global__ void __launch_bounds__(128) k_ba_cam_update(BundleDev d) {
  __shared__ double sh[32];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double ss = 0.0;
  if (c < d.n_cams) {
    const int row = d.cam_row[c];
    if (row < 0) {
      for (int k = 0; k < 12; k++) d.cam_se3_new[12 * c + k] = d.cam_se3[12 * c + k];
    } else {
      double mu[6], ex[12], np[12];
      for (int k = 0; k < 6; k++) { mu[k] = d.upd[row + k]; if (d.add_cam_update) ss += mu[k] * mu[k]; }
      se3_exp(mu, ex);
      se3_mul(ex, d.cam_se3 + 12 * c, np);
      for (int k = 0; k < 12; k++) d.cam_se3_new[12 * c + k] = np[k];
    }
  }
  const double t = block_sum(ss, sh);
  if (threadIdx.x == 0) atomicAdd(&d.scal[4], t);
}

__global__ void __launch_bounds__(256) k_ba_new_error(BundleDev d) {
  __shared__ double sh[32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (m < d.n_meas && d.m_state[m] != M_ERASED) {
    double v3[3];
    se3_apply(d.cam_se3_new + 12 * d.m_cam[m], d.pt_pos_new + 3 * d.m_pt[m], v3);
    if (v3[2] <= 0) e = 1.0;
    else {
      const CamProj q = cam_project(d.cam, v3[0] / v3[2], v3[1] / v3[2]);
      const double s = d.m_sin[m];
      const double e0 = s * (d.m_found[2 * m] - q.im[0]), e1 = s * (d.m_found[2 * m + 1] - q.im[1]);
      e = mest_objective(e0 * e0 + e1 * e1, d.scal[1], d.est);
    }
  }
  const double t = block_sum(e, sh);
  if (threadIdx.x == 0) atomicAdd(&d.scal[3], t);
}

// end of an LM step: erase the bad measurements, appending (point, camera) in list order
__global__ void __launch_bounds__(1024) k_ba_erase(BundleDev d, int step) {
  __shared__ int wcnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = d.counters[1];
  __syncthreads();
  for (int m0 = 0; m0 < d.n_meas; m0 += blockDim.x) {
    const int m = m0 + threadIdx.x;
    const bool bad = m < d.n_meas && d.m_state[m] == M_BAD;
    const unsigned b = __ballot_sync(kFull, bad);
    if (lane == 0) wcnt[warp] = __popc(b);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < warp; w++) before += wcnt[w];
    if (bad) {
      const int o = before + __popc(b & ((1u << lane) - 1));
      d.outliers[2 * o] = d.m_pt[m]; d.outliers[2 * o + 1] = d.m_cam[m];
      d.m_state[m] = M_ERASED;
      d.m_erase_step[m] = step;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wcnt[w]; base_s += t; }
    __syncthreads();
  }
  if (threadIdx.x == 0) d.counters[1] = base_s;
}

// Large graphs: the same ordered erase with many CTAs.  k_ba_erase_count: bad measurements per
// 1024-measurement block; k_ba_erase_scan: exclusive scan of the block counts (one CTA) on top of the
// running total; k_ba_erase_write: ordered positions inside each block by ballot / popc.
__global__ void __launch_bounds__(1024) k_ba_erase_count(BundleDev d, int* block_cnt) {
  __shared__ int wcnt[32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const bool bad = m < d.n_meas && d.m_state[m] == M_BAD;
  const unsigned b = __ballot_sync(kFull, bad);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wcnt[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    const int v = warp_sum_int(wcnt[lane]);
    if (lane == 0) block_cnt[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(1024) k_ba_erase_scan(BundleDev d, int* block_cnt, int n_blocks) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = d.counters[1];
  __syncthreads();
  for (int base = 0; base < n_blocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? block_cnt[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int ws = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(kFull, ws, o);
        if (lane >= o) ws += n;
      }
      wsum[lane] = ws;
    }
    __syncthreads();
    if (i < n_blocks) block_cnt[i] = carry + (warp ? wsum[warp - 1] : 0) + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) d.counters[1] = carry;
}
__global__ void __launch_bounds__(1024) k_ba_erase_write(BundleDev d, const int* block_cnt, int step) {
  __shared__ int wcnt[32];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const bool bad = m < d.n_meas && d.m_state[m] == M_BAD;
  const unsigned b = __ballot_sync(kFull, bad);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wcnt[warp] = __popc(b);
  __syncthreads();
  if (bad) {
    int before = block_cnt[blockIdx.x];
    for (int w = 0; w < warp; w++) before += wcnt[w];
    const int o = before + __popc(b & ((1u << lane) - 1));
    d.outliers[2 * o] = d.m_pt[m]; d.outliers[2 * o + 1] = d.m_cam[m];
    d.m_state[m] = M_ERASED;
    d.m_erase_step[m] = step;
  }
}

// sharded handles: local erase marks -> global measurement order (merged with an all-reduce max)
__global__ void __launch_bounds__(256) k_ba_scatter_steps(const int* step, const int* gid, int* out, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n && step[m] > 0) out[gid[m]] = step[m];
}

}  // namespace ptam
