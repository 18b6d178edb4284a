// sm_100a kernels of path B (Bundle::Compute, reference src/Bundle.cc:209-551).
//
// per LM step      k_ba_project      ProjectAndFindSquaredError per measurement (Bundle.cc:219-225)
//                  k_ba_select       exact order statistic of the squared errors (sigma^2, :230-237)
//                  k_ba_jacobian     weights, outlier marks, B (2x3), W = A^T B (:251-332)
//                  k_ba_acc_cam      U_j, epsA_j, current error: warp per camera, IN LIST ORDER (no atomics)
//                  k_ba_acc_pt       V_i, epsB_i: thread per point, in list order
// per lambda trial k_ba_vinv         V*_i^-1 by 3x3 LDL^T (:341-359)
//                  k_ba_schur_diag   S_jj = U*_j - sum_i ..., vE_j: warp per camera, in point order (:374-406)
//                  k_ba_schur_off    S_jk = - sum_i W_ij V*^-1 W_ik^T: warp per camera pair, in point order (:410-446)
//                  (ldlt.cu)         blocked square-root-free LDL^T of S (DMMA trailing update) + solve (:457-458)
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
  double* Vinv;         // [P][12]: V*_i^-1 (9, row-major) and V*_i^-1 epsB_i (3): one 96-byte record, read as 16-byte pairs
  const int* pt_off;    // [P+1] CSR by point
  const int* pt_meas;   // [M] a point's measurements by ascending camera id (the reference's std::set<int> order)
  const int* pt_meas_ins;  // [M] the same in list (insertion) order; aliases pt_meas when the two agree
  const int* pt_cam;    // [M] camera id of pt_meas[o]
  // CSR by camera
  const int* cam_off;       // [C+1]
  const int* cam_meas_ins;  // [M] a camera's measurements in list order
  const int* cam_meas_pt;   // [M] the same by ascending point id; aliases cam_meas_ins when the two agree
  // camera pairs (j > k, both free) in block order b = jf (jf - 1) / 2 + kf (jf, kf = start row / 6)
  long long n_blocks;
  const int* blk_off;   // [n_blocks + 1] triples of block b: [blk_off[b], blk_off[b + 1]), ascending point id
  const int* cam_order; // cameras by descending measurement count (launch order of the per-camera kernels)
  const int* pr_mj;     // measurement of the point in camera j
  const int* pr_mk;     // ... and in camera k
  const int* nz_blocks; // the pairs with common points, long ones (>= kLongBlock triples) first
  const int* pair_info; // [0] number of such pairs, [1] long ones among them, [2] triples
  const int* free_cam;  // [n / 6] camera id of free camera jf
  // measurements (insertion order)
  const int* m_cam; const int* m_pt;
  const double* m_found;  // [M][2]
  const double* m_sin;    // [M] dSqrtInvNoise
  int* m_state;           // [M]
  double* m_v3cam;        // [M][4] (x, y, z, pad: 32-byte records)
  double* m_derivs;       // [M][4]
  double* m_eps;          // [M][2]
  double* m_e2;           // [M]
  double* m_W;            // [M][18]
  double* m_B;            // [M][6]
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
  const double* sel_keys;        // sharded handles: the squared errors of ALL shards (all-gathered, ~0 = not a key), else null
  int sel_n;                     // entries of sel_keys
  int* counters;  // 0 n_valid, 1 n_outliers_total, 2 n_bad_this_step
  int* outliers;  // [M][2] (point, camera) in erase order
  int* m_erase_step;  // [M] LM step (1-based) at which the measurement was erased, 0 = still in the graph
  // deterministic grid-wide sums
  double* err_cam;    // [C] each camera's share of the current robust error
  double* partials;   // [max grid] per-block partial sums of the kernel in flight
  unsigned* tickets;  // [8] arrival counters (self-resetting)
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

// 16-byte accesses to the per-measurement records (W 144 B, B 48 B, derivatives 32 B, eps 16 B per measurement: all
// multiples of 16 from 256-byte aligned arrays): half the memory instructions of the latency-bound passes.
template <int N>
PTAM_DEV void ld_pairs(const double* p, double (&v)[N]) {
  static_assert(N % 2 == 0, "pairs");
#pragma unroll
  for (int q = 0; q < N; q += 2) { const double2 t = *reinterpret_cast<const double2*>(p + q); v[q] = t.x; v[q + 1] = t.y; }
}
template <int N>
PTAM_DEV void st_pairs(double* p, const double (&v)[N]) {
  static_assert(N % 2 == 0, "pairs");
#pragma unroll
  for (int q = 0; q < N; q += 2) *reinterpret_cast<double2*>(p + q) = make_double2(v[q], v[q + 1]);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_project(BundleDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.n_meas) return;
  if (d.m_state[m] == M_ERASED) return;
  const int c = d.m_cam[m], p = d.m_pt[m];
  double v3[3];
  se3_apply(d.cam_se3 + 12 * c, d.pt_pos + 3 * p, v3);
  {
    const double rec[4] = {v3[0], v3[1], v3[2], 0.0};
    st_pairs(d.m_v3cam + 4 * (size_t)m, rec);
  }
  if (v3[2] <= 0) { d.m_state[m] = M_BAD; return; }
  d.m_state[m] = M_ALIVE;
  const CamProj q = cam_project(d.cam, v3[0] / v3[2], v3[1] / v3[2]);
  double dv[4];
  cam_derivs(d.cam, q, dv);
  st_pairs(d.m_derivs + 4 * (size_t)m, dv);
  const double s = d.m_sin[m];
  double fd[2];
  ld_pairs(d.m_found + 2 * (size_t)m, fd);
  const double e0 = s * (fd[0] - q.im[0]), e1 = s * (fd[1] - q.im[1]);
  const double ep[2] = {e0, e1};
  st_pairs(d.m_eps + 2 * (size_t)m, ep);
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
// patterns, six passes with digit widths 11,11,11,11,11,9.  Every pass is ONE launch, k_ba_hist_pick: all CTAs
// add their digits (aggregated per warp with match.any, per-CTA histogram in shared memory, non-zero bins flushed
// with global atomics); the last CTA to arrive finds the bin holding the wanted rank, extends the prefix and
// leaves the histogram zeroed for the next pass (three launches per pass before: memset, histogram, pick).
// Sharded handles run the same passes on the all-gathered keys of every shard.  After the last pass the prefix IS
// the bit pattern of the floor(n/2)-th smallest squared error (Tools.h:152-162 sorts and takes element n/2),
// identical on every shard.
// ---------------------------------------------------------------------------------------------
constexpr int kSelPasses = 6;
constexpr int kSelBins = 2048;
PTAM_HD int sel_shift(int pass) { return pass < 5 ? 53 - 11 * pass : 0; }
PTAM_HD int sel_width(int pass) { return pass < 5 ? 11 : 9; }

__global__ void __launch_bounds__(256) k_ba_hist_pick(BundleDev d, int pass, double min_sigma_sq) {
  __shared__ int h[kSelBins];
  __shared__ int wsum[8];
  __shared__ bool s_last;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int shift = sel_shift(pass), width = sel_width(pass);
  const unsigned long long prefix = d.sel_state[0];
  const int lane = threadIdx.x & 31;
  const int n_keys = d.sel_keys ? d.sel_n : d.n_meas;
  for (int m0 = blockIdx.x * blockDim.x; m0 < n_keys; m0 += gridDim.x * blockDim.x) {
    const int m = m0 + threadIdx.x;
    bool take = false;
    unsigned digit = 0;
    if (m < n_keys) {
      unsigned long long key = ~0ull;
      if (d.sel_keys) key = (unsigned long long)__double_as_longlong(d.sel_keys[m]);
      else if (d.m_state[m] == M_ALIVE) key = (unsigned long long)__double_as_longlong(d.m_e2[m]);
      if (key != ~0ull) {
        take = pass == 0 || (key >> (shift + width)) == (prefix >> (shift + width));
        digit = (unsigned)(key >> shift) & ((1u << width) - 1u);
      }
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
  // ---- the last CTA to arrive picks the bin of the k-th key (the counter wraps to zero for the next pass)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicInc(d.tickets + 4, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int t = threadIdx.x, warp = t >> 5;
  constexpr int kPer = kSelBins / 256;
  int c[kPer], tot = 0;
#pragma unroll
  for (int q = 0; q < kPer; q++) { c[q] = __ldcg(&d.hist16[kPer * t + q]); tot += c[q]; }
#pragma unroll
  for (int q = 0; q < kPer; q++) d.hist16[kPer * t + q] = 0;  // the next pass (and the next LM step) start from zero
  int inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  int before = 0, n_all = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) { if (w < warp) before += wsum[w]; n_all += wsum[w]; }
  // pass 0: n_all = number of valid measurements over all shards
  const long long kk = pass == 0 ? (long long)(n_all / 2) : (long long)d.sel_state[1];
  const int excl = before + inc - tot;
  if (n_all > 0 && kk >= excl && kk < excl + tot) {  // exactly one thread
    int r = (int)(kk - excl), q = 0;
    while (r >= c[q]) { r -= c[q]; q++; }
    const unsigned long long prefix2 = (pass == 0 ? 0ull : prefix) | ((unsigned long long)(kPer * t + q) << shift);
    d.sel_state[0] = prefix2;
    d.sel_state[1] = (unsigned long long)r;
    if (pass == kSelPasses - 1) {
      const double med = __longlong_as_double((long long)prefix2);
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

// Sharded handles: this shard's squared errors into its slot of the all-gather buffer (erased / outlier
// measurements and the padding of the slot as ~0, which is no key: a NaN pattern no squared error has).
__global__ void __launch_bounds__(256) k_ba_sel_keys(BundleDev d, double* slot, int slot_n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= slot_n) return;
  long long key = ~0ll;
  if (m < d.n_meas && d.m_state[m] == M_ALIVE) key = __double_as_longlong(d.m_e2[m]);
  slot[m] = __longlong_as_double(key);
}

// The lower triangle of S row by row (row r: r + 1 values at r (r + 1) / 2) with vE behind it: what the shards
// exchange per lambda trial (n (n + 1) / 2 + n doubles instead of n^2 + n), and back.
// (`extra`: a few scalars that ride along, e.g. the error sums of the LM step with its first lambda trial.)
__global__ void __launch_bounds__(256) k_ba_pack_lower(const double* S, const double* vE, int n, double* out, const double* extra, int n_extra) {
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const double* src = S + (size_t)r * n;
    double* dst = out + (size_t)r * (r + 1) / 2;
    for (int c = threadIdx.x; c <= r; c += blockDim.x) dst[c] = src[c];
  }
  if (blockIdx.x == 0) {
    double* dst = out + (size_t)n * (n + 1) / 2;
    for (int c = threadIdx.x; c < n; c += blockDim.x) dst[c] = vE[c];
    if ((int)threadIdx.x < n_extra) dst[n + threadIdx.x] = extra[threadIdx.x];
  }
}
__global__ void __launch_bounds__(256) k_ba_unpack_lower(const double* in, int n, double* S, double* vE, double* extra, int n_extra) {
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    double* dst = S + (size_t)r * n;
    const double* src = in + (size_t)r * (r + 1) / 2;
    for (int c = threadIdx.x; c <= r; c += blockDim.x) dst[c] = src[c];
  }
  if (blockIdx.x == 0) {
    const double* src = in + (size_t)n * (n + 1) / 2;
    for (int c = threadIdx.x; c < n; c += blockDim.x) vE[c] = src[c];
    if ((int)threadIdx.x < n_extra) extra[threadIdx.x] = src[n + threadIdx.x];
  }
}


// ---------------------------------------------------------------------------------------------
// Deterministic accumulation.  Every sum of the reference that runs over measurements is taken IN THE
// REFERENCE'S ORDER, one term after the other, with the reference's expressions (this file is compiled
// without FMA contraction): no floating-point atomics anywhere on path B.
//   U_j, epsA_j   over camera j's measurements in list order            (Bundle.cc:251-332)
//   V_i, epsB_i   over point i's measurements in list order             (same loop)
//   S_jj, vE_j    U*_j, epsA_j minus the terms of camera j's points in point order   (:374-406)
//   S_jk          minus the terms of the points seen by both, in point order         (:410-446)
// A segment (one camera's measurements, one camera pair's points) belongs to ONE WARP: 32 lanes compute the
// terms of 32 consecutive items, stage them in shared memory (value-major, pitch 33: conflict-free both
// ways), then lane v adds value v of the 32 items one after the other.  Items that contribute nothing
// (erased / outlier measurements, the tail of the last chunk) stage +0.0, which leaves a sum unchanged.
// ---------------------------------------------------------------------------------------------
// Items per round of the per-camera segment kernels (CTA per camera).  The sums of a camera are taken in order by
// ONE thread per value, so the longest camera (3 778 measurements at C4) is the critical path: with 512 items per
// round its rounds halve while a round's staging, which is latency, takes as long as with 256 (k_ba_acc_cam
// 70 -> 54 us, k_ba_schur_diag 115 -> 86 us at C4; one CTA per SM then, which is what the shared memory allows).
constexpr int kSegThreads = 512;
constexpr int kOffThreads = 128;  // ... of the per-camera-pair kernel
constexpr int kLongBlock = 512;   // camera pairs with at least this many common points are scheduled first

// Every block stores its partial; the last block to arrive adds them in index order (lane-strided,
// then the fixed xor tree): the same bits whatever the block schedule.  `t` is valid in thread 0.
PTAM_DEV void grid_sum_finish(double t, double* partials, unsigned* ticket, double* out, bool add_to_out) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: ready for the next launch
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) s += __ldcg(&partials[i]);
    s = warp_sum(s);
    if (threadIdx.x == 0) *out = add_to_out ? *out + s : s;
  }
}

// One round of a segment sum: thread t has staged the NV terms of item t at sv[q * (R + 1) + t] (value-major,
// pitch R + 1: conflict-free both ways; zeros when the item contributes nothing); thread v < NV then adds (or
// subtracts) value v of the round's `cnt` items one after the other.
template <int NV, int R>
PTAM_DEV void seg_accumulate(const double* sv, int cnt, double& acc, bool subtract) {
  const int t = threadIdx.x;
  __syncthreads();
  if (t < NV) {
    const double* row = sv + t * (R + 1);
    if (subtract) {
#pragma unroll 8
      for (int i = 0; i < cnt; i++) acc -= row[i];
    } else {
#pragma unroll 8
      for (int i = 0; i < cnt; i++) acc += row[i];
    }
  }
  __syncthreads();
}

// A (2x6) of one measurement: meas.dSqrtInvNoise * m2CamDerivs * v2CamFrameMotion per generator (Bundle.cc:287-301)
PTAM_DEV void ba_cam_jacobian(double X, double Y, double Z, double ooz, double d0, double d1, double d2, double d3, double* A) {
  const double gx[6] = {1, 0, 0, 0, Z, -Y}, gy[6] = {0, 1, 0, -Z, 0, X}, gz[6] = {0, 0, 1, Y, -X, 0};
#pragma unroll
  for (int q = 0; q < 6; q++) {
    const double a0 = (gx[q] - X * gz[q] * ooz) * ooz, a1 = (gy[q] - Y * gz[q] * ooz) * ooz;
    A[q] = d0 * a0 + d1 * a1;
    A[6 + q] = d2 * a0 + d3 * a1;
  }
}

// k_ba_jacobian — one thread per observation: weight, outlier mark, B (2x3), W = A^T B (Bundle.cc:251-332).
// The accumulators are summed by k_ba_acc_cam / k_ba_acc_pt.
__global__ void __launch_bounds__(128) k_ba_jacobian(BundleDev d) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.n_meas || d.m_state[m] != M_ALIVE) return;
  const double sigma2 = d.scal[1];
  const double e2 = d.m_e2[m];
  const double w = mest_sqrt_weight(e2, sigma2, d.est);
  double ep[2];
  ld_pairs(d.m_eps + 2 * (size_t)m, ep);
  const double eps0 = w * ep[0], eps1 = w * ep[1];
  ep[0] = eps0; ep[1] = eps1;
  st_pairs(d.m_eps + 2 * (size_t)m, ep);
  if (w == 0) { d.m_state[m] = M_BAD; return; }
  const int c = d.m_cam[m];
  const double s = d.m_sin[m];
  double dv[4];
  ld_pairs(d.m_derivs + 4 * (size_t)m, dv);
  const double d0 = s * (w * dv[0]), d1 = s * (w * dv[1]);
  const double d2 = s * (w * dv[2]), d3 = s * (w * dv[3]);
  double pc[4];
  ld_pairs(d.m_v3cam + 4 * (size_t)m, pc);
  const double X = pc[0], Y = pc[1], Z = pc[2];
  const double ooz = 1.0 / Z;
  double A[12];
  if (d.cam_fixed[c]) {
#pragma unroll
    for (int i = 0; i < 12; i++) A[i] = 0.0;
  } else ba_cam_jacobian(X, Y, Z, ooz, d0, d1, d2, d3, A);
  double B[6];
  const double* R = d.cam_se3 + 12 * c;
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const double a0 = (R[q] - X * R[6 + q] * ooz) * ooz, a1 = (R[3 + q] - Y * R[6 + q] * ooz) * ooz;
    B[q] = d0 * a0 + d1 * a1;
    B[3 + q] = d2 * a0 + d3 * a1;
  }
  st_pairs(d.m_B + 6 * (size_t)m, B);
  double W[18];  // W = A^T B (6x3), zero for a fixed camera
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) W[3 * r + cc] = A[r] * B[cc] + A[6 + r] * B[3 + cc];
  st_pairs(d.m_W + 18 * (size_t)m, W);
}

// k_ba_acc_cam — CTA per camera, its measurements in list order: U_j (lower, packed 21), epsA_j (6) and the
// camera's share of the robust error (value 27).  A is recomputed from the stored projection (76 B per
// measurement instead of a 96-byte round trip).  The last block adds the per-camera errors in camera order.
constexpr int kAccCamSmem = 28 * (kSegThreads + 1) * (int)sizeof(double);
__global__ void __launch_bounds__(kSegThreads) k_ba_acc_cam(BundleDev d) {
  extern __shared__ __align__(16) double seg_sv[];
  __shared__ bool s_last;
  const int c = d.cam_order[blockIdx.x], t = threadIdx.x;  // longest camera first: the tail of the grid is short work
  const double sigma2 = d.scal[1];
  const bool cfree = !d.cam_fixed[c];
  const int o0 = d.cam_off[c], o1 = d.cam_off[c + 1];
  double acc = 0.0;
  // the index and the state of the NEXT round's measurement are fetched a round ahead: one dependent load
  // (index -> data) instead of three (index -> state -> data) in front of every round's arithmetic (79 -> 70 us
  // at C4; the same in k_ba_schur_diag / k_ba_schur_off changed nothing: their rounds are bound by the sums)
  int m_nx = -1, st_nx = 0;
  if (o0 + t < o1) { m_nx = d.cam_meas_ins[o0 + t]; st_nx = d.m_state[m_nx]; }
  for (int base = o0; base < o1; base += kSegThreads) {
    const int o = base + t;
    double* sv = seg_sv + t;
    constexpr int kP = kSegThreads + 1;
    bool filled = false;
    double err = 0.0;
    const int m = m_nx, st = st_nx;
    if (o + kSegThreads < o1) { m_nx = d.cam_meas_ins[o + kSegThreads]; st_nx = d.m_state[m_nx]; }
    if (o < o1) {
      if (st == M_BAD) err = 1.0;
      else if (st == M_ALIVE) {
        const double e2 = d.m_e2[m];
        err = mest_objective(e2, sigma2, d.est);
        if (cfree) {
          filled = true;
          const double w = mest_sqrt_weight(e2, sigma2, d.est);
          const double s = d.m_sin[m];
          double dv[4];
          ld_pairs(d.m_derivs + 4 * (size_t)m, dv);
          const double d0 = s * (w * dv[0]), d1 = s * (w * dv[1]);
          const double d2 = s * (w * dv[2]), d3 = s * (w * dv[3]);
          double pc[4];
          ld_pairs(d.m_v3cam + 4 * (size_t)m, pc);
          const double X = pc[0], Y = pc[1], Z = pc[2];
          double A[12];
          ba_cam_jacobian(X, Y, Z, 1.0 / Z, d0, d1, d2, d3, A);
          double ep[2];
          ld_pairs(d.m_eps + 2 * (size_t)m, ep);
          const double eps0 = ep[0], eps1 = ep[1];  // already weighted
          int q = 0;
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = 0; cc <= r; cc++) sv[kP * q++] = A[r] * A[cc] + A[6 + r] * A[6 + cc];
#pragma unroll
          for (int r = 0; r < 6; r++) sv[kP * (21 + r)] = A[r] * eps0 + A[6 + r] * eps1;
        }
      }
    }
    if (!filled) {
#pragma unroll
      for (int q = 0; q < 27; q++) sv[kP * q] = 0.0;
    }
    sv[kP * 27] = err;
    seg_accumulate<28, kSegThreads>(seg_sv, min(kSegThreads, o1 - base), acc, false);
  }
  if (t < 21) d.U[21 * c + t] = cfree ? acc : 0.0;
  else if (t < 27) d.epsA[6 * c + t - 21] = cfree ? acc : 0.0;
  else if (t == 27) {
    d.err_cam[c] = acc;
    __threadfence();
    s_last = atomicInc(d.tickets + 0, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && t < 32) {  // current error = the cameras' shares, in camera order
    __threadfence();
    double s = 0.0;
    for (int i = t; i < d.n_cams; i += 32) s += __ldcg(&d.err_cam[i]);
    s = warp_sum(s);
    if (t == 0) d.scal[2] = s;
  }
}

// k_ba_acc_pt — thread per point, its measurements in list order: V_i (lower, packed), epsB_i.
__global__ void __launch_bounds__(256) k_ba_acc_pt(BundleDev d) {
  const int i = d.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.p_hi) return;
  double V[6] = {0, 0, 0, 0, 0, 0}, eB[3] = {0, 0, 0};
  for (int o = d.pt_off[i]; o < d.pt_off[i + 1]; o++) {
    const int m = d.pt_meas_ins[o];
    if (d.m_state[m] != M_ALIVE) continue;
    double B[6], ep[2];
    ld_pairs(d.m_B + 6 * (size_t)m, B);
    ld_pairs(d.m_eps + 2 * (size_t)m, ep);
    const double b0 = B[0], b1 = B[1], b2 = B[2], b3 = B[3], b4 = B[4], b5 = B[5];
    const double eps0 = ep[0], eps1 = ep[1];
    V[0] += b0 * b0 + b3 * b3;
    V[1] += b1 * b0 + b4 * b3;
    V[2] += b1 * b1 + b4 * b4;
    V[3] += b2 * b0 + b5 * b3;
    V[4] += b2 * b1 + b5 * b4;
    V[5] += b2 * b2 + b5 * b5;
    eB[0] += b0 * eps0 + b3 * eps1;
    eB[1] += b1 * eps0 + b4 * eps1;
    eB[2] += b2 * eps0 + b5 * eps1;
  }
#pragma unroll
  for (int q = 0; q < 6; q++) d.V[6 * (size_t)i + q] = V[q];
#pragma unroll
  for (int q = 0; q < 3; q++) d.epsB[3 * (size_t)i + q] = eB[q];
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_vinv(BundleDev d) {
  const int i = d.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.p_hi) return;
  const double lambda = d.scal[6];
  const double* v = d.V + 6 * (size_t)i;
  double Vs[9] = {v[0], v[1], v[3], v[1], v[2], v[4], v[3], v[4], v[5]};
  double inv[9];
  if (Vs[0] * Vs[4] * Vs[8] == 0) {
#pragma unroll
    for (int k = 0; k < 9; k++) inv[k] = 0;
  } else {
    Vs[0] *= (1.0 + lambda); Vs[4] *= (1.0 + lambda); Vs[8] *= (1.0 + lambda);
    ldlt_inverse<3>(Vs, inv);
  }
  double rec[12];
#pragma unroll
  for (int k = 0; k < 9; k++) rec[k] = inv[k];
  const double* e = d.epsB + 3 * (size_t)i;
#pragma unroll
  for (int r = 0; r < 3; r++) rec[9 + r] = inv[3 * r] * e[0] + inv[3 * r + 1] * e[1] + inv[3 * r + 2] * e[2];
  st_pairs(d.Vinv + 12 * (size_t)i, rec);
}

// Schur complement (Bundle.cc:365-453): k_ba_zero_lower clears the lower triangle (the solve left its factors
// there), then two segment kernels write the diagonal blocks + vE and the camera pairs with common points.
__global__ void __launch_bounds__(256) k_ba_zero_lower(double* S, int n, unsigned* work_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0u;  // k_ba_schur_off (next in the stream) fetches its pairs through it
  for (int i = blockIdx.x; i < n; i += gridDim.x)
    for (int j = threadIdx.x; j <= i; j += blockDim.x) S[(size_t)i * n + j] = 0.0;
}

// k_ba_schur_diag — CTA per camera j: S_jj = U*_j - sum_i W_ij V*_i^-1 W_ij^T,  vE_j = epsA_j - sum_i W_ij V*_i^-1 epsB_i
// over camera j's points in POINT order (the reference scans i = 0 .. P-1, Bundle.cc:396-405).
constexpr int kSchurDiagSmem = 27 * (kSegThreads + 1) * (int)sizeof(double);
__global__ void __launch_bounds__(kSegThreads, 1) k_ba_schur_diag(BundleDev d) {
  extern __shared__ __align__(16) double seg_sv[];
  const int c = d.cam_order[blockIdx.x], t = threadIdx.x;  // longest camera first
  const int row = d.cam_row[c];
  if (row < 0) return;
  const double lambda = d.scal[6];
  int r6 = 0, c6 = 0;  // t < 21: packed lower-triangle index -> (row, column)
  if (t < 21) { while ((r6 + 1) * (r6 + 2) / 2 <= t) r6++; c6 = t - r6 * (r6 + 1) / 2; }
  double acc = 0.0;
  if (t < 21) { acc = d.U[21 * c + t]; if (r6 == c6) acc *= (1.0 + lambda); }
  else if (t < 27) acc = d.epsA[6 * c + t - 21];
  const int o0 = d.cam_off[c], o1 = d.cam_off[c + 1];
  for (int base = o0; base < o1; base += kSegThreads) {
    const int o = base + t;
    double* sv = seg_sv + t;
    constexpr int kP = kSegThreads + 1;
    bool filled = false;
    if (o < o1) {
      const int m = d.cam_meas_pt[o];
      if (d.m_state[m] == M_ALIVE) {
        filled = true;
        const int i = d.m_pt[m];
        double vv[12];
        ld_pairs(d.Vinv + 12 * (size_t)i, vv);
        const double* Vi = vv;
        const double* ve = vv + 9;
        double Wr[18], WV[18];
        ld_pairs(d.m_W + 18 * (size_t)m, Wr);
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) WV[3 * r + cc] = Wr[3 * r] * Vi[cc] + Wr[3 * r + 1] * Vi[3 + cc] + Wr[3 * r + 2] * Vi[6 + cc];
        int q = 0;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int cc = 0; cc <= r; cc++) sv[kP * q++] = WV[3 * r] * Wr[3 * cc] + WV[3 * r + 1] * Wr[3 * cc + 1] + WV[3 * r + 2] * Wr[3 * cc + 2];
#pragma unroll
        for (int r = 0; r < 6; r++) sv[kP * (21 + r)] = Wr[3 * r] * ve[0] + Wr[3 * r + 1] * ve[1] + Wr[3 * r + 2] * ve[2];
      }
    }
    if (!filled) {
#pragma unroll
      for (int q = 0; q < 27; q++) sv[kP * q] = 0.0;
    }
    seg_accumulate<27, kSegThreads>(seg_sv, min(kSegThreads, o1 - base), acc, true);
  }
  if (t < 21) {
    d.S[(size_t)(row + r6) * d.n + row + c6] = acc;
    d.S[(size_t)(row + c6) * d.n + row + r6] = acc;
  } else if (t < 27) d.vE[row + t - 21] = acc;
}

// block index b = jf (jf - 1) / 2 + kf  ->  (jf, kf), kf < jf
PTAM_DEV void pair_decode(long long b, int& jf, int& kf) {
  jf = (int)((1.0 + sqrt(1.0 + 8.0 * (double)b)) * 0.5);
  while ((long long)jf * (jf - 1) / 2 > b) jf--;
  while ((long long)(jf + 1) * jf / 2 <= b) jf++;
  kf = (int)(b - (long long)jf * (jf - 1) / 2);
}

// k_ba_schur_off — CTA per camera pair with common points (j > k, both free): S_jk = - sum_i W_ij V*_i^-1 W_ik^T
// over the points seen by both, in POINT order (the reference walks the points and their scripts,
// Bundle.cc:410-446).  The pair-major list (blk_off, pr_mj, pr_mk) and the list of non-empty pairs (longest
// first) are built once per Compute.  Pairs are fetched through an atomic counter (the order of the fetches
// does not matter: every block is written by exactly one CTA, in its own fixed order).
__global__ void __launch_bounds__(kOffThreads) k_ba_schur_off(BundleDev d) {
  __shared__ double sv[36 * (kOffThreads + 1)];
  __shared__ int s_idx;
  const int t = threadIdx.x;
  const int n_nz = d.pair_info[0];
  while (true) {
    if (t == 0) s_idx = (int)atomicAdd(d.tickets + 3, 1u);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= n_nz) break;
    const int b = d.nz_blocks[idx];
    int jf, kf;
    pair_decode(b, jf, kf);
    double acc = 0.0;
    const int o0 = d.blk_off[b], o1 = d.blk_off[b + 1];
    for (int base = o0; base < o1; base += kOffThreads) {
      const int o = base + t;
      double* st = sv + t;
      constexpr int kP = kOffThreads + 1;
      bool filled = false;
      if (o < o1) {
        const int mj = d.pr_mj[o], mk = d.pr_mk[o];
        if (d.m_state[mj] == M_ALIVE && d.m_state[mk] == M_ALIVE) {
          filled = true;
          // (W_ij V*_i^-1 kept from k_ba_schur_diag instead of recomputed here was tried: a second 18-double array
          // per measurement pushes the working set of the pairs out of the L2, 219 -> 367 us at C4)
          double Vi[10];
          ld_pairs(d.Vinv + 12 * (size_t)d.m_pt[mj], Vi);
          double Wj[18];
          ld_pairs(d.m_W + 18 * (size_t)mj, Wj);
          const double* Wk = d.m_W + 18 * (size_t)mk;
          double WV[18];
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = 0; cc < 3; cc++) WV[3 * r + cc] = Wj[3 * r] * Vi[cc] + Wj[3 * r + 1] * Vi[3 + cc] + Wj[3 * r + 2] * Vi[6 + cc];
          double wk[18];
          ld_pairs(Wk, wk);
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int cc = 0; cc < 6; cc++) st[kP * (6 * r + cc)] = WV[3 * r] * wk[3 * cc] + WV[3 * r + 1] * wk[3 * cc + 1] + WV[3 * r + 2] * wk[3 * cc + 2];
        }
      }
      if (!filled) {
#pragma unroll
        for (int q = 0; q < 36; q++) st[kP * q] = 0.0;
      }
      seg_accumulate<36, kOffThreads>(sv, min(kOffThreads, o1 - base), acc, true);
    }
    if (t < 36) d.S[(size_t)(6 * jf + t / 6) * d.n + 6 * kf + t % 6] = acc;
    __syncthreads();  // s_idx is rewritten next
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
// CSR by point on the device (ptam_bundle_begin, the usual case: the measurement list is sorted by (camera,
// point), MapMaker.cc:871-882).  Then list order, ascending camera id and ascending list index coincide inside
// every point, so the slots of a point can be claimed in any order (integer atomics) and put in order afterwards:
// the result does not depend on the schedule.
//   k_ba_csr_count  thread per measurement: measurements per point
//   k_ba_csr_scan   one CTA: pt_off = exclusive scan; the counters are zeroed again for the fill; also
//                   sum k (k - 1) / 2, an upper bound of the co-visible triples (overflow check on the host)
//   k_ba_csr_fill   thread per measurement: claims a slot of its point
//   k_ba_csr_sort   thread per point: its slots by ascending list index (insertion sort, heap sort for long
//                   lists), then the camera of every slot
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_csr_count(const int* m_pt, int M, int* cnt) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) atomicAdd(&cnt[m_pt[m]], 1);
}

__global__ void __launch_bounds__(1024) k_ba_csr_scan(int* cnt, int* off, int P, long long* pairs_bound) {
  __shared__ int ws[32];
  __shared__ long long wp[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kPer = 8;
  if (threadIdx.x == 0) carry = 0;
  long long pairs = 0;
  __syncthreads();
  for (int base = 0; base < P; base += blockDim.x * kPer) {
    const int i0 = base + threadIdx.x * kPer;
    int c[kPer], s0 = 0;
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      c[q] = i0 + q < P ? cnt[i0 + q] : 0;
      if (i0 + q < P) cnt[i0 + q] = 0;
      s0 += c[q];
      pairs += (long long)c[q] * (c[q] - 1) / 2;
    }
    int inc = s0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = ws[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFull, w, o);
        if (lane >= o) w += u;
      }
      ws[lane] = w;
    }
    __syncthreads();
    int run = carry + (warp ? ws[warp - 1] : 0) + inc - s0;
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      if (i0 + q < P) off[i0 + q] = run;
      run += c[q];
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[P] = carry;
#pragma unroll
  for (int o = 16; o; o >>= 1) pairs += __shfl_xor_sync(kFull, pairs, o);
  if (lane == 0) wp[warp] = pairs;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < 32; w++) t += wp[w];
    *pairs_bound = t;
  }
}

__global__ void __launch_bounds__(256) k_ba_csr_fill(const int* m_pt, int M, const int* off, int* cur, int* pt_meas) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int i = m_pt[m];
  pt_meas[off[i] + atomicAdd(&cur[i], 1)] = m;
}

__global__ void __launch_bounds__(256) k_ba_csr_sort(const int* off, int P, int* pt_meas, const int* m_cam, int* pt_cam) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  int* a = pt_meas + off[i];
  const int k = off[i + 1] - off[i];
  if (k <= 32) {
    for (int q = 1; q < k; q++) {
      const int v = a[q];
      int r = q - 1;
      while (r >= 0 && a[r] > v) { a[r + 1] = a[r]; r--; }
      a[r + 1] = v;
    }
  } else {  // heap sort
    auto sift = [&](int root, int end) {
      for (;;) {
        int child = 2 * root + 1;
        if (child >= end) break;
        if (child + 1 < end && a[child] < a[child + 1]) child++;
        if (a[root] >= a[child]) break;
        const int t = a[root]; a[root] = a[child]; a[child] = t;
        root = child;
      }
    };
    for (int st = k / 2 - 1; st >= 0; st--) sift(st, k);
    for (int end = k - 1; end > 0; end--) {
      const int t = a[0]; a[0] = a[end]; a[end] = t;
      sift(0, end);
    }
  }
  for (int q = 0; q < k; q++) pt_cam[off[i] + q] = m_cam[a[q]];
}

__global__ void __launch_bounds__(256) k_ba_iota(int* a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// ---------------------------------------------------------------------------------------------
// The pair-major list behind k_ba_schur_off (GenerateOffDiagScripts regrouped by camera pair,
// Bundle.cc:572-599), built on the device once per Compute, without a sort:
//   k_ba_pair_count  thread per point: +1 (integer atomics: exact) for every pair of free cameras observing it
//   k_ba_pair_scan   one CTA: blk_off = exclusive scan of the counts; the non-empty pairs, long ones first
//   k_ba_pair_fill   CTA per non-empty pair (j, k): walks the shorter of the two cameras' point-ordered lists,
//                    keeps the points the other camera observes too (a point's list has <= a few entries) and
//                    compacts them in order: the triples of a pair come out by ascending point id by construction.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ba_pair_count(BundleDev d, int* blk_cnt) {
  const int i = d.p_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.p_hi) return;
  const int o0 = d.pt_off[i], o1 = d.pt_off[i + 1];
  for (int a = o0 + 1; a < o1; a++) {
    const int jrow = d.cam_row[d.pt_cam[a]];
    if (jrow < 0) continue;
    const int jf = jrow / 6;
    for (int b = o0; b < a; b++) {
      const int krow = d.cam_row[d.pt_cam[b]];
      if (krow >= 0) atomicAdd(&blk_cnt[(size_t)jf * (jf - 1) / 2 + krow / 6], 1);
    }
  }
}

// info[0] = non-empty pairs, info[1] = long ones among them, info[2] = triples
__global__ void __launch_bounds__(1024) k_ba_pair_scan(const int* cnt, int* off, int* nz, int* info, long long n) {
  __shared__ int ws[3][32];
  __shared__ int carry[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kPer = 8;
  for (int pass = 0; pass < 2; pass++) {  // pass 0: offsets + long pairs; pass 1: the short pairs behind the long ones
    if (threadIdx.x == 0) { carry[0] = 0; carry[1] = pass ? info[1] : 0; carry[2] = 0; }
    __syncthreads();
    for (long long base = 0; base < n; base += (long long)blockDim.x * kPer) {
      const long long i0 = base + (long long)threadIdx.x * kPer;
      int c[kPer], s0 = 0, s1 = 0;
#pragma unroll
      for (int q = 0; q < kPer; q++) {
        c[q] = i0 + q < n ? cnt[i0 + q] : 0;
        s0 += c[q];
        s1 += pass ? (c[q] > 0 && c[q] < kLongBlock) : (c[q] >= kLongBlock);
      }
      int i0s = s0, i1s = s1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u0 = __shfl_up_sync(kFull, i0s, o), u1 = __shfl_up_sync(kFull, i1s, o);
        if (lane >= o) { i0s += u0; i1s += u1; }
      }
      if (lane == 31) { ws[0][warp] = i0s; ws[1][warp] = i1s; }
      __syncthreads();
      if (warp == 0) {
        int w0 = ws[0][lane], w1 = ws[1][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u0 = __shfl_up_sync(kFull, w0, o), u1 = __shfl_up_sync(kFull, w1, o);
          if (lane >= o) { w0 += u0; w1 += u1; }
        }
        ws[0][lane] = w0; ws[1][lane] = w1;
      }
      __syncthreads();
      int e0 = carry[0] + (warp ? ws[0][warp - 1] : 0) + i0s - s0;
      int e1 = carry[1] + (warp ? ws[1][warp - 1] : 0) + i1s - s1;
#pragma unroll
      for (int q = 0; q < kPer; q++) {
        if (i0 + q < n) {
          if (!pass) off[i0 + q] = e0;
          const bool take = pass ? (c[q] > 0 && c[q] < kLongBlock) : (c[q] >= kLongBlock);
          if (take) nz[e1++] = (int)(i0 + q);
          e0 += c[q];
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) { carry[0] += ws[0][31]; carry[1] += ws[1][31]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (!pass) { off[n] = carry[0]; info[2] = carry[0]; info[1] = carry[1]; }
      else info[0] = carry[1];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k_ba_pair_fill(BundleDev d, int* pr_mj, int* pr_mk) {
  __shared__ int wcnt[4];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n_nz = d.pair_info[0];
  for (int idx = blockIdx.x; idx < n_nz; idx += gridDim.x) {
    const int b = d.nz_blocks[idx];
    int jf, kf;
    pair_decode(b, jf, kf);
    const int cj = d.free_cam[jf], ck = d.free_cam[kf];
    const bool walk_j = d.cam_off[cj + 1] - d.cam_off[cj] <= d.cam_off[ck + 1] - d.cam_off[ck];
    const int cw = walk_j ? cj : ck, co = walk_j ? ck : cj;  // walked camera, other camera
    const int o0 = d.cam_off[cw], o1 = d.cam_off[cw + 1];
    int at = d.blk_off[b];
    for (int base = o0; base < o1; base += blockDim.x) {
      int mw = -1, mo = -1;
      if (base + t < o1) {
        mw = d.cam_meas_pt[base + t];
        const int i = d.m_pt[mw];
        for (int o = d.pt_off[i]; o < d.pt_off[i + 1]; o++)
          if (d.pt_cam[o] == co) { mo = d.pt_meas[o]; break; }
      }
      const unsigned bal = __ballot_sync(kFull, mo >= 0);
      if (lane == 0) wcnt[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < 4; w++) { if (w < warp) before += wcnt[w]; total += wcnt[w]; }
      if (mo >= 0) {
        const int p = at + before + __popc(bal & ((1u << lane) - 1));
        pr_mj[p] = walk_j ? mw : mo;
        pr_mk[p] = walk_j ? mo : mw;
      }
      at += total;
      __syncthreads();
    }
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
      double W[18];
      ld_pairs(d.m_W + 18 * (size_t)m, W);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double a = 0;
#pragma unroll
        for (int q = 0; q < 6; q++) a += W[3 * q + r] * d.upd[row + q];
        sum[r] += a;
      }
    }
    const double v0 = d.epsB[3 * i] - sum[0], v1 = d.epsB[3 * i + 1] - sum[1], v2 = d.epsB[3 * i + 2] - sum[2];
    double Vi[10];
    ld_pairs(d.Vinv + 12 * (size_t)i, Vi);
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double u = Vi[3 * r] * v0 + Vi[3 * r + 1] * v1 + Vi[3 * r + 2] * v2;
      d.pt_pos_new[3 * i + r] = d.pt_pos[3 * i + r] + u;
      ss += u * u;
    }
  }
  const double t = block_sum(ss, sh);
  grid_sum_finish(t, d.partials, d.tickets + 1, &d.scal[4], true);  // on top of the cameras' share (k_ba_cam_update)
}

// one CTA (launched before k_ba_point_update): the new camera poses and the cameras' share of |delta|^2
__global__ void __launch_bounds__(256) k_ba_cam_update(BundleDev d) {
  __shared__ double sh[32];
  double ss = 0.0;
  for (int c = threadIdx.x; c < d.n_cams; c += blockDim.x) {
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
  if (threadIdx.x == 0) d.scal[4] = t;
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
      double fd[2];
      ld_pairs(d.m_found + 2 * (size_t)m, fd);
      const double e0 = s * (fd[0] - q.im[0]), e1 = s * (fd[1] - q.im[1]);
      e = mest_objective(e0 * e0 + e1 * e1, d.scal[1], d.est);
    }
  }
  const double t = block_sum(e, sh);
  grid_sum_finish(t, d.partials, d.tickets + 2, &d.scal[3], false);
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

// ---------------------------------------------------------------------------------------------
// Sharded handles on one node: the cross-camera reduction of S, vE over NVLink peer memory, one kernel instead of
// pack -> ncclAllReduce -> unpack.  Every rank's S / vE live in a window the other ranks map (CUDA IPC).  The lower
// triangle is cut into row ranges of equal size, one per rank; the owner of an element reads every rank's copy
// (P2P loads), adds them in rank order — so all ranks end up with the same bits — and stores the sum into every
// rank's copy (P2P stores).  Two flag barriers in peer memory frame it: "S is built everywhere" before the first
// load, "every owner has stored" before the solve.  A few scalars (the LM step's error sums and votes) ride along:
// every rank writes its own into every rank's window before signalling, and each rank adds them up itself.
// Spins are bounded: a rank that never arrives raises `err` instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
constexpr int kPeerMax = 8;
constexpr int kPeerExtra = 8;
struct PeerWin {
  double* S[kPeerMax];          // every rank's window: S [n][n], vE [n], extras [kPeerMax][kPeerExtra]
  unsigned* flags[kPeerMax];    // every rank's flag block: [2][kPeerMax] (barrier A, barrier B), written by the peers
  int rank, world, n;
  int* err;                     // local: [0] time-out code, [1] CTAs of k_ba_peer_reduce that have finished (monotonic)
};
PTAM_DEV double* peer_vE(const PeerWin& w, int p) { return w.S[p] + (size_t)w.n * w.n; }
PTAM_DEV double* peer_extra(const PeerWin& w, int p) { return w.S[p] + (size_t)w.n * w.n + w.n; }

PTAM_DEV void peer_signal(const PeerWin& w, int which, unsigned epoch, int p) {  // thread p < world
  __threadfence_system();
  *reinterpret_cast<volatile unsigned*>(&w.flags[p][which * kPeerMax + w.rank]) = epoch;
}

PTAM_DEV bool peer_wait(const PeerWin& w, int which, unsigned epoch) {  // one thread
  const volatile unsigned* f = w.flags[w.rank] + which * kPeerMax;
  for (int p = 0; p < w.world; p++) {
    long long spins = 0;
    while ((int)(f[p] - epoch) < 0) {
      if (++spins > 15000000ll) { w.err[0] = 1 + which; return false; }  // ~10 s: a peer is gone
      __nanosleep(64);
    }
  }
  __threadfence_system();
  return true;
}

// barrier B, second half: every owner has stored (the solve may read S)
__global__ void __launch_bounds__(32) k_ba_peer_wait(PeerWin w, unsigned epoch) {
  if (threadIdx.x == 0) peer_wait(w, 1, epoch);
}

// rows [row_lo, row_hi) of the lower triangle belong to this rank; rank 0 also owns vE; every rank sums the extras.
// CTA 0 first hands this rank's extras to the peers and signals barrier A ("my S is built": everything before this
// kernel in the stream); every CTA waits for A of all ranks; the CTA that finishes last signals barrier B.
__global__ void __launch_bounds__(256) k_ba_peer_reduce(PeerWin w, unsigned epoch_a, unsigned epoch_b, int row_lo, int row_hi,
                                                        const double* extra_in, double* extra_out, int n_extra) {
  __shared__ int ok;
  if (blockIdx.x == 0 && (int)threadIdx.x < w.world) {
    const int p = threadIdx.x;
    for (int q = 0; q < n_extra; q++) peer_extra(w, p)[w.rank * kPeerExtra + q] = extra_in[q];
    peer_signal(w, 0, epoch_a, p);
  }
  if (threadIdx.x == 0) ok = peer_wait(w, 0, epoch_a) ? 1 : 0;
  __syncthreads();
  const int n = w.n, W = w.world;
  if (ok) {
  // NVLink loads take microseconds: 16-byte accesses, four of them in flight per thread and peer (rows are 16-byte
  // aligned when n is even, which 6 x cameras is)
  const bool vec = (n & 1) == 0;
  for (int r = row_lo + blockIdx.x; r < row_hi; r += gridDim.x) {
    const size_t o = (size_t)r * n;
    const int pairs = vec ? (r + 1) >> 1 : 0;
    for (int i0 = threadIdx.x; i0 < pairs; i0 += 4 * blockDim.x) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < pairs ? *reinterpret_cast<const double2*>(w.S[0] + o + 2 * i) : make_double2(0.0, 0.0);
      }
      for (int p = 1; p < W; p++) {
        double2 t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * blockDim.x;
          t[u] = i < pairs ? *reinterpret_cast<const double2*>(w.S[p] + o + 2 * i) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { v[u].x += t[u].x; v[u].y += t[u].y; }
      }
      for (int p = 0; p < W; p++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * blockDim.x;
          if (i < pairs) *reinterpret_cast<double2*>(w.S[p] + o + 2 * i) = v[u];
        }
      }
    }
    for (int c = 2 * pairs + threadIdx.x; c <= r; c += blockDim.x) {  // the odd element of the row (or the whole row)
      double v = w.S[0][o + c];
      for (int p = 1; p < W; p++) v += w.S[p][o + c];
      for (int p = 0; p < W; p++) w.S[p][o + c] = v;
    }
  }
  if (w.rank == 0 && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
      double v = peer_vE(w, 0)[c];
      for (int p = 1; p < W; p++) v += peer_vE(w, p)[c];
      for (int p = 0; p < W; p++) peer_vE(w, p)[c] = v;
    }
  }
  if (blockIdx.x == gridDim.x - 1 && (int)threadIdx.x < n_extra) {
    const double* e = peer_extra(w, w.rank);
    double v = e[threadIdx.x];
    for (int p = 1; p < W; p++) v += e[p * kPeerExtra + threadIdx.x];
    extra_out[threadIdx.x] = v;
  }
  }
  // the last CTA to finish tells the peers that this rank's stores are on their way (fence, then flag)
  __shared__ int last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&w.err[1], 1) + 1) % (int)gridDim.x == 0;
  __syncthreads();
  if (last && (int)threadIdx.x < W) peer_signal(w, 1, epoch_b, threadIdx.x);
}

// (debug probe of the peer window: dst[i] = src[i] over 16-byte words, grid-stride)
__global__ void __launch_bounds__(256) k_ba_peer_copy(const double2* src, double2* dst, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// sharded handles: local erase marks -> global measurement order (merged with an all-reduce max)
__global__ void __launch_bounds__(256) k_ba_scatter_steps(const int* step, const int* gid, int* out, int n) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n && step[m] > 0) out[gid[m]] = step[m];
}

// ... and the marked measurements as (global index, step) pairs, in any order (the host sorts the few of them)
__global__ void __launch_bounds__(256) k_ba_marks_compact(const int* steps, int n, int* pairs, int* count) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int st = m < n ? steps[m] : 0;
  const unsigned bal = __ballot_sync(kFull, st > 0);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(count, __popc(bal));
  base = __shfl_sync(kFull, base, 0);
  if (st > 0) {
    const int o = base + __popc(bal & ((1u << lane) - 1u));
    pairs[2 * o] = m; pairs[2 * o + 1] = st;
  }
}

}  // namespace ptam
