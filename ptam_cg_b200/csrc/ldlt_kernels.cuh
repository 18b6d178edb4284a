// sm_100a kernels of the dense solve of path B:  delta_a = S^-1 vE  (reference src/Bundle.cc:457-458,
// TooN Cholesky<>: square-root-free LDL^T).  Own translation unit (ldlt.cu) because it is compiled WITH FMA
// contraction, while the per-measurement passes of bundle.cu are compiled without (they mirror the
// reference's expression order bit for bit).
#pragma once
#include "common.cuh"

namespace ptam {

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

// Programmatic dependent launch: the steps of the factorisation are launched with programmatic stream
// serialisation, so the CTAs of step k + 1 are already resident (shared memory carved, mbarrier initialised) when
// step k retires; nothing step k wrote is read before griddep_wait().
PTAM_DEV void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PTAM_DEV void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// mma.m8n8k4 (f64): one 8x8 tile of  C -= L U^T  over the eight columns of a sub-panel.  C, L in `a`, U in `u`.
PTAM_DEV void tile_rank8(double (*a)[kLda], const double (*u)[8], int r0, int c0, int k0, int lane) {
  double2* cp = reinterpret_cast<double2*>(&a[r0 + (lane >> 2)][c0 + 2 * (lane & 3)]);
  double2 c = *cp;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const double fa = -a[r0 + (lane >> 2)][k0 + 4 * h + (lane & 3)];
    const double fb = u[c0 + (lane >> 2)][4 * h + (lane & 3)];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c.x), "+d"(c.y) : "d"(fa), "d"(fb));
  }
  *cp = c;
}

// Warp-specialised, with look-ahead: warps 0..1 (one thread per row) factor sub-panel s while warps 2..7 still apply
// sub-panel s - 1 to the columns further right (8x8 tiles on the f64 tensor pipe, the matrix stays in shared
// memory).  The update warps first bring the NEXT sub-panel's eight columns up to date and release the row warps
// (named barrier 1), then do the rest; the row warps hand a finished sub-panel over through named barrier 2.
// `us` is double-buffered by sub-panel parity ([2][64][8]).
PTAM_DEV void block_ldlt64(double (*a)[kLda], double (*us)[8], double* dinv, double* ysh, int nb) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool row_role = tid < kNB;
  double yr = row_role ? ysh[tid] : 0.0;  // threads 0..63 own one row of every sub-panel
  if (row_role) dinv[tid] = 1.0;
  __syncthreads();
  const int nsub = (min(nb, kNB) + 7) >> 3;  // the identity padding of a short last block needs no work
  if (row_role) {
    const int r = tid;
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
      const int c0 = 8 * s;
      double (*u)[8] = us + (s & 1) * kNB;
      if (s > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // the columns of this sub-panel carry sub-panel s - 1
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
      asm volatile("bar.sync 3, 64;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double rcp = 1.0 / dg[j][j];
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
      } else {  // rows of the finished sub-panels: their entries here are scratch, keep them finite for the tiles
        double2* row = reinterpret_cast<double2*>(&a[r][c0]);
#pragma unroll
        for (int j = 0; j < 8; j += 2) row[j >> 1] = make_double2(0.0, 0.0);
      }
      double2* urow = reinterpret_cast<double2*>(&u[r][0]);
#pragma unroll
      for (int j = 0; j < 8; j += 2) urow[j >> 1] = make_double2(uv[j], uv[j + 1]);
      ysh[r] = yr;
      asm volatile("bar.arrive 2, 256;" ::: "memory");  // sub-panel s is in shared memory
    }
  } else {
    const int uw = warp - 2;  // 0..5
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
      const int c0 = 8 * s;
      const double (*u)[8] = us + (s & 1) * kNB;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (s + 1 < nsub) {  // the next sub-panel's columns first: they are the chain
        for (int ti = s + 1 + uw; ti < nsub; ti += 6) tile_rank8(a, u, 8 * ti, c0 + 8, c0, lane);
        asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
      int q = 0;
      for (int tj = s + 2; tj < nsub; tj++)
        for (int ti = tj; ti < nsub; ti++, q++)
          if (q % 6 == uw) tile_rank8(a, u, 8 * ti, 8 * tj, c0, lane);
    }
  }
  __syncthreads();
}

// Shared memory of k_ldlt_panel (dynamic): the diagonal block, the rank-8 operand, the right-hand side and
// the reciprocals, plus three 64x64 operands of the PENDING update (see below); the third is reused for
// the updated rows of this CTA.
constexpr int kPanelSmem = (4 * kNB * kLda + 2 * kNB * 8 + 2 * kNB) * (int)sizeof(double) + 16;  // + the mbarrier of the bulk loads

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
  double* y1 = reinterpret_cast<double*>(us + 2 * kNB);  // `us` is double-buffered
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
    griddep_wait();
    griddep_launch_dependents();
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
    griddep_wait();
    griddep_launch_dependents();
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
  if (j0 >= n) { griddep_launch_dependents(); return; }  // the last row block may be short of its second diagonal column block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned mbar_a = smem_u32(mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  griddep_wait();
  griddep_launch_dependents();
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


}  // namespace ptam
