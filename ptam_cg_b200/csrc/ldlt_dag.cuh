// Dense solve of path B as ONE persistent launch (reference src/Bundle.cc:457-458, TooN Cholesky<>).
//
// The blocked LDL^T of ldlt_kernels.cuh walks the panels as a sequence of grids: every step waits for ALL CTAs of
// the previous one.  Here the same block operations are tasks of a dependency graph, handed to resident CTAs and
// ordered by release / acquire flags in global memory, so that only the chain
//     F(k)  factor diagonal block k            ->  R(k+1,k)  solve block row k+1 against it
//                                              ->  U(k+1,k+1;k)  bring diagonal block k+1 up to date  ->  F(k+1)
// is sequential, inside ONE CTA (block 0) with all of its operands staying in shared memory.  Everything else
// happens beside the chain on the other SMs:
//     RU(i,k), i >= k+2   row solve of block (i,k) against the published L11 / D, y_i -= L_ik y_k, W_ik = -L_ik D_k,
//                         then the update of block (i,k+1) (the next panel's column: the chain needs it first)
//     D(i,k), i >= k+2    update of the diagonal block (i,i) (D(k+2,k) is the other block the chain needs first)
//     T(I,J;k), J >= k+2  128x64 tile of the trailing update below the diagonal blocks, on the f64 tensor pipe
//                         (DMMA), as in k_ldlt_update
// Tasks are numbered in rounds and fetched through one atomic ticket.  Round k: D(.,k) and the tiles of panel k's
// first tail column (J = k+2: the chain and the next row solves need them), then RU(.,k+1), then the other tiles
// of panel k, so that the row solves of a panel are long done when a CTA reaches its tiles.  A task only ever waits for tasks
// with a smaller number or for the chain, and the chain only for tasks of earlier panels, so with all CTAs
// resident (cooperative launch) the earliest unfinished task can always run.
// Flags (ints, set to k_start before the launch):
//     fdone        panels factored and published (L11, D, 1/D, y_k)
//     rflag[i]     panels block row i has been solved against (L_ik, W_ik stored, y_i updated)
//     uflag[i][j]  panels whose update block (i,j) has received
// The updates of one block are applied in panel order whoever runs them, so the result does not depend on the
// schedule (repeat runs are bit-identical).  W is kept per panel (n x 64 each): no buffer is reused while a tile
// of an earlier panel may still read it.
#pragma once
#include "ldlt_kernels.cuh"

namespace ptam {

struct LdltDagArgs {
  double* A;        // n x n row-major, lower triangle
  double* W;        // [panel][n][64]   -(L D) rows
  double* y;        // right-hand side, consumed in place
  double* gd;       // 1 / D, published with each panel
  int* flags;       // ticket, error, fdone, pad, rflag[nblk], uflag[nblk][nblk]
  const int* task_off;  // first task number of every round (nblk + 1 entries, from k_start on)
  int n, nblk, n_tasks;
  int k_start;      // panels before it are factored and fully applied already
};
constexpr int kDagTicket = 0, kDagErr = 1, kDagFdone = 2, kDagRflag = 4;
constexpr long long kDagSpinCycles = 4000000000ll;  // ~2 s: a wait that long is a lost dependency, not a slow one
constexpr int kDagSmem = (5 * kNB * kLda + 2 * kNB * 8 + 3 * kNB) * (int)sizeof(double) + 32;
static_assert((kUTM + kUTN) * kLds <= 4 * kNB * kLda, "the tile operands alias the four panel buffers");

#ifdef PTAM_DAG_CLOCKS
__device__ long long g_dag_clk[32];
#define DAG_T(k) if (threadIdx.x == 0) { const long long now = clock64(); g_dag_clk[k] += now - t_prev; t_prev = now; }
#else
#define DAG_T(k)
#endif

PTAM_DEV int* dag_rflag(const LdltDagArgs& p, int i) { return p.flags + kDagRflag + i; }
PTAM_DEV int* dag_uflag(const LdltDagArgs& p, int i, int j) { return p.flags + kDagRflag + p.nblk + i * p.nblk + j; }

PTAM_DEV int dag_ld_relaxed(const int* f) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}

// One thread waits until *f >= v.  Gives up (and makes every other wait of the grid give up) after
// kDagSpinCycles: the solve then reports an error instead of hanging the device.
PTAM_DEV void dag_spin(const int* f, int v, int* gerr, volatile int* s_abort) {
  // relaxed polls, ONE acquire fence at the end (an acquire load per poll costs a cache invalidation each time)
  if (dag_ld_relaxed(f) < v) {
    const long long t0 = clock64();
    unsigned it = 0;
    while (dag_ld_relaxed(f) < v) {
      if ((++it & 255u) == 0 && (dag_ld_relaxed(gerr) != 0 || clock64() - t0 > kDagSpinCycles)) {
        atomicExch(gerr, 1);
        *s_abort = 1;
        return;
      }
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// Up to five flags, polled by the first lanes of five different warps at once; true = all reached.
PTAM_DEV bool dag_wait(const LdltDagArgs& p, volatile int* s_abort, const int* f0, int v0, const int* f1 = nullptr, int v1 = 0,
                       const int* f2 = nullptr, int v2 = 0, const int* f3 = nullptr, int v3 = 0, const int* f4 = nullptr, int v4 = 0) {
  const int tid = threadIdx.x;
  int* gerr = p.flags + kDagErr;
  if (tid == 0) { if (f0) dag_spin(f0, v0, gerr, s_abort); }
  else if (tid == 32) { if (f1) dag_spin(f1, v1, gerr, s_abort); }
  else if (tid == 64) { if (f2) dag_spin(f2, v2, gerr, s_abort); }
  else if (tid == 96) { if (f3) dag_spin(f3, v3, gerr, s_abort); }
  else if (tid == 128) { if (f4) dag_spin(f4, v4, gerr, s_abort); }
  __syncthreads();
  return *s_abort == 0;
}

// All stores of the CTA before this call are visible to whoever acquires the flag.
PTAM_DEV void dag_publish(int* f, int v, int* f2 = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) {  // the barrier orders the CTA's stores before this thread's release (cumulative)
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(v) : "memory");
    if (f2) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f2), "r"(v) : "memory");
  }
}

// Rows below a factored block, w L11^T = a: four threads per row (r = tid / 4), thread q = tid % 4 owns the 16
// columns c = q (mod 4); see ldlt_panel_body for the scheme (same expressions, same order per element).
PTAM_DEV void dag_row_solve(const double (*a)[kLda], const double (*wo)[kLda], double (&x)[kNB / 4]) {
  const int tid = threadIdx.x, r = tid >> 2, q = tid & 3, lane = tid & 31;
#pragma unroll
  for (int j = 0; j < kNB / 4; j++) x[j] = wo[r][4 * j + q];
#pragma unroll
  for (int c = 0; c < kNB; c += 2) {
    const int jc = c >> 2;
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
}

// c[i][j] += sum_q w[ty + 16 i][q] l[tx + 16 j][q]   (4x4 register tiles over K = 64)
template <bool kLowerOnly>
PTAM_DEV void dag_rank64(double (&c)[4][4], const double (*w)[kLda], const double (*l)[kLda]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll 2
  for (int q = 0; q < kNB; q += 2) {
    double2 vw[4], vl[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      vw[i] = *reinterpret_cast<const double2*>(&w[ty + 16 * i][q]);
      vl[i] = *reinterpret_cast<const double2*>(&l[tx + 16 * i][q]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (!kLowerOnly || j <= i) { c[i][j] += vw[i].x * vl[j].x; c[i][j] += vw[i].y * vl[j].y; }
  }
}

struct DagSmem {
  double (*a)[kLda];
  double (*lh)[kLda];
  double (*wd)[kLda];
  double (*wo)[kLda];
  double (*nx)[kLda];  // the chain's next diagonal block
  double (*us)[8];
  double *y1, *y2, *dinv;
  unsigned long long* mbar;
  volatile int* ctl;  // [0] ticket, [1] abort
};

// ---- the chain (block 0)
PTAM_DEV void dag_chain(const LdltDagArgs& p, const DagSmem& s) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, n = p.n;
  const unsigned mbar_a = smem_u32(s.mbar);
  unsigned phase = 0;
#ifdef PTAM_DAG_CLOCKS
  long long t_prev = clock64();
#endif
  {
    const int f0 = p.k_start * kNB, nb = min(kNB, n - f0);
    for (int e = tid; e < kNB * kNB; e += kPanelThreads) {
      const int r = e >> 6, c = e & 63;
      s.a[r][c] = (r < nb && c <= r) ? p.A[(size_t)(f0 + r) * n + f0 + c] : (r == c ? 1.0 : 0.0);
    }
    if (tid < kNB) s.y1[tid] = tid < nb ? p.y[f0 + tid] : 0.0;
  }
  __syncthreads();
  for (int k = p.k_start; k < p.nblk; k++) {
    const int k0 = k * kNB, nb = min(kNB, n - k0);
    // The factorisation; then the next block row -- block (k+1,k) must carry panel k-1, the next diagonal block
    // too -- is fetched by the TMA engine while the last thread publishes F(k).  (Waiting for the flags and issuing
    // the copies from the idle last warp DURING the last sub-panel was tried: two fences and the issue take 2.5 us,
    // the slot is 0.9 us long.)
    const int i0 = k0 + kNB, rows = min(kNB, n - i0);
    const bool more = k + 1 < p.nblk;
    auto fetch_next = [&](int t, int nthreads) {  // threads t = 0 .. nthreads-1 of one group, after the flags were acquired
      if (t == 0) {
        const unsigned bytes = (unsigned)rows * (unsigned)((kNB + rows) * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
      }
      asm volatile("fence.proxy.async;" ::: "memory");  // written by other SMs (generic proxy), read by the copy engine
      for (int q = t; q < 2 * kNB; q += nthreads) {
        const int r = q & 63;
        const bool diag = q >= kNB;
        double* dst = diag ? s.nx[r] : s.wo[r];
        if (r < rows) {
          const double* src = p.A + (size_t)(i0 + r) * n + (diag ? i0 : k0);
          const unsigned bytes = (unsigned)((diag ? rows : kNB) * sizeof(double));
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mbar_a) : "memory");
        } else {
          for (int c = 0; c < kNB; c++) dst[c] = 0.0;
        }
      }
    };
    block_ldlt64(s.a, s.us, s.dinv, s.y1, nb);
    DAG_T(0)
    for (int e = tid; e < nb * kNB; e += kPanelThreads) {
      const int r = e >> 6, c = e & 63;
      if (c <= r) p.A[(size_t)(k0 + r) * n + k0 + c] = s.a[r][c];
    }
    if (tid < nb) { p.y[k0 + tid] = s.y1[tid]; p.gd[k0 + tid] = s.dinv[tid]; }
    __syncthreads();
    // F(k) is published by the last thread; the others go on (its warp joins the row solve after the fence)
    if (tid == kPanelThreads - 1) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags + kDagFdone), "r"(k + 1) : "memory");
    if (!more) break;
    if (tid < 192) {
      if (tid == 0) dag_spin(dag_uflag(p, k + 1, k), k, p.flags + kDagErr, s.ctl + 1);
      else if (tid == 32) dag_spin(dag_uflag(p, k + 1, k + 1), k, p.flags + kDagErr, s.ctl + 1);
      asm volatile("bar.sync 5, 192;" ::: "memory");
      if (s.ctl[1] == 0) {
        if (tid < 2 * kNB) fetch_next(tid, 2 * kNB);
        if (tid >= 2 * kNB) { const int r = tid - 2 * kNB; s.y2[r] = r < rows ? __ldcg(p.y + i0 + r) : 0.0; }  // complete through panel k - 1
      }
    }
    __syncthreads();
    if (s.ctl[1] != 0) return;
    DAG_T(1)
    {
      unsigned ok = 0;
      while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(mbar_a), "r"(phase) : "memory");
      phase ^= 1u;
    }
    DAG_T(2)
    double x[kNB / 4];
    dag_row_solve(s.a, s.wo, x);
    {
      const int r = tid >> 2, q = tid & 3;
      double dot = 0.0;
#pragma unroll
      for (int j = 0; j < kNB / 4; j++) {
        const int c = 4 * j + q;
        const double l = x[j] * s.dinv[c];  // value * (1 / d), as TooN does
        s.wd[r][c] = -x[j];
        s.lh[r][c] = l;
        dot += l * s.y1[c];
      }
      dot += __shfl_xor_sync(kFull, dot, 1);
      dot += __shfl_xor_sync(kFull, dot, 2);
      if (q == 0) s.y2[r] -= dot;  // y_(k+1) - L y_k
    }
    __syncthreads();
    DAG_T(3)
    // the solved rows to global memory, whole rows per warp.  rflag[k+1] is released by the last warp once all warps
    // have issued their stores (they only ARRIVE at the named barrier and go on with the update; the last warp
    // joins them after the fence of its release)
    for (int e = tid; e < rows * kNB; e += kPanelThreads) {
      const int r = e >> 6, c = e & 63;
      p.A[(size_t)(i0 + r) * n + k0 + c] = s.lh[r][c];  // W_(k+1)k stays here: no tile below the diagonal reads it
    }
    if (tid < kPanelThreads - 32) asm volatile("bar.arrive 6, 256;" ::: "memory");
    else {
      asm volatile("bar.sync 6, 256;" ::: "memory");
      if (tid == kPanelThreads - 1) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(dag_rflag(p, k + 1)), "r"(k + 1) : "memory");
    }
    DAG_T(4)
    // U(k+1,k+1;k) and hand-over: the diagonal block of the next panel
    double cd[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        cd[i][j] = (r < rows && c <= r) ? s.nx[r][c] : (r == c ? 1.0 : 0.0);
      }
    dag_rank64<true>(cd, s.wd, s.lh);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = ty + 16 * i, c = tx + 16 * j;
        s.a[r][c] = c <= r ? cd[i][j] : 0.0;  // nobody reads `a` any more: the row solve ended before the last barrier
      }
    if (tid < kNB) s.y1[tid] = s.y2[tid];
    __syncthreads();
    DAG_T(6)
  }
}

// ---- RU(i,k): block row i against panel k, then the update of block (i,k+1).  RU(k+2,k) is what the chain waits
// for next, so the task is laid out like the chain itself: operands by the TMA engine, results to global memory as
// whole rows, the first release (rflag) by the last warp while the first warps already poll the flags of the update.
PTAM_DEV void dag_mbar_wait(unsigned mbar_a, unsigned& phase) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mbar_a), "r"(phase) : "memory");
  phase ^= 1u;
}

// threads 0..127: 64 full rows of block A into dst_a, `rows_b` rows of block B into dst_b (the others zeroed)
PTAM_DEV void dag_fetch2(double (*dst_a)[kLda], const double* src_a, double (*dst_b)[kLda], const double* src_b, int n, int rows_b, unsigned mbar_a) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    const unsigned bytes = (unsigned)(kNB + rows_b) * kNB * (unsigned)sizeof(double);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
  }
  if (tid < 2 * kNB) {
    const int r = tid & 63;
    const bool second = tid >= kNB;
    double* dst = second ? dst_b[r] : dst_a[r];
    if (!second || r < rows_b) {
      const double* src = (second ? src_b : src_a) + (size_t)r * n;
      asm volatile("fence.proxy.async;" ::: "memory");  // written by other SMs (generic proxy), read by the copy engine
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dst)), "l"(src), "r"((unsigned)(kNB * sizeof(double))), "r"(mbar_a) : "memory");
    } else {
      for (int c = 0; c < kNB; c++) dst[c] = 0.0;
    }
  }
}

PTAM_DEV void dag_task_ru(const LdltDagArgs& p, const DagSmem& s, int k, int i, unsigned& phase) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, n = p.n;
  const int k0 = k * kNB, i0 = i * kNB, j0 = k0 + kNB, rows = min(kNB, n - i0);
  const unsigned mbar_a = smem_u32(s.mbar);
  if (!dag_wait(p, s.ctl + 1, p.flags + kDagFdone, k + 1, dag_uflag(p, i, k), k)) return;
#ifdef PTAM_DAG_CLOCKS
  long long t_prev = clock64();
#define RU_T(q) if (i == k + 2 && threadIdx.x == 0) { const long long now = clock64(); atomicAdd((unsigned long long*)&g_dag_clk[16 + q], (unsigned long long)(now - t_prev)); t_prev = now; }
#else
#define RU_T(q)
#endif
  dag_fetch2(s.a, p.A + (size_t)k0 * n + k0, s.wo, p.A + (size_t)i0 * n + k0, n, rows, mbar_a);
  if (tid >= 2 * kNB && tid < 3 * kNB) { const int r = tid - 2 * kNB; s.dinv[r] = __ldcg(p.gd + k0 + r); s.y1[r] = __ldcg(p.y + k0 + r); }
  dag_mbar_wait(mbar_a, phase);
  __syncthreads();
  RU_T(0)
  double x[kNB / 4];
  dag_row_solve(s.a, s.wo, x);
  {
    const int r = tid >> 2, q = tid & 3, row = i0 + r;
    double dot = 0.0;
#pragma unroll
    for (int j = 0; j < kNB / 4; j++) {
      const int c = 4 * j + q;
      const double l = x[j] * s.dinv[c];  // value * (1 / d), as TooN does
      s.wd[r][c] = -x[j];
      s.nx[r][c] = l;
      dot += l * s.y1[c];
    }
    dot += __shfl_xor_sync(kFull, dot, 1);
    dot += __shfl_xor_sync(kFull, dot, 2);
    if (q == 0 && row < n) __stcg(p.y + row, __ldcg(p.y + row) - dot);
  }
  __syncthreads();
  RU_T(1)
  for (int e = tid; e < rows * kNB; e += kPanelThreads) {
    const int r = e >> 6, c = e & 63;
    p.A[(size_t)(i0 + r) * n + k0 + c] = s.nx[r][c];
    p.W[((size_t)k * n + i0 + r) * kNB + c] = s.wd[r][c];
  }
  if (tid < kPanelThreads - 32) asm volatile("bar.arrive 6, 256;" ::: "memory");
  else {
    asm volatile("bar.sync 6, 256;" ::: "memory");
    if (tid == kPanelThreads - 1) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(dag_rflag(p, i)), "r"(k + 1) : "memory");
  }
  // block (i,k+1) += W_ik L_(k+1)k^T: it must carry panel k-1, and the chain must have stored L_(k+1)k
  RU_T(2)
  if (!dag_wait(p, s.ctl + 1, dag_rflag(p, k + 1), k + 1, dag_uflag(p, i, k + 1), k)) return;
  RU_T(3)
  dag_fetch2(s.lh, p.A + (size_t)j0 * n + k0, s.wo, p.A + (size_t)i0 * n + j0, n, rows, mbar_a);
  dag_mbar_wait(mbar_a, phase);
  __syncthreads();
  RU_T(4)
  double co[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ii++)
#pragma unroll
    for (int j = 0; j < 4; j++) co[ii][j] = s.wo[ty + 16 * ii][tx + 16 * j];
  dag_rank64<false>(co, s.wd, s.lh);
#pragma unroll
  for (int ii = 0; ii < 4; ii++)
#pragma unroll
    for (int j = 0; j < 4; j++) s.wo[ty + 16 * ii][tx + 16 * j] = co[ii][j];  // the thread's own elements
  __syncthreads();
  for (int e = tid; e < rows * kNB; e += kPanelThreads) {
    const int r = e >> 6, c = e & 63;
    p.A[(size_t)(i0 + r) * n + j0 + c] = s.wo[r][c];
  }
  RU_T(5)
  dag_publish(dag_uflag(p, i, k + 1), k + 1);
  RU_T(6)
}

// ---- T(I,J;k): rows of blocks I, I+1 x columns of block J,  C += W_k[rows] L_Jk^T  (DMMA, see ldlt_update_body)
PTAM_DEV void dag_task_tile(const LdltDagArgs& p, const DagSmem& s, int k, int bi, int bj, unsigned& phase) {
  // row tile bi >= 1 (block rows I, I+1) has column blocks bj = 1 .. 2 bi: the blocks strictly below the diagonal
  // (the diagonal blocks are D tasks); with bj = 2 bi (J == I) only block (I+1,I) is left of the tile
  const int n = p.n, k0 = k * kNB, r0 = k0 + kNB;
  const int I = k + 1 + 2 * bi, J = k + 1 + bj;
  const int i0 = r0 + bi * kUTM, j0 = r0 + bj * kUTN;
  const bool two = I + 1 < p.nblk, low = J < I;
  if (!low && !two) return;
#ifdef PTAM_DAG_CLOCKS
  const long long t_in = clock64();
#endif
  if (!dag_wait(p, s.ctl + 1, dag_rflag(p, I), k + 1, two ? dag_rflag(p, I + 1) : nullptr, k + 1, dag_rflag(p, J), k + 1,
                low ? dag_uflag(p, I, J) : nullptr, k, two ? dag_uflag(p, I + 1, J) : nullptr, k))
    return;
#ifdef PTAM_DAG_CLOCKS
  if (blockIdx.x == 1 && threadIdx.x == 0) g_dag_clk[12] += clock64() - t_in;
#endif
  double* sW = reinterpret_cast<double*>(s.a);  // [kUTM][kLds]
  double* sL = sW + kUTM * kLds;                // [kUTN][kLds]
  const double* Wp = p.W + (size_t)k * n * kNB;
  double* A = p.A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned mbar_a = smem_u32(s.mbar);
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
        // the rows were written by other SMs (generic proxy) and are read by the copy engine (async proxy)
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"((unsigned)(kNB * sizeof(double))), "r"(mbar_a) : "memory");
      } else {
        for (int c = 0; c < kNB; c++) dst[c] = 0.0;
      }
    }
  }
  const int wm = warp & 3, wn = warp >> 2;
  const bool active = low || wm >= 2;  // the warp's 32 rows lie in a block row below block column J
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
        if (i < n && j + 1 <= i) v = __ldcg(reinterpret_cast<const double2*>(Ci + j));
        else if (i < n && j <= i) v.x = __ldcg(Ci + j);
        acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
      }
    }
  }
  {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(mbar_a), "r"(phase) : "memory");
    phase ^= 1u;
  }
  __syncthreads();
  if (active) {
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
  dag_publish(low ? dag_uflag(p, I, J) : dag_uflag(p, I + 1, J), k + 1, (low && two) ? dag_uflag(p, I + 1, J) : nullptr);
}

// ---- D(i,k): diagonal block (i,i) += W_ik L_ik^T, lower triangle (i >= k+2; block (k+1,k+1) is the chain's own).
// D(k+2,k) is the other thing the chain waits for: a 64x64 task of its own instead of a corner of a 128x64 tile.
PTAM_DEV void dag_task_diag(const LdltDagArgs& p, const DagSmem& s, int k, int i, unsigned& phase) {
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, n = p.n;
  const int k0 = k * kNB, i0 = i * kNB, rows = min(kNB, n - i0);
  const unsigned mbar_a = smem_u32(s.mbar);
  if (!dag_wait(p, s.ctl + 1, dag_rflag(p, i), k + 1, dag_uflag(p, i, i), k)) return;
  if (tid == 0) {
    const unsigned bytes = (unsigned)rows * (unsigned)((2 * kNB + rows) * sizeof(double));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
  }
  if (tid < 3 * kNB) {
    const int r = tid & 63, which = tid >> 6;  // 0: W_ik, 1: L_ik, 2: the block itself
    double* dst = which == 0 ? s.wd[r] : which == 1 ? s.lh[r] : s.wo[r];
    if (r < rows) {
      const double* src = which == 0 ? p.W + ((size_t)k * n + i0 + r) * kNB
                        : which == 1 ? p.A + (size_t)(i0 + r) * n + k0 : p.A + (size_t)(i0 + r) * n + i0;
      const unsigned bytes = (unsigned)((which == 2 ? rows : kNB) * sizeof(double));
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mbar_a) : "memory");
    } else {
      for (int c = 0; c < kNB; c++) dst[c] = 0.0;
    }
  }
  dag_mbar_wait(mbar_a, phase);
  __syncthreads();
  double cd[4][4];
#pragma unroll
  for (int ii = 0; ii < 4; ii++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * ii, c = tx + 16 * j;
      cd[ii][j] = (r < rows && c <= r) ? s.wo[r][c] : 0.0;
    }
  dag_rank64<true>(cd, s.wd, s.lh);
#pragma unroll
  for (int ii = 0; ii < 4; ii++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = ty + 16 * ii, c = tx + 16 * j;
      if (r < rows && c <= r) p.A[(size_t)(i0 + r) * n + i0 + c] = cd[ii][j];  // 16 consecutive columns per half warp
    }
  dag_publish(dag_uflag(p, i, i), k + 1);
}

// ticket and error flag cleared, every progress flag at k_start
__global__ void k_ldlt_dag_init(int* flags, int n_ints, int k_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ints) flags[i] = i < kDagFdone ? 0 : k_start;
}

__global__ void __launch_bounds__(kPanelThreads, 1) k_ldlt_dag(LdltDagArgs p) {
  extern __shared__ __align__(16) unsigned char dag_smem[];
  DagSmem s;
  s.a = reinterpret_cast<double (*)[kLda]>(dag_smem);
  s.lh = s.a + kNB; s.wd = s.lh + kNB; s.wo = s.wd + kNB;
  s.nx = s.wo + kNB;
  s.us = reinterpret_cast<double (*)[8]>(s.nx + kNB);
  s.y1 = reinterpret_cast<double*>(s.us + 2 * kNB);
  s.y2 = s.y1 + kNB; s.dinv = s.y2 + kNB;
  s.mbar = reinterpret_cast<unsigned long long*>(s.dinv + kNB);
  s.ctl = reinterpret_cast<volatile int*>(s.mbar + 1);
  const int tid = threadIdx.x;
  if (tid == 0) {
    s.ctl[0] = 0; s.ctl[1] = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s.mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (blockIdx.x == 0) { dag_chain(p, s); return; }
  // The ticket of the next task is fetched by thread 32 (thread 0 may still sit in the fence of its last release).
  unsigned phase = 0;
  int k = p.k_start;
  const int t_first = p.task_off[p.k_start];  // the tasks before it: RU(., k_start)
  if (tid == 32) s.ctl[0] = atomicAdd(p.flags + kDagTicket, 1);
  while (true) {
    __syncthreads();
    const int t = s.ctl[0];
    const bool stop = t >= p.n_tasks || s.ctl[1] != 0;
    __syncthreads();  // everybody has read the slot: thread 32 rewrites it right after the task
    if (stop) break;
    int ru_k = -1, ru_i = 0, dg_i = -1, bi = 0, bj = 0;
    if (t < t_first) { ru_k = p.k_start; ru_i = p.k_start + 2 + t; }
    else {
      while (t >= p.task_off[k + 1]) k++;
      int local = t - p.task_off[k];
      const int r0 = (k + 1) * kNB;
      const int nt1 = max(0, (r0 < p.n ? (p.n - r0 + kUTM - 1) / kUTM : 0) - 1);  // row tiles with blocks below the diagonal
      const int n_dg = max(0, p.nblk - k - 2);                                    // D(., k)
      const int n_ru = max(0, p.nblk - k - 3);                                    // RU(., k+1)
      if (local < n_dg) dg_i = k + 2 + local;
      else if (local < n_dg + nt1) { bi = 1 + local - n_dg; bj = 1; }
      else if (local < n_dg + nt1 + n_ru) { ru_k = k + 1; ru_i = k + 3 + (local - n_dg - nt1); }
      else {
        local -= n_dg + nt1 + n_ru;  // row tile bi >= 1 has 2 bi - 1 tiles with bj >= 2; (bi - 1)^2 of them before it
        bi = 1 + (int)sqrt((double)local);
        while ((bi - 1) * (bi - 1) > local) bi--;
        while (bi * bi <= local) bi++;
        bj = local - (bi - 1) * (bi - 1) + 2;
      }
    }
#ifdef PTAM_DAG_CLOCKS
    const long long t_task = clock64();
#endif
    if (ru_k >= 0) dag_task_ru(p, s, ru_k, ru_i, phase);
    else if (dg_i >= 0) dag_task_diag(p, s, k, dg_i, phase);
    else dag_task_tile(p, s, k, bi, bj, phase);
    if (tid == 32) s.ctl[0] = atomicAdd(p.flags + kDagTicket, 1);  // overlaps the fence of thread 0's release
#ifdef PTAM_DAG_CLOCKS
    if (blockIdx.x == 1 && tid == 0) { const int w = ru_k >= 0 ? 8 : 10; g_dag_clk[w] += clock64() - t_task; g_dag_clk[w + 1] += 1; }
#endif
  }
}

}  // namespace ptam
