// Launch schedule of the blocked LDL^T solve (kernels: ldlt_kernels.cuh).  Compiled with FMA contraction.
#include "ldlt.h"
#include "ldlt_kernels.cuh"
#include "ldlt_dag.cuh"

#include <algorithm>
#include <cstdlib>

namespace ptam {

static int dag_flag_ints(int nblk) { return ((kDagRflag + nblk + nblk * nblk) + 3) & ~3; }

size_t ldlt_workspace_doubles(int n) {
  const int nblk = (n + kNB - 1) / kNB;
  const size_t w = (size_t)std::max(nblk, 2) * n * kNB;  // the per-panel launch schedule uses the first two panels' worth
  return w + (size_t)((n + 1) & ~1) + (size_t)(dag_flag_ints(nblk) + nblk + 4) / 2 + 2;
}

// Kernel launch, optionally with programmatic stream serialisation (see griddep_wait in ldlt_kernels.cuh)
template <class... P, class... A>
static cudaError_t launch_k(void (*kern)(P...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

#define LDLT_TRY(expr)                          \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) { err = #expr; return _e; } \
  } while (0)

cudaError_t LdltSolver::init(cudaStream_t main_stream) {
  stream = main_stream;
  use_pdl = !(std::getenv("PTAM_B200_NO_PDL"));
  LDLT_TRY(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
  LDLT_TRY(cudaFuncSetAttribute(k_ldlt_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdateSmem));
  LDLT_TRY(cudaFuncSetAttribute(k_ldlt_step, cudaFuncAttributeMaxDynamicSharedMemorySize, kPanelSmem));
  LDLT_TRY(cudaFuncSetAttribute(k_ldlt_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPanelSmem));
  LDLT_TRY(cudaFuncSetAttribute(k_ldlt_dag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDagSmem));
  use_dag = !(std::getenv("PTAM_B200_LDLT_STEPS"));
  {
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    LDLT_TRY(cudaGetDevice(&dev));
    LDLT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LDLT_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    LDLT_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ldlt_dag, kPanelThreads, kDagSmem));
    dag_max_ctas = sms * per_sm;
    if (const char* e = std::getenv("PTAM_B200_LDLT_DAG_TILES")) dag_tail_tiles = std::max(0, std::atoi(e));
    if (const char* e = std::getenv("PTAM_B200_LDLT_CTAS")) dag_max_ctas = std::min(dag_max_ctas, std::max(2, std::atoi(e)));
    if (!coop || dag_max_ctas < 2) use_dag = false;
  }
  return cudaSuccess;
}

void LdltSolver::destroy() {
  for (auto e : ev_panel) cudaEventDestroy(e);
  for (auto e : ev_tail) cudaEventDestroy(e);
  ev_panel.clear(); ev_tail.clear();
  if (stream2) cudaStreamDestroy(stream2);
  stream2 = nullptr;
}

// Blocked LDL^T on two streams.  Panel k's kernel first applies the part of panel k-1's trailing update
// that falls on its own 64 columns (fused: k_ldlt_panel), so the main stream is ONE kernel per panel;
// the rest of panel k's trailing update (column blocks from k+2 on: the tail) runs on the second stream
// after panel k.  Panel k+1 does not touch those tiles and overlaps it; panel k+2 waits for it (it reads
// tiles that tail updates and overwrites the Wp buffer it reads).  Wp is double-buffered by panel parity.
// Once the tail is small (<= kFuseTailTiles tiles) it is not launched on its own: it rides in the NEXT
// panel's launch (k_ldlt_step: panel k + tail k-1 in one grid), so the late, latency-bound part of the
// factorisation is a plain sequence of kernels on one stream without event records / waits in between.
cudaError_t LdltSolver::solve(double* S, double* y, double* x, double* ws, int n) {
  if (n & 1) { err = "odd system size (16-byte row segments are assumed; the reduced camera system has 6 rows per camera)"; return cudaErrorInvalidValue; }
  return use_dag ? solve_dag(S, y, x, ws, n) : solve_steps(S, y, x, ws, n);
}

// The factorisation (with the forward substitution) from panel k_start on as ONE cooperative launch, see
// ldlt_dag.cuh; then z = D^-1 y and the backward substitution as before.  Large systems start with the per-panel
// launch schedule: while the trailing update is what takes the time (hundreds of tiles per panel) its grids keep
// two CTAs per SM busy; the persistent kernel takes over where the chain of panels is the limit.
static int dag_first_panel(int n, int nblk, int tail_tiles) {
  int k_start = 0;
  for (int k = 0; k < nblk; k++) {  // first panel whose tail is small enough
    const int rem = n - (k + 1) * kNB;
    const int nt = rem > 0 ? (rem + kUTM - 1) / kUTM : 0;
    if (nt * nt <= tail_tiles) { k_start = k; break; }
  }
  if (k_start == 1) k_start = 2;  // the W rows of the two schedules must not overlap
  if (k_start >= nblk - 1) k_start = 0;
  return k_start;
}

int ldlt_dag_schedule(int n, int tail_tiles, std::vector<int>& off) {
  const int nblk = (n + kNB - 1) / kNB;
  const int k_start = dag_first_panel(n, nblk, tail_tiles);
  // round k (k >= k_start): D(., k), tiles of column k+2, RU(., k+1), the other tiles of panel k; before them RU(., k_start)
  off.assign(nblk + 1, 0);
  off[k_start] = std::max(0, nblk - k_start - 2);
  for (int k = k_start; k < nblk; k++) {
    const int r0 = (k + 1) * kNB;
    const int nt1 = std::max(0, (r0 < n ? (n - r0 + kUTM - 1) / kUTM : 0) - 1);
    off[k + 1] = off[k] + std::max(0, nblk - k - 2) + nt1 + std::max(0, nblk - k - 3) + nt1 * nt1;
  }
  return k_start;
}

cudaError_t LdltSolver::prepare(double* ws, int n) {
  if (n <= 0) return cudaSuccess;
  const int nblk = (n + kNB - 1) / kNB;
  if ((int)ev_panel.size() < nblk) {
    const size_t old = ev_panel.size();
    ev_panel.resize(nblk); ev_tail.resize(nblk); tail_of.resize(nblk, false);
    for (size_t k = old; k < ev_panel.size(); k++) {
      LDLT_TRY(cudaEventCreateWithFlags(&ev_panel[k], cudaEventDisableTiming));
      LDLT_TRY(cudaEventCreateWithFlags(&ev_tail[k], cudaEventDisableTiming));
    }
  }
  if (!use_dag) return cudaSuccess;
  const int k_start = dag_first_panel(n, nblk, dag_tail_tiles);
  if (dag_n != n || dag_ws != ws || dag_k_start != k_start) {
    ldlt_dag_schedule(n, dag_tail_tiles, dag_off);
    double* gd = ws + (size_t)std::max(nblk, 2) * n * kNB;
    int* task_off = reinterpret_cast<int*>(gd + ((n + 1) & ~1)) + dag_flag_ints(nblk);
    LDLT_TRY(cudaMemcpyAsync(task_off, dag_off.data(), sizeof(int) * (nblk + 1), cudaMemcpyHostToDevice, stream));
    LDLT_TRY(cudaStreamSynchronize(stream));  // the source is pageable
    dag_n = n; dag_ws = ws; dag_k_start = k_start; dag_tasks = dag_off[nblk];
  }
  return cudaSuccess;
}

cudaError_t LdltSolver::solve_dag(double* S, double* y, double* x, double* ws, int n) {
  if (n == 0) return cudaSuccess;
  const int nblk = (n + kNB - 1) / kNB;
  const int k_start = dag_first_panel(n, nblk, dag_tail_tiles);
  {
    const cudaError_t e = prepare(ws, n);
    if (e != cudaSuccess) return e;
  }
  if (k_start > 0) {
    const cudaError_t e = factor_steps(S, y, ws, n, k_start);
    if (e != cudaSuccess) return e;
  }
  LdltDagArgs a;
  a.A = S; a.y = y; a.n = n; a.nblk = nblk; a.k_start = k_start;
  a.W = ws;
  a.gd = ws + (size_t)std::max(nblk, 2) * n * kNB;
  a.flags = reinterpret_cast<int*>(a.gd + ((n + 1) & ~1));
  a.task_off = a.flags + dag_flag_ints(nblk);
  a.n_tasks = dag_tasks;
  dag_err = a.flags + kDagErr;
  const int n_flags = dag_flag_ints(nblk);
  k_ldlt_dag_init<<<(n_flags + 255) / 256, 256, 0, stream>>>(a.flags, n_flags, k_start);
  const int grid = std::min(dag_max_ctas, 1 + a.n_tasks);
  void* kargs[] = {&a};
  LDLT_TRY(cudaLaunchCooperativeKernel((const void*)k_ldlt_dag, dim3((unsigned)grid), dim3(kPanelThreads), kargs, (size_t)kDagSmem, stream));
  k_ldlt_scale<<<(n + 255) / 256, 256, 0, stream>>>(S, y, y, n);
  k_ldlt_back<<<kBackCtas, kBackThreads, 0, stream>>>(S, y, x, n);
  launches += 4;
  LDLT_TRY(cudaGetLastError());
  return cudaSuccess;
}

cudaError_t LdltSolver::solve_steps(double* S, double* y, double* x, double* Wp, int n) {
  if (n == 0) return cudaSuccess;
  dag_err = nullptr;
  const cudaError_t e = factor_steps(S, y, Wp, n, (n + kNB - 1) / kNB);
  if (e != cudaSuccess) return e;
  k_ldlt_scale<<<(n + 255) / 256, 256, 0, stream>>>(S, y, y, n);
  launches++;
  k_ldlt_back<<<kBackCtas, kBackThreads, 0, stream>>>(S, y, x, n);  // one cluster, all panels
  launches++;
  LDLT_TRY(cudaGetLastError());
  return cudaSuccess;
}

// Panels 0 .. k_end-1 by the per-panel launch schedule.  With k_end short of the last panel the trailing update of
// panel k_end-1 is completed as well: the matrix from block row / column k_end on then carries every panel before it.
cudaError_t LdltSolver::factor_steps(double* S, double* y, double* Wp, int n, int k_end) {
  const int n_panels = (n + kNB - 1) / kNB;
  if ((int)ev_panel.size() < n_panels) {
    const size_t old = ev_panel.size();
    ev_panel.resize(n_panels); ev_tail.resize(n_panels); tail_of.resize(n_panels, false);
    for (size_t k = old; k < ev_panel.size(); k++) {
      LDLT_TRY(cudaEventCreateWithFlags(&ev_panel[k], cudaEventDisableTiming));
      LDLT_TRY(cudaEventCreateWithFlags(&ev_tail[k], cudaEventDisableTiming));
    }
  }
  // y is consumed in place as the right-hand side (forward substitution rides with the panels)
  int last_tail = -1;
  bool deferred = false;      // the tail of panel k-1 waits to be launched together with panel k (k_ldlt_step)
  for (int k = 0, k0 = 0; k0 < n && k < k_end; k++, k0 += kNB) {
    const int nb = std::min(kNB, n - k0);
    const int rem = n - k0 - nb;
    double* wp = Wp + (size_t)(k & 1) * n * kNB;
    double* wprev = k > 0 ? Wp + (size_t)((k - 1) & 1) * n * kNB : nullptr;
    const int n_ctas = std::max(1, (rem + kPanelRows - 1) / kPanelRows);
    // panel k reads tiles the tail of panel k-2 updated, and overwrites the Wp buffer that tail read
    if (k >= 2 && tail_of[k - 2]) LDLT_TRY(cudaStreamWaitEvent(stream, ev_tail[k - 2], 0));
    // programmatic dependent launch along the single-stream part of the chain (no event wait in front of this step)
    const bool pdl = use_pdl && k > 0 && !(k >= 2 && tail_of[k - 2]) && !tail_of[k - 1];
    if (deferred) {
      const int nt = (n - k0 + kUTM - 1) / kUTM;  // tail of panel k-1: its trailing matrix starts at k0
      LDLT_TRY(launch_k(k_ldlt_step, n_ctas + nt * nt, kPanelThreads, kPanelSmem, stream, pdl, S, wp, wprev, y, n, k0, n_ctas));
    } else {
      LDLT_TRY(launch_k(k_ldlt_panel, n_ctas, kPanelThreads, kPanelSmem, stream, pdl, S, wp, (const double*)wprev, y, n, k0));
    }
    launches++;
    tail_of[k] = false;
    deferred = false;
    if (rem > kNB) {  // column blocks from k+2 on exist: the tail of the trailing update
      const int nt = (rem + kUTM - 1) / kUTM;
      const int n_tail = nt * (nt + 1) - nt;
      if (n_tail <= kFuseTailTiles) {
        deferred = true;  // small enough to hide behind panel k+1 at one CTA per SM: same launch, same stream
      } else {            // large: its own launch (two CTAs per SM) on the second stream
        LDLT_TRY(cudaEventRecord(ev_panel[k], stream));
        LDLT_TRY(cudaStreamWaitEvent(stream2, ev_panel[k], 0));
        k_ldlt_update<<<n_tail, 256, kUpdateSmem, stream2>>>(S, wp, n, k0, 2);
        LDLT_TRY(cudaEventRecord(ev_tail[k], stream2));
        launches++;
        tail_of[k] = true;
        last_tail = k;
      }
    }
  }
  if (last_tail >= 0) LDLT_TRY(cudaStreamWaitEvent(stream, ev_tail[last_tail], 0));
  if (k_end < n_panels && k_end > 0) {  // what panel k_end-1 still owes the trailing matrix
    const int k = k_end - 1, k0 = k * kNB;
    const double* wp = Wp + (size_t)(k & 1) * n * kNB;
    const int nt = (n - k0 - kNB + kUTM - 1) / kUTM;
    if (deferred) k_ldlt_update<<<nt * (nt + 1), 256, kUpdateSmem, stream>>>(S, wp, n, k0, 0);  // every tile
    else k_ldlt_update<<<nt, 256, kUpdateSmem, stream>>>(S, wp, n, k0, 1);                      // the tail ran on its own: column k_end only
    launches++;
  }
  LDLT_TRY(cudaGetLastError());
  return cudaSuccess;
}

}  // namespace ptam

#ifdef PTAM_DAG_CLOCKS
extern "C" int ptam_debug_dag_clocks(long long* out, int reset) {
  cudaDeviceSynchronize();
  if (reset) { long long z[32] = {}; return (int)cudaMemcpyToSymbol(ptam::g_dag_clk, z, sizeof(z)); }
  return (int)cudaMemcpyFromSymbol(out, ptam::g_dag_clk, sizeof(long long) * 32);
}
#endif
#ifdef PTAM_PANEL_DEBUG
extern "C" int ptam_debug_read(long long* out) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(out, ptam::g_dbg, sizeof(long long) * 8); }
#endif
