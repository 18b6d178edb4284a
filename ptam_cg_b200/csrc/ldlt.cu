// Launch schedule of the blocked LDL^T solve (kernels: ldlt_kernels.cuh).  Compiled with FMA contraction.
#include "ldlt.h"
#include "ldlt_kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace ptam {

size_t ldlt_workspace_doubles(int n) { return 2 * (size_t)n * kNB; }

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
cudaError_t LdltSolver::solve(double* S, double* y, double* x, double* Wp, int n) {
  if (n == 0) return cudaSuccess;
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
  for (int k = 0, k0 = 0; k0 < n; k++, k0 += kNB) {
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
  k_ldlt_scale<<<(n + 255) / 256, 256, 0, stream>>>(S, y, y, n);
  launches++;
  k_ldlt_back<<<kBackCtas, kBackThreads, 0, stream>>>(S, y, x, n);  // one cluster, all panels
  launches++;
  LDLT_TRY(cudaGetLastError());
  return cudaSuccess;
}

}  // namespace ptam

#ifdef PTAM_PANEL_DEBUG
extern "C" int ptam_debug_read(long long* out) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(out, ptam::g_dbg, sizeof(long long) * 8); }
#endif
