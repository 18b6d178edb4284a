// Host interface of the dense solve of path B (reference src/Bundle.cc:457-458:
// Cholesky<>(mS).backsub(vE), TooN's square-root-free LDL^T).  Implemented in ldlt.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ptam {

// Doubles of workspace the solver needs for an n x n system: the W = -(L21 D1) rows of every panel, the
// reciprocals of D, the dependency flags and the task table of the persistent factorisation.
size_t ldlt_workspace_doubles(int n);

// Task table of the persistent factorisation (ldlt_dag.cuh): first panel of the persistent kernel and, from it on,
// the first ticket of every round.  Pure host arithmetic.
int ldlt_dag_schedule(int n, int tail_tiles, std::vector<int>& task_off);

struct LdltSolver {
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream: the caller's; stream2: the solver's own
  std::vector<cudaEvent_t> ev_panel, ev_tail;
  std::vector<char> tail_of;
  std::string err;
  int64_t launches = 0;
  bool use_pdl = true;  // programmatic dependent launch between the steps of the factorisation
  // the factorisation as one persistent launch (ldlt_dag.cuh); off: one launch per panel (PTAM_B200_LDLT_STEPS=1)
  bool use_dag = true;
  int dag_max_ctas = 0;             // CTAs of k_ldlt_dag that are resident at once (cooperative launch)
  int dag_n = -1, dag_k_start = -1, dag_tasks = 0; // the system the task table in the workspace was written for
  int dag_tail_tiles = 120;         // panels whose tail has more 128x64 tiles than this keep the per-panel schedule
  const double* dag_ws = nullptr;
  std::vector<int> dag_off;
  const int* dag_err = nullptr;     // device flag: a dependency of the last persistent factorisation timed out

  // `main_stream` is the stream the rest of the LM step runs on.  Returns cudaSuccess or the failing call's code.
  cudaError_t init(cudaStream_t main_stream);
  void destroy();
  // Solves S x = y for the symmetric S (n x n row-major, LOWER triangle read, overwritten by the L / D
  // factors) with y consumed in place.  A non-positive-definite S yields inf / NaN as in the reference.
  cudaError_t solve(double* S, double* y, double* x, double* workspace, int n);
  // Everything solve() would do only once for this (workspace, n) -- the upload of the task table, the events of the
  // per-panel schedule -- so that the solve itself is nothing but launches (it can then be captured into a CUDA graph).
  cudaError_t prepare(double* workspace, int n);
  cudaError_t solve_steps(double* S, double* y, double* x, double* workspace, int n);
  cudaError_t solve_dag(double* S, double* y, double* x, double* workspace, int n);
  cudaError_t factor_steps(double* S, double* y, double* workspace, int n, int k_end);
};

}  // namespace ptam
