// Host interface of the dense solve of path B (reference src/Bundle.cc:457-458:
// Cholesky<>(mS).backsub(vE), TooN's square-root-free LDL^T).  Implemented in ldlt.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ptam {

// Doubles of workspace the solver needs for an n x n system (the W = -(L21 D1) panels, double-buffered).
size_t ldlt_workspace_doubles(int n);

struct LdltSolver {
  cudaStream_t stream = nullptr, stream2 = nullptr;  // stream: the caller's; stream2: the solver's own
  std::vector<cudaEvent_t> ev_panel, ev_tail;
  std::vector<char> tail_of;
  std::string err;
  int64_t launches = 0;
  bool use_pdl = true;  // programmatic dependent launch between the steps of the factorisation

  // `main_stream` is the stream the rest of the LM step runs on.  Returns cudaSuccess or the failing call's code.
  cudaError_t init(cudaStream_t main_stream);
  void destroy();
  // Solves S x = y for the symmetric S (n x n row-major, LOWER triangle read, overwritten by the L / D
  // factors) with y consumed in place.  A non-positive-definite S yields inf / NaN as in the reference.
  cudaError_t solve(double* S, double* y, double* x, double* workspace, int n);
};

}  // namespace ptam
