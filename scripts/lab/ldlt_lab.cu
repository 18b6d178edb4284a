// Lab harness of the dense solve: per-panel launch schedule against the persistent dependency-graph kernel on the
// same systems (agreement of x and of the factors, residual, repeatability, time per solve).
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a [-DPTAM_DAG_CLOCKS] \
//        -o scripts/lab/ldlt_lab scripts/lab/ldlt_lab.cu ptam_cg_b200/csrc/ldlt.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../ptam_cg_b200/csrc/ldlt.h"

#ifdef PTAM_DAG_CLOCKS
extern "C" int ptam_debug_dag_clocks(long long* out, int reset);
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

static double urand(unsigned long long& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

int main(int argc, char** argv) {
  std::vector<int> sizes;
  for (int i = 1; i < argc; i++) sizes.push_back(std::atoi(argv[i]));
  if (sizes.empty()) sizes = {50, 64, 100, 128, 294, 700, 2994};
  const int reps = std::getenv("LAB_REPS") ? std::atoi(std::getenv("LAB_REPS")) : 10;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  ptam::LdltSolver sol;
  CK(sol.init(st));
  std::printf("dag_max_ctas %d use_dag %d\n", sol.dag_max_ctas, (int)sol.use_dag);
  for (int n : sizes) {
    const size_t nn = (size_t)n * n;
    std::vector<double> S(nn), y(n);
    unsigned long long seed = 1234567ull + n;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < i; j++) { const double v = urand(seed); S[(size_t)i * n + j] = v; S[(size_t)j * n + i] = v; }
      S[(size_t)i * n + i] = 0.6 * n + 2.0 + urand(seed);
      y[i] = urand(seed) * 10.0;
    }
    double *dS0, *dS, *dy0, *dy, *dx, *dws;
    CK(cudaMalloc(&dS0, nn * 8)); CK(cudaMalloc(&dS, nn * 8)); CK(cudaMalloc(&dy0, n * 8)); CK(cudaMalloc(&dy, n * 8)); CK(cudaMalloc(&dx, n * 8));
    CK(cudaMalloc(&dws, ptam::ldlt_workspace_doubles(n) * 8));
    CK(cudaMemcpy(dS0, S.data(), nn * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy0, y.data(), n * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<double> xs[3], Ls[3];
    double ms_mode[2] = {0, 0};
    for (int mode = 0; mode < 3; mode++) {  // 0 steps, 1 dag, 2 dag again
      const bool dag = mode > 0;
      double best = 1e30, sum = 0;
      const int R = mode == 2 ? 1 : reps;
#ifdef PTAM_DAG_CLOCKS
      ptam_debug_dag_clocks(nullptr, 1);
#endif
      for (int r = 0; r < R; r++) {
        CK(cudaMemcpyAsync(dS, dS0, nn * 8, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dy, dy0, n * 8, cudaMemcpyDeviceToDevice, st));
        CK(cudaEventRecord(e0, st));
        cudaError_t e = dag ? sol.solve_dag(dS, dy, dx, dws, n) : sol.solve_steps(dS, dy, dx, dws, n);
        if (e != cudaSuccess) { std::printf("n %d mode %d: solve failed: %s (%s)\n", n, mode, sol.err.c_str(), cudaGetErrorString(e)); return 1; }
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 || R == 1) { best = std::min(best, (double)ms); sum += ms; }
        if (dag && sol.dag_err) {
          int err = 0; CK(cudaMemcpy(&err, sol.dag_err, 4, cudaMemcpyDeviceToHost));
          if (err) { std::printf("n %d: persistent factorisation reported a TIMEOUT\n", n); return 2; }
        }
      }
      if (mode < 2) ms_mode[mode] = R > 1 ? sum / (R - 1) : sum;
      xs[mode].resize(n); Ls[mode].resize(nn);
      CK(cudaMemcpy(xs[mode].data(), dx, n * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(Ls[mode].data(), dS, nn * 8, cudaMemcpyDeviceToHost));
      if (mode < 2) std::printf("n %5d %s: mean %.4f ms  best %.4f ms\n", n, dag ? "dag  " : "steps", ms_mode[mode], best);
#ifdef PTAM_DAG_CLOCKS
      if (mode == 1) {
        long long c[32]; ptam_debug_dag_clocks(c, 0);
        const double per = 1.0 / (1.9e3 * reps * ((n + 63) / 64));
        std::printf("   chain us/panel: factor %.2f store+flags+issue %.2f load-wait %.2f solve %.2f store-rows %.2f update %.2f\n",
                    c[0] * per, c[1] * per, c[2] * per, c[3] * per, c[4] * per, c[6] * per);
        std::printf("   RU(k+2,k) us: loads %.2f solve %.2f stores %.2f wait-L %.2f loads %.2f update+stores %.2f publish %.2f\n",
                    c[16] * per, c[17] * per, c[18] * per, c[19] * per, c[20] * per, c[21] * per, c[22] * per);
        std::printf("   chain waits inside the factor: block row (RU) %.2f us, diagonal block (D) %.2f us\n", c[24] * per, c[25] * per);
        std::printf("   worker 1: %lld RU tasks %.2f us each, %lld tiles %.2f us each (%.2f us of it waiting for flags)\n", c[9],
                    c[9] ? c[8] / 1.9e3 / c[9] : 0.0, c[11], c[11] ? c[10] / 1.9e3 / c[11] : 0.0, c[11] ? c[12] / 1.9e3 / c[11] : 0.0);
      }
#endif
    }
    // agreement
    double dx_max = 0, x_max = 0, dl_max = 0, l_max = 0;
    for (int i = 0; i < n; i++) { dx_max = std::max(dx_max, std::fabs(xs[0][i] - xs[1][i])); x_max = std::max(x_max, std::fabs(xs[0][i])); }
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        dl_max = std::max(dl_max, std::fabs(Ls[0][(size_t)i * n + j] - Ls[1][(size_t)i * n + j]));
        l_max = std::max(l_max, std::fabs(Ls[0][(size_t)i * n + j]));
      }
    const bool same = std::memcmp(xs[1].data(), xs[2].data(), n * 8) == 0;
    double rmax = 0, ymax = 0;
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int j = 0; j < n; j++) s += S[(size_t)i * n + j] * xs[1][j];
      rmax = std::max(rmax, std::fabs(s - y[i])); ymax = std::max(ymax, std::fabs(y[i]));
    }
    std::printf("n %5d: |x_dag - x_steps| %.3e (|x| %.3e)  |L_dag - L_steps| %.3e (|L| %.3e)  residual %.3e  repeat %s  speed-up %.3f\n",
                n, dx_max, x_max, dl_max, l_max, rmax / ymax, same ? "bit-identical" : "DIFFERENT", ms_mode[0] / ms_mode[1]);
    cudaFree(dS0); cudaFree(dS); cudaFree(dy0); cudaFree(dy); cudaFree(dx); cudaFree(dws);
  }
  return 0;
}
