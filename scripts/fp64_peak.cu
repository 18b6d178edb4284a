// Micro-benchmark of the B200's FP64 paths ("the build must measure a DMMA peak").
//   latencies (clock64, one warp): dependent DADD / DMUL / DFMA / IEEE division / DMMA m8n8k4 chains
//   throughput (all SMs, CUDA events): independent DFMA chains, independent mma.sync.m8n8k4.f64 from registers
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp64_peak scripts/fp64_peak.cu ; prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k_lat(double* out, long long* cyc, double x, double y, int iters) {
  double a = x, b = y;
  long long t[7];
  t[0] = clock64();
  for (int i = 0; i < iters; i++) a = __dadd_rn(a, b);
  t[1] = clock64();
  for (int i = 0; i < iters; i++) a = __dmul_rn(a, b);
  t[2] = clock64();
  for (int i = 0; i < iters; i++) a = __fma_rn(a, b, b);
  t[3] = clock64();
  for (int i = 0; i < iters; i++) a = 1.0 / a + b;   // division + add
  t[4] = clock64();
  double c0 = a, c1 = b;
  for (int i = 0; i < iters; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(x), "d"(y));
  t[5] = clock64();
  for (int i = 0; i < iters; i++) a = __dadd_rn(1.0 / a, b) ;
  t[6] = clock64();
  if (threadIdx.x == 0) { for (int i = 0; i < 6; i++) cyc[i] = t[i + 1] - t[i]; }
  out[threadIdx.x] = a + c0 + c1;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, double x, double y, int iters) {
  double a[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) a[c] = x + c + threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = __fma_rn(a[c], y, x);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) k_dmma(double* out, double x, double y, int iters) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { c0[c] = x + c; c1[c] = y + c + threadIdx.x; }
  const double fa = x + threadIdx.x * 1e-3, fb = y - threadIdx.x * 1e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[c]), "+d"(c1[c]) : "d"(fa), "d"(fb));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += c0[c] + c1[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 256 * 4);
  cudaMalloc(&cyc, sizeof(long long) * 8);
  const int it = 4096;
  k_lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, it);
  k_lat<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, it);
  long long h[6];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, p.multiProcessorCount);
  printf(" \"latency_cycles\": {\"dadd\": %.1f, \"dmul\": %.1f, \"dfma\": %.1f, \"ddiv_plus_dadd\": %.1f, \"dmma_m8n8k4\": %.1f},\n",
         (double)h[0] / it, (double)h[1] / it, (double)h[2] / it, (double)h[3] / it, (double)h[4] / it);
  const int iters = 20000;
  const int grid = p.multiProcessorCount * 8;
  double best_fma = 0, best_mma = 0;
  {
    const double ms = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, 1.0, 0.999, iters); }, 5);
    best_fma = 2.0 * 8 * iters * (double)grid * 256 / (ms * 1e-3) / 1e12;
  }
  {
    const double ms4 = time_ms([&] { k_dmma<4><<<grid, 256>>>(out, 1.0, 0.999, iters); }, 5);
    const double ms8 = time_ms([&] { k_dmma<8><<<grid, 256>>>(out, 1.0, 0.999, iters); }, 5);
    const double t4 = 2.0 * 256 * 4 * iters * (double)grid * 8 / (ms4 * 1e-3) / 1e12;  // 8 warps per CTA, 256 MAC per mma
    const double t8 = 2.0 * 256 * 8 * iters * (double)grid * 8 / (ms8 * 1e-3) / 1e12;
    best_mma = t4 > t8 ? t4 : t8;
    printf(" \"dmma_tflops_4chains\": %.2f, \"dmma_tflops_8chains\": %.2f,\n", t4, t8);
  }
  printf(" \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f,\n", best_fma, best_mma);
  printf(" \"how\": \"dependent chains of %d ops timed with clock64 in one warp; throughput: %d CTAs x 256 threads, independent register chains, best of 5 CUDA-event timings\"}\n", it, grid);
  return 0;
}
