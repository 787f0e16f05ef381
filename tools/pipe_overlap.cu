// Do DMMA (FP64 tensor sub-pipe) and DFMA (FP64 pipe) overlap on B200, or share one datapath?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_mix(double* out, int iters, double a, double b, int dmma_warps) {
  const int warp = threadIdx.x >> 5;
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  if (warp < dmma_warps) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 8192, grid = p.multiProcessorCount * 2;
  for (int dw = 0; dw <= 8; dw += 2) {
    k_mix<<<grid, 256>>>(out, iters, 1.0000001, 1e-9, dw); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k_mix<<<grid, 256>>>(out, iters, 1.0000001, 1e-9, dw); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double dmma_tf = (double)grid * dw * iters * 8 * 512.0 / (best * 1e-3) / 1e12;
    double dfma_tf = (double)grid * (8 - dw) * 32 * iters * 16 * 2.0 / (best * 1e-3) / 1e12;
    printf("{\"dmma_warps\": %d, \"dfma_warps\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sum\": %.2f}\n", dw, 8 - dw, best, dmma_tf, dfma_tf, dmma_tf + dfma_tf);
  }
  return 0;
}
