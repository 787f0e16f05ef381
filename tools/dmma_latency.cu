// dmma_latency.cu -- FP64 tensor-pipe issue / dependency model of a B200 SM sub-partition (tuning aid, r2):
// cycles per mma.sync.m8n8k4.f64 per scheduler as a function of the number of INDEPENDENT accumulator chains per warp
// and of the warps per scheduler, plus the same for dependent DFMA chains.  One CTA per SM, warps = 4 * wps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_latency tools/dmma_latency.cu && tools/dmma_latency
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double* out, int iters, double a, double b, long long* clk) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) dmma(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int CH>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* clk) {
  double c[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) c[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// chain phase of the adjoint: CH independent DMMA chains plus one dependent DFMA chain of the same length interleaved
template <int CH>
__global__ void k_mix(double* out, int iters, double a, double b, long long* clk) {
  double c[CH][2], f = threadIdx.x;
#pragma unroll
  for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) dmma(c[i][0], c[i][1], a, b);
    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = f;
#pragma unroll
  for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int CH>
static void run(const char* what, int kind, double* out, long long* clk) {
  const int iters = 4096;
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 32 * 4 * wps;
    for (int rep = 0; rep < 2; rep++) {
      if (kind == 0) k_dmma<CH><<<148, threads>>>(out, iters, 1.0000001, 1e-9, clk);
      else if (kind == 1) k_dfma<CH><<<148, threads>>>(out, iters, 1.0000001, 1e-9, clk);
      else k_mix<CH><<<148, threads>>>(out, iters, 1.0000001, 1e-9, clk);
      cudaDeviceSynchronize();
    }
    long long c;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double per_instr_per_sched = (double)c / ((double)iters * CH * wps);
    printf("{\"kind\": \"%s\", \"chains_per_warp\": %d, \"warps_per_scheduler\": %d, \"clk_per_instr_per_scheduler\": %.2f, \"clk_per_chain_step\": %.1f}\n",
           what, CH, wps, per_instr_per_sched, (double)c / iters);
  }
}

int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 64); cudaMalloc(&clk, 64);
  run<1>("dmma", 0, out, clk); run<2>("dmma", 0, out, clk); run<3>("dmma", 0, out, clk); run<4>("dmma", 0, out, clk);
  run<6>("dmma", 0, out, clk); run<7>("dmma", 0, out, clk); run<8>("dmma", 0, out, clk); run<14>("dmma", 0, out, clk);
  run<1>("dfma", 1, out, clk); run<2>("dfma", 1, out, clk); run<4>("dfma", 1, out, clk); run<8>("dfma", 1, out, clk); run<14>("dfma", 1, out, clk);
  run<6>("dmma6+dfma1", 2, out, clk); run<7>("dmma7+dfma1", 2, out, clk);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
