// pipe_mix.cu -- does interleaving DFMA (FP64 vector) and DMMA (FP64 tensor) instructions from DIFFERENT warps of one
// scheduler cost more than the sum of their stand-alone pipe times on B200?  (tuning aid, r2)
// One CTA of 16 warps per SM (4 per scheduler); `kd` of the 4 warps of every scheduler issue independent DMMA chains,
// the others independent DFMA chains, all for a fixed number of clock cycles; the counts give the effective cost.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_mix tools/pipe_mix.cu && tools/pipe_mix
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(512) k_mix(unsigned long long* cnt, long long window, int kd, int warps_per_sched, double a, double b, double* out) {
  const int warp = threadIdx.x >> 5;
  const int slot = warp >> 2;          // warps w, w+4, w+8, w+12 share scheduler w & 3
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  unsigned long long n = 0;
  if (slot >= warps_per_sched) {
  } else if (slot < kd) {
    while (clock64() - t0 < window) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
      n += 32;
    }
  } else {
    while (clock64() - t0 < window) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i][0]) : "d"(a), "d"(b));
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i][1]) : "d"(a), "d"(b));
        }
      n += 64;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cnt[warp] = n;
}

int main() {
  unsigned long long* cnt; double* out;
  cudaMalloc(&cnt, 16 * 8); cudaMalloc(&out, 64);
  const long long window = 4000000;
  for (int wps = 2; wps <= 4; wps += 2)
    for (int kd = 0; kd <= wps; kd++) {
      cudaMemset(cnt, 0, 16 * 8);
      k_mix<<<148, 512>>>(cnt, window, kd, wps, 1.0000001, 1e-9, out);
      cudaDeviceSynchronize();
      unsigned long long h[16];
      cudaMemcpy(h, cnt, sizeof(h), cudaMemcpyDeviceToHost);
      // scheduler 0 hosts warps 0, 4, 8, 12
      double nd = 0, nf = 0;
      for (int s = 0; s < wps; s++) { if (s < kd) nd += (double)h[4 * s]; else nf += (double)h[4 * s]; }
      const double standalone = (16.0 * nd + 2.18 * nf) / window;
      printf("{\"warps_per_scheduler\": %d, \"dmma_warps\": %d, \"dfma_warps\": %d, \"dmma_per_kclk\": %.2f, \"dfma_per_kclk\": %.2f, "
             "\"pipe_time_at_standalone_costs_over_window\": %.3f}\n", wps, kd, wps - kd, 1e3 * nd / window, 1e3 * nf / window, standalone);
    }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
