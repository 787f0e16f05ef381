// Microbenchmark: measured FP64 peaks of this B200 (DFMA, DMMA shapes, tanh, fp64 atomics).
// These are the denominators for the FP64 roofline (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// m16n8k4: A 16x4 (2 regs), B 4x8 (1 reg), C 16x8 (4 regs)
template<int NT>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

// m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs)
template<int NT>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

// m16n8k16: A 16x16 (8 regs), B 16x8 (4 regs)
template<int NT>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

// TF32 m16n8k8
template<int NT>
__global__ void __launch_bounds__(256) k_tf32(float* out, int iters, unsigned a, unsigned b) {
  float c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a), "r"(b), "r"(a), "r"(b), "r"(a), "r"(b));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_tanh(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += tanh(x); s2 += tanh(y); x += 1e-4; y -= 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += exp(x); s2 += exp(y); x += 1e-4; y -= 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_div(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += 1.0 / x; s2 += 1.0 / y; x += 1e-4; y += 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_log(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += log(x); s2 += log(y); x += 1e-4; y += 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_pow(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += pow(x, -2.0/3.0); s2 += pow(y, -2.0/3.0); x += 1e-4; y += 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_cbrt(double* out, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-3, y = x0 * 0.5 + threadIdx.x * 2e-3, s = 0, s2 = 0;
  for (int it = 0; it < iters; it++) { s += cbrt(x); s2 += cbrt(y); x += 1e-4; y += 1e-4; }
  if (s + s2 == 12345.678) out[0] = s;
}

// fp64 atomicAdd scatter: each thread does `per` atomics at pseudo-random addresses in [0,n)
__global__ void __launch_bounds__(256) k_atomic(double* buf, size_t n, int per, int local) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  unsigned long long s = tid * 0x9E3779B97F4A7C15ull + 12345;
  for (int i = 0; i < per; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    size_t idx = local ? ((tid * 3 + (s >> 33) % 64) % n) : ((s >> 20) % n);
    atomicAdd(&buf[idx], 1.0);
  }
}

template<class F> float timeit(F f, int rep = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < rep; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096;
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = nsm * bps; double thr = (double)grid * 256;
    float ms;
    ms = timeit([&]{ k_dfma<8><<<grid,256>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\":\"dfma_ilp8\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, thr*iters*8*2/ms/1e9);
    ms = timeit([&]{ k_dmma884<8><<<grid,256>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\":\"dmma_m8n8k4_x8\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, (double)grid*8*iters*8*(8*8*4*2.0)/ms/1e9);
    ms = timeit([&]{ k_dmma1684<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\":\"dmma_m16n8k4_x4\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, (double)grid*8*iters*4*(16*8*4*2.0)/ms/1e9);
    ms = timeit([&]{ k_dmma1688<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\":\"dmma_m16n8k8_x4\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, (double)grid*8*iters*4*(16*8*8*2.0)/ms/1e9);
    ms = timeit([&]{ k_dmma16816<4><<<grid,256>>>(out, iters, 1.0000001, 1e-9); });
    printf("{\"bench\":\"dmma_m16n8k16_x4\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, (double)grid*8*iters*4*(16*8*16*2.0)/ms/1e9);
    ms = timeit([&]{ k_tf32<8><<<grid,256>>>((float*)out, iters, 0x3f800000u, 0x3f800000u); });
    printf("{\"bench\":\"tf32_m16n8k8_x8\",\"blocks_per_sm\":%d,\"tflops\":%.3f}\n", bps, (double)grid*8*iters*8*(16*8*8*2.0)/ms/1e9);
  }
  {
    int grid = nsm * 8; double thr = (double)grid * 256; float ms;
    ms = timeit([&]{ k_tanh<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"tanh_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
    ms = timeit([&]{ k_exp<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"exp_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
    ms = timeit([&]{ k_div<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"div_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
    ms = timeit([&]{ k_log<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"log_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
    ms = timeit([&]{ k_pow<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"pow_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
    ms = timeit([&]{ k_cbrt<<<grid,256>>>(out, 1024, 0.3); });
    printf("{\"bench\":\"cbrt_f64\",\"gops\":%.2f}\n", thr*1024*2/ms/1e6);
  }
  {
    size_t n = 6500000; double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    int grid = 8192, per = 24; float ms;
    ms = timeit([&]{ k_atomic<<<grid,256>>>(buf, n, per, 0); });
    printf("{\"bench\":\"atomicAdd_f64_random_52MB\",\"gatomics_per_s\":%.2f}\n", (double)grid*256*per/ms/1e6);
    ms = timeit([&]{ k_atomic<<<grid,256>>>(buf, n, per, 1); });
    printf("{\"bench\":\"atomicAdd_f64_local\",\"gatomics_per_s\":%.2f}\n", (double)grid*256*per/ms/1e6);
  }
  return 0;
}
