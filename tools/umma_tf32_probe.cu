// umma_tf32_probe.cu -- does tcgen05.mma kind::tf32 accept the operand layout the TF32 adjoint's weight-gradient
// contraction would feed it?  ("next step".)  Standalone, one CTA:
//   D[64 x 56] (TMEM, fp32) = sum over 128 rows r of A[r][m] * B[r][n]      (dW = d^T h: rows on K)
// Both operands are MN-major (for a fixed row the slots are contiguous), stored as 4-slot-wide planes
//   A: [16 planes][128 rows][4 floats],  B: [14 planes][128 rows][4 floats]
// which is the canonical no-swizzle MN-major UMMA layout ((T,1,m),(8,k)):((1,T,SBO),(T,LBO)) with T = 4 floats:
// consecutive K rows 16 B apart, SBO = plane stride (2 048 B), LBO = 128 B (CUTLASS cute/atom/mma_traits_sm100.hpp).
// 16 MMAs of K = 8 rows; the TMEM tile is read back with tcgen05.ld by all four warps, dumped whole, and the host
// locates the M = 64 accumulator rows in it (which lanes hold which rows is part of what is being probed).
// Every wait is a bounded spin: a wrong descriptor must end in an error code, not in a hung GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_tf32_probe tools/umma_tf32_probe.cu && tools/umma_tf32_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 64, N = 56, K = 128, NPA = M / 4, NPB = N / 4, TCOLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in 16 B
// units, version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version_
  return d;                 // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE
}

// mode 0: MN-major planes (LBO = 128 B, SBO = plane stride); mode 1: the same with LBO / SBO swapped;
// mode 2: K-major canonical ((8,m),(T,2)):((T,SBO),(1,LBO)) -- blocks [k/8][k half][m/8][m%8][4 floats]
__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out,
                                               int* __restrict__ status, int mode) {
  extern __shared__ __align__(128) float smem_dyn[];
  float* sA = smem_dyn;
  float* sB = smem_dyn + NPA * K * 4;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < NPA * K * 4; i += 128) sA[i] = A[i];
  for (int i = tid; i < NPB * K * 4; i += 128) sB[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> tensor-core (async proxy) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) status[1] = (int)tmem;
  {   // pre-fill the accumulator tile with a recognisable pattern (lane + column / 1024): checks the tcgen05.st/ld path
      // and tells "the MMA wrote nothing" apart from "the MMA wrote zeros"
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c++) {
      const uint32_t val = __float_as_uint((float)tid + (float)c / 1024.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(val) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (tid == 0 && status[2] == 0) {
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at bit 4, a/b_format TF32 (2) at bits 7 / 10,
    // a_major / b_major = MN (1) at bits 15 / 16, n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24
    const uint32_t major = mode == 2 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (major << 15) | (major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ks++) {
      uint64_t da, db;
      if (mode == 0) {
        da = make_desc(smem_u32(sA) + ks * 128, 128, K * 16);
        db = make_desc(smem_u32(sB) + ks * 128, 128, K * 16);
      } else if (mode == 1) {
        da = make_desc(smem_u32(sA) + ks * 128, K * 16, 128);
        db = make_desc(smem_u32(sB) + ks * 128, K * 16, 128);
      } else {
        da = make_desc(smem_u32(sA) + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
        db = make_desc(smem_u32(sB) + ks * 2 * (N / 8) * 128, (N / 8) * 128, 128);
      }
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  } else if (tid == 0) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait for the commit
  uint32_t done = 0;
  for (long long spin = 0; spin < (1LL << 24) && !done; spin++)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(&bar))
                 : "memory");
  if (!done) {
    if (tid == 0) *status = -1;   // the MMA never completed
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every warp reads its own 32 lanes x 64 columns
    uint32_t v[64];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 8)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[c0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]), "=r"(v[c0 + 6]),
                     "=r"(v[c0 + 7])
                   : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 64; c++) out[tid * 64 + c] = __uint_as_float(v[c]);   // out[lane 0..127][column 0..63]
    if (tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS));
}

static float tf32_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  std::vector<float> a(K * M), b(K * N);   // logical a[r][m], b[r][n]
  srand(1);
  for (auto& x : a) x = tf32_round((float)rand() / RAND_MAX - 0.5f);
  for (auto& x : b) x = tf32_round((float)rand() / RAND_MAX - 0.5f);
  std::vector<float> pa(NPA * K * 4), pb(NPB * K * 4);
  const int mode = getenv("PROBE_MODE") ? atoi(getenv("PROBE_MODE")) : 0;
  for (int r = 0; r < K; r++) {
    if (mode != 2) {
      for (int m = 0; m < M; m++) pa[((m / 4) * K + r) * 4 + m % 4] = a[r * M + m];
      for (int n = 0; n < N; n++) pb[((n / 4) * K + r) * 4 + n % 4] = b[r * N + n];
    } else {   // [k/8][k half][m/8][m%8][k%4]
      for (int m = 0; m < M; m++) pa[((((r / 8) * 2 + (r % 8) / 4) * (M / 8) + m / 8) * 8 + m % 8) * 4 + r % 4] = a[r * M + m];
      for (int n = 0; n < N; n++) pb[((((r / 8) * 2 + (r % 8) / 4) * (N / 8) + n / 8) * 8 + n % 8) * 4 + r % 4] = b[r * N + n];
    }
  }
  std::vector<double> ref(M * N, 0.0);
  for (int r = 0; r < K; r++)
    for (int m = 0; m < M; m++)
      for (int n = 0; n < N; n++) ref[m * N + n] += (double)a[r * M + m] * b[r * N + n];
  float *dA, *dB, *dO;
  int* dS;
  cudaMalloc(&dA, pa.size() * 4); cudaMalloc(&dB, pb.size() * 4); cudaMalloc(&dO, 128 * 64 * 4); cudaMalloc(&dS, 16);
  cudaMemcpy(dA, pa.data(), pa.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 128 * 64 * 4); cudaMemset(dS, 0, 16);
  if (getenv("PROBE_NO_MMA")) { int one = 1; cudaMemcpy(dS + 2, &one, 4, cudaMemcpyHostToDevice); }
  const int smem = (NPA + NPB) * K * 4 * (int)sizeof(float);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(dA, dB, dO, dS, mode);
  cudaError_t e = cudaDeviceSynchronize();
  int sts[4] = {0, 0, 0, 0};
  std::vector<float> out(128 * 64);
  cudaMemcpy(sts, dS, 16, cudaMemcpyDeviceToHost);
  const int st = sts[0];
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
  printf("{\"mode\": %d, \"cuda\": \"%s\", \"status\": %d, \"tmem_base\": %d, \"out[0][0..3]\": [%g, %g, %g, %g], \"out[33][5]\": %g", mode, cudaGetErrorString(e), st,
         sts[1], out[0], out[1], out[2], out[3], out[33 * 64 + 5]);
  if (e == cudaSuccess && st == 1) {
    // where is accumulator row m?  try the candidate lane maps and report the best
    const char* names[3] = {"lane = m", "lane = 32 (m / 16) + m % 16", "lane = 2 m"};
    double best = 1e300;
    int bi = -1;
    for (int cand = 0; cand < 3; cand++) {
      double err = 0.0;
      for (int m = 0; m < M; m++) {
        const int lane = cand == 0 ? m : cand == 1 ? 32 * (m / 16) + m % 16 : 2 * m;
        for (int n = 0; n < N; n++) err = fmax(err, fabs((double)out[lane * 64 + n] - ref[m * N + n]));
      }
      if (err < best) { best = err; bi = cand; }
    }
    int nz = 0;
    for (int l = 0; l < 128; l++) { bool any = false; for (int c = 0; c < 64; c++) any |= out[l * 64 + c] != 0.f; nz += any; }
    printf(", \"best_lane_map\": \"%s\", \"max_abs_err\": %.3e, \"nonzero_lanes\": %d, \"ref_max\": %.3f", names[bi], best, nz,
           fabs(ref[0]) > 0 ? 0.0 + [&] { double mx = 0; for (double v : ref) mx = fmax(mx, fabs(v)); return mx; }() : 0.0);
  }
  printf("}\n");
  return 0;
}
