// pcx_mlp.cuh -- Field (input normalisation + tanh MLP + Dirichlet wrapper) forward and adjoint
// on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64; tcgen05 has no f64 kind).
//
// Replaces: pancax/networks/fields.py:79-101 + equinox.nn.MLP (pancax/networks/mlp.py:72-85) vmapped
// over nodes (pancax/physics_kernels/base.py:248-250) and the reverse-mode pass JAX derives for it.
//
// Layout idea (B200-first, not a translation of the vmap):
//  * a warp owns tiles of 8 rows (nodes); the whole layer chain of a tile stays in registers.
//    The m8n8k4 accumulator fragment of layer i (lane (r=l/4, c=l%4) holds row r, "slots" 2c,2c+1 of
//    each 8-wide tile) IS the A fragment of layer i+1 if the K index is permuted consistently:
//    accumulator register x of tile j, lane c  <->  feature 8j + 4x + c, so k-step ks of the next
//    layer covers features 4ks..4ks+3.  The weights are pre-packed once per step into B fragments
//    that follow this permutation (k_pack_weights), so no shuffle / shared-memory round trip sits
//    between layers.
//  * activations are written for the backward pass in that same fragment order
//    act[layer][tile][row][8 slots]  (one fully coalesced 512 B store per warp per tile) instead of
//    being recomputed (3.4 GB per time step at 2.1 M nodes: HBM is cheap, FP64 flops are not).
//  * backward: the delta chain runs like the forward pass (DMMA with transposed weight fragments);
//    the weight-gradient contraction sum_rows d_i^T h_i -- the only true dense contraction of the
//    path -- needs rows on the K axis, so d_i and h_i tiles are staged through shared memory
//    (bank-conflict-free stride) and each warp owns a fixed 8-row block of every layer's dW in
//    registers for the CTA's whole persistent loop; bias gradients come for free from a column of
//    ones in a padding slot.  Per-CTA partials are summed in fixed order by k_reduce_wgrad.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcx {

constexpr int MLP_MAX_LAYERS = 8;  // depth <= 7 hidden layers

struct MlpShape {
  int n_in, n_out, width, depth, use_final_bias;
  int NT;      // 8-wide feature tiles (width + 1 <= 8*NT: one spare slot carries the bias column)
  int KS;      // k-steps of 4 features: ceil(width / 4)
  // offsets (in doubles) into the flat equinox-order weight vector
  long long offW[MLP_MAX_LAYERS], offB[MLP_MAX_LAYERS];  // offB = -1 if the layer has no bias
  long long nweights;
  // offsets into the packed-fragment buffer
  long long pF0, pFB, pFH, pFO, pFOB, pFE, pBO, pBH, npacked;
  // input normalisation (fields.py:80-86)
  double x_min[3], x_scale[3];  // x_hat = (x - x_min) / (x_max - x_min); x_scale = x_max - x_min
  double t_min, t_scale;
  int normalize_time;
  int nd;
};

__host__ __device__ inline int slot_to_feat(int s) { return 8 * (s >> 3) + 4 * (s & 1) + ((s & 7) >> 1); }
__host__ __device__ inline int feat_to_slot(int f) { return 8 * (f >> 3) + 2 * (f & 3) + ((f & 7) >> 2); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// the same, pinned in program order relative to other volatile asm (the scheduler-token acquire / release below)
__device__ __forceinline__ void dmma_v(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void token_acquire(int* tok) {   // one lane spins, the warp follows
  if ((threadIdx.x & 31) == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(tok);
    int old;
    do {
      asm volatile("atom.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(a) : "memory");
      if (old != 0) __nanosleep(40);
    } while (old != 0);
  }
  __syncwarp();
}
__device__ __forceinline__ void token_release(int* tok) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(tok);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(0) : "memory");
  }
}

__device__ __forceinline__ void st_cs_f64x2(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_cs_f32x2_fwd(float* p, float a, float b) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// read-only global load pinned in program order relative to the other volatile asm of a kernel (the streaming
// stores): ptxas otherwise sinks a load whose value is only needed an iteration later down to its use
__device__ __forceinline__ double ld_nc_early(const double* p) {
  double r;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
// (raw - lo) / scale of fields.py:80-86, bit-identical to the plain division; a ZERO numerator (t = 0, nodes on the
// lower faces) would send the whole warp through the division's slow-path subroutine (ncu r2: a CALL on half of the row
// groups), so those lanes divide 1 instead and select 0
__device__ __forceinline__ double normalise(double raw, double lo, double scale) {
  const double num = raw - lo;
  const bool z = num == 0.0;
  const double q = (z ? 1.0 : num) / scale;
  return z ? num : q;      // keeps the sign of a zero numerator, like the division
}
__device__ __forceinline__ double2 ld_cs_f64x2(const double* p) {
  double2 r;
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Weight packing: flat equinox vector -> B fragments (one double per lane per fragment).
//   F0  [NT][32]            layer 0:   lane(r',c') = W0[feat(8jn+r')][c']                (c' < n_in)
//   FB  [depth][NT][32][2]  biases in accumulator layout: b_i[feat(8jn+2c+x)]
//   FH  [depth-1][KSP][NT][32] hidden i>=1: W_i[feat(8jn+r')][4ks+c']
//   FO  [KSP][32]           output:    W_out[r'][4ks+c']                                  (r' < n_out)
//   FOB [32][2]             output bias (if any) in accumulator layout: b_out[2c+x]
//   FE  [depth-1][KSP][4][2] hidden i>=1, rows of features 8(NT-1)+f, f < 2: W_i[8(NT-1)+f][4ks+c]  (edge tile on the
//                           FP64 vector pipe, see k_mlp_forward NEDGE)
//   BO  [NT][32]            backward top: W_out[c'][feat(8jn+r')]                          (c' < n_out)
//   BH  [depth-1][KSP][NT][32] backward hidden i>=1: W_i[4ks+c'][feat(8jn+r')]
//   BE  [depth-1][KSP][4][2] backward hidden i>=1, columns of input features 8(NT-1)+f, f < 2: W_i[4ks+c][8(NT-1)+f]
//                           (edge tile of the delta chain on the FP64 vector pipe, see k_mlp_backward BEDGE).  It has no
//                           section of its own: when the width leaves k-step KSP-1 of BH unused (KS < KSP, all zeros),
//                           BE of layer i sits in the first 8*KSP doubles of that k-step, so the adjoint kernel finds it
//                           in the shared-memory copy of BH it stages anyway.
// KSP = 2*NT (padded k-step count so offsets do not depend on width).
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_weights(MlpShape s, const double* __restrict__ w, double* __restrict__ pk) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (tid >= s.npacked) return;
  const int NT = s.NT, KSP = 2 * s.NT, W = s.width, D = s.depth;
  double val = 0.0;
  long long t = tid;
  if (t < s.pFB) {  // F0
    int lane = t & 31, jn = (int)(t >> 5);
    int r = lane >> 2, c = lane & 3;
    int o = slot_to_feat(8 * jn + r);
    if (o < W && c < s.n_in) val = w[s.offW[0] + (long long)o * s.n_in + c];
  } else if (t < s.pFH) {  // FB
    t -= s.pFB;
    int x = t & 1, lane = (t >> 1) & 31;
    long long q = t >> 6;
    int jn = (int)(q % NT), i = (int)(q / NT);
    int c = lane & 3;
    int o = slot_to_feat(8 * jn + 2 * c + x);
    if (o < W && s.offB[i] >= 0) val = w[s.offB[i] + o];
  } else if (t < s.pFO) {  // FH
    t -= s.pFH;
    int lane = t & 31;
    long long q = t >> 5;
    int jn = (int)(q % NT);
    q /= NT;
    int ks = (int)(q % KSP), i = (int)(q / KSP) + 1;
    int r = lane >> 2, c = lane & 3;
    int o = slot_to_feat(8 * jn + r), k = 4 * ks + c;
    if (o < W && k < W) val = w[s.offW[i] + (long long)o * W + k];
  } else if (t < s.pFOB) {  // FO
    t -= s.pFO;
    int lane = t & 31, ks = (int)(t >> 5);
    int r = lane >> 2, c = lane & 3;
    int k = 4 * ks + c;
    if (r < s.n_out && k < W) val = w[s.offW[D] + (long long)r * W + k];
  } else if (t < s.pFE) {  // FOB
    t -= s.pFOB;
    int x = t & 1, lane = (int)(t >> 1);
    int c = lane & 3;
    int o = 2 * c + x;
    if (o < s.n_out && s.offB[D] >= 0) val = w[s.offB[D] + o];
  } else if (t < s.pBO) {  // FE: rows of the last tile's (at most two) real features, [depth-1][KSP][4 c][2 f]
    t -= s.pFE;
    int f = t & 1, c = (int)(t >> 1) & 3;
    long long q = t >> 3;
    int ks = (int)(q % KSP), i = (int)(q / KSP) + 1;
    int o = 8 * (NT - 1) + f, k = 4 * ks + c;
    if (o < W && k < W) val = w[s.offW[i] + (long long)o * W + k];
  } else if (t < s.pBH) {  // BO
    t -= s.pBO;
    int lane = t & 31, jn = (int)(t >> 5);
    int r = lane >> 2, c = lane & 3;
    int k = slot_to_feat(8 * jn + r);
    if (c < s.n_out && k < W) val = w[s.offW[D] + (long long)c * W + k];
  } else {  // BH (+ BE in its unused last k-step)
    t -= s.pBH;
    int lane = t & 31;
    long long q = t >> 5;
    int jn = (int)(q % NT);
    q /= NT;
    int ks = (int)(q % KSP), i = (int)(q / KSP) + 1;
    int r = lane >> 2, c = lane & 3;
    int o = 4 * ks + c, k = slot_to_feat(8 * jn + r);
    if (o < W && k < W) val = w[s.offW[i] + (long long)o * W + k];
    const int e = jn * 32 + lane;       // position inside the k-step
    if (s.KS < KSP && ks == KSP - 1 && e < KSP * 8) {
      const int f = e & 1, cc = (e >> 1) & 3, kb = e >> 3;
      const int ob = 4 * kb + cc, kk = 8 * (NT - 1) + f;
      val = (ob < W && kk < W) ? w[s.offW[i] + (long long)ob * W + kk] : 0.0;
    }
  }
  pk[tid] = val;
}

struct MlpIO {
  // rows to evaluate
  const double* x;      // row r input coordinates at x[r*x_stride + j], j < nd
  long long x_stride;
  const double* trow;   // per-row time at trow[r*x_stride] or nullptr -> use `t`
  double t;
  long long n;          // number of rows
  long long nn_mod;     // > 0: rows are (time step, node) pairs: x row = r % nn_mod, t = times[r / nn_mod]
  const double* times;  // device, used when nn_mod > 0
  const double* bc_a;   // [n*n_out] or nullptr
  const double* bc_b;   // [n*n_out] or nullptr
  double* u;            // forward: [n*n_out] out (may be nullptr)
  const double* g_u;    // backward: [n*n_out] cotangent on u
  double* act;          // activations [depth][NT][rows_cap][8] (nullptr: do not store)
  float* act32;         // TF32-adjoint mode: the same planes stored as fp32 instead (pcx_mlp_tf32.cuh)
  int act32_quad;       // act32 in tcgen05 operand order [layer][row/4][tile j][slot%8][row%4] (pcx_mlp_umma.cuh)
  long long rows_cap;
  const double* pk;     // packed weights
  const float* pk32;    // TF32-adjoint mode: packed hi/lo chain weights (k_pack_weights_tf32)
  double* part;         // backward: per-CTA partial weight gradients [gridDim][part_stride]
  long long part_stride;
  unsigned int* sched;       // forward: group counter of the dynamic scheduler (zeroed before the launch) or nullptr
  int l2pf;                  // adjoint: L2 prefetch of the next tile's activations
  unsigned long long* dbg;   // tuning: per-CTA (start, end, smid) stamps [3*gridDim] or nullptr (PCX_OPT_TUNE_STAMPS)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ double mlp_input(const MlpShape& s, const MlpIO& io, long long row, int c) {
  // fields.py:80-88: inputs = hstack(x_norm, t_norm)
  if (row >= io.n) return 0.0;
  const long long xr = io.nn_mod > 0 ? row % io.nn_mod : row;
  if (c < s.nd) return (io.x[xr * io.x_stride + c] - s.x_min[c]) / s.x_scale[c];
  if (c == s.nd) {
    double t = io.nn_mod > 0 ? io.times[row / io.nn_mod] : (io.trow ? io.trow[row * io.x_stride] : io.t);
    return s.normalize_time ? (t - s.t_min) / s.t_scale : t;
  }
  return 0.0;
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk copy (TMA engine, SASS UBLKCP) and cp.async (LDGSTS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// one elected thread stages `ndbl` doubles of packed weights into shared memory; everybody waits
__device__ __forceinline__ void stage_weights(double* dst, const double* src, long long ndbl, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long bytes = ndbl * 8;
    mbar_expect_tx(bar, (uint32_t)bytes);
    for (long long o = 0; o < bytes; o += 32768) {
      const uint32_t nb = (uint32_t)((bytes - o) < 32768 ? (bytes - o) : 32768);
      bulk_g2s(reinterpret_cast<char*>(dst) + o, reinterpret_cast<const char*>(src) + o, nb, bar);
    }
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
// L2 prefetch of a contiguous global range (TMA engine, no shared memory involved)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tanh, branch-free, table driven: 15 FP64-pipe instructions (the CUDA library tanh is ~45, and on
// sm_100a every FP64 instruction is paid for in tensor-pipe time -- DMMA and DFMA share one datapath).
//   tanh|x| = 1 - 2E / (1 + E),  E = exp(-2|x|) = 2^n * 2^(j/128) * 2^(s/128),  |s| <= 1/2
// k = 128 n + j = round(-2 log2(e) 128 |x|) by the magic-number trick; 2^(j/128) from a 128-entry
// table in shared memory; 2^(s/128) by a degree-5 Taylor polynomial (truncation 5e-19); the
// reciprocal by MUFU.RCP64H (rcp.approx.ftz.f64, ~2^-20) + one cubic Newton step (e^3 ~ 1e-18).
// |abs error| <= 3e-16 on [-40, 40] (tests/test_host_and_abi.py replays the arithmetic in numpy):
// the hidden activations are O(1) and enter sums, so absolute accuracy is what matters.
// ---------------------------------------------------------------------------------------------
constexpr int EXP2_TAB_N = 128;
__device__ const double g_exp2_tab[EXP2_TAB_N] = {
    0x1.0000000000000p+0, 0x1.0163da9fb3335p+0, 0x1.02c9a3e778061p+0, 0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0, 0x1.0706b29ddf6dep+0, 0x1.0874518759bc8p+0, 0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0, 0x1.0cc922b7247f7p+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0, 0x1.12abdc06c31ccp+0, 0x1.1429aaea92de0p+0, 0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0, 0x1.18af9388c8deap+0, 0x1.1a35beb6fcb75p+0, 0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0, 0x1.1ed5022fcd91dp+0, 0x1.2063b88628cd6p+0, 0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0, 0x1.251ce4fb2a63fp+0, 0x1.26b4565e27cddp+0, 0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0, 0x1.2b87fd0dad990p+0, 0x1.2d285a6e4030bp+0, 0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0, 0x1.32170fc4cd831p+0, 0x1.33c08b26416ffp+0, 0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0, 0x1.38cae6d05d866p+0, 0x1.3a7db34e59ff7p+0, 0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0, 0x1.3fa4504ac801cp+0, 0x1.4160a21f72e2ap+0, 0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0, 0x1.46a41ed1d0057p+0, 0x1.486a2b5c13cd0p+0, 0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0, 0x1.4dcb299fddd0dp+0, 0x1.4f9b2769d2ca7p+0, 0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0, 0x1.551a4ca5d920fp+0, 0x1.56f4736b527dap+0, 0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0, 0x1.5c9268a5946b7p+0, 0x1.5e76f15ad2148p+0, 0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0, 0x1.6434634ccc320p+0, 0x1.6623882552225p+0, 0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6c012750bdabfp+0, 0x1.6dfb23c651a2fp+0, 0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0, 0x1.73f9a48a58174p+0, 0x1.75feb564267c9p+0, 0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0, 0x1.7c1ed0130c132p+0, 0x1.7e2f336cf4e62p+0, 0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0, 0x1.8471a4623c7adp+0, 0x1.868d99b4492edp+0, 0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8cf3216b5448cp+0, 0x1.8f1ae99157736p+0, 0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0, 0x1.95a44cbc8520fp+0, 0x1.97d829fde4e50p+0, 0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0, 0x1.9e86319e32323p+0, 0x1.a0c667b5de565p+0, 0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0, 0x1.a799e1330b358p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0, 0x1.b0e07298db666p+0, 0x1.b33a2b84f15fbp+0, 0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.ba5b030a1064ap+0, 0x1.bcc1e904bc1d2p+0, 0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c67f12e57d14bp+0, 0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0, 0x1.cdf0b555dc3fap+0, 0x1.d072d4a07897cp+0, 0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0, 0x1.d80e316c98398p+0, 0x1.da9e603db3285p+0, 0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e264614f5a129p+0, 0x1.e502ee78b3ff6p+0, 0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0, 0x1.ecf482d8e67f1p+0, 0x1.efa1bee615a27p+0, 0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0, 0x1.f7bfdad9cbe14p+0, 0x1.fa7c1819e90d8p+0, 0x1.fd3c22b8f71f1p+0,
};

__device__ __forceinline__ double rcp_approx_f64(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

__device__ __forceinline__ double tanh_tb(double x, const double* __restrict__ T) {
  const double a = fmin(fabs(x), 22.0);
  const double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: round-to-nearest-integer trick
  const double C = -0x1.71547652b82fep+8;         // -2 log2(e) * 128
  const double t = fma(a, C, MAGIC);
  const int k = __double2loint(t);                // <= 0
  const double s = fma(a, C, MAGIC - t);          // exact integer subtracted inside one fma
  double p = 0x1.5d87fe78a6731p-45;               // (ln2/128)^5/5!
  p = fma(p, s, 0x1.3b2ab6fba4e77p-35);
  p = fma(p, s, 0x1.c6b08d704a0c0p-26);
  p = fma(p, s, 0x1.ebfbdff82c58fp-17);
  p = fma(p, s, 0x1.62e42fefa39efp-8);
  p = fma(p, s, 1.0);
  const double m = T[k & (EXP2_TAB_N - 1)] * p;   // in [0.99, 2.01)
  const int n = k >> 7;
  const int mh = __double2hiint(m), ml = __double2loint(m);
  const double E = __hiloint2double(mh + (n << 20), ml);
  const double m2E = __hiloint2double((mh + ((n + 1) << 20)) ^ 0x80000000, ml);   // -2E
  const double den = 1.0 + E;
  const double r0 = rcp_approx_f64(den);
  const double e = fma(-den, r0, 1.0);
  const double r1 = fma(r0, fma(e, e, e), r0);
  const double q = fma(m2E, r1, 1.0);
  return copysign(q, x);
}

// Forward-kernel variant of the same tanh: 13 FP64-pipe instructions instead of 15 and a conflict-free table.
//  * the 2^(j/64) table is replicated 16 times, entry j of copy l at T[16 j + l]: the 16 lanes of a half warp (one
//    shared-memory wavefront of an LDS.64) read 16 different bank pairs whatever their j -- the 128-entry single copy
//    cost 2.75 extra wavefronts per lookup on random j (ncu r01: 73.9 M bank conflicts per forward launch);
//    64 entries: |s| <= 1/2 in units of ln2/64, degree-5 truncation (ln2/128)^6/720 = 3.5e-17 relative;
//  * the clamp |x| <= 22 is an integer min on the high word (ALU pipe), not a DSETP/select pair on the FP64 pipe;
//  * s = a C - k takes k back from the integer pipe (I2F, conversion unit) instead of a DADD on the FP64 pipe.
constexpr int EXP2_REP_N = 64, EXP2_REP_COPIES = 16;
template <bool CHEAP = true>   // false (tuning only): the table alone, clamp and s as in tanh_tb
__device__ __forceinline__ double tanh_rep(double x, const double* __restrict__ T /* + (lane & 15) */) {
  const int xh = __double2hiint(x);
  const int ah = min(xh & 0x7fffffff, 0x40360000);                 // |x| clamped to [0, 22 + 2^-16): E underflows to ~0 anyway
  const double a = CHEAP ? __hiloint2double(ah, __double2loint(x)) : fmin(fabs(x), 22.0);
  const double MAGIC = 6755399441055744.0;
  const double C = -0x1.71547652b82fep+7;                          // -2 log2(e) * 64
  const double t = fma(a, C, MAGIC);
  const int k = __double2loint(t);                                 // <= 0
  const double s = CHEAP ? fma(a, C, -(double)k) : fma(a, C, MAGIC - t);   // exact: k is the rounded product
  double p = 0x1.5d87fe78a6731p-40;                                // (ln2/64)^5/5!
  p = fma(p, s, 0x1.3b2ab6fba4e77p-31);                            // (ln2/64)^4/4!
  p = fma(p, s, 0x1.c6b08d704a0c0p-23);                            // (ln2/64)^3/3!
  p = fma(p, s, 0x1.ebfbdff82c58fp-15);                            // (ln2/64)^2/2!
  p = fma(p, s, 0x1.62e42fefa39efp-7);                             // ln2/64
  p = fma(p, s, 1.0);
  const double m = T[(k & (EXP2_REP_N - 1)) * EXP2_REP_COPIES] * p;
  const int n = k >> 6;
  const int mh = __double2hiint(m), ml = __double2loint(m);
  const double E = __hiloint2double(mh + (n << 20), ml);
  const double m2E = __hiloint2double((mh + ((n + 1) << 20)) ^ 0x80000000, ml);   // -2E
  const double den = 1.0 + E;
  const double r0 = rcp_approx_f64(den);
  const double e = fma(-den, r0, 1.0);
  const double r1 = fma(r0, fma(e, e, e), r0);
  const double q = fma(m2E, r1, 1.0);
  return __hiloint2double((__double2hiint(q) & 0x7fffffff) | (xh & 0x80000000), __double2loint(q));
}

// The same arithmetic for N independent arguments, written STAGE BY STAGE (all range reductions, all table loads, one
// Horner step for every argument at a time, all reciprocal seeds, ...): tanh is a ~20-deep dependent chain (DFMA latency
// 8.4 cycles at 2.2 cycles issue, one LDS, one MUFU, one I2F), a warp issues in order, and left to itself ptxas
// interleaved only two or three of the 14 chains of a layer -- the tanh phases of the forward kernel ran the FP64
// pipe at 38 % (knock-out timing, profiles/r02_kernel_ab_8).  Results are bit-identical to tanh_rep.
template <int N>
__device__ __forceinline__ void tanh_rep_batch(const double (&x)[N], double (&y)[N], const double* __restrict__ T /* + (lane & 15) */) {
  const double MAGIC = 6755399441055744.0;
  const double C = -0x1.71547652b82fep+7;                          // -2 log2(e) * 64
  int k[N];
  double s[N], tb[N], p[N];
#pragma unroll
  for (int e = 0; e < N; e++) {
    const int ah = min(__double2hiint(x[e]) & 0x7fffffff, 0x40360000);
    const double a = __hiloint2double(ah, __double2loint(x[e]));
    const double t = fma(a, C, MAGIC);
    k[e] = __double2loint(t);
    s[e] = fma(a, C, -(double)k[e]);
  }
#pragma unroll
  for (int e = 0; e < N; e++) tb[e] = T[(k[e] & (EXP2_REP_N - 1)) * EXP2_REP_COPIES];
#pragma unroll
  for (int e = 0; e < N; e++) p[e] = fma(0x1.5d87fe78a6731p-40, s[e], 0x1.3b2ab6fba4e77p-31);
#pragma unroll
  for (int e = 0; e < N; e++) p[e] = fma(p[e], s[e], 0x1.c6b08d704a0c0p-23);
#pragma unroll
  for (int e = 0; e < N; e++) p[e] = fma(p[e], s[e], 0x1.ebfbdff82c58fp-15);
#pragma unroll
  for (int e = 0; e < N; e++) p[e] = fma(p[e], s[e], 0x1.62e42fefa39efp-7);
#pragma unroll
  for (int e = 0; e < N; e++) p[e] = fma(p[e], s[e], 1.0);
  double E2[N], den[N], r0[N];     // -2E, 1 + E, reciprocal seed
#pragma unroll
  for (int e = 0; e < N; e++) {
    const double m = tb[e] * p[e];
    const int n = k[e] >> 6;
    const int mh = __double2hiint(m), ml = __double2loint(m);
    const double E = __hiloint2double(mh + (n << 20), ml);
    E2[e] = __hiloint2double((mh + ((n + 1) << 20)) ^ 0x80000000, ml);
    den[e] = 1.0 + E;
  }
#pragma unroll
  for (int e = 0; e < N; e++) r0[e] = rcp_approx_f64(den[e]);
#pragma unroll
  for (int e = 0; e < N; e++) den[e] = fma(-den[e], r0[e], 1.0);          // e
#pragma unroll
  for (int e = 0; e < N; e++) den[e] = fma(den[e], den[e], den[e]);       // e + e^2
#pragma unroll
  for (int e = 0; e < N; e++) r0[e] = fma(r0[e], den[e], r0[e]);          // r1
#pragma unroll
  for (int e = 0; e < N; e++) {
    const double q = fma(E2[e], r0[e], 1.0);
    y[e] = __hiloint2double((__double2hiint(q) & 0x7fffffff) | (__double2hiint(x[e]) & 0x80000000), __double2loint(q));
  }
}

// ---------------------------------------------------------------------------------------------
// Forward.  One warp = MT tiles of 8 rows; grid-stride over row groups.  The forward weight
// fragments [F0 | FB | FH | FO | FOB] are staged ONCE per CTA into shared memory with a bulk
// (TMA) copy and read conflict-free (lane-contiguous); activations are streamed to HBM (.cs).
// ---------------------------------------------------------------------------------------------
// NEDGE = 1 or 2: the last 8-wide tile holds only NEDGE real features (width 50: features 48, 49).  Its 13 DMMAs per
// hidden layer would be 6/8 padding, and DMMA and DFMA share one datapath, so that tile is computed with DFMAs
// instead: every lane already holds A-fragment element (row r, feature 4ks+c), multiplies it with the two weight
// rows W[48..49][4ks+c] (one broadcast LDS.128 per k-step), and a 2-step butterfly over the 4 lanes of a row
// finishes the sums -- 26 DFMA + 4 DADD + 8 SHFL instead of 13 DMMA (= 104 DFMA issue slots).  NEDGE = 0: off.
//
// Per-group prologue (r2): raw coordinate / time loads are issued a whole group ahead and normalised at use (the
// load -> subtract -> divide chain sat exposed in front of layer 0: 3.4 % of the samples on one DADD), the lane's
// normalisation constants are hoisted, and (time step, node) of a row is tracked incrementally instead of a 64-bit
// division and modulo per group.
// TV = 0: tanh_tb (single 128-entry table); TV = 1: tanh_rep (replicated conflict-free table, 13 FP64 instructions);
// TV = 2 (tuning): replicated table with tanh_tb's clamp / range reduction
constexpr int fwd_table_doubles(int tv) { return tv ? EXP2_REP_N * EXP2_REP_COPIES : EXP2_TAB_N; }
// LOCK = 1 (one CTA of 16 warps per SM): at most ONE warp per scheduler is inside a hidden layer's DMMA loop at a time.
// Measured (tools/pipe_mix.cu): the FP64 datapath is arbitrated in favour of DMMA -- two warps issuing DMMAs starve the
// DFMA (tanh) warps of the same scheduler to 0.4 % of their stand-alone rate, ONE leaves them 37 % of the pipe -- so
// without the token the warps of a scheduler fall into lock step (all in the tanh phase, all in the DMMA phase) and
// nothing overlaps the latency-bound parts (loads, stores, address arithmetic) of a row group.
// QUAD = 1: the fp32 activations of the TF32 mode in tcgen05 operand order (MlpIO::act32_quad); a compile-time variant so
// that the fp64 path keeps its register allocation (as a run-time branch it cost the default kernel 24 B of spills and
// 0.18 ms at C3 size)
template <int NT, int KODD, int MT, int MINB, int NEDGE, int TV = 1, int NTHR = 256, int LOCK = 0, int QUAD = 0>
__global__ void __launch_bounds__(NTHR, MINB) k_mlp_forward(const __grid_constant__ MlpShape s, const __grid_constant__ MlpIO io) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t wbar;
  __shared__ int dtoken[4];
  if (threadIdx.x < 4) dtoken[threadIdx.x] = 0;
  int* const mytoken = &dtoken[(threadIdx.x >> 5) & 3];   // warp w of a CTA runs on scheduler w & 3
  if (io.dbg && threadIdx.x == 0) { io.dbg[3 * blockIdx.x] = globaltimer_ns(); unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); io.dbg[3 * blockIdx.x + 2] = sm; }
  double* sw = reinterpret_cast<double*>(smem_raw);
  double* stab = sw + (s.pBO - s.pF0);          // exp2 table behind the weight fragments
  if constexpr (TV == 0) {
    if (threadIdx.x < EXP2_TAB_N) stab[threadIdx.x] = g_exp2_tab[threadIdx.x];
  } else {
    for (int i = threadIdx.x; i < EXP2_REP_N * EXP2_REP_COPIES; i += blockDim.x)
      stab[i] = g_exp2_tab[(i / EXP2_REP_COPIES) * (EXP2_TAB_N / EXP2_REP_N)];
  }
  stage_weights(sw, io.pk + s.pF0, s.pBO - s.pF0, &wbar);   // (contains a __syncthreads)
  const double* __restrict__ pk = sw - s.pF0;   // same offsets as the global buffer

  const int lane = threadIdx.x & 31;
  if constexpr (TV != 0) stab += lane & (EXP2_REP_COPIES - 1);
  auto act_fn = [&](double z) { if constexpr (TV == 0) return tanh_tb(z, stab); else return tanh_rep<TV == 1>(z, stab); };
  const int r = lane >> 2, c = lane & 3;
  constexpr int KSP = 2 * NT;
  constexpr int KS = 2 * NT - KODD;   // compile-time k-step count: no per-step branch, loads pipeline freely
  constexpr int NE = 2 * NT;          // A-fragment elements (k-steps' worth of features) per row tile
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ngroups = (io.n + 8 * MT - 1) / (8 * MT);
  // slots of the last tile whose features are all padding need no tanh (their pre-activation is 0)
  const int last_real_x1 = (8 * (NT - 1) + 4 < s.width);   // reg x=1 of tile NT-1 holds features 8(NT-1)+4+c

  // fields.py:80-88: inputs = hstack(x_norm, t_norm); lane c owns input column c.  (raw - in_min) / in_scale with
  // in_min = 0, in_scale = 1 where nothing is normalised (exact); padding columns read raw = 0.
  const bool is_x = c < s.nd, is_t = c == s.nd;
  const double in_min = is_x ? s.x_min[c] : (is_t && s.normalize_time) ? s.t_min : 0.0;
  const double in_scale = is_x ? s.x_scale[c] : (is_t && s.normalize_time) ? s.t_scale : 1.0;
  // Dynamic group scheduler (r2).  The two CTAs resident on an SM do not progress at the same rate -- with the static
  // grid-stride split, on EVERY SM one CTA finished after 78 % of the kernel and the other ran the last 22 % alone, at
  // half the warps per scheduler (per-CTA %globaltimer stamps, tools/cta_stamps.py) -- so warps take their next group of
  // 8*MT rows from a global counter (io.sched, zeroed before the launch), one group ahead so the atomic's latency never
  // shows.  Rows are independent: results do not depend on who computes them.  io.sched == nullptr: static stride.
  auto grab = [&](long long prev) -> long long {
    if (!io.sched) return prev + warps_total;
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(io.sched, 1u);
    return (long long)__shfl_sync(0xffffffffu, v, 0) + warps_total;   // the first warps_total groups are handed out statically
  };
  // rows of the NEXT group of this lane: row, and (time step, node) when rows are (time step, node) pairs
  long long rowN[MT], nodeN[MT];
  int tqN[MT];
  const bool small = io.n < 0x7fffffffLL;
  auto seek = [&](long long gnext) {
#pragma unroll
    for (int m = 0; m < MT; m++) {
      rowN[m] = gnext * (8 * MT) + 8 * m + r;
      if (io.nn_mod > 0 && rowN[m] < io.n) {
        if (small) tqN[m] = (int)((unsigned)rowN[m] / (unsigned)io.nn_mod);
        else tqN[m] = (int)(rowN[m] / io.nn_mod);
        nodeN[m] = rowN[m] - (long long)tqN[m] * io.nn_mod;
      } else {
        tqN[m] = 0;
        nodeN[m] = rowN[m];
      }
    }
  };
  auto load_raw = [&](int m) -> double {
    if (rowN[m] >= io.n) return in_min;      // normalises to 0
    if (is_x) return ld_nc_early(io.x + nodeN[m] * io.x_stride + c);
    if (is_t) return io.nn_mod > 0 ? ld_nc_early(io.times + tqN[m]) : (io.trow ? ld_nc_early(io.trow + rowN[m] * io.x_stride) : io.t);
    return 0.0;
  };
  long long g = warp_id;
  long long gnext = grab(g);
  seek(g);
  double rawn[MT];
#pragma unroll
  for (int m = 0; m < MT; m++) rawn[m] = load_raw(m);

  for (; g < ngroups;) {
    const long long gnext2 = grab(gnext);   // two groups ahead: the atomic's round trip is covered by a whole group
    const long long row0 = g * (8 * MT);
    double h[MT][NT][2], acc[MT][NT][2];
    // ---- layer 0
    double a0[MT];
    double bcA[MT][2], bcB[MT][2];
    seek(gnext);   // next group's raw input: a whole layer chain to arrive
#pragma unroll
    for (int m = 0; m < MT; m++) {
      const double rw = rawn[m];
      rawn[m] = load_raw(m);
      a0[m] = normalise(rw, in_min, in_scale);
      const long long row = row0 + 8 * m + r;
#pragma unroll
      for (int x = 0; x < 2; x++) {
        const int oi = 2 * c + x;
        const bool ok = io.u && row < io.n && oi < s.n_out;
        bcA[m][x] = (ok && io.bc_a) ? ld_nc_early(io.bc_a + row * s.n_out + oi) : 0.0;
        bcB[m][x] = (ok && io.bc_b) ? ld_nc_early(io.bc_b + row * s.n_out + oi) : 1.0;
      }
    }
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
      const double2 b2 = *reinterpret_cast<const double2*>(pk + s.pFB + ((long long)(0 * NT + jn) * 32 + lane) * 2);
      const double bf = pk[s.pF0 + jn * 32 + lane];
#pragma unroll
      for (int m = 0; m < MT; m++) {
        acc[m][jn][0] = b2.x;
        acc[m][jn][1] = b2.y;
        dmma(acc[m][jn][0], acc[m][jn][1], a0[m], bf);
      }
    }
    // activation of element e of every row tile (element e = register e&1 of tile e>>1 = features 4e..4e+3)
    auto activate = [&](int e) {
#pragma unroll
      for (int m = 0; m < MT; m++) {
        if (e < NE - 1 || last_real_x1) h[m][e >> 1][e & 1] = act_fn(acc[m][e >> 1][e & 1]);
        else h[m][e >> 1][e & 1] = 0.0;
      }
    };
    // stored for the adjoint once both elements of a tile exist (one coalesced 512 B row block per warp and tile)
    auto store_tile = [&](int i, int jn) {
      if (io.act32) {   // TF32-adjoint mode: fp32 planes, same slot order (half the store traffic)
#pragma unroll
        for (int m = 0; m < MT; m++) {
          const long long row = row0 + 8 * m + r;
          if (row < io.n) {
            if constexpr (QUAD) {   // K-major tcgen05 operand order: 4 consecutive rows of one slot are 16 contiguous bytes
              float* q = io.act32 + (((long long)i * (io.rows_cap >> 2) + (row >> 2)) * NT + jn) * 32 + 8 * c + (row & 3);
              __stcs(q, (float)h[m][jn][0]);
              __stcs(q + 4, (float)h[m][jn][1]);
            } else {
              st_cs_f32x2_fwd(io.act32 + (((long long)i * NT + jn) * io.rows_cap + row) * 8 + 2 * c, (float)h[m][jn][0],
                              (float)h[m][jn][1]);
            }
          }
        }
      } else if (io.act) {
#pragma unroll
        for (int m = 0; m < MT; m++) {
          const long long row = row0 + 8 * m + r;
          if (row < io.n) st_cs_f64x2(io.act + (((long long)i * NT + jn) * io.rows_cap + row) * 8 + 2 * c, h[m][jn][0], h[m][jn][1]);
        }
      }
    };
#pragma unroll 1
    for (int i = 0;; i++) {
      // activation of layer i: a phase of its own.  (Tried in r2 and measured slower, 3.64 -> 3.93 ms: computing the
      // activation of element e two k-steps ahead of the DMMAs that consume it, in one interleaved stream -- a warp
      // issues in order, so the DMMAs then queue behind the dependent DFMA chain of the tanh in front of them; as
      // separate phases one warp's DMMA phase is a clean stream of independent instructions that fills the pipe while
      // the other warps of the scheduler sit in their tanh chains.)
      if constexpr (TV == 1 && MT == 1 && NE >= 4) {
        // two stage-by-stage batches (register budget: one batch keeps ~8 registers per element in flight)
        constexpr int NA = NE / 2, NB = NE - 1 - NA;      // elements [0, NA) and [NA, NE - 1); the last one separately
        {
          double xa[NA], ya[NA];
#pragma unroll
          for (int e = 0; e < NA; e++) xa[e] = acc[0][e >> 1][e & 1];
          tanh_rep_batch<NA>(xa, ya, stab);
#pragma unroll
          for (int e = 0; e < NA; e++) h[0][e >> 1][e & 1] = ya[e];
        }
        {
          double xb[NB], yb[NB];
#pragma unroll
          for (int e = 0; e < NB; e++) xb[e] = acc[0][(NA + e) >> 1][(NA + e) & 1];
          tanh_rep_batch<NB>(xb, yb, stab);
#pragma unroll
          for (int e = 0; e < NB; e++) h[0][(NA + e) >> 1][(NA + e) & 1] = yb[e];
        }
        activate(NE - 1);     // all padding for width 50: no tanh
      } else {
#pragma unroll
        for (int e = 0; e < NE; e++) activate(e);
      }
#pragma unroll
      for (int jn = 0; jn < NT; jn++) store_tile(i, jn);
      if (i == s.depth - 1) break;
      // ---- hidden layer i+1
      const double* fb = pk + s.pFB + (long long)(i + 1) * NT * 64;
      const double* fh = pk + s.pFH + (long long)i * KSP * NT * 32;
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        const double2 b2 = *reinterpret_cast<const double2*>(fb + (jn * 32 + lane) * 2);
#pragma unroll
        for (int m = 0; m < MT; m++) {
          acc[m][jn][0] = b2.x;
          acc[m][jn][1] = b2.y;
        }
      }
      constexpr int NTD = NEDGE ? NT - 1 : NT;   // tiles computed on the tensor pipe
      double pe[MT][2];
#pragma unroll
      for (int m = 0; m < MT; m++) pe[m][0] = pe[m][1] = 0.0;
      const double2* fe = reinterpret_cast<const double2*>(pk + s.pFE + (long long)i * KSP * 8) + c;
      if constexpr (LOCK) token_acquire(mytoken);
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        if constexpr (LOCK) { if (ks == KS - 1) token_release(mytoken); }   // the last k-step overlaps the next holder's start
#pragma unroll
        for (int jn = 0; jn < NTD; jn++) {
          const double bf = fh[(ks * NT + jn) * 32 + lane];
#pragma unroll
          for (int m = 0; m < MT; m++) {
            if constexpr (LOCK) dmma_v(acc[m][jn][0], acc[m][jn][1], h[m][ks >> 1][ks & 1], bf);
            else dmma(acc[m][jn][0], acc[m][jn][1], h[m][ks >> 1][ks & 1], bf);
          }
        }
        if constexpr (NEDGE > 0) {
          const double2 we = fe[ks * 4];
#pragma unroll
          for (int m = 0; m < MT; m++) {
            pe[m][0] = fma(h[m][ks >> 1][ks & 1], we.x, pe[m][0]);
            if (NEDGE > 1) pe[m][1] = fma(h[m][ks >> 1][ks & 1], we.y, pe[m][1]);
          }
        }
      }
      if constexpr (NEDGE > 0) {
#pragma unroll
        for (int m = 0; m < MT; m++) {
#pragma unroll
          for (int f = 0; f < NEDGE; f++) {
            pe[m][f] += __shfl_xor_sync(0xffffffffu, pe[m][f], 1);
            pe[m][f] += __shfl_xor_sync(0xffffffffu, pe[m][f], 2);
          }
          // lane c keeps feature 8(NT-1)+c in register 0 (the bias fragment is already there, 0 beyond NEDGE)
          const double mine = (c == 0) ? pe[m][0] : (NEDGE > 1 && c == 1) ? pe[m][1] : 0.0;
          acc[m][NT - 1][0] += mine;
          acc[m][NT - 1][1] = 0.0;
        }
      }
    }
    // ---- output layer
    if (io.u) {
      double o[MT][2];
      const double2 ob = *reinterpret_cast<const double2*>(pk + s.pFOB + lane * 2);
#pragma unroll
      for (int m = 0; m < MT; m++) {
        o[m][0] = ob.x;
        o[m][1] = ob.y;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        const double bf = pk[s.pFO + ks * 32 + lane];
#pragma unroll
        for (int m = 0; m < MT; m++) dmma(o[m][0], o[m][1], h[m][ks >> 1][ks & 1], bf);
      }
      // lane (r,c) holds z[row r][2c+x]; u = a + b*z (fields.py:101 with a tabulated affine wrapper)
#pragma unroll
      for (int m = 0; m < MT; m++) {
        const long long row = row0 + 8 * m + r;
        if (row < io.n) {
#pragma unroll
          for (int x = 0; x < 2; x++) {
            const int oi = 2 * c + x;
            if (oi < s.n_out) {
              const long long k = row * s.n_out + oi;
              double z = o[m][x];
              if (io.bc_b) z *= bcB[m][x];
              if (io.bc_a) z += bcA[m][x];
              io.u[k] = z;
            }
          }
        }
      }
    }
    g = gnext;
    gnext = gnext2;
  }
  if (io.dbg) {
    __syncthreads();
    if (threadIdx.x == 0) io.dbg[3 * blockIdx.x + 1] = globaltimer_ns();
  }
}

// ---------------------------------------------------------------------------------------------
// Backward.  CTA = NT+1 warps, persistent over tiles of ROWS = 8*(NT+1) rows.
//   * every warp runs the delta chain of its own 8 rows in registers (DMMA with the transposed weight
//     fragments BH/BO, staged once per CTA in shared memory by a bulk copy);
//   * the stored activations h_L of the tile are prefetched one layer ahead with cp.async into a
//     double-buffered, XOR-swizzled shared tile (no registers, DRAM latency hidden behind a whole
//     layer of DMMA work);
//   * the weight-gradient contraction dW_L += d_L^T h_L needs rows on the K axis: d_L goes to a
//     second swizzled tile; warp w < NT owns rows 8w..8w+7 (out slots) of every hidden dW_L and of
//     dW_0 in registers for the whole persistent loop, the last warp owns dW_out;
//   * bias gradients come from a column of ones patched into a padding slot of the h tile.
// Swizzle: element (row, col) of a 64-wide tile lives at row*64 + (col ^ g(row)),
// g(row) = ((row&1)<<3)|((row&2)<<1): the 64-bit fragment reads (rows 4ks+c, cols 8j+r) and the
// 128-bit row writes (rows r, cols 8j+2c) are both bank-conflict free.
// DH = depth-1 (hidden->hidden layers) is a template parameter so accumulators index statically.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int swz(int row) { return ((row & 1) << 3) | ((row & 2) << 1); }

template <int NT, int DH, bool WSMEM>
struct BwdSmem {
  static constexpr int NW = NT + 1;
  static constexpr int ROWS = 8 * NW;
  static constexpr int S = 64;           // tile row stride (8*NT <= 64)
  static constexpr int SN = 12;          // narrow tiles (inputs / zbar): conflict-free for both patterns
  static constexpr int KSP = 2 * NT;
  static constexpr int NWB = NT * 32 + DH * KSP * NT * 32;   // BO + BH fragments
  double W[WSMEM ? NWB : 2];   // WSMEM = false: too large for shared memory, read through L1 instead
  double H[2][ROWS * S];
  double D[2][ROWS * S];
  double Zb[ROWS * SN];
  double In[2][ROWS * SN];
  uint64_t bar;
  uint64_t sync;   // split-phase CTA barrier (arrive early, wait late): count = blockDim.x
};

// BEDGE = 1 or 2: the last 8-wide tile holds only BEDGE real features (width 50: 48, 49).  The 13 DMMAs per hidden layer
// that produce the delta chain's entries for that tile would be 6/8 padding; like the forward kernel's NEDGE they are
// replaced by DFMAs on the lane-local A-fragment elements (row r, out feature 4ks+c) times the two weight columns
// W[4ks+c][48..49] (one 16-byte read-only load per k-step, L1-resident) and a 2-step butterfly over the 4 lanes of a row:
// 26 DFMA + 4 DADD + 8 SHFL instead of 13 DMMA.  BEDGE = 0: off.
//
// CE = 1 (needs NT = 7, BEDGE = 2, i.e. width 50): the same treatment for the weight-gradient contraction.  Of the 7 x 7
// blocks of a hidden layer's dW the 13 that touch the last tile are 5/8 or 6/8 padding (in slots 48 / 50 / ones = features
// 48, 49 and the bias column; out slots 48 / 50).  They leave the tensor pipe: the 36 full blocks are shared 7 per warp
// on warps 0-3 (out block w x in tiles 0..5 plus one block of out blocks 4 / 5) and 2 per warp on warps 4-7 -- 9 per
// scheduler -- and warps 4-7 sum the edge entries with DFMAs, rows 16 (w-4) .. +15 each, lane l on the slot pair
// (2l, 2l+1): dW[o][48|50|ones] += d[row][o] * (h48 | h50 | 1) and dW[48|50][i] += (d48 | d50) * h[row][i], one row per
// k-step of the DMMA loop next to it.  Per hidden layer and 64-row tile 576 DMMA + 640 DFMA warp instructions instead of
// 784 DMMA (algorithmic: 637.5 DMMA-equivalents); the accumulators reuse the registers of the dropped blocks.  The four
// row-group partials of the edge sums are combined in fixed order through shared memory when the CTA finishes.
template <int NT, int KODD, int DH, bool WSMEM, int BEDGE = 0, int CE = 0>
__global__ void __launch_bounds__(32 * (NT + 1), 1) k_mlp_backward(const __grid_constant__ MlpShape s, const __grid_constant__ MlpIO io) {
  using SM = BwdSmem<NT, DH, WSMEM>;
  constexpr int ROWS = SM::ROWS, S = SM::S, SN = SM::SN;
  constexpr int KSP = 2 * NT;
  constexpr int KS = 2 * NT - KODD;
  constexpr int KR = ROWS / 4;  // k-steps over rows in the weight-gradient contraction (even: ROWS = 8 (NT + 1))
  static_assert(BEDGE == 0 || KODD == 1, "the edge weights live in the unused last k-step of BH");
  static_assert(CE == 0 || (NT == 7 && BEDGE == 2 && DH > 0), "contraction edge: width 50 only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int lrow = warp * 8 + r;   // this lane's local row in the CTA tile
  const int gl = swz(lrow);        // its swizzle (row writes / own-row reads)
  const int gc = swz(c);           // swizzle of fragment reads: rows 4ks + c  => (row & 3) == c
  const int ones_slot = feat_to_slot(s.width);
  if (io.dbg && threadIdx.x == 0) { io.dbg[3 * blockIdx.x] = globaltimer_ns(); unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm)); io.dbg[3 * blockIdx.x + 2] = sm; }

  // the layer-0 contraction of a tile runs together with the top-layer contraction of the CTA's NEXT tile (below): the
  // first tile then contracts these zeros
  for (int i = threadIdx.x; i < ROWS * S; i += blockDim.x) sm.D[1][i] = 0.0;
  for (int i = threadIdx.x; i < ROWS * SN; i += blockDim.x) sm.In[1][i] = 0.0;
  if constexpr (WSMEM) stage_weights(sm.W, io.pk + s.pBO, SM::NWB, &sm.bar);
  const double* __restrict__ wBO = WSMEM ? sm.W : io.pk + s.pBO;
  const double* __restrict__ wBH = wBO + NT * 32;

  // weight-gradient accumulators (accumulator layout: lane(r,c): [out slot 8*jr + r][in slot 8*jc+2c+x])
  // Work split of the weight-gradient contraction (balanced to within 5 %: 320 DMMAs per tile for warps
  // 0..NT-1, 336 for warp NT, so no tensor pipe of the four SM sub-partitions idles):
  //   warp w < NT : hidden layers: out-slot block w x in-slot tiles 0..NT-2 (gH[l][jc]); output layer:
  //                 in-slot tile w (gO); layer 0: out-slot block w (g0)
  //   warp NT     : hidden layers: in-slot tile NT-1 of EVERY out-slot block (gH[l][jr])
  double gH[DH > 0 ? DH : 1][NT][2];
  double g0[2];                       // warp < NT : layer 0 (out slots 8*warp.., in cols 2c+x of [inputs,1])
  double gO[2];                       // warp < NT : output layer (out r, in slots 8*warp + 2c + x)
  gO[0] = gO[1] = 0.0;
#pragma unroll
  for (int l = 0; l < (DH > 0 ? DH : 1); l++)
#pragma unroll
    for (int j = 0; j < NT; j++) gH[l][j][0] = gH[l][j][1] = 0.0;
  g0[0] = g0[1] = 0.0;

  const long long ntiles = (io.n + ROWS - 1) / ROWS;

  // prefetch of layer `layer`'s stored activations of tile `tile` into H[buf] (own row, 16 B pieces)
  auto prefetch_h = [&](long long tile, int layer, int buf) {
    const long long row = tile * ROWS + lrow;
    const bool ok = tile < ntiles && row < io.n;
    const long long rr = ok ? row : 0;
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
      cp_async16(&sm.H[buf][lrow * S + ((8 * jn + 2 * c) ^ gl)],
                 io.act + (((long long)layer * NT + jn) * io.rows_cap + rr) * 8 + 2 * c, ok);
    cp_async_commit();
  };

  // Split-phase CTA barrier: every thread ARRIVEs as soon as its own rows of the D / H tiles are
  // published and WAITs only right before it reads other warps' rows (the weight-gradient
  // contraction).  The delta chain of the layer -- which needs nothing from other warps -- runs
  // between the two, so the skew between warps (and the DRAM latency of the slowest prefetch) is
  // absorbed by a whole layer of DMMA work instead of stalling at a __syncthreads().
  if (threadIdx.x == 0) mbar_init(&sm.sync, blockDim.x);
  __syncthreads();
  uint32_t sphase = 0;
#define SYNC_ARRIVE() mbar_arrive(&sm.sync)
#define SYNC_WAIT() do { mbar_wait(&sm.sync, sphase); sphase ^= 1; } while (0)

  // per-tile top inputs (cotangent column and normalised inputs): the raw global loads are issued
  // a whole tile ahead (4 registers in flight), the arithmetic happens when the tile starts.  (time step, node) of
  // the lane's row is tracked incrementally from tile to tile (no 64-bit division / modulo per tile).
  const long long tstride = (long long)gridDim.x * ROWS;
  long long rowT = (long long)blockIdx.x * ROWS + lrow;            // this lane's row in the next tile to be issued
  int tqT = io.nn_mod > 0 ? (int)(rowT / io.nn_mod) : 0;
  long long nodeT = io.nn_mod > 0 ? rowT - (long long)tqT * io.nn_mod : rowT;
  const double in_min = c < s.nd ? s.x_min[c] : (c == s.nd && s.normalize_time) ? s.t_min : 0.0;
  const double in_scale = c < s.nd ? s.x_scale[c] : (c == s.nd && s.normalize_time) ? s.t_scale : 1.0;
  struct TopRaw { double g, b, x; bool live; };
  auto issue_top = [&]() {      // for the tile whose row is rowT; then advances to the CTA's following tile
    TopRaw q = {0.0, 1.0, in_min, false};
    if (rowT < io.n) {
      q.live = true;
      if (c < s.n_out) {
        const long long k = rowT * s.n_out + c;
        q.g = io.g_u[k];
        if (io.bc_b) q.b = io.bc_b[k];
      }
      if (c < s.nd) q.x = io.x[nodeT * io.x_stride + c];
      else if (c == s.nd) q.x = io.nn_mod > 0 ? io.times[tqT] : (io.trow ? io.trow[rowT * io.x_stride] : io.t);
      else q.x = 0.0;
    }
    rowT += tstride;
    if (io.nn_mod > 0) {
      nodeT += tstride;
      while (nodeT >= io.nn_mod) { nodeT -= io.nn_mod; tqT++; }
    }
    return q;
  };
  auto finish_top = [&](const TopRaw& q, double& zA, double& v0, double& v1) {
    zA = q.g * q.b;
    // fields.py:80-88: inputs = hstack(x_norm, t_norm)  (same arithmetic as the forward kernel)
    v0 = normalise(q.x, in_min, in_scale);
    v1 = (4 + c == s.n_in) ? 1.0 : 0.0;
    if (c == s.n_in) v0 = 1.0;  // n_in < 4
    if (!q.live) { zA = 0.0; v0 = 0.0; v1 = 0.0; }
  };

  int hb = 0, db = 0;
  prefetch_h(blockIdx.x, s.depth - 1, hb);
  double zA, in0, in1;
  {
    const TopRaw q0 = issue_top();
    finish_top(q0, zA, in0, in1);
  }
  int parity = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1) {
    const long long row = tile * ROWS + lrow;
    const bool live = row < io.n;

    // ---- top: zbar = g_u * bc_b  -> A fragment (k = out index c) and Zb tile; inputs tile
    sm.Zb[lrow * SN + c] = zA;
    sm.Zb[lrow * SN + 4 + c] = 0.0;
    sm.In[parity][lrow * SN + c] = in0;
    sm.In[parity][lrow * SN + 4 + c] = in1;
    cp_async_wait_all();
    if (live && c == ((ones_slot & 7) >> 1)) sm.H[hb][lrow * S + (ones_slot ^ gl)] = 1.0;
    __syncwarp();      // the ones patch is read back by the other lanes of this row below
    SYNC_ARRIVE();     // own rows of H[hb] = h_depth (+ones), Zb, In -- and of D[db^1] = d_0 of the previous tile -- published
    const TopRaw qn = issue_top();   // next tile's global loads fly during this whole tile
    // the stored activations of the CTA's NEXT tile (depth x NT planes of ROWS x 64 B, contiguous each) are pulled into L2
    // a whole tile ahead: the cp.async prefetches below have half a stage to land and then find them there instead of
    // in HBM (PCX_OPT_TUNE_L2PF)
    if (io.l2pf && threadIdx.x < s.depth * NT) {
      const long long nrow = (tile + gridDim.x) * ROWS;
      if (nrow + ROWS <= io.n)
        bulk_prefetch_l2(io.act + ((long long)threadIdx.x * io.rows_cap + nrow) * 8, ROWS * 64);
    }

    // d_{depth-1} = (zbar W_out) * (1 - h^2), h read back from the own row of the tile
    double dC[NT][2];
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
      double a0 = 0.0, a1 = 0.0;
      dmma(a0, a1, zA, wBO[jn * 32 + lane]);
      const double2 hv = *reinterpret_cast<const double2*>(&sm.H[hb][lrow * S + ((8 * jn + 2 * c) ^ gl)]);
      const bool p0 = (8 * jn + 2 * c) == ones_slot, p1 = (8 * jn + 2 * c + 1) == ones_slot;
      dC[jn][0] = p0 ? 0.0 : a0 * (1.0 - hv.x * hv.x);
      dC[jn][1] = p1 ? 0.0 : a1 * (1.0 - hv.y * hv.y);
    }
    SYNC_WAIT();       // everybody is done with the previous tile: H[hb^1] is free, Zb / H[hb] / D[db^1] complete
    // prefetch the next h: layer DH-1 of this tile, or (depth == 1) the next tile's top layer
    if (DH > 0) prefetch_h(tile, DH - 1, hb ^ 1);
    else prefetch_h(tile + gridDim.x, s.depth - 1, hb ^ 1);
    if (warp < NT) {
      // dW_out[o][slot] += sum_rows zbar[row][o] * h_depth[row][slot], in-slot tile `warp`, TOGETHER with the layer-0
      // contraction of the CTA's previous tile, dW_0 += d_0^T [inputs, 1]: each is one dependent accumulator chain of
      // KR DMMAs -- latency-bound on its own (ncu r1: the two stages held 16 % of the time for 6 % of the DMMAs) -- so
      // they are interleaved and split over even / odd row k-steps: four independent chains, one barrier round less
      double gOb[2] = {0.0, 0.0}, g0b[2] = {0.0, 0.0};
      const double* Dp = sm.D[db ^ 1];
      const double* Ip = sm.In[parity ^ 1];
#pragma unroll 2
      for (int ks = 0; ks < KR; ks += 2) {
        dmma(gO[0], gO[1], sm.Zb[(4 * ks + c) * SN + r], sm.H[hb][(4 * ks + c) * S + ((8 * warp + r) ^ gc)]);
        dmma(g0[0], g0[1], Dp[(4 * ks + c) * S + ((8 * warp + r) ^ gc)], Ip[(4 * ks + c) * SN + r]);
        dmma(gOb[0], gOb[1], sm.Zb[(4 * ks + 4 + c) * SN + r], sm.H[hb][(4 * ks + 4 + c) * S + ((8 * warp + r) ^ gc)]);
        dmma(g0b[0], g0b[1], Dp[(4 * ks + 4 + c) * S + ((8 * warp + r) ^ gc)], Ip[(4 * ks + 4 + c) * SN + r]);
      }
      gO[0] += gOb[0]; gO[1] += gOb[1];
      g0[0] += g0b[0]; g0[1] += g0b[1];
    }
    // ---- hidden layers L = DH .. 1 : d_{L-1} = (d_L W_L) * (1 - h_L^2) ; dW_L = d_L^T h_L
#pragma unroll
    for (int L = DH; L >= 1; L--) {
      // dC = d_L.  Publish it together with the own rows of h_L (prefetched into H[hb^1]).
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
        *reinterpret_cast<double2*>(&sm.D[db][lrow * S + ((8 * jn + 2 * c) ^ gl)]) = make_double2(dC[jn][0], dC[jn][1]);
      cp_async_wait_all();
      hb ^= 1;
      if (live && c == ((ones_slot & 7) >> 1)) sm.H[hb][lrow * S + (ones_slot ^ gl)] = 1.0;
      SYNC_ARRIVE();
      // delta chain (own rows only)
      double dn[NT][2];
#pragma unroll
      for (int jn = 0; jn < NT; jn++) dn[jn][0] = dn[jn][1] = 0.0;
      const double* bh = wBH + (long long)(L - 1) * KSP * NT * 32;
      constexpr int NTD = BEDGE ? NT - 1 : NT;   // tiles computed on the tensor pipe
      double pe0 = 0.0, pe1 = 0.0;
      const double2* be = reinterpret_cast<const double2*>(bh + (KSP - 1) * NT * 32) + c;   // BE: see k_pack_weights
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
#pragma unroll
        for (int jn = 0; jn < NTD; jn++) dmma(dn[jn][0], dn[jn][1], dC[ks >> 1][ks & 1], bh[(ks * NT + jn) * 32 + lane]);
        if constexpr (BEDGE > 0) {
          const double2 we = be[ks * 4];
          pe0 = fma(dC[ks >> 1][ks & 1], we.x, pe0);
          if (BEDGE > 1) pe1 = fma(dC[ks >> 1][ks & 1], we.y, pe1);
        }
      }
      if constexpr (BEDGE > 0) {
        pe0 += __shfl_xor_sync(0xffffffffu, pe0, 1);
        pe0 += __shfl_xor_sync(0xffffffffu, pe0, 2);
        if (BEDGE > 1) {
          pe1 += __shfl_xor_sync(0xffffffffu, pe1, 1);
          pe1 += __shfl_xor_sync(0xffffffffu, pe1, 2);
        }
        // lane c keeps input feature 8(NT-1)+c in register 0; everything else in the tile is padding
        dn[NT - 1][0] = (c == 0) ? pe0 : (BEDGE > 1 && c == 1) ? pe1 : 0.0;
        dn[NT - 1][1] = 0.0;
      }
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        const double2 hv = *reinterpret_cast<const double2*>(&sm.H[hb][lrow * S + ((8 * jn + 2 * c) ^ gl)]);
        const bool p0 = (8 * jn + 2 * c) == ones_slot, p1 = (8 * jn + 2 * c + 1) == ones_slot;
        dC[jn][0] = p0 ? 0.0 : dn[jn][0] * (1.0 - hv.x * hv.x);
        dC[jn][1] = p1 ? 0.0 : dn[jn][1] * (1.0 - hv.y * hv.y);
      }
      SYNC_WAIT();     // all rows of D[db] = d_L and H[hb] = h_L are in; H[hb^1] is free
      // prefetch h for the next layer (or the next tile's top layer) into the buffer just released
      if (L > 1) prefetch_h(tile, L - 2, hb ^ 1);
      else prefetch_h(tile + gridDim.x, s.depth - 1, hb ^ 1);
      if constexpr (CE) {
        const double* Dt = sm.D[db];
        const double* Ht = sm.H[hb];
        if (warp < 4) {
          // out block `warp` x in tiles 0..5, plus block (4 + warp/2, 4 + warp%2)
          const int jrx = 4 + (warp >> 1);
          const bool odd = warp & 1;
#pragma unroll 4
          for (int ks = 0; ks < KR; ks++) {
            const double a = Dt[(4 * ks + c) * S + ((8 * warp + r) ^ gc)];
            const double ax = Dt[(4 * ks + c) * S + ((8 * jrx + r) ^ gc)];
            double b[NT - 1];
#pragma unroll
            for (int jc = 0; jc < NT - 1; jc++) b[jc] = Ht[(4 * ks + c) * S + ((8 * jc + r) ^ gc)];
#pragma unroll
            for (int jc = 0; jc < NT - 1; jc++) dmma(gH[L - 1][jc][0], gH[L - 1][jc][1], a, b[jc]);
            dmma(gH[L - 1][NT - 1][0], gH[L - 1][NT - 1][1], ax, odd ? b[NT - 2] : b[NT - 3]);
          }
        } else {
          // blocks (4 + (w-4)/2, 2 ((w-4)%2) + {0, 1}) and the edge sums of rows 16 (w-4) .. +15
          const int w4 = warp - 4;
          const int jr = 4 + (w4 >> 1), jc0 = 2 * (w4 & 1);
          const int ecol = lane < 4 * (NT - 1) + 4 ? 2 * lane : 0;   // slot pair (2 lane, 2 lane + 1) of the 8 NT slots
          constexpr int E0 = 8 * (NT - 1), E1 = 8 * (NT - 1) + 2;     // slots of features 48 and 49
#pragma unroll 4
          for (int ks = 0; ks < KR; ks++) {
            const double a = Dt[(4 * ks + c) * S + ((8 * jr + r) ^ gc)];
            dmma(gH[L - 1][0][0], gH[L - 1][0][1], a, Ht[(4 * ks + c) * S + ((8 * jc0 + r) ^ gc)]);
            dmma(gH[L - 1][1][0], gH[L - 1][1][1], a, Ht[(4 * ks + c) * S + ((8 * jc0 + 8 + r) ^ gc)]);
            const int row = 16 * w4 + ks, g = swz(row);
            const double2 dv = *reinterpret_cast<const double2*>(&Dt[row * S + (ecol ^ g)]);
            const double2 hv = *reinterpret_cast<const double2*>(&Ht[row * S + (ecol ^ g)]);
            const double h48 = Ht[row * S + (E0 ^ g)], h50 = Ht[row * S + (E1 ^ g)];
            const double d48 = Dt[row * S + (E0 ^ g)], d50 = Dt[row * S + (E1 ^ g)];
            gH[L - 1][2][0] = fma(dv.x, h48, gH[L - 1][2][0]); gH[L - 1][2][1] = fma(dv.y, h48, gH[L - 1][2][1]);
            gH[L - 1][3][0] = fma(dv.x, h50, gH[L - 1][3][0]); gH[L - 1][3][1] = fma(dv.y, h50, gH[L - 1][3][1]);
            gH[L - 1][4][0] += dv.x;                           gH[L - 1][4][1] += dv.y;
            gH[L - 1][5][0] = fma(hv.x, d48, gH[L - 1][5][0]); gH[L - 1][5][1] = fma(hv.y, d48, gH[L - 1][5][1]);
            gH[L - 1][6][0] = fma(hv.x, d50, gH[L - 1][6][0]); gH[L - 1][6][1] = fma(hv.y, d50, gH[L - 1][6][1]);
          }
        }
      } else if (warp < NT) {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++) {
          const double a = sm.D[db][(4 * ks + c) * S + ((8 * warp + r) ^ gc)];
#pragma unroll
          for (int jc = 0; jc < NT - 1; jc++)
            dmma(gH[L - 1][jc][0], gH[L - 1][jc][1], a, sm.H[hb][(4 * ks + c) * S + ((8 * jc + r) ^ gc)]);
        }
      } else {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++) {
          const double b = sm.H[hb][(4 * ks + c) * S + ((8 * (NT - 1) + r) ^ gc)];
#pragma unroll
          for (int jr = 0; jr < NT; jr++)
            dmma(gH[L - 1][jr][0], gH[L - 1][jr][1], sm.D[db][(4 * ks + c) * S + ((8 * jr + r) ^ gc)], b);
        }
      }
      db ^= 1;
    }
    // ---- layer 0: publish d_0; its contraction dW_0 = d_0^T [inputs, 1] runs in the next tile's top stage (or after
    // the loop).  D[db] was last read by the contraction two stages back, which every warp has left (they all arrived at
    // the stage in between).
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
      *reinterpret_cast<double2*>(&sm.D[db][lrow * S + ((8 * jn + 2 * c) ^ gl)]) = make_double2(dC[jn][0], dC[jn][1]);
    finish_top(qn, zA, in0, in1);
    db ^= 1;
    // the next tile's top-layer h is already in flight into H[hb^1]
    hb ^= 1;
  }
  // ---- the last tile's layer-0 contraction
  SYNC_ARRIVE();
  cp_async_wait_all();
  SYNC_WAIT();
  if (warp < NT) {
    double g0b[2] = {0.0, 0.0};
    const double* Dp = sm.D[db ^ 1];
    const double* Ip = sm.In[parity ^ 1];
#pragma unroll 2
    for (int ks = 0; ks < KR; ks += 2) {
      dmma(g0[0], g0[1], Dp[(4 * ks + c) * S + ((8 * warp + r) ^ gc)], Ip[(4 * ks + c) * SN + r]);
      dmma(g0b[0], g0b[1], Dp[(4 * ks + 4 + c) * S + ((8 * warp + r) ^ gc)], Ip[(4 * ks + 4 + c) * SN + r]);
    }
    g0[0] += g0b[0]; g0[1] += g0b[1];
  }
#undef SYNC_ARRIVE
#undef SYNC_WAIT
  if (io.dbg && threadIdx.x == 0) io.dbg[3 * blockIdx.x + 1] = globaltimer_ns();

  // ---- write per-CTA partials (padded slot layout):
  //   part[cta] = [ L0: NT*8 x 8 | hidden l: (8NT x 8NT) each | out: 8 x 8NT ]
  double* P = io.part + (long long)blockIdx.x * io.part_stride;
  constexpr int WPd = 8 * NT;
  double* Po = P + WPd * 8 + (long long)DH * WPd * WPd;
  if constexpr (CE) {
    if (warp < NT) {
      P[(8 * warp + r) * 8 + 2 * c] = g0[0];
      P[(8 * warp + r) * 8 + 2 * c + 1] = g0[1];
      Po[r * WPd + 8 * warp + 2 * c] = gO[0];
      Po[r * WPd + 8 * warp + 2 * c + 1] = gO[1];
    }
    // edge sums: the four row-group partials of warps 4-7 meet in shared memory (the D tiles are free now)
    double* scr = &sm.D[0][0];     // [w4][l][k = 0..4][64 slots]
    static_assert(4 * DH * 5 * 64 <= 2 * ROWS * S, "edge scratch fits the D tiles");
    __syncthreads();
    if (warp >= 4) {
#pragma unroll
      for (int l = 0; l < DH; l++)
#pragma unroll
        for (int k = 0; k < 5; k++) {
          scr[(((warp - 4) * DH + l) * 5 + k) * 64 + 2 * lane] = gH[l][2 + k][0];
          scr[(((warp - 4) * DH + l) * 5 + k) * 64 + 2 * lane + 1] = gH[l][2 + k][1];
        }
    }
    __syncthreads();
    constexpr int E0 = 8 * (NT - 1), E1 = 8 * (NT - 1) + 2;
    for (int i = threadIdx.x; i < DH * 5 * WPd; i += blockDim.x) {
      const int col = i % WPd, k = (i / WPd) % 5, l = i / (5 * WPd);
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) v += scr[((q * DH + l) * 5 + k) * 64 + col];
      double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
      if (k < 3) Pl[col * WPd + (k == 0 ? E0 : k == 1 ? E1 : ones_slot)] = v;     // dW[o = col][48 | 50 | ones]
      else if (col < E0) Pl[(k == 3 ? E0 : E1) * WPd + col] = v;                  // dW[48 | 50][i = col]
    }
#pragma unroll
    for (int l = 0; l < DH; l++) {
      double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
      if (warp < 4) {
#pragma unroll
        for (int jc = 0; jc < NT - 1; jc++) {
          Pl[(8 * warp + r) * WPd + 8 * jc + 2 * c] = gH[l][jc][0];
          Pl[(8 * warp + r) * WPd + 8 * jc + 2 * c + 1] = gH[l][jc][1];
        }
        const int jrx = 4 + (warp >> 1), jcx = 4 + (warp & 1);
        Pl[(8 * jrx + r) * WPd + 8 * jcx + 2 * c] = gH[l][NT - 1][0];
        Pl[(8 * jrx + r) * WPd + 8 * jcx + 2 * c + 1] = gH[l][NT - 1][1];
      } else {
        const int w4 = warp - 4;
        const int jr = 4 + (w4 >> 1), jc0 = 2 * (w4 & 1);
#pragma unroll
        for (int j = 0; j < 2; j++) {
          Pl[(8 * jr + r) * WPd + 8 * (jc0 + j) + 2 * c] = gH[l][j][0];
          Pl[(8 * jr + r) * WPd + 8 * (jc0 + j) + 2 * c + 1] = gH[l][j][1];
        }
      }
    }
  } else if (warp < NT) {
    P[(8 * warp + r) * 8 + 2 * c] = g0[0];
    P[(8 * warp + r) * 8 + 2 * c + 1] = g0[1];
#pragma unroll
    for (int l = 0; l < DH; l++) {
      double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
#pragma unroll
      for (int jc = 0; jc < NT - 1; jc++) {
        Pl[(8 * warp + r) * WPd + 8 * jc + 2 * c] = gH[l][jc][0];
        Pl[(8 * warp + r) * WPd + 8 * jc + 2 * c + 1] = gH[l][jc][1];
      }
    }
    Po[r * WPd + 8 * warp + 2 * c] = gO[0];
    Po[r * WPd + 8 * warp + 2 * c + 1] = gO[1];
  } else {
#pragma unroll
    for (int l = 0; l < DH; l++) {
      double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
#pragma unroll
      for (int jr = 0; jr < NT; jr++) {
        Pl[(8 * jr + r) * WPd + 8 * (NT - 1) + 2 * c] = gH[l][jr][0];
        Pl[(8 * jr + r) * WPd + 8 * (NT - 1) + 2 * c + 1] = gH[l][jr][1];
      }
    }
  }
}

// Fixed-order sum over CTAs + un-padding into the flat equinox-order gradient.
// 64 parameters per block, 4 threads per parameter: thread (p, g) sums the partials q = g, g + 4, g + 8, ... with two
// independent accumulators (even / odd terms), the four group sums are combined in the order g = 0..3 through shared
// memory -- a fixed order, so the gradient stays bit-reproducible, at 4 x the memory-level parallelism of one thread
// per parameter walking all 148 partials (r1: 19 us per step, 1.2 % of the step of a 2 M-point shard).
__global__ void __launch_bounds__(256) k_reduce_wgrad(MlpShape s, const double* __restrict__ part, long long part_stride, int nparts,
                                                      double* __restrict__ g) {
  __shared__ double sh[4][64];
  const int pl = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const long long p = blockIdx.x * 64LL + pl;
  double acc = 0.0;
  if (p < s.nweights) {
    const int WPd = 8 * s.NT, W = s.width, D = s.depth;
    // locate the layer
    int layer = 0;
    bool is_bias = false;
    long long local = 0;
    for (int i = 0; i <= D; i++) {
      const int fin = (i == 0) ? s.n_in : W, fout = (i == D) ? s.n_out : W;
      if (p >= s.offW[i] && p < s.offW[i] + (long long)fin * fout) { layer = i; is_bias = false; local = p - s.offW[i]; break; }
      if (s.offB[i] >= 0 && p >= s.offB[i] && p < s.offB[i] + fout) { layer = i; is_bias = true; local = p - s.offB[i]; break; }
    }
    const int fin = (layer == 0) ? s.n_in : W;
    const int o = is_bias ? (int)local : (int)(local / fin);
    const int k = is_bias ? -1 : (int)(local % fin);
    long long off;
    if (layer == 0) {
      off = (long long)feat_to_slot(o) * 8 + (is_bias ? s.n_in : k);
    } else if (layer < D) {
      off = (long long)WPd * 8 + (long long)(layer - 1) * WPd * WPd + (long long)feat_to_slot(o) * WPd +
            feat_to_slot(is_bias ? W : k);
    } else {
      off = (long long)WPd * 8 + (long long)(D - 1) * WPd * WPd + (long long)o * WPd + feat_to_slot(is_bias ? W : k);
    }
    double a0 = 0.0, a1 = 0.0;
    int q = grp;
    for (; q + 4 < nparts; q += 8) {
      a0 += part[(long long)q * part_stride + off];
      a1 += part[(long long)(q + 4) * part_stride + off];
    }
    if (q < nparts) a0 += part[(long long)q * part_stride + off];
    acc = a0 + a1;
  }
  sh[grp][pl] = acc;
  __syncthreads();
  if (grp == 0 && p < s.nweights) g[p] = ((sh[0][pl] + sh[1][pl]) + sh[2][pl]) + sh[3][pl];
}

}  // namespace pcx
