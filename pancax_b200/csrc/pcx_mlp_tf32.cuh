// pcx_mlp_tf32.cuh -- TF32 tensor-core adjoint of the Field (BASELINE.json north_star item 4: "DMMA for
// fp64 or TF32 with a stated tolerance"; config C5 "fp64 vs TF32 tensor-core adjoint").
//
// Replaces the same reverse-mode pass as k_mlp_backward (what JAX derives for pancax/networks/fields.py:79-101
// + equinox.nn.MLP vmapped over the nodes, pancax/physics_kernels/base.py:248-250), selected per handle with
// PCX_OPT_TF32_ADJOINT.  The FORWARD pass stays on the FP64 pipe in either mode: the FE sweep differentiates u
// over an element (h = 1/128), which amplifies any error in u by 1/h, so a TF32 forward cannot hold 1e-4.
// The adjoint is different: it is linear in the cotangent, every product enters a long sum, and fp32-level
// accuracy is ample.
//
// Arithmetic: error-compensated 3xTF32.  Every operand x (fp32) is split as x = hi + lo with hi = the tf32 part of x,
// lo = x - hi, and a*b is evaluated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi on mma.sync.m16n8k8.tf32
// with fp32 accumulators (the dropped a_lo*b_lo term is 2^-20 relative at most).  Single-pass TF32 was rejected on
// measured numbers: rounding the weights is a systematic 2^-11 perturbation of the delta chain that does not
// average out over rows, and the bias gradients cancel heavily -- on the oracle's cases single-pass TF32 gives
// 1.5e-4 .. 1.6e-3 relative gradient error, 3xTF32 gives 1e-7 (tools/tf32_adjoint_sim.py).  Stated tolerance
// of this mode: 1e-5 relative on every gradient leaf (tests/test_gpu_tf32.py), inside north_star's 1e-4.
//
// Data flow (B200-first):
//  * the forward kernel stores the hidden activations a second way for this mode: fp32, still in accumulator
//    slot order, act32[layer][8-wide tile j][row][8 slots] (896 B per row instead of 1 792);
//  * CTA = 8 warps, persistent over tiles of 128 rows; one elected thread pulls the (layer, tile j) planes of
//    a row tile -- 4 KB contiguous each -- into a ring of three shared-memory buffers with bulk (TMA) copies TWO
//    stages ahead (a stage is ~3 000 cycles, DRAM latency + 28 KB at one SM's share of the bandwidth is about as
//    long: one stage ahead left the mbarrier wait exposed), completion on an mbarrier per ring slot; no swizzle is needed: with [j][row][8] planes both access patterns (own-row
//    float2 in accumulator layout, and the transposed fragment read rows k0+t / column g of the contraction)
//    touch 32 distinct banks;
//  * the m16n8k8 accumulator fragment of layer L (lane (g,t): rows g, g+8; slots 2t, 2t+1 of tile j) is the
//    A fragment of layer L-1 once K is permuted (k = t <-> slot 8j+2t, k = t+4 <-> slot 8j+2t+1); the chain
//    weights are pre-packed as {hi(b0), hi(b1), lo(b0), lo(b1)} per lane in that order (one LDS.128 per
//    fragment), so the delta chain never leaves registers;
//  * weight-gradient contraction dW_L += d_L^T h_L (rows on K): warp w owns the 16 out slots of M tile w & 3 and half
//    of the in-slot tiles (4 + 3 blocks per warp scheduler and layer at width 50: the tensor pipes stay balanced);
//    dW_out and dW_0 are one block per M tile, on warps 0..3 and 4..7; bias gradients come from a column of ones
//    patched into a padding slot of the h tile;
//  * fp32 accumulators are flushed into the CTA's fp64 partial (same padded slot layout as the fp64 kernel,
//    so k_reduce_wgrad is shared) every TF32_FLUSH_TILES tiles: fp32 summation never spans more than 1 024 rows.
#pragma once
#include "pcx_mlp.cuh"

namespace pcx {

constexpr int TF32_FLUSH_TILES = 8;
constexpr int TF32_KC_UNROLL = 2;   // chunks of two k-steps unrolled together in the contraction loop

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// x = hi + lo in two instructions (LOP3 + FADD): hi = x with the 13 low mantissa bits cleared (what the tensor core
// reads of an fp32 register anyway), lo = x - hi exactly; lo goes to the MMA as raw fp32 bits (the hardware drops its
// low 13 bits: 2^-21 of x).  cvt.rna.tf32.f32 is NOT used on the device: on sm_100a it expands to a ~5-instruction
// sequence (add, mask, two NaN/Inf compares, select), which made the operand splits the bottleneck of the first version.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (c_hi, c_lo) += a * b with a, b split.  The tensor core accumulates with truncation (measured: one fp32 accumulator
// over 384 MMAs drifts by 1e-5 relative), so the two correction products go to their own accumulator -- its magnitude is
// 2^-10 of the sum, its truncation error irrelevant -- and the main accumulator is kept short-lived by the callers.
__device__ __forceinline__ void mma_3xtf32(float (&ch)[4], float (&cl)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(cl, al[0], al[1], al[2], al[3], bh0, bh1);
  mma_tf32(ch, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
  mma_tf32(cl, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
}

// ---------------------------------------------------------------------------------------------
// Weight packing for the TF32 adjoint (floats):
//   Wc [depth-1][NT kj][NT jn][32 lanes][4] = {hi(b0), hi(b1), lo(b0), lo(b1)} with
//        b0 = W_l[feat(8kj+2t)][feat(8jn+g)], b1 = W_l[feat(8kj+2t+1)][feat(8jn+g)]     (l = 1..depth-1)
//   Wo [NT jn][32 lanes][2]               = {hi, lo} of W_out[t][feat(8jn+g)]            (t < n_out)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline long long tf32_pack_floats(int NT, int depth) {
  return (long long)(depth - 1) * NT * NT * 32 * 4 + (long long)NT * 32 * 2;
}

__global__ void k_pack_weights_tf32(MlpShape s, const double* __restrict__ w, float* __restrict__ pk) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int NT = s.NT, W = s.width, D = s.depth;
  const long long nc = (long long)(D - 1) * NT * NT * 32, no = (long long)NT * 32;
  if (tid >= nc + no) return;
  if (tid < nc) {
    const int lane = (int)(tid & 31);
    long long q = tid >> 5;
    const int jn = (int)(q % NT);
    q /= NT;
    const int kj = (int)(q % NT), l = (int)(q / NT) + 1;
    const int g = lane >> 2, t = lane & 3;
    const int n = slot_to_feat(8 * jn + g);
    float out[4];
#pragma unroll
    for (int x = 0; x < 2; x++) {
      const int k = slot_to_feat(8 * kj + 2 * t + x);
      const double v = (k < W && n < W) ? w[s.offW[l] + (long long)k * W + n] : 0.0;
      const float hi = __uint_as_float(f2tf32((float)v));
      out[x] = hi;
      out[2 + x] = __uint_as_float(f2tf32((float)(v - (double)hi)));
    }
    reinterpret_cast<float4*>(pk)[tid] = make_float4(out[0], out[1], out[2], out[3]);
  } else {
    const long long u = tid - nc;
    const int lane = (int)(u & 31), jn = (int)(u >> 5);
    const int g = lane >> 2, t = lane & 3;
    const int n = slot_to_feat(8 * jn + g);
    const double v = (t < s.n_out && n < W) ? w[s.offW[D] + (long long)t * W + n] : 0.0;
    const float hi = __uint_as_float(f2tf32((float)v));
    reinterpret_cast<float2*>(pk + nc * 4)[u] = make_float2(hi, __uint_as_float(f2tf32((float)(v - (double)hi))));
  }
}

template <int NT, int DH>
struct BwdTfSmem {
  static constexpr int NW = 8;
  static constexpr int ROWS = 16 * NW;       // 128 rows per CTA tile
  static constexpr int TS = ROWS * 8;        // floats per (tile j) plane
  static constexpr int NWC = (DH > 0 ? DH : 0) * NT * NT * 32;
  // bytes of everything but the h ring; the ring takes 3 buffers (loads issued two stages ahead) when that fits
  // into the 227 KB a CTA may use, else 2 (one stage ahead)
  static constexpr long long FIXED = (long long)(NWC > 0 ? NWC : 1) * 16 + NT * 32 * 8 + 2LL * NT * TS * 4 + ROWS * 4 * 4 +
                                     2LL * TS * 4 + 64;
  static constexpr int NH = (FIXED + 3LL * NT * TS * 4 <= 227 * 1024) ? 3 : 2;
  float4 Wc[NWC > 0 ? NWC : 1];
  float2 Wo[NT * 32];
  float H[NH][NT * TS];
  float D[2][NT * TS];
  float Zb[ROWS * 4];       // zbar, 4 columns (n_out <= 3)
  float In[2][TS];
  uint64_t wbar, hbar[NH], sync;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int NT, int DH>
__global__ void __launch_bounds__(256, 1) k_mlp_backward_tf32(const __grid_constant__ MlpShape s, const __grid_constant__ MlpIO io) {
  using SM = BwdTfSmem<NT, DH>;
  constexpr int ROWS = SM::ROWS, TS = SM::TS, NW = SM::NW;
  constexpr int NM = (NT + 1) / 2;   // 16-slot M tiles over the out slots
  constexpr int KR = ROWS / 8;       // k-steps over rows in the weight-gradient contraction
  static_assert(NM <= 4 && NW == 8, "work split below: M tile = warp & 3, in-slot half = warp >> 2");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int lr[2] = {warp * 16 + g, warp * 16 + g + 8};   // this lane's two local rows
  const int ones_slot = feat_to_slot(s.width);
  const int ones_off = (ones_slot >> 3) * TS + (ones_slot & 7);
  const long long ntiles = (io.n + ROWS - 1) / ROWS;

  if (threadIdx.x == 0) {
    mbar_init(&sm.wbar, 1);
    for (int i = 0; i < SM::NH; i++) mbar_init(&sm.hbar[i], 1);
    mbar_init(&sm.sync, blockDim.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // stage [Wc | Wo] once per CTA
    const uint32_t bytes = (uint32_t)(SM::NWC * sizeof(float4) + NT * 32 * sizeof(float2));
    mbar_expect_tx(&sm.wbar, bytes);
    const char* src = reinterpret_cast<const char*>(io.pk32);
    if constexpr (SM::NWC > 0) {
      const uint32_t nb = (uint32_t)(SM::NWC * sizeof(float4));
      for (uint32_t o = 0; o < nb; o += 32768) bulk_g2s(reinterpret_cast<char*>(sm.Wc) + o, src + o, (nb - o) < 32768 ? (nb - o) : 32768, &sm.wbar);
      src += nb;
    }
    bulk_g2s(sm.Wo, src, (uint32_t)(NT * 32 * sizeof(float2)), &sm.wbar);
  }

  // stored fp32 activations of (tile, layer) -> H[buf]: NT contiguous 4 KB planes, one elected thread
  auto prefetch_h = [&](long long tile, int layer, int buf) {
    if (threadIdx.x == 0 && tile < ntiles) {
      mbar_expect_tx(&sm.hbar[buf], (uint32_t)(NT * TS * sizeof(float)));
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
        bulk_g2s(&sm.H[buf][jn * TS], io.act32 + (((long long)layer * NT + jn) * io.rows_cap + tile * ROWS) * 8,
                 (uint32_t)(TS * sizeof(float)), &sm.hbar[buf]);
    }
  };
  // The h loads of a CTA form one sequence q = 0, 1, 2, ...: (tile it, layer depth-1), (it, depth-2), ..., (it, 0),
  // (it+1, depth-1), ...; load q lands in ring buffer q % NH and is issued NH-1 consuming stages ahead.
  constexpr int NH = SM::NH;
  const int depth = DH + 1;
  auto prefetch_q = [&](long long q) {
    const long long it = q / depth;
    prefetch_h(blockIdx.x + it * gridDim.x, depth - 1 - (int)(q - it * depth), (int)(q % NH));
  };
  uint32_t hph = 0;   // bit b = phase parity of hbar[b]
  auto wait_h = [&](int buf) {
    mbar_wait(&sm.hbar[buf], (hph >> buf) & 1u);
    hph ^= 1u << buf;
  };
  uint32_t sphase = 0;
#define SYNC_ARRIVE() mbar_arrive(&sm.sync)
#define SYNC_WAIT() do { mbar_wait(&sm.sync, sphase); sphase ^= 1; } while (0)

  // per-tile top inputs of this lane's two rows: raw loads issued a tile ahead
  struct TopRaw { double g, b, x, tt; };
  auto issue_top = [&](long long tile, int lrow) {
    TopRaw q = {0.0, 1.0, 0.0, 0.0};
    const long long row = tile * ROWS + lrow;
    if (tile < ntiles && row < io.n) {
      if (t < s.n_out) {
        const long long k = row * s.n_out + t;
        q.g = io.g_u[k];
        if (io.bc_b) q.b = io.bc_b[k];
      }
      const long long xr = io.nn_mod > 0 ? row % io.nn_mod : row;
      if (t < s.nd) q.x = io.x[xr * io.x_stride + t];
      else if (t == s.nd) q.tt = io.nn_mod > 0 ? io.times[row / io.nn_mod] : (io.trow ? io.trow[row * io.x_stride] : io.t);
    }
    return q;
  };
  // -> zbar[row][t], inputs[row][t], inputs[row][4 + t]   (fields.py:80-88, same arithmetic as mlp_input)
  auto finish_top = [&](long long tile, int lrow, const TopRaw& q, float& z, float& v0, float& v1) {
    const long long row = tile * ROWS + lrow;
    const bool live = tile < ntiles && row < io.n;
    double a0 = 0.0;
    if (t < s.nd) a0 = (q.x - s.x_min[t]) / s.x_scale[t];
    else if (t == s.nd) a0 = s.normalize_time ? (q.tt - s.t_min) / s.t_scale : q.tt;
    double a1 = (4 + t == s.n_in) ? 1.0 : 0.0;
    if (t == s.n_in) a0 = 1.0;   // n_in < 4: the bias column sits right behind the inputs
    z = live ? (float)(q.g * q.b) : 0.f;
    v0 = live ? (float)a0 : 0.f;
    v1 = live ? (float)a1 : 0.f;
  };

  // weight-gradient accumulators (fp32, flushed to the fp64 partial every TF32_FLUSH_TILES tiles).
  // Work split of the contraction, balanced per warp scheduler (warps w and w+4 share one, and the tensor pipe is
  // per scheduler): warp w owns out-slot M tile mi = w & 3 (slots 16mi .. 16mi+15); warps 0..3 take the in-slot
  // tiles [0, NB0), warps 4..7 the tiles [NB0, NT) -- 4 + 3 blocks of 16x8 per scheduler and hidden layer at NT = 7.
  // The two outer layers are one block per M tile each: dW_out^T goes to warps 0..3, dW_0 to warps 4..7.
  //   G[l][b] = dW_{l+1}[out slots 16mi + {g, g+8}][in slots 8(nj0+b) + {2t, 2t+1}]
  //   gX      = warps 0..3: dW_out^T[h slots 16mi + {g, g+8}][o = 2t, 2t+1];  warps 4..7: dW_0[out slots ..][in col 2t, 2t+1]
  constexpr int NB0 = (NT + 1) / 2;
  const int mi = warp & 3, half = warp >> 2;
  const int nj0 = half ? NB0 : 0, nb = half ? NT - NB0 : NB0;
  const bool m_active = mi < NM;
  const bool m_upper = 2 * mi + 1 < NT;   // the M tile's upper 8 slots exist
  float G[DH > 0 ? DH : 1][NB0][4];
  float gX[1][4] = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int l = 0; l < (DH > 0 ? DH : 1); l++)
#pragma unroll
    for (int bb = 0; bb < NB0; bb++)
#pragma unroll
      for (int x = 0; x < 4; x++) G[l][bb][x] = 0.f;
  double* P = io.part + (long long)blockIdx.x * io.part_stride;
  constexpr int WPd = 8 * NT;
  // the partial is zeroed by the host before the launch; every entry has exactly one owner thread, whose reductions
  // to it are issued in program order -> a fire-and-forget RED.ADD.F64 (no load latency) keeps the sum deterministic
  auto flush1 = [&](double* p, float& acc) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"((double)acc) : "memory");
    acc = 0.f;
  };
  auto flush = [&]() {
    if (m_active) {
#pragma unroll
      for (int x = 0; x < 4; x++) {
        if (x >= 2 && !m_upper) continue;
        const int ms = 16 * mi + g + 8 * (x >> 1);   // M slot of this accumulator register
        const int nc = 2 * t + (x & 1);              // column inside the 8-wide N tile
#pragma unroll
        for (int l = 0; l < DH; l++) {
          double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
#pragma unroll
          for (int bb = 0; bb < NB0; bb++)
            if (bb < nb) flush1(&Pl[ms * WPd + 8 * (nj0 + bb) + nc], G[l][bb][x]);
        }
        if (half == 0) flush1(&P[WPd * 8 + (long long)DH * WPd * WPd + nc * WPd + ms], gX[0][x]);   // dW_out[o][slot]
        else flush1(&P[ms * 8 + nc], gX[0][x]);                                                     // dW_0[slot][col]
      }
    }
  };

  // own rows of a freshly landed h tile: the ones column (bias gradients) on live rows, zeros on rows past the end
  auto patch_h = [&](long long tile, int buf) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const bool live = tile * ROWS + lr[e] < io.n;
      if (live) {
        if (t == 0) sm.H[buf][ones_off + lr[e] * 8] = 1.f;
      } else {
#pragma unroll
        for (int jn = 0; jn < NT; jn++) *reinterpret_cast<float2*>(&sm.H[buf][jn * TS + lr[e] * 8 + 2 * t]) = make_float2(0.f, 0.f);
      }
    }
    __syncwarp();
  };
  // dC <- dn * (1 - h^2) with h from the own rows of H[buf]
  auto apply_dtanh = [&](float (&dC)[NT][4], const float (&dn)[NT][4], int buf) {
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
      const float2 h0 = *reinterpret_cast<const float2*>(&sm.H[buf][jn * TS + lr[0] * 8 + 2 * t]);
      const float2 h1 = *reinterpret_cast<const float2*>(&sm.H[buf][jn * TS + lr[1] * 8 + 2 * t]);
      dC[jn][0] = dn[jn][0] * fmaf(-h0.x, h0.x, 1.f);
      dC[jn][1] = dn[jn][1] * fmaf(-h0.y, h0.y, 1.f);
      dC[jn][2] = dn[jn][2] * fmaf(-h1.x, h1.x, 1.f);
      dC[jn][3] = dn[jn][3] * fmaf(-h1.y, h1.y, 1.f);
    }
  };
  auto publish_d = [&](const float (&dC)[NT][4], int buf) {
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
      *reinterpret_cast<float2*>(&sm.D[buf][jn * TS + lr[0] * 8 + 2 * t]) = make_float2(dC[jn][0], dC[jn][1]);
      *reinterpret_cast<float2*>(&sm.D[buf][jn * TS + lr[1] * 8 + 2 * t]) = make_float2(dC[jn][2], dC[jn][3]);
    }
  };
  // acc[b] += sum_rows A[row][16mi + ..] * B_b[row][..]: A = planes (2mi, 2mi+1) of Ap, B_b = plane b of Bp, b < nblk
  auto contract = [&](auto& acc, int nblk, const float* __restrict__ Ap, const float* __restrict__ Bp, bool b_narrow = false) {
    constexpr int NB = sizeof(acc) / sizeof(acc[0]);
    const float* __restrict__ A0 = Ap + (2 * mi) * TS;
    const float* __restrict__ A1 = A0 + TS;
    float cl[NB][4];
#pragma unroll
    for (int bb = 0; bb < NB; bb++) cl[bb][0] = cl[bb][1] = cl[bb][2] = cl[bb][3] = 0.f;
#pragma unroll TF32_KC_UNROLL
    for (int kc = 0; kc < KR; kc += 2) {   // main accumulator lives for two k-steps, then joins the sum with a rounded add
      float ch[NB][4];
#pragma unroll
      for (int bb = 0; bb < NB; bb++) ch[bb][0] = ch[bb][1] = ch[bb][2] = ch[bb][3] = 0.f;
#pragma unroll
      for (int ks = kc; ks < kc + 2; ks++) {
        const int o0 = (8 * ks + t) * 8 + g, o1 = o0 + 32;
        uint32_t ah[4], al[4];
        split_tf32(A0[o0], ah[0], al[0]);
        split_tf32(A0[o1], ah[2], al[2]);
        if (m_upper) {
          split_tf32(A1[o0], ah[1], al[1]);
          split_tf32(A1[o1], ah[3], al[3]);
        } else {
          ah[1] = al[1] = ah[3] = al[3] = 0u;
        }
#pragma unroll
        for (int bb = 0; bb < NB; bb++) {
          if (bb < nblk) {
            uint32_t bh0, bl0, bh1, bl1;
            if (b_narrow) {   // a [ROWS][4] plane (zbar): columns 4..7 are zero
              const int r0 = (8 * ks + t) * 4 + g;
              split_tf32(g < 4 ? Bp[r0] : 0.f, bh0, bl0);
              split_tf32(g < 4 ? Bp[r0 + 16] : 0.f, bh1, bl1);
            } else {
              split_tf32(Bp[bb * TS + o0], bh0, bl0);
              split_tf32(Bp[bb * TS + o1], bh1, bl1);
            }
            mma_3xtf32(ch[bb], cl[bb], ah, al, bh0, bh1, bl0, bl1);
          }
        }
      }
#pragma unroll
      for (int bb = 0; bb < NB; bb++)
#pragma unroll
        for (int x = 0; x < 4; x++) acc[bb][x] += ch[bb][x];
    }
#pragma unroll
    for (int bb = 0; bb < NB; bb++)
#pragma unroll
      for (int x = 0; x < 4; x++) acc[bb][x] += cl[bb][x];
  };

  int hb = 0, db = 0, parity = 0, since_flush = 0;
  long long hq = 0;   // sequence number of the h load the current stage consumes; hb = hq % NH
  auto next_h = [&]() { hq++; hb = (hb + 1 == NH) ? 0 : hb + 1; };
#pragma unroll
  for (int q = 0; q < NH - 1; q++) prefetch_q(q);
  float zA[2], in0[2], in1[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const TopRaw q0 = issue_top(blockIdx.x, lr[e]);
    finish_top(blockIdx.x, lr[e], q0, zA[e], in0[e], in1[e]);
  }
  mbar_wait(&sm.wbar, 0);

  // top of a tile, publishing half: zbar = g_u * bc_b and the inputs tile into Zb / In[par], the landed h_depth patched
  auto top_publish = [&](long long tile, int par) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      sm.Zb[lr[e] * 4 + t] = zA[e];
      sm.In[par][lr[e] * 8 + t] = in0[e];
      sm.In[par][lr[e] * 8 + 4 + t] = in1[e];
    }
    wait_h(hb);
    patch_h(tile, hb);
  };
  // top of a tile, register half: d_{depth-1} = (zbar W_out) * (1 - h_depth^2) for the own rows
  auto top_chain = [&](float (&dC)[NT][4]) {
    uint32_t ah[4], al[4];
    split_tf32(zA[0], ah[0], al[0]);
    split_tf32(zA[1], ah[1], al[1]);
    ah[2] = al[2] = ah[3] = al[3] = 0u;
    float dn[NT][4];
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
      float dl[4] = {0.f, 0.f, 0.f, 0.f};
      dn[jn][0] = dn[jn][1] = dn[jn][2] = dn[jn][3] = 0.f;
      const float2 w = sm.Wo[jn * 32 + lane];
      mma_3xtf32(dn[jn], dl, ah, al, __float_as_uint(w.x), 0u, __float_as_uint(w.y), 0u);
#pragma unroll
      for (int x = 0; x < 4; x++) dn[jn][x] += dl[x];
    }
    apply_dtanh(dC, dn, hb);
  };

  // The layer-0 stage of a tile and the top stage of the CTA's next tile are ONE barrier stage: both contractions are
  // a single block per M tile (dW_0 on warps 4..7, dW_out on warps 0..3), so fused they keep all four tensor pipes
  // busy instead of serialising twice behind the barrier.  Prologue: the top stage of the first tile alone.
  float dC[NT][4];
  TopRaw qn[2];
  top_publish(blockIdx.x, parity);
  fence_proxy_async_smem();
  SYNC_ARRIVE();
#pragma unroll
  for (int e = 0; e < 2; e++) qn[e] = issue_top(blockIdx.x + gridDim.x, lr[e]);
  top_chain(dC);
  SYNC_WAIT();
  prefetch_q(hq + NH - 1);
  if (m_active && half == 0) contract(gX, 1, sm.H[hb], sm.Zb, true);   // dW_out^T[slot][o] += sum_rows h_depth[row][slot] zbar[row][o]

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1) {
    if constexpr (DH == 0) {   // no hidden stage separates two fused stages: Zb / In / D of the stage before last
      SYNC_ARRIVE();           // must have been consumed before they are overwritten below
      SYNC_WAIT();
    }
    // ---- hidden layers L = DH .. 1 : d_{L-1} = (d_L W_L) * (1 - h_L^2) ; dW_L = d_L^T h_L
#pragma unroll
    for (int L = DH; L >= 1; L--) {
      publish_d(dC, db);
      next_h();
      wait_h(hb);
      patch_h(tile, hb);
      fence_proxy_async_smem();
      SYNC_ARRIVE();
      // delta chain (own 16 rows, registers only)
      float dn[NT][4], dl[NT][4];
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
#pragma unroll
        for (int x = 0; x < 4; x++) dn[jn][x] = dl[jn][x] = 0.f;
      const float4* __restrict__ wc = sm.Wc + (long long)(L - 1) * NT * NT * 32 + lane;
#pragma unroll
      for (int kj = 0; kj < NT; kj++) {
        uint32_t ah[4], al[4];
        split_tf32(dC[kj][0], ah[0], al[0]);   // (row g,   k = t)     <- slot 8kj + 2t
        split_tf32(dC[kj][2], ah[1], al[1]);   // (row g+8, k = t)
        split_tf32(dC[kj][1], ah[2], al[2]);   // (row g,   k = t + 4) <- slot 8kj + 2t + 1
        split_tf32(dC[kj][3], ah[3], al[3]);   // (row g+8, k = t + 4)
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          const float4 w = wc[(kj * NT + jn) * 32];
          mma_3xtf32(dn[jn], dl[jn], ah, al, __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w));
        }
      }
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
#pragma unroll
        for (int x = 0; x < 4; x++) dn[jn][x] += dl[jn][x];
      apply_dtanh(dC, dn, hb);
      SYNC_WAIT();     // all rows of D[db] = d_L and H[hb] = h_L are in; the ring slot of the stage before is free
      prefetch_q(hq + NH - 1);
      if (m_active) contract(G[L - 1], nb, sm.D[db], sm.H[hb] + nj0 * TS);
      db ^= 1;
    }
    // ---- layer 0 of this tile (dW_0 = d_0^T [inputs, 1]) fused with the top stage of the next tile
    const long long next = tile + gridDim.x;
    const bool has_next = next < ntiles;
    publish_d(dC, db);
#pragma unroll
    for (int e = 0; e < 2; e++) finish_top(next, lr[e], qn[e], zA[e], in0[e], in1[e]);
    next_h();          // the next tile's top-layer h is in flight into this ring slot
    if (has_next) {
      top_publish(next, parity ^ 1);
      fence_proxy_async_smem();
    }
    SYNC_ARRIVE();
    if (has_next) {
#pragma unroll
      for (int e = 0; e < 2; e++) qn[e] = issue_top(next + gridDim.x, lr[e]);
      top_chain(dC);
    }
    SYNC_WAIT();
    prefetch_q(hq + NH - 1);
    if (m_active) {
      if (half == 1) contract(gX, 1, sm.D[db], sm.In[parity]);
      else if (has_next) contract(gX, 1, sm.H[hb], sm.Zb, true);
    }
    db ^= 1;
    if (++since_flush == TF32_FLUSH_TILES) { flush(); since_flush = 0; }
  }
#undef SYNC_ARRIVE
#undef SYNC_WAIT
  flush();
}

}  // namespace pcx
