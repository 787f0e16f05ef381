// pcx_mlp_umma.cuh -- TF32 adjoint of the Field with the weight-gradient contractions on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32, operands read from shared memory through matrix descriptors, accumulators in tensor
// memory (TMEM), completion through tcgen05.commit -> mbarrier, read-back with tcgen05.ld.  PCX_OPT_TF32_ADJOINT = 2
// (1 = the mma.sync kernel of pcx_mlp_tf32.cuh; same math, same stated tolerance, same partial layout).
//
// Same reverse-mode pass as k_mlp_backward / k_mlp_backward_tf32 (what JAX derives for pancax/networks/fields.py:79-101 +
// equinox.nn.MLP vmapped over the nodes, pancax/physics_kernels/base.py:248-250).  What moves to tcgen05 is every
// contraction over rows -- the only true dense contractions of the path (north_star item 4):
//     dW_L   += d_L^T h_L        M = 64 (56 out slots), N = 56 in slots, K = 128 rows per tile   (L = 1 .. depth-1)
//     dW_out += h_top^T zbar     M = 64 (h slots),      N = 8  (n_out <= 3 real),  K = 128
//     dW_0   += d_0^T [x, 1]     M = 64 (out slots),    N = 8  (n_in + 1 real),    K = 128
// as error-compensated 3xTF32: per 8-row k-step  main += A_hi B_hi  and  corr += A_lo B_hi + A_hi B_lo  (two TMEM
// accumulator sets; the tensor core accumulates with truncation, so the main set is flushed into the CTA's fp64 partial
// every UMMA_FLUSH_TILES tiles; the corrections are 2^-11 of it and are flushed once).  The delta chain
// d_{L-1} = (d_L W_L)(1 - h_L^2) stays in registers on mma.sync (its A operand is the previous layer's accumulator
// fragment; UMMA would need it staged through shared memory in a second, row-major layout).
//
// Operand layout (established on the hardware by tools/umma_tf32_probe.cu, mode 2): both operands K-major, canonical
// no-swizzle  ((8,m),(4,2)):((4,SBO),(1,LBO))  in floats -- for a tile with MN slots and K = rows:
//     float index(row, slot) = ((row / 4) * (MN / 8) + slot / 8) * 32 + (slot % 8) * 4 + row % 4
// i.e. 16-byte core-matrix rows = 4 consecutive K rows of one slot, SBO = 128 B between 8-slot groups, LBO = (MN/8) 128 B
// between the two 4-row halves of a k-step.  The forward kernel writes the fp32 activations in exactly this order
// (act32[layer][row/4][tile j][slot%8][row%4], MlpIO::act32_quad), so a 128-row tile of one layer is ONE contiguous
// 28 KB bulk (TMA) copy straight into operand position; d_L, zbar and the inputs are published in the same order by
// the warps that produce them.  "hi" operands are the raw fp32 tiles (the tensor core ignores the 13 low mantissa
// bits, as mma.sync does), "lo" tiles x - trunc(x) are written next to them.
// M = 64 accumulator row m lands in TMEM lane 32 (m / 16) + m % 16 (probe); the A tiles hold 56 slots, rows 56..63
// of the accumulators come from the 128 bytes behind each 7-group block and are never read.
#pragma once
#include "pcx_mlp_tf32.cuh"

namespace pcx {

constexpr int UMMA_FLUSH_TILES = 4;

// float index inside a K-major operand tile with G = MN / 8 slot groups
__device__ __forceinline__ int umma_idx(int row, int slot, int G) { return (((row >> 2) * G + (slot >> 3)) << 5) + ((slot & 7) << 2) + (row & 3); }

template <int DH>
struct BwdUmmaSmem {
  static constexpr int NT = 7, ROWS = 128;
  static constexpr int TF = ROWS * 8 * NT;       // floats per 56-slot tile (28 KB)
  static constexpr int TN = ROWS * 8;            // floats per 8-column tile (4 KB)
  static constexpr int NWC = DH * NT * NT * 32;
  float4 Wc[NWC];
  float2 Wo[NT * 32];
  float H[2][TF];          // stored activations of the stage (hi operand), ring of two
  float Hl[TF];            // their lo parts
  float D[TF], Dl[TF];     // d_L: hi (raw) and lo
  float Nb[2][TN];         // the narrow B operand of the top / layer-0 stage: zbar or [inputs, 1]; hi, lo
  uint64_t wbar, hbar[2], mbar[2], pub;   // weights landed | h tile landed | MMAs of a stage complete | stage published
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); base offset 0, no swizzle
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(u[i]);
}

// TMEM columns: main set at 0, correction set at 256; inside a set hidden layer l at 64 l, dW_out at 192, dW_0 at 200
constexpr int UMMA_TCOLS = 512, UMMA_CORR = 256, UMMA_COL_OUT = 192, UMMA_COL_L0 = 200;

// CTA = 8 compute warps (16 rows each: publication of the operands, delta chain) + 1 issuer warp whose lane 0 waits for a
// stage to be published (mbarrier `pub`, 256 arrivals), issues its 48 MMAs + tcgen05.commit and the bulk copy of the next
// h tile.  Stage k commits to mbar[k & 1]; a compute thread waits for stage k-1 before it overwrites what that stage's
// MMAs read, and for nothing else: there is no CTA-wide barrier in the loop.
constexpr int UMMA_THREADS = 288;

template <int DH>
__global__ void __launch_bounds__(UMMA_THREADS, 1) k_mlp_backward_umma(const __grid_constant__ MlpShape s, const __grid_constant__ MlpIO io) {
  using SM = BwdUmmaSmem<DH>;
  constexpr int NT = SM::NT, ROWS = SM::ROWS, TF = SM::TF;
  static_assert(DH >= 1 && DH <= 3, "TMEM column map / shared memory budget");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int lr[2] = {warp * 16 + g, warp * 16 + g + 8};   // this lane's two local rows
  const int ones_slot = feat_to_slot(s.width);
  const long long ntiles = (io.n + ROWS - 1) / ROWS;
  const int depth = DH + 1;

  if (threadIdx.x == 0) {
    mbar_init(&sm.wbar, 1);
    mbar_init(&sm.hbar[0], 1);
    mbar_init(&sm.hbar[1], 1);
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    mbar_init(&sm.pub, 256);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(UMMA_TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  if (threadIdx.x == 0) {   // stage [Wc | Wo] once per CTA
    const uint32_t bytes = (uint32_t)(SM::NWC * sizeof(float4) + NT * 32 * sizeof(float2));
    mbar_expect_tx(&sm.wbar, bytes);
    const char* src = reinterpret_cast<const char*>(io.pk32);
    const uint32_t nb = (uint32_t)(SM::NWC * sizeof(float4));
    for (uint32_t o = 0; o < nb; o += 32768) bulk_g2s(reinterpret_cast<char*>(sm.Wc) + o, src + o, (nb - o) < 32768 ? (nb - o) : 32768, &sm.wbar);
    bulk_g2s(sm.Wo, src + nb, (uint32_t)(NT * 32 * sizeof(float2)), &sm.wbar);
  }

  // The h loads of a CTA form one sequence q = 0, 1, 2, ...: (tile it, layer depth-1), ..., (it, 0), (it+1, depth-1), ...;
  // load q lands in H[q & 1] -- one contiguous 28 KB block of act32 -- and is issued when the stage consuming q-1 starts.
  const long long quads_cap = io.rows_cap / 4;
  auto prefetch_q = [&](long long q) {   // thread 0 only
    const long long it = q / depth;
    const long long tile = blockIdx.x + it * gridDim.x;
    if (tile < ntiles) {
      const int layer = depth - 1 - (int)(q - it * depth);
      const int buf = (int)(q & 1);
      mbar_expect_tx(&sm.hbar[buf], (uint32_t)(TF * sizeof(float)));
      bulk_g2s(sm.H[buf], io.act32 + ((long long)layer * quads_cap + tile * (ROWS / 4)) * (NT * 32), (uint32_t)(TF * sizeof(float)), &sm.hbar[buf]);
    }
  };
  uint32_t hph = 0;
  auto wait_h = [&](int buf) {
    mbar_wait(&sm.hbar[buf], (hph >> buf) & 1u);
    hph ^= 1u << buf;
  };
  // stage k (counted over the CTA's whole run) commits to mbar[k & 1]: its phase parity there is (k >> 1) & 1
  auto wait_stage = [&](long long k) {
    if (k >= 0) {
      mbar_wait(&sm.mbar[k & 1], (uint32_t)((k >> 1) & 1));
      tc_fence_after();
    }
  };

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
  auto finish_top = [&](long long tile, int lrow, const TopRaw& q, float& z, float& v0, float& v1) {
    const long long row = tile * ROWS + lrow;
    const bool live = tile < ntiles && row < io.n;
    double a0 = 0.0;
    if (t < s.nd) a0 = (q.x - s.x_min[t]) / s.x_scale[t];
    else if (t == s.nd) a0 = s.normalize_time ? (q.tt - s.t_min) / s.t_scale : q.tt;
    double a1 = (4 + t == s.n_in) ? 1.0 : 0.0;
    if (t == s.n_in) a0 = 1.0;
    z = live ? (float)(q.g * q.b) : 0.f;
    v0 = live ? (float)a0 : 0.f;
    v1 = live ? (float)a1 : 0.f;
  };

  // own rows of a freshly landed h tile: the ones column (bias gradients) on live rows, zeros on rows past the end; then
  // the lo parts of the warp's 16 rows (4 row quads = 896 contiguous floats)
  auto patch_and_split_h = [&](long long tile, int buf) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const bool live = tile * ROWS + lr[e] < io.n;
      if (live) {
        if (t == 0) sm.H[buf][umma_idx(lr[e], ones_slot, NT)] = 1.f;
      } else {
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          sm.H[buf][umma_idx(lr[e], 8 * jn + 2 * t, NT)] = 0.f;
          sm.H[buf][umma_idx(lr[e], 8 * jn + 2 * t + 1, NT)] = 0.f;
        }
      }
    }
    __syncwarp();
    const float4* src = reinterpret_cast<const float4*>(&sm.H[buf][warp * 4 * NT * 32]);
    float4* dst = reinterpret_cast<float4*>(&sm.Hl[warp * 4 * NT * 32]);
#pragma unroll
    for (int i = 0; i < NT; i++) {
      const float4 v = src[lane + 32 * i];
      float4 o;
      o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
      o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
      o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
      o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
      dst[lane + 32 * i] = o;
    }
  };
  auto apply_dtanh = [&](float (&dC)[NT][4], const float (&dn)[NT][4], int buf) {
#pragma unroll
    for (int jn = 0; jn < NT; jn++) {
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const float h = sm.H[buf][umma_idx(lr[e], 8 * jn + 2 * t + x, NT)];
          dC[jn][2 * e + x] = dn[jn][2 * e + x] * fmaf(-h, h, 1.f);
        }
    }
  };
  auto publish_d = [&](const float (&dC)[NT][4]) {
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const float v = dC[jn][2 * e + x];
          const int o = umma_idx(lr[e], 8 * jn + 2 * t + x, NT);
          sm.D[o] = v;
          sm.Dl[o] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        }
  };
  // narrow B operand: column t <- a, column 4 + t <- b of this lane's two rows
  // narrow B operand (N = 8) of the top / layer-0 stage in Nb: column t <- a, column 4 + t <- b of this lane's two rows; a
  // tile with two slot groups: group 0 = hi, group 1 = lo
  auto publish_narrow = [&](const float (&a)[2], const float (&b)[2]) {
    float* dst = &sm.Nb[0][0];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int o0 = umma_idx(lr[e], t, 2), o1 = umma_idx(lr[e], 4 + t, 2);
      dst[o0] = a[e];
      dst[32 + o0] = a[e] - __uint_as_float(__float_as_uint(a[e]) & 0xFFFFE000u);
      dst[o1] = b[e];
      dst[32 + o1] = b[e] - __uint_as_float(__float_as_uint(b[e]) & 0xFFFFE000u);
    }
  };

  // ---- the contraction of a stage: 16 k-steps x 3 MMAs, issued by thread 0 after the stage barrier
  // A: 56-slot tile pair (hi, lo); B: 56-slot pair (GB = 7, N = 56) or narrow pair (GB = 1, N = 8); col = TMEM column
  bool fresh_main = true, fresh_corr = true;   // thread 0: the accumulators hold nothing yet (after start / flush)
  auto issue_contraction = [&](const float* Ahi, const float* Alo, const float* Bhi, const float* Blo, int GB, int N, int col) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    const uint32_t lbo_b = (uint32_t)GB * 128u;
    // The k-step loop is deliberately NOT unrolled / strength-reduced: the compute warps' mma.sync chain runs on the same
    // tensor cores for the next ~5 000 cycles, and 48 MMAs issued back to back hold it up in one block (measured at C3
    // size: descriptors rebuilt per k-step 4.14 ms, unrolled with 64-bit descriptor increments 4.54 ms; a __nanosleep
    // between k-steps 5.4 ms).
    const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bhi), b_lo = smem_u32(Blo);
    const uint32_t d_main = tmem + (uint32_t)col, d_corr = tmem + (uint32_t)(UMMA_CORR + col);
#pragma unroll 1
    for (int ks = 0; ks < ROWS / 8; ks++) {
      const uint32_t oa = (uint32_t)ks * 2u * NT * 128u, ob = (uint32_t)ks * 2u * lbo_b;
      const uint64_t dah = umma_desc(a_hi + oa, NT * 128, 128), dal = umma_desc(a_lo + oa, NT * 128, 128);
      const uint64_t dbh = umma_desc(b_hi + ob, lbo_b, 128), dbl = umma_desc(b_lo + ob, lbo_b, 128);
      umma_tf32(d_main, dah, dbh, idesc, (ks > 0 || !fresh_main) ? 1u : 0u);
      umma_tf32(d_corr, dal, dbh, idesc, (ks > 0 || !fresh_corr) ? 1u : 0u);
      umma_tf32(d_corr, dah, dbl, idesc, 1u);
    }
  };

  // ---- flush of TMEM accumulators into the CTA's fp64 partial (zeroed by the host; one owner thread per entry, its
  // reductions issued in program order: deterministic).  Warps 0..3 read their lane quadrant; lane i < 16 holds row 16 q + i.
  double* P = io.part + (long long)blockIdx.x * io.part_stride;
  constexpr int WPd = 8 * NT;
  auto flush_set = [&](int set_col) {
    if (warp < 4) {
      const int m = 16 * warp + (lane & 15);
      const bool own = lane < 16 && m < WPd;
      const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)set_col;
      float v[8];
#pragma unroll 1
      for (int l = 0; l < DH; l++) {
        double* Pl = P + WPd * 8 + (long long)l * WPd * WPd;
#pragma unroll 1
        for (int c0 = 0; c0 < WPd; c0 += 8) {
          tmem_ld8(tl + 64 * l + c0, v);
          if (own) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&Pl[m * WPd + c0 + i]), "d"((double)v[i]) : "memory");
          }
        }
      }
      tmem_ld8(tl + UMMA_COL_OUT, v);   // dW_out^T[slot m][o]
      if (own) {
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&P[WPd * 8 + (long long)DH * WPd * WPd + i * WPd + m]), "d"((double)v[i]) : "memory");
      }
      tmem_ld8(tl + UMMA_COL_L0, v);    // dW_0[slot m][col]
      if (own) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(&P[m * 8 + i]), "d"((double)v[i]) : "memory");
      }
    }
    tc_fence_before();
  };

  // ---- chain pieces (registers, mma.sync; as in k_mlp_backward_tf32)
  auto top_chain = [&](float (&dC)[NT][4], const float (&zA)[2], int buf) {
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
    apply_dtanh(dC, dn, buf);
  };
  auto hidden_chain = [&](float (&dC)[NT][4], int L, int buf) {
    float dn[NT][4], dl[NT][4];
#pragma unroll
    for (int jn = 0; jn < NT; jn++)
#pragma unroll
      for (int x = 0; x < 4; x++) dn[jn][x] = dl[jn][x] = 0.f;
    const float4* __restrict__ wc = sm.Wc + (long long)(L - 1) * NT * NT * 32 + lane;
#pragma unroll
    for (int kj = 0; kj < NT; kj++) {
      uint32_t ah[4], al[4];
      split_tf32(dC[kj][0], ah[0], al[0]);
      split_tf32(dC[kj][2], ah[1], al[1]);
      split_tf32(dC[kj][1], ah[2], al[2]);
      split_tf32(dC[kj][3], ah[3], al[3]);
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
    apply_dtanh(dC, dn, buf);
  };

  const long long quads_l2 = quads_cap;
  if (warp == 8) {
    // ======================================================== issuer warp (lane 0)
    if (lane == 0) {
      long long hq = 0, k = 0;   // h load / stage counters
      int since_flush = 0;
      prefetch_q(0);
      mbar_wait(&sm.wbar, 0);
      long long issued = 1;   // h loads issued so far (load 0 above)
      auto stage = [&](const float* Ahi, const float* Alo, const float* Bhi, const float* Blo, int GB, int N, int col, bool consumes_h) {
        mbar_wait(&sm.pub, (uint32_t)(k & 1));   // all 256 compute threads have published stage k (and fenced)
        tc_fence_after();
        // Every compute thread waited for the MMAs of stage k-1 and finished that stage's chain before it published: all
        // of stages < k is history, so of the two ring buffers only the one holding load hq (consumed by this stage or
        // by the next top stage) is in use -- load hq + 1 may go out now.
        while (issued < hq + 2) prefetch_q(issued++);
        issue_contraction(Ahi, Alo, Bhi, Blo, GB, N, col);
        umma_commit(&sm.mbar[k & 1]);
        if (consumes_h) hq++;
        k++;
      };
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (since_flush == UMMA_FLUSH_TILES) { since_flush = 0; fresh_main = true; }   // the compute warps drained the main set
        if (io.l2pf) {   // the next tile's stored activations into L2 a tile ahead
          const long long nt2 = tile + gridDim.x;
          if (nt2 < ntiles)
            for (int l = 0; l < depth; l++)
              bulk_prefetch_l2(io.act32 + ((long long)l * quads_l2 + nt2 * (ROWS / 4)) * (NT * 32), (uint32_t)(TF * sizeof(float)));
        }
        stage(sm.H[hq & 1], sm.Hl, &sm.Nb[0][0], &sm.Nb[0][0] + 32, 2, 8, UMMA_COL_OUT, true);
#pragma unroll 1
        for (int L = DH; L >= 1; L--) stage(sm.D, sm.Dl, sm.H[hq & 1], sm.Hl, NT, WPd, 64 * (L - 1), true);
        stage(sm.D, sm.Dl, &sm.Nb[0][0], &sm.Nb[0][0] + 32, 2, 8, UMMA_COL_L0, false);
        fresh_main = false;
        fresh_corr = false;
        since_flush++;
      }
    }
  } else {
    // ======================================================== compute warps
    long long hq = 0, k = 0;
    int since_flush = 0;
    float zA[2], in0[2], in1[2];
    TopRaw qn[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const TopRaw q0 = issue_top(blockIdx.x, lr[e]);
      finish_top(blockIdx.x, lr[e], q0, zA[e], in0[e], in1[e]);
    }
    mbar_wait(&sm.wbar, 0);
    auto published = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&sm.pub);
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      float dC[NT][4];
      // ================= top stage: dW_out += h_top^T zbar ; d_DH = (zbar W_out)(1 - h_top^2)
      wait_stage(k - 1);      // layer 0 of the previous tile: Nb is free
      if (since_flush == UMMA_FLUSH_TILES) {   // every MMA issued so far has completed: drain the main set
        flush_set(0);
        since_flush = 0;
      }
      {
        const int buf = (int)(hq & 1);
        const float zero[2] = {0.f, 0.f};
        publish_narrow(zA, zero);
        wait_h(buf);
        patch_and_split_h(tile, buf);
        published();
#pragma unroll
        for (int e = 0; e < 2; e++) qn[e] = issue_top(tile + gridDim.x, lr[e]);   // next tile's global loads fly during this tile
        top_chain(dC, zA, buf);
        hq++; k++;
      }
      // ================= hidden stages L = DH .. 1: dW_L += d_L^T h_L ; d_{L-1} = (d_L W_L)(1 - h_L^2)
#pragma unroll
      for (int L = DH; L >= 1; L--) {
        wait_stage(k - 1);          // D / Dl / Hl are free again
        const int buf = (int)(hq & 1);
        publish_d(dC);
        wait_h(buf);
        patch_and_split_h(tile, buf);
        published();
        hidden_chain(dC, L, buf);
        hq++; k++;
      }
      // ================= layer 0: dW_0 += d_0^T [inputs, 1]
      wait_stage(k - 1);
      publish_d(dC);
      publish_narrow(in0, in1);
      published();
      k++;
#pragma unroll
      for (int e = 0; e < 2; e++) finish_top(tile + gridDim.x, lr[e], qn[e], zA[e], in0[e], in1[e]);
      since_flush++;
    }
    if (k > 0) {
      wait_stage(k - 1);
      flush_set(0);
      flush_set(UMMA_CORR);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(UMMA_TCOLS));
}

}  // namespace pcx
