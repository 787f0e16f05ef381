// pcx_jet.cuh -- strong-form / collocation path: the jet (u, du/dx_j, laplacian u, du/dt) of
// the Field at arbitrary points, fused per point on the FP64 tensor pipe, and its adjoint.
//
// Replaces what pancax derives with nested jacfwd / hessian through equinox.nn.MLP
// (pancax/physics_kernels/base.py:205-229: field_values, field_gradients, field_hessians -> field_laplacians,
// field_time_derivatives) inside strong-form residuals (poisson.py:26-29, heat_equation.py:11-17) and the
// reverse-mode pass JAX derives for loss_functions/strong_form_loss_functions.py:8-27.
//
// Forward-mode coordinate tangents ("streams") through a tanh layer z = W h + b, s = tanh z, s' = 1 - s^2,
// s'' = -2 s s':
//     h+   = s
//     g_j+ = s' (W g_j)                                   j < nd   (d/dx_j)
//     q+   = s' (W q) + s'' sum_j (W g_j)^2               (sum_j d^2/dx_j^2: the Laplacian stream)
//     tau+ = s' (W tau)                                   (d/dt)
// S = nd + 3 streams share every weight fragment; the coupling between them is lane-local because lane (r,c)
// holds the same row / slots of every stream in the m8n8k4 accumulator layout of pcx_mlp.cuh.  A warp owns 8
// points and walks all layers; between layers the streams round-trip through warp-private planes in global
// memory (they have to be stored for the adjoint anyway), so the per-thread state is one stream in, one
// accumulator, s and sum_j (W g_j)^2: 56 doubles.
//
// Planes ("act layout" of pcx_mlp.cuh): plane = [NT tiles][rows_cap][8 slots].
//   POST[l][s]  post-activation streams of hidden layer l (inputs of layer l+1 / of the output layer)
//   PRE[l][s]   forward: pre-activation streams W g_j, W q, W tau (s >= 1; s = 0 unused: s itself is POST[l][0])
//               backward: overwritten in place by the adjoints zbar_s of the pre-activations (all s)
//   ADJ         adjoint of the post-activation streams of the layer being processed: scratch, one
//               [S][NT][8 rows][8 slots] slot per resident warp of k_jet_backward (L2-resident)
// Adjoint of one layer (given hbar+, gbar_j+, qbar+, taubar+), s''' = -2 s' (s' - 2 s^2):
//     zbar_q   = s' qbar+                     zbar_tau = s' taubar+
//     zbar_g_j = s' gbar_j+ + 2 s'' (W g_j) qbar+
//     zbar     = s' hbar+ + s'' (sum_j gbar_j+ (W g_j) + qbar+ (W q) + taubar+ (W tau)) + s''' qbar+ sum_j (W g_j)^2
//     Wbar += sum_streams zbar_s^T p_s  (p_s = POST[l-1][s]),  bbar += sum_rows zbar,  pbar_s = zbar_s W
// The contraction over rows x streams runs in k_jet_wgrad (8-warp CTAs, swizzled shared tiles, DMMA).
#pragma once
#include "pcx_mlp.cuh"
#include "pcx_element.cuh"

namespace pcx {

constexpr int JET_MAX_S = 6;   // nd + 3 streams, nd <= 3

struct JetIO {
  const double* pts;      // [n][nd+1] rows (x..., t)
  long long n;            // points in this launch
  long long rows_cap;     // plane row capacity
  double* post;           // POST planes: [depth][S] planes
  double* pre;            // PRE / ZBAR planes: [depth][S] planes
  double* adj;            // ADJ scratch: [warps of k_jet_backward][S][NT][64]
  double* zjet;           // forward out: [n][S][n_out]   raw network jet (before the Dirichlet wrapper)
  const double* obar;     // backward in: [n][S][n_out]   adjoint of zjet
  const double* pk;       // packed weights
  int S, nd;
};

__device__ __forceinline__ long long plane_elems(int NT, long long rows_cap) { return (long long)NT * rows_cap * 8; }

// stream index helpers: 0 = value, 1..nd = d/dx_j, nd+1 = Laplacian, nd+2 = d/dt.  S = nd + 3, or nd + 2 when the caller
// needs no time derivative (Poisson: c_t = 0): the d/dt stream, a sixth of the work in 3D, is then not carried at all
// ---------------------------------------------------------------------------------------------
// Forward jet.  256 threads, grid-stride over groups of 8 points; forward fragments + exp2 table in smem.
// ---------------------------------------------------------------------------------------------
#ifndef PCX_JET_FWD_THREADS
#define PCX_JET_FWD_THREADS 256   // measured: 384 threads (12 warps per SM, 168 registers, no spills) is neutral -- 18.5 vs 18.4 ms per 1 M points
#endif
#ifndef PCX_JET_FWD_MINB
#define PCX_JET_FWD_MINB 1     // measured: 2 CTAs per SM (128 registers, 88 B of spills) 21.6 ms per 1 M points against 20.0 ms
#endif
template <int NT, int KODD>
__global__ void __launch_bounds__(PCX_JET_FWD_THREADS, PCX_JET_FWD_MINB) k_jet_forward(const __grid_constant__ MlpShape s, const __grid_constant__ JetIO io) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t wbar;
  double* sw = reinterpret_cast<double*>(smem_raw);
  double* stab = sw + (s.pBO - s.pF0);
  if (threadIdx.x < EXP2_TAB_N) stab[threadIdx.x] = g_exp2_tab[threadIdx.x];
  stage_weights(sw, io.pk + s.pF0, s.pBO - s.pF0, &wbar);
  const double* __restrict__ pk = sw - s.pF0;

  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  constexpr int KSP = 2 * NT;
  constexpr int KS = 2 * NT - KODD;
  const int S = io.S, nd = io.nd;
  const long long PL = plane_elems(NT, io.rows_cap);
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ngroups = (io.n + 7) / 8;

  for (long long g = warp_id; g < ngroups; g += warps_total) {
    const long long row = g * 8 + r;
    const bool live = row < io.n;
    const long long rr = live ? row : 0;
    // element offset of (tile jn, this lane's two slots) inside a plane
    auto off = [&](int jn) { return ((long long)jn * io.rows_cap + row) * 8 + 2 * c; };
    double sv[NT][2], sq[NT][2], acc[NT][2], hin[NT][2];

    for (int l = 0; l < s.depth; l++) {
      double* postL = io.post + (long long)l * S * PL;
      double* preL = io.pre + (long long)l * S * PL;
      const double* postP = io.post + (long long)(l - 1) * S * PL;
      // ---------------- stream 0: z = W h + b, s = tanh z
      if (l == 0) {
        double a0 = 0.0;
        if (live) {
          const double v = io.pts[rr * (nd + 1) + (c <= nd ? c : 0)];
          if (c < nd) a0 = (v - s.x_min[c]) / s.x_scale[c];
          else if (c == nd) a0 = s.normalize_time ? (v - s.t_min) / s.t_scale : v;
        }
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          const double2 b2 = *reinterpret_cast<const double2*>(pk + s.pFB + ((long long)jn * 32 + lane) * 2);
          acc[jn][0] = b2.x; acc[jn][1] = b2.y;
          dmma(acc[jn][0], acc[jn][1], a0, pk[s.pF0 + jn * 32 + lane]);
        }
      } else {
        const double* fb = pk + s.pFB + (long long)l * NT * 64;
        const double* fh = pk + s.pFH + (long long)(l - 1) * KSP * NT * 32;
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          const double2 h2 = live ? *reinterpret_cast<const double2*>(postP + off(jn)) : make_double2(0.0, 0.0);
          hin[jn][0] = h2.x; hin[jn][1] = h2.y;
          const double2 b2 = *reinterpret_cast<const double2*>(fb + (jn * 32 + lane) * 2);
          acc[jn][0] = b2.x; acc[jn][1] = b2.y;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
          for (int jn = 0; jn < NT; jn++) dmma(acc[jn][0], acc[jn][1], hin[ks >> 1][ks & 1], fh[(ks * NT + jn) * 32 + lane]);
      }
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        // padding slots have zero weights and zero bias: tanh(0) = 0 keeps them zero
        sv[jn][0] = tanh_tb(acc[jn][0], stab);
        sv[jn][1] = tanh_tb(acc[jn][1], stab);
        sq[jn][0] = sq[jn][1] = 0.0;
        if (live) *reinterpret_cast<double2*>(postL + off(jn)) = make_double2(sv[jn][0], sv[jn][1]);
      }
      // ---------------- derivative streams (order: g_0.., then tau, then q which needs sum_j (W g_j)^2)
      for (int pass = 1; pass < S; pass++) {
        const int st = pass <= nd ? pass : (S == nd + 3 && pass == nd + 1 ? nd + 2 : nd + 1);   // g_j..., (tau,) q
        if (l == 0) {
          // inputs of layer 0: d x_hat_j / d x_j = 1 / x_scale_j, d t_hat / d t = 1 / t_scale, Laplacian stream 0
          double a = 0.0;
          if (st <= nd) a = (c == st - 1) ? 1.0 / s.x_scale[st - 1] : 0.0;
          else if (st == nd + 2) a = (c == nd) ? (s.normalize_time ? 1.0 / s.t_scale : 1.0) : 0.0;
#pragma unroll
          for (int jn = 0; jn < NT; jn++) {
            acc[jn][0] = acc[jn][1] = 0.0;
            if (st != nd + 1) dmma(acc[jn][0], acc[jn][1], a, pk[s.pF0 + jn * 32 + lane]);
          }
        } else {
          const double* fh = pk + s.pFH + (long long)(l - 1) * KSP * NT * 32;
          const double* src = postP + (long long)st * PL;
#pragma unroll
          for (int jn = 0; jn < NT; jn++) {
            const double2 h2 = live ? *reinterpret_cast<const double2*>(src + off(jn)) : make_double2(0.0, 0.0);
            hin[jn][0] = h2.x; hin[jn][1] = h2.y;
            acc[jn][0] = acc[jn][1] = 0.0;
          }
#pragma unroll
          for (int ks = 0; ks < KS; ks++)
#pragma unroll
            for (int jn = 0; jn < NT; jn++) dmma(acc[jn][0], acc[jn][1], hin[ks >> 1][ks & 1], fh[(ks * NT + jn) * 32 + lane]);
        }
        double* preS = preL + (long long)st * PL;
        double* postS = postL + (long long)st * PL;
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          double o[2];
#pragma unroll
          for (int x = 0; x < 2; x++) {
            const double z = acc[jn][x], sx = sv[jn][x];
            const double s1 = fma(-sx, sx, 1.0);
            if (st <= nd) { sq[jn][x] = fma(z, z, sq[jn][x]); o[x] = s1 * z; }
            else if (st == nd + 2) o[x] = s1 * z;
            else o[x] = s1 * fma(-2.0 * sx, sq[jn][x], z);        // s' z + s'' sumsq, s'' = -2 s s'
          }
          if (live) {
            __stcs(reinterpret_cast<double2*>(preS + off(jn)), make_double2(acc[jn][0], acc[jn][1]));   // read next by the adjoint only: streamed
            *reinterpret_cast<double2*>(postS + off(jn)) = make_double2(o[0], o[1]);
          }
        }
      }
    }
    // ---------------- output layer: zjet[row][st][o] = W_out POST[depth-1][st] (+ b_out for the value stream)
    const double* postT = io.post + (long long)(s.depth - 1) * S * PL;
    for (int st = 0; st < S; st++) {
      const double* src = postT + (long long)st * PL;
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        const double2 h2 = live ? *reinterpret_cast<const double2*>(src + off(jn)) : make_double2(0.0, 0.0);
        hin[jn][0] = h2.x; hin[jn][1] = h2.y;
      }
      double o0 = 0.0, o1 = 0.0;
      if (st == 0) {
        const double2 ob = *reinterpret_cast<const double2*>(pk + s.pFOB + lane * 2);
        o0 = ob.x; o1 = ob.y;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ks++) dmma(o0, o1, hin[ks >> 1][ks & 1], pk[s.pFO + ks * 32 + lane]);
      if (live) {
        if (2 * c < s.n_out) io.zjet[(row * S + st) * s.n_out + 2 * c] = o0;
        if (2 * c + 1 < s.n_out) io.zjet[(row * S + st) * s.n_out + 2 * c + 1] = o1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Adjoint of the jet through the layers (per warp: 8 points).  Writes zbar_s of every layer into the PRE
// planes (for k_jet_wgrad) and leaves nothing else behind.  Transposed fragments [BO | BH] in smem.
// ---------------------------------------------------------------------------------------------
template <int NT, int KODD>
__global__ void __launch_bounds__(256, 2) k_jet_backward(const __grid_constant__ MlpShape s, const __grid_constant__ JetIO io) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t wbar;
  double* sw = reinterpret_cast<double*>(smem_raw);
  stage_weights(sw, io.pk + s.pBO, s.npacked - s.pBO, &wbar);
  const double* __restrict__ wBO = sw;
  const double* __restrict__ wBH = sw + NT * 32;

  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  constexpr int KSP = 2 * NT;
  constexpr int KS = 2 * NT - KODD;
  const int S = io.S, nd = io.nd;
  const long long PL = plane_elems(NT, io.rows_cap);
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const long long warp_id = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long ngroups = (io.n + 7) / 8;

  for (long long g = warp_id; g < ngroups; g += warps_total) {
    const long long row = g * 8 + r;
    const bool live = row < io.n;
    auto off = [&](int jn) { return ((long long)jn * io.rows_cap + row) * 8 + 2 * c; };
    // PRE / POST planes stream through once (evict-first), the ADJ scratch is kept (L2 evict-last policy): without the
    // hints the 9 GB of plane traffic pushed the scratch out of L2 between its write and its read half of the time
    auto ld2 = [&](const double* plane, int jn) {
      return live ? __ldcs(reinterpret_cast<const double2*>(plane + off(jn))) : make_double2(0.0, 0.0);
    };
    auto st2 = [&](double* plane, int jn, double a, double b) { __stcs(reinterpret_cast<double2*>(plane + off(jn)), make_double2(a, b)); };
    // ADJ scratch (r2): one slot per RESIDENT warp, [stream][tile][8 rows][8 slots], reused for every group the warp
    // walks -- 21.5 KB x 2 368 warps = 51 MB that stay in the 126 MB L2 instead of S planes over all rows written to and
    // read back from HBM once per layer (ncu: the kernel moved 9.8 GB per 262 144 points at 5.5 TB/s, more than half of
    // it this scratch).  Each lane only ever reads what it wrote itself.
    double* const adjw = io.adj + warp_id * ((long long)S * NT * 64) + r * 8 + 2 * c;
    auto adjp = [&](int st, int jn) { return adjw + ((long long)st * NT + jn) * 64; };
    uint64_t keep;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    auto ldadj = [&](int st, int jn) {
      double2 v = make_double2(0.0, 0.0);
      if (live) asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(adjp(st, jn)), "l"(keep));
      return v;
    };
    auto stadj = [&](int st, int jn, double a, double b) {
      asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(adjp(st, jn)), "d"(a), "d"(b), "l"(keep) : "memory");
    };
    // ---- top: adjoint of the last hidden layer's post-activation streams  pbar_s = obar_s W_out
    for (int st = 0; st < S; st++) {
      double zA = 0.0;
      if (live && c < s.n_out) zA = io.obar[(row * S + st) * s.n_out + c];
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        double a0 = 0.0, a1 = 0.0;
        dmma(a0, a1, zA, wBO[jn * 32 + lane]);
        if (live) stadj(st, jn, a0, a1);
      }
    }
    for (int l = s.depth - 1; l >= 0; l--) {
      const double* postL = io.post + (long long)l * S * PL;
      double* preL = io.pre + (long long)l * S * PL;
      double* preQ = preL + (long long)(nd + 1) * PL;
      double* preT = preL + (long long)(nd + 2) * PL;
      const bool has_t = S == nd + 3;
      // The coupling between the streams is local to a (tile, slot): one 8-wide tile at a time, so that nothing but
      // the tile's own values is live here (r2: the arrays over all NT tiles kept the kernel at one 8-warp CTA per SM
      // with exposed global-memory latency; per-tile scoping fits 128 registers = two CTAs per SM)
#pragma unroll 1
      for (int jn = 0; jn < NT; jn++) {
        const double2 s2 = ld2(postL, jn), q2 = ldadj(nd + 1, jn), h2 = ldadj(0, jn);
        const double2 zq = ld2(preQ, jn);
        const double2 zt = has_t ? ld2(preT, jn) : make_double2(0.0, 0.0), tb = has_t ? ldadj(nd + 2, jn) : make_double2(0.0, 0.0);
        double2 zg[3], gb[3];
#pragma unroll
        for (int j = 0; j < 3; j++)
          if (j < nd) { zg[j] = ld2(preL + (long long)(1 + j) * PL, jn); gb[j] = ldadj(1 + j, jn); }
        double oq[2], ot[2], zb[2], og[3][2];
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const double sx = x ? s2.y : s2.x, qx = x ? q2.y : q2.x, hx = x ? h2.y : h2.x;
          const double zqx = x ? zq.y : zq.x, ztx = x ? zt.y : zt.x, tbx = x ? tb.y : tb.x;
          const double s1 = fma(-sx, sx, 1.0), s2d = -2.0 * sx * s1;
          double z0 = fma(s1, hx, s2d * fma(qx, zqx, tbx * ztx));
          double sq = 0.0;
          oq[x] = s1 * qx;
          ot[x] = s1 * tbx;
#pragma unroll
          for (int j = 0; j < 3; j++)
            if (j < nd) {
              const double z = x ? zg[j].y : zg[j].x, gx = x ? gb[j].y : gb[j].x;
              z0 = fma(s2d * gx, z, z0);
              sq = fma(z, z, sq);
              og[j][x] = fma(s1, gx, 2.0 * s2d * z * qx);
            }
          const double s3 = -2.0 * s1 * fma(-2.0 * sx, sx, s1);      // s''' = -2 s' (s' - 2 s^2)
          zb[x] = fma(s3 * qx, sq, z0);
        }
        if (live) {
          st2(preQ, jn, oq[0], oq[1]);
          if (has_t) st2(preT, jn, ot[0], ot[1]);
#pragma unroll
          for (int j = 0; j < 3; j++)
            if (j < nd) st2(preL + (long long)(1 + j) * PL, jn, og[j][0], og[j][1]);
          st2(preL, jn, zb[0], zb[1]);
        }
      }
      if (l == 0) break;
      // chain: adjoint of the previous layer's post-activation streams  pbar_s = zbar_s W_l
      const double* bh = wBH + (long long)(l - 1) * KSP * NT * 32;
      for (int st = 0; st < S; st++) {
        double dC[NT][2], dn[NT][2];
        const double* srcp = preL + (long long)st * PL;
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          const double2 z2 = ld2(srcp, jn);       // written by this same thread above
          dC[jn][0] = z2.x; dC[jn][1] = z2.y;
          dn[jn][0] = dn[jn][1] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ks++)
#pragma unroll
          for (int jn = 0; jn < NT; jn++) dmma(dn[jn][0], dn[jn][1], dC[ks >> 1][ks & 1], bh[(ks * NT + jn) * 32 + lane]);
        if (live) {
#pragma unroll
          for (int jn = 0; jn < NT; jn++) stadj(st, jn, dn[jn][0], dn[jn][1]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-gradient contraction over rows x streams for ONE layer:
//     G[out slot][in slot] += sum_st sum_rows Z_st[row][out slot] * P_st[row][in slot]
// Z_st / P_st are planes (act layout) or, for the output layer / layer 0, small row-major operands:
//   MODE 0 hidden layer : Z = ZBAR[l][st] planes, P = POST[l-1][st] planes (ones slot patched for st = 0: bias)
//   MODE 1 output layer : Z = obar [row][S][n_out] (out index = slot), P = POST[depth-1][st] (+ ones for st = 0)
//   MODE 2 layer 0      : Z = ZBAR[0][st] planes, P = input vector of stream st ([x_hat, t_hat, 1] / e_j / 0 / e_t)
// CTA = NT+1 warps, 64-row tiles, persistent over (stream, tile) pairs, cp.async double buffering into the same
// XOR-swizzled tiles as k_mlp_backward; per-CTA partials go to the section of the standard partial record
// (k_reduce_wgrad layout) that belongs to the layer.
// ---------------------------------------------------------------------------------------------
struct JetWgradArgs {
  const double* Z;        // planes of stream 0 (stream stride = plane) or obar
  const double* P;        // planes of stream 0 or nullptr (MODE 2)
  const double* pts;      // MODE 2
  long long n, rows_cap;
  int S, nd, n_out, n_in, ones_slot;
  double* part;           // per-CTA partial records
  long long part_stride;
  long long sec_off;      // offset of this layer's section inside a record
};

// tile row strides: 64 doubles for a full plane tile; the 8-column operands (MODE 1: Z = obar, MODE 2: P = input vector)
// take 16 (columns 0..7 XOR the swizzle {0, 4, 8, 12}), which lets two of those CTAs share an SM (k_jet_wgrad_io)
template <int MODE> __host__ __device__ constexpr int jet_wgrad_sz() { return MODE == 1 ? 16 : 64; }
template <int MODE> __host__ __device__ constexpr int jet_wgrad_sp() { return MODE == 2 ? 16 : 64; }
// NBUF = 2: tiles double-buffered inside the CTA (one CTA per SM for full tiles: 131 KB).  NBUF = 1 (hidden layers,
// PCX_JET_WGRAD_HIDDEN_NBUF): one buffer, 64 KB, TWO CTAs per SM -- the other CTA's DMMAs cover this one's loads.
#ifndef PCX_JET_WGRAD_HIDDEN_NBUF
#define PCX_JET_WGRAD_HIDDEN_NBUF 1
#endif
template <int NT, int MODE, int NBUF = 2> __host__ __device__ constexpr size_t jet_wgrad_smem() {
  return (size_t)NBUF * 8 * (NT + 1) * (jet_wgrad_sz<MODE>() + jet_wgrad_sp<MODE>()) * sizeof(double);
}

template <int NT, int MODE, int NBUF = 2>
__device__ __forceinline__ void jet_wgrad_body(const MlpShape& s, const JetWgradArgs& A, unsigned char* smem_raw) {
  constexpr int NW = NT + 1, ROWS = 8 * NW, SZ = jet_wgrad_sz<MODE>(), SP = jet_wgrad_sp<MODE>(), KR = ROWS / 4;
  double* Zt = reinterpret_cast<double*>(smem_raw);            // [NBUF][ROWS*SZ]
  double* Pt = Zt + NBUF * ROWS * SZ;                          // [NBUF][ROWS*SP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  const int lrow = warp * 8 + r;
  const int gl = swz(lrow), gc = swz(c);
  const long long PL = (long long)NT * A.rows_cap * 8;
  const long long ntiles = (A.n + ROWS - 1) / ROWS;
  const long long nwork = ntiles * A.S;

  // accumulators: MODE 0: warp<NT: out block `warp` x in tiles 0..NT-2 ; warp NT: in tile NT-1 x all out blocks
  //               MODE 1: warp<NT: in tile `warp` (out rows = output index) ; MODE 2: warp<NT: out block `warp` x 8 inputs
  double acc[NT][2];
#pragma unroll
  for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = 0.0;

  auto load = [&](long long w, int buf) {
    const int st = (int)(w / ntiles);
    const long long tile = w - (long long)st * ntiles;
    const long long row = tile * ROWS + lrow;
    const bool ok = w < nwork && row < A.n;
    const long long rr = ok ? row : 0;
    double* zt = Zt + buf * ROWS * SZ;
    double* pt = Pt + buf * ROWS * SP;
    if constexpr (MODE != 1) {
      const double* zp = A.Z + (long long)st * PL;
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
        cp_async16(&zt[lrow * SZ + ((8 * jn + 2 * c) ^ gl)], zp + ((long long)jn * A.rows_cap + rr) * 8 + 2 * c, ok);
    }
    if constexpr (MODE != 2) {
      const double* pp = A.P + (long long)st * PL;
#pragma unroll
      for (int jn = 0; jn < NT; jn++)
        cp_async16(&pt[lrow * SP + ((8 * jn + 2 * c) ^ gl)], pp + ((long long)jn * A.rows_cap + rr) * 8 + 2 * c, ok);
    }
    cp_async_commit();
    if constexpr (MODE == 1) {        // Z tile columns 0..7 = obar[row][st][o]
      double v0 = 0.0, v1 = 0.0;
      if (ok) {
        if (2 * c < A.n_out) v0 = A.Z[(row * A.S + st) * A.n_out + 2 * c];
        if (2 * c + 1 < A.n_out) v1 = A.Z[(row * A.S + st) * A.n_out + 2 * c + 1];
      }
      *reinterpret_cast<double2*>(&zt[lrow * SZ + ((2 * c) ^ gl)]) = make_double2(v0, v1);
    }
    if constexpr (MODE == 2) {        // P tile columns 0..7 = input vector of stream st
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int x = 0; x < 2; x++) {
        const int col = 2 * c + x;
        double v = 0.0;
        if (ok) {
          if (st == 0) {
            if (col < A.nd) v = (A.pts[row * (A.nd + 1) + col] - s.x_min[col]) / s.x_scale[col];
            else if (col == A.nd) { const double t = A.pts[row * (A.nd + 1) + A.nd]; v = s.normalize_time ? (t - s.t_min) / s.t_scale : t; }
            else if (col == A.n_in) v = 1.0;                          // bias column
          } else if (st <= A.nd) {
            if (col == st - 1) v = 1.0 / s.x_scale[st - 1];
          } else if (st == A.nd + 2) {
            if (col == A.nd) v = s.normalize_time ? 1.0 / s.t_scale : 1.0;
          }
        }
        if (x == 0) v0 = v; else v1 = v;
      }
      *reinterpret_cast<double2*>(&pt[lrow * SP + ((2 * c) ^ gl)]) = make_double2(v0, v1);
    }
  };

  int buf = 0;
  if constexpr (NBUF == 2) load(blockIdx.x, 0);
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x, buf ^= (NBUF - 1)) {
    const int st = (int)(w / ntiles);
    const long long tile = w - (long long)st * ntiles;
    const bool live = tile * ROWS + lrow < A.n;
    if constexpr (NBUF == 1) {
      __syncthreads();        // everybody is done reading the tile of the previous work item
      load(w, 0);
    }
    cp_async_wait_all();
    double* zt = Zt + buf * ROWS * SZ;
    double* pt = Pt + buf * ROWS * SP;
    if (MODE != 2 && st == 0 && live && c == ((A.ones_slot & 7) >> 1)) pt[lrow * SP + (A.ones_slot ^ gl)] = 1.0;
    __syncthreads();          // tile `buf` complete; everybody is done reading tile `buf^1`
    if constexpr (NBUF == 2) load(w + gridDim.x, buf ^ 1);
    if constexpr (MODE == 0) {
      if (warp < NT) {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++) {
          const double a = zt[(4 * ks + c) * SZ + ((8 * warp + r) ^ gc)];
#pragma unroll
          for (int jc = 0; jc < NT - 1; jc++) dmma(acc[jc][0], acc[jc][1], a, pt[(4 * ks + c) * SP + ((8 * jc + r) ^ gc)]);
        }
      } else {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++) {
          const double b = pt[(4 * ks + c) * SP + ((8 * (NT - 1) + r) ^ gc)];
#pragma unroll
          for (int jr = 0; jr < NT; jr++) dmma(acc[jr][0], acc[jr][1], zt[(4 * ks + c) * SZ + ((8 * jr + r) ^ gc)], b);
        }
      }
    } else if constexpr (MODE == 1) {
      if (warp < NT) {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++)
          dmma(acc[0][0], acc[0][1], zt[(4 * ks + c) * SZ + (r ^ gc)], pt[(4 * ks + c) * SP + ((8 * warp + r) ^ gc)]);
      }
    } else {
      if (warp < NT) {
#pragma unroll 4
        for (int ks = 0; ks < KR; ks++)
          dmma(acc[0][0], acc[0][1], zt[(4 * ks + c) * SZ + ((8 * warp + r) ^ gc)], pt[(4 * ks + c) * SP + (r ^ gc)]);
      }
    }
  }
  cp_async_wait_all();
  double* P = A.part + (long long)blockIdx.x * A.part_stride + A.sec_off;
  constexpr int WPd = 8 * NT;
  if constexpr (MODE == 0) {
    if (warp < NT) {
#pragma unroll
      for (int jc = 0; jc < NT - 1; jc++) {
        P[(8 * warp + r) * WPd + 8 * jc + 2 * c] = acc[jc][0];
        P[(8 * warp + r) * WPd + 8 * jc + 2 * c + 1] = acc[jc][1];
      }
    } else {
#pragma unroll
      for (int jr = 0; jr < NT; jr++) {
        P[(8 * jr + r) * WPd + 8 * (NT - 1) + 2 * c] = acc[jr][0];
        P[(8 * jr + r) * WPd + 8 * (NT - 1) + 2 * c + 1] = acc[jr][1];
      }
    }
  } else if constexpr (MODE == 1) {
    if (warp < NT) {
      P[r * WPd + 8 * warp + 2 * c] = acc[0][0];
      P[r * WPd + 8 * warp + 2 * c + 1] = acc[0][1];
    }
  } else {
    if (warp < NT) {
      P[(8 * warp + r) * 8 + 2 * c] = acc[0][0];
      P[(8 * warp + r) * 8 + 2 * c + 1] = acc[0][1];
    }
  }
}

template <int NT, int MODE, int NBUF = 2>
__global__ void __launch_bounds__(32 * (NT + 1), NBUF == 1 ? 2 : 1) k_jet_wgrad(const __grid_constant__ MlpShape s, const __grid_constant__ JetWgradArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  jet_wgrad_body<NT, MODE, NBUF>(s, A, smem_raw);
}

// Layer 0 (MODE 2, blockIdx.y = 0) and the output layer (MODE 1, blockIdx.y = 1) in ONE launch: both are small contractions
// that stream whole planes for a handful of DMMAs (ncu r2: 19 % / 15 % DMMA, 8 warps per SM, 0.21 + 0.27 ms back to back
// at 262 144 points); with their narrow operand tiles they fit two to an SM and run side by side.
template <int NT>
__global__ void __launch_bounds__(32 * (NT + 1), 2) k_jet_wgrad_io(const __grid_constant__ MlpShape s, const __grid_constant__ JetWgradArgs A2,
                                                                   const __grid_constant__ JetWgradArgs A1) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (blockIdx.y == 0) jet_wgrad_body<NT, 2>(s, A2, smem_raw);
  else jet_wgrad_body<NT, 1>(s, A1, smem_raw);
}

// ---------------------------------------------------------------------------------------------
// Per point: Dirichlet wrapper on the jet, strong-form residual, loss partial and adjoint seeds.
//   tab [n][n_out][2*(nd+3)] = (a, d_j a.., lap a, a_t, b, d_j b.., lap b, b_t) or nullptr (a = 0, b = 1)
//   u = a + b z;  d_j u = d_j a + d_j b z + b d_j z;  lap u = lap a + lap b z + 2 sum_j d_j b d_j z + b lap z;
//   u_t = a_t + b_t z + b z_t                                           (fields.py:101 differentiated)
//   r = c_t u_t - c_lap lap u - src - target                            (poisson.py:26-29, heat_equation.py:11-17)
//   loss partial = scale * r^2;  obar = d(scale r^2)/d zjet
// ---------------------------------------------------------------------------------------------
struct JetPointArgs {
  const double* zjet;     // [n][S][n_out]
  const double* tab;      // or nullptr
  const double* src;      // [n][n_out] or nullptr
  const double* target;   // [n][n_out] or nullptr
  long long n;
  int S, nd, n_out;
  double c_t, c_lap, scale;
  double* u; double* grad_u; double* lap_u; double* u_t;   // outputs (any may be nullptr)
  double* resid;          // [n][n_out] or nullptr
  double* obar;           // [n][S][n_out] or nullptr
  double* part;           // [gridDim] loss partials or nullptr
};

__global__ void __launch_bounds__(256) k_jet_point(const __grid_constant__ JetPointArgs A) {
  __shared__ double sh[8];
  double lsum = 0.0;
  const long long total = A.n * A.n_out;
  const int S = A.S, nd = A.nd, K = nd + 3;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long row = k / A.n_out;
    const int i = (int)(k - row * A.n_out);
    const double* zj = A.zjet + row * S * A.n_out + i;     // stride n_out between streams
    const bool has_t = S == nd + 3;
    const double z = zj[0], lz = zj[(long long)(nd + 1) * A.n_out], zt = has_t ? zj[(long long)(nd + 2) * A.n_out] : 0.0;
    double a = 0.0, b = 1.0, la = 0.0, lb = 0.0, at = 0.0, bt = 0.0;
    double da[3] = {0.0, 0.0, 0.0}, db[3] = {0.0, 0.0, 0.0};
    if (A.tab) {
      const double* t = A.tab + k * 2 * K;
      a = t[0]; la = t[1 + nd]; at = t[2 + nd];
      b = t[K]; lb = t[K + 1 + nd]; bt = t[K + 2 + nd];
      for (int j = 0; j < nd; j++) { da[j] = t[1 + j]; db[j] = t[K + 1 + j]; }
    }
    double lap = fma(lb, z, la) + b * lz;
    for (int j = 0; j < nd; j++) {
      const double gz = zj[(long long)(1 + j) * A.n_out];
      lap = fma(2.0 * db[j], gz, lap);
      if (A.grad_u) A.grad_u[k * nd + j] = fma(db[j], z, da[j]) + b * gz;
    }
    const double ut = fma(bt, z, at) + b * zt;
    if (A.u) A.u[k] = fma(b, z, a);
    if (A.lap_u) A.lap_u[k] = lap;
    if (A.u_t) A.u_t[k] = ut;
    double rres = A.c_t * ut - A.c_lap * lap;
    if (A.src) rres -= A.src[k];
    if (A.resid) A.resid[k] = rres;
    if (A.target) rres -= A.target[k];
    lsum = fma(rres, rres, lsum);
    if (A.obar) {
      const double rb = 2.0 * A.scale * rres;             // d loss / d r
      const double lapb = -A.c_lap * rb, utb = A.c_t * rb;
      double* ob = A.obar + row * S * A.n_out + i;
      ob[0] = fma(lb, lapb, bt * utb);
      for (int j = 0; j < nd; j++) ob[(long long)(1 + j) * A.n_out] = 2.0 * db[j] * lapb;
      ob[(long long)(nd + 1) * A.n_out] = b * lapb;
      if (has_t) ob[(long long)(nd + 2) * A.n_out] = b * utb;
    }
  }
  if (A.part) {
    const double t = block_sum<256>(lsum, sh);
    if (threadIdx.x == 0) A.part[blockIdx.x] = t * A.scale;
  }
}

}  // namespace pcx
