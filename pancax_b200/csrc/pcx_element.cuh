// pcx_element.cuh -- element sweeps: gather, geometry, constitutive update, energy reduction,
// internal-force / tangent-action scatter.
//
// Replaces pancax/physics_kernels/base.py:297-361 (element_integral, element_interpolants,
// potential_energy_on_block, potential_energy_and_internal_force) with
// pancax/fem/function_space.py:63-89 inlined, for all elements of the mesh in one launch.
//
// One thread = one element (all its quadrature points), everything in registers:
//   J[k][j]   = sum_a dN[q][a][k] X[a][j]                 (function_space.py:64,83)
//   Gu[i][k]  = sum_a U[a][i] dN[q][a][k]                 parent-space gradient of u
//   H         = Gu Jinv^T   (= U^T gradN, base.py:323-325 with gradN = (Jinv dN^T)^T, fs:87-89)
//   fe[a][i] += sum_k (JxW * P Jinv)[i][k] dN[q][a][k]     (the VJP JAX derives for base.py:356-361)
// so the (nq, nnpe, nd) shape-gradient tensor XLA materialises (3.2 GB at 16.7 M qp) never exists.
// Energy: per-thread fixed-order sum over q, warp shuffle tree, fixed-order block partial;
// k_reduce_partials finishes in fixed order -> bit-reproducible Pi.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pcx_models.cuh"

namespace pcx {

constexpr int TAB_MAX_NQ = 9;
constexpr int TAB_MAX_NNPE = 8;

struct ParentTab {
  int nq, nnpe, nd;
  double w[TAB_MAX_NQ];
  double N[TAB_MAX_NQ * TAB_MAX_NNPE];        // [q][a]
  double dN[TAB_MAX_NQ * TAB_MAX_NNPE * 3];   // [q][a][k]
};

// blockIdx.y = time step within the batch: per-step arrays are offset by t * (their per-step size).
struct ElemArgs {
  long long ne;
  long long nn;
  const double* coords;    // [nn][ND]
  const int32_t* conns;    // [ne][NNPE]
  const double* us;        // [nn][NDOF]
  const double* v;         // tangent direction [nn][NDOF] (TANGENT mode)
  const double* src_q;     // Poisson source at qps [ne][nq] or nullptr
  double* fe;              // [ne][NNPE][NDOF] element vectors (deterministic mode) or nullptr
  double* f_atomic;        // [nn][NDOF] atomic target (fast mode) or nullptr
  double* epart;           // [gridDim] energy partials (FORCE mode) or nullptr
  double* ppart;           // [gridDim][PCX_MAX_PROPS] property-gradient partials or nullptr
  int* flags;              // device flag word: bit0 = det(F) <= 0 or non-finite energy seen
  const double* props_dev; // effective properties [PCX_MAX_PROPS] (device: trainable ones change every step)
  int nprops;
  int want_force;
  int plane_stress;        // PCX_FORM_PLANE_STRESS: F_33 from the per-point root of sigma_33 = 0 (2D solid only)
  double* f33_out;         // [T][ne][nq] the roots found by the force sweep, kept for ...
  const double* f33_in;    // ... the tangent sweep at the same displacements (nullptr: solve again)
  // SimpleFeFv (hyperviscoelasticity/simple_fefv.py): state Fv [ne][nq][n_prony][9] per quadrature point.
  // Forward sweep of one time step: z_old -> z_new and the energy.  Reverse sweep: z_old and the adjoint lam_in of
  // the step's new state -> the stress-like S (scattered like P) and lam_out, the adjoint of the old state.
  int n_prony;
  double prony_kappa[PCX_MAX_PRONY], prony_c[PCX_MAX_PRONY];
  const double* z_old;
  double* z_new;
  const double* lam_in;
  double* lam_out;
  double* qbuf;            // [ne][nq][NDOF][ND] parent-space point fluxes (k_visco_qp -> k_visco_scatter)
  double wpsi;             // k_visco_qp MODE 3: weight of the step's energy in the loss (the flux then carries every term)
  double* vscr;            // k_visco_qp MODE 5: scratch [ne][nq][n_prony][9] (tangent of the old-state adjoint along E_33)
  double* pp_P;            // k_visco_qp MODE 2 / 5: [ne][nq][9] PK1 stress d psi / d grad_u at fixed old state, or nullptr (pcx_state_step)
};

template <int ND>
__device__ __forceinline__ double inv_det(const double* J, double* Ji) {
  if constexpr (ND == 2) {
    const double det = J[0] * J[3] - J[1] * J[2];
    const double id = 1.0 / det;
    Ji[0] = J[3] * id; Ji[1] = -J[1] * id; Ji[2] = -J[2] * id; Ji[3] = J[0] * id;
    return det;
  } else {
    double cof[9];
    cofactor3(J, cof);
    const double det = J[0] * cof[0] + J[1] * cof[1] + J[2] * cof[2];
    const double id = 1.0 / det;
    // inverse = cof^T / det
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Ji[3 * i + j] = cof[3 * j + i] * id;
    return det;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block-wide deterministic sum; result valid in thread 0
template <int NTHREADS>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NTHREADS / 32; i++) t += sh[i];
  }
  return t;
}

// MODEL: 0 = Poisson, else PCX_MODEL_*.  TANGENT: false -> energy + internal force,
// true -> stiffness action K v (forward-mode duals through the same constitutive source).
// VAR: 0 = plain; 1 = PlaneStress (2D solid: per-point root find).  SimpleFeFv has its own per-point kernels below
// (k_visco_qp).  Compile-time variants: as run-time branches they cost the plain kernels half of their speed in register spills.
// Register budget of the 3D sweeps (measured, tools/elem_ab.cu, profiles/r02_elem_ab.jsonl; Hex8 128^3 x 2 steps, NeoHookean):
//   r1  everything in registers (X, U, accumulators: 144 registers + the constitutive chain): 255 registers, 8 warps per SM,
//       every DFMA waits out the latency of the one before it (ncu: FP64 pipe 49 % active)                        1.64 ms
//   r2a accumulators in shared memory [slot][thread], X / U in registers (236 registers, 8 warps)                     1.61 ms
//   r2b FORCE sweep: X / U in shared memory, re-read per quadrature point as 24 conflict-free LDS.128 ([pair][thread][2]
//       planes, asm volatile so that the compiler does not hoist them back into registers), accumulators in registers:
//       166 registers = 3 CTAs = 12 warps per SM                                                                      1.30 ms
//       (same with LDS.64 planes 1.35; U only + accumulators in shared memory 1.46; X, U AND accumulators in shared memory
//       1.55: the read-modify-write of the accumulators is what hurts; 4 CTAs per SM at 128 registers spill: 1.73;
//       two quadrature points per iteration: 1.54).  Results are bit-identical across the variants.
//       Other models, same change: Gent 1.99 -> 1.65, Swanson 2.72 -> 2.32 (340 B of spills at 168 registers and still ahead),
//       Hencky 3.17 -> 2.52 ms.
//   r2b TANGENT sweep (dual numbers; X, U and the direction V in shared memory, 72 KB per CTA, 3 CTAs per SM; one time step):
//       NeoHookean 1.075 -> 1.00 ms, Hencky 2.27 -> 1.92 ms (2 CTAs per SM: 0.97 / 2.11).
//   2D: everything in registers; the tangent sweep is 22 % faster with 3 CTAs per SM.
#ifndef PCX_ELEM_QUNROLL
#define PCX_ELEM_QUNROLL 1
#endif
#ifndef PCX_ELEM3D_MINB
#define PCX_ELEM3D_MINB 3          // force sweep (PCX_ELEM3D_MINB_T: tangent sweep)
#endif
// PCX_ELEM3D_XU_SMEM (force sweep): 0 = X / U in registers; 1 = U in shared memory; 2 = X and U ([slot][thread] planes,
// LDS.64); 3 = X and U in [pair][thread][2] planes read with LDS.128
#ifndef PCX_ELEM3D_XU_SMEM
#define PCX_ELEM3D_XU_SMEM 3
#endif
#ifndef PCX_ELEM3D_ACC_SMEM
#define PCX_ELEM3D_ACC_SMEM 0      // force sweep: nodal accumulators in shared memory (PCX_ELEM3D_ACC_SMEM_T: tangent sweep)
#endif
#ifndef PCX_ELEM3D_ACC_SMEM_T
#define PCX_ELEM3D_ACC_SMEM_T 0
#endif
__host__ __device__ constexpr bool elem_acc_smem(int nd, bool tangent) { return nd == 3 && (tangent ? PCX_ELEM3D_ACC_SMEM_T : PCX_ELEM3D_ACC_SMEM); }
#ifndef PCX_ELEM3D_XU_SMEM_T
#define PCX_ELEM3D_XU_SMEM_T 3     // tangent sweep: 0 or 3 (X, U and V in [pair][thread][2] planes)
#endif
#ifndef PCX_ELEM3D_MINB_T
#define PCX_ELEM3D_MINB_T 3
#endif
__host__ __device__ constexpr int elem_xu_smem(int nd, bool tangent) { return nd == 3 ? (tangent ? PCX_ELEM3D_XU_SMEM_T : PCX_ELEM3D_XU_SMEM) : 0; }
// bytes of dynamic shared memory a k_element launch needs (128 threads per CTA)
__host__ __device__ constexpr int elem_dyn_smem_bytes(int nnpe, int nd, int ndof, bool tangent) {
  return ((elem_acc_smem(nd, tangent) ? nnpe * ndof : 0) + (elem_xu_smem(nd, tangent) >= 1 ? nnpe * ndof : 0) +
          (elem_xu_smem(nd, tangent) >= 2 ? nnpe * nd : 0) + (elem_xu_smem(nd, tangent) == 3 && tangent ? nnpe * ndof : 0)) * 128 * (int)sizeof(double);
}
constexpr int elem_min_blocks(int nd, bool tangent) { return nd == 3 ? (tangent ? PCX_ELEM3D_MINB_T : PCX_ELEM3D_MINB) : (tangent ? 3 : 0); }   // 0 = no request
// 128-bit shared-memory load of two consecutive doubles; asm volatile: re-issued per quadrature point, never hoisted
__device__ __forceinline__ void lds_v2(const double* p, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(p)));
}
template <int NNPE, int ND, int NDOF, int MODEL, bool TANGENT, int VAR = 0>
__global__ void __launch_bounds__(128, elem_min_blocks(ND, TANGENT)) k_element(const __grid_constant__ ParentTab tab, const __grid_constant__ ElemArgs A0) {
  constexpr int NTH = 128;
  constexpr bool ACC_SMEM = elem_acc_smem(ND, TANGENT);
  __shared__ double sh[NTH / 32];
  constexpr int XU = elem_xu_smem(ND, TANGENT);
  // dynamic shared memory (elem_dyn_smem_bytes): [slot][thread] planes of the accumulators, then U, then X
  extern __shared__ __align__(16) double elem_dyn_s[];
  double (*acc_s)[NTH] = reinterpret_cast<double (*)[NTH]>(elem_dyn_s);
  double (*u_s)[NTH] = reinterpret_cast<double (*)[NTH]>(elem_dyn_s + (ACC_SMEM ? NNPE * NDOF : 0) * NTH);
  double (*x_s)[NTH] = reinterpret_cast<double (*)[NTH]>(elem_dyn_s + ((ACC_SMEM ? NNPE * NDOF : 0) + (XU >= 1 ? NNPE * NDOF : 0)) * NTH);
  constexpr int UOFF = (ACC_SMEM ? NNPE * NDOF : 0) * NTH, XOFF = UOFF + (XU >= 1 ? NNPE * NDOF : 0) * NTH;   // XU == 3: [pair][thread][2]
  constexpr int VOFF = XOFF + (XU >= 2 ? NNPE * ND : 0) * NTH;                                                  // tangent, XU == 3
  constexpr bool V_SMEM = TANGENT && XU == 3;
  const long long e = blockIdx.x * (long long)NTH + threadIdx.x;
  // per-time-step views
  struct {
    long long ne; const double* coords; const int32_t* conns; const double* us; const double* v; const double* src_q;
    double* fe; double* f_atomic; double* epart; double* ppart; int* flags; Props props; int nprops; int want_force;
  } A;
  {
    const long long ts = blockIdx.y;
    A.ne = A0.ne; A.coords = A0.coords; A.conns = A0.conns; A.src_q = A0.src_q; A.flags = A0.flags;
    A.nprops = A0.nprops; A.want_force = A0.want_force;
    A.us = A0.us + ts * A0.nn * NDOF;
    A.v = A0.v ? A0.v + ts * A0.nn * NDOF : nullptr;
    A.fe = A0.fe ? A0.fe + ts * A0.ne * (NNPE * NDOF) : nullptr;
    A.f_atomic = A0.f_atomic ? A0.f_atomic + ts * A0.nn * NDOF : nullptr;
    A.epart = A0.epart ? A0.epart + ts * gridDim.x : nullptr;
    A.ppart = A0.ppart ? A0.ppart + ts * (long long)gridDim.x * PCX_MAX_PROPS : nullptr;
#pragma unroll
    for (int k = 0; k < PCX_MAX_PROPS; k++) A.props.p[k] = (MODEL != 0 && k < A0.nprops) ? __ldg(A0.props_dev + k) : 0.0;
    props_derive<MODEL>(A.props);
  }
  const bool live = e < A.ne;
  double Wsum = 0.0;
  double dps[PCX_MAX_PROPS];
#pragma unroll
  for (int k = 0; k < PCX_MAX_PROPS; k++) dps[k] = 0.0;
  const bool want_dp = A.ppart != nullptr;
  bool bad = false;

  if (live) {
    int32_t cn[NNPE];
    if constexpr (NNPE % 4 == 0) {
      const int4* cp = reinterpret_cast<const int4*>(A.conns + e * NNPE);
#pragma unroll
      for (int q = 0; q < NNPE / 4; q++) {
        int4 t = __ldg(cp + q);
        cn[4 * q] = t.x; cn[4 * q + 1] = t.y; cn[4 * q + 2] = t.z; cn[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int a = 0; a < NNPE; a++) cn[a] = __ldg(A.conns + e * NNPE + a);
    }
    double X[XU >= 2 ? 1 : NNPE][XU >= 2 ? 1 : ND], U[XU >= 1 ? 1 : NNPE][XU >= 1 ? 1 : NDOF], acc[ACC_SMEM ? 1 : NNPE][ACC_SMEM ? 1 : NDOF];
    double V[TANGENT && !V_SMEM ? NNPE : 1][NDOF];
#define PCX_X(a, j) (XU >= 2 ? *reinterpret_cast<const volatile double*>(&x_s[(a) * ND + (j)][threadIdx.x]) : X[XU >= 2 ? 0 : (a)][XU >= 2 ? 0 : (j)])
#define PCX_U(a, i) (XU == 3 ? *reinterpret_cast<const volatile double*>(&elem_dyn_s[UOFF + ((((a) * NDOF + (i)) >> 1) * NTH + threadIdx.x) * 2 + (((a) * NDOF + (i)) & 1)]) : XU >= 1 ? *reinterpret_cast<const volatile double*>(&u_s[(a) * NDOF + (i)][threadIdx.x]) : U[XU >= 1 ? 0 : (a)][XU >= 1 ? 0 : (i)])
#pragma unroll
    for (int a = 0; a < NNPE; a++) {
#pragma unroll
      for (int j = 0; j < ND; j++) {
        const double xv = __ldg(A.coords + (long long)cn[a] * ND + j);
        if constexpr (XU == 3) elem_dyn_s[XOFF + (((a * ND + j) >> 1) * NTH + threadIdx.x) * 2 + ((a * ND + j) & 1)] = xv;
        else if constexpr (XU == 2) x_s[a * ND + j][threadIdx.x] = xv;
        else X[a][j] = xv;
      }
#pragma unroll
      for (int i = 0; i < NDOF; i++) {
        const double uv = __ldg(A.us + (long long)cn[a] * NDOF + i);
        if constexpr (XU == 3) elem_dyn_s[UOFF + (((a * NDOF + i) >> 1) * NTH + threadIdx.x) * 2 + ((a * NDOF + i) & 1)] = uv;
        else if constexpr (XU >= 1) u_s[a * NDOF + i][threadIdx.x] = uv;
        else U[a][i] = uv;
        if constexpr (ACC_SMEM) acc_s[a * NDOF + i][threadIdx.x] = 0.0;
        else acc[a][i] = 0.0;
        if constexpr (V_SMEM) elem_dyn_s[VOFF + (((a * NDOF + i) >> 1) * NTH + threadIdx.x) * 2 + ((a * NDOF + i) & 1)] = __ldg(A.v + (long long)cn[a] * NDOF + i);
        else if constexpr (TANGENT) V[a][i] = __ldg(A.v + (long long)cn[a] * NDOF + i);
      }
    }

    constexpr int QU = (ND == 3) ? PCX_ELEM_QUNROLL : 1;
#pragma unroll QU
    for (int q = 0; q < tab.nq; q++) {
      const double* dN = tab.dN + q * (TAB_MAX_NNPE * 3);
      double J[ND * ND], Ji[ND * ND];
#pragma unroll
      for (int k = 0; k < ND * ND; k++) J[k] = 0.0;
      if constexpr (XU == 3) {
#pragma unroll
        for (int p = 0; p < NNPE * ND / 2; p++) {
          double xv[2];
          lds_v2(elem_dyn_s + XOFF + (p * NTH + threadIdx.x) * 2, xv[0], xv[1]);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int a = (2 * p + h) / ND, j = (2 * p + h) % ND;
#pragma unroll
            for (int k = 0; k < ND; k++) J[k * ND + j] = fma(dN[a * 3 + k], xv[h], J[k * ND + j]);
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < NNPE; a++)
#pragma unroll
          for (int j = 0; j < ND; j++) {
            // shared-memory copies are re-read per quadrature point (volatile: not hoisted back into registers)
            const double xv = PCX_X(a, j);
#pragma unroll
            for (int k = 0; k < ND; k++) J[k * ND + j] = fma(dN[a * 3 + k], xv, J[k * ND + j]);
          }
      }
      const double detJ = inv_det<ND>(J, Ji);
      const double JxW = detJ * tab.w[q];
      // parent-space gradient of u (and of v)
      double Gu[NDOF][ND];
      double Gv[TANGENT ? NDOF : 1][ND];
#pragma unroll
      for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int k = 0; k < ND; k++) { Gu[i][k] = 0.0; if constexpr (TANGENT) Gv[i][k] = 0.0; }
      if constexpr (XU == 3) {
#pragma unroll
        for (int p = 0; p < NNPE * NDOF / 2; p++) {
          double uv[2], vv[2] = {0.0, 0.0};
          lds_v2(elem_dyn_s + UOFF + (p * NTH + threadIdx.x) * 2, uv[0], uv[1]);
          if constexpr (V_SMEM) lds_v2(elem_dyn_s + VOFF + (p * NTH + threadIdx.x) * 2, vv[0], vv[1]);
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int a = (2 * p + h) / NDOF, i = (2 * p + h) % NDOF;
#pragma unroll
            for (int k = 0; k < ND; k++) {
              Gu[i][k] = fma(uv[h], dN[a * 3 + k], Gu[i][k]);
              if constexpr (V_SMEM) Gv[i][k] = fma(vv[h], dN[a * 3 + k], Gv[i][k]);
            }
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < NNPE; a++)
#pragma unroll
          for (int i = 0; i < NDOF; i++) {
            const double uv = PCX_U(a, i);
#pragma unroll
            for (int k = 0; k < ND; k++) {
              Gu[i][k] = fma(uv, dN[a * 3 + k], Gu[i][k]);
              if constexpr (TANGENT) Gv[i][k] = fma(V[a][i], dN[a * 3 + k], Gv[i][k]);
            }
          }
      }
      // H[i][j] = sum_k Gu[i][k] Ji[j][k]
      double H[NDOF][ND], dH[TANGENT ? NDOF : 1][ND];
#pragma unroll
      for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int j = 0; j < ND; j++) {
          double s = 0.0, sv = 0.0;
#pragma unroll
          for (int k = 0; k < ND; k++) {
            s = fma(Gu[i][k], Ji[j * ND + k], s);
            if constexpr (TANGENT) sv = fma(Gv[i][k], Ji[j * ND + k], sv);
          }
          H[i][j] = s;
          if constexpr (TANGENT) dH[i][j] = sv;
        }

      // physics: stress-like output S[i][j] (P or dP) and, for Poisson, the u-conjugate term
      double S[NDOF][ND];
      double su = 0.0;  // d(density)/du_q (Poisson: -f)
      if constexpr (MODEL == 0) {
        // poisson.py:16-19: 0.5 grad_u.grad_u - f u
        const double fq = A.src_q ? A.src_q[e * tab.nq + q] : 0.0;
        if constexpr (!TANGENT) {
          double uq = 0.0, w = 0.0;
#pragma unroll
          for (int a = 0; a < NNPE; a++) uq = fma(tab.N[q * TAB_MAX_NNPE + a], PCX_U(a, 0), uq);
#pragma unroll
          for (int j = 0; j < ND; j++) { w = fma(H[0][j], H[0][j], w); S[0][j] = H[0][j]; }
          Wsum = fma(JxW, 0.5 * w - fq * uq, Wsum);
          su = -fq;
        } else {
#pragma unroll
          for (int j = 0; j < ND; j++) S[0][j] = dH[0][j];
        }
      } else {
        // solid_mechanics.py:102-109: plane strain embeds the 2x2 gradient in a 3x3 of zeros
        // (tensor_math.py:64-65); F = grad_u + I (mechanics/base.py:19-21)
        if constexpr (!TANGENT) {
          double F[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
          for (int i = 0; i < NDOF; i++)
#pragma unroll
            for (int j = 0; j < ND; j++) F[3 * i + j] += H[i][j];
          if constexpr (VAR == 1) {
            if (!plane_stress_solve<MODEL>(F, A.props)) bad = true;
            if (A0.f33_out) A0.f33_out[((long long)blockIdx.y * A0.ne + e) * tab.nq + q] = F[8];
          }
          double W, P[9], dp[PCX_MAX_PROPS];
          model_eval<MODEL, double>(F, A.props, W, P, dp, want_dp);
          if (!(W == W) || fabs(W) > 1.0e300) bad = true;
          Wsum = fma(JxW, W, Wsum);
          if (want_dp) {
#pragma unroll
            for (int k = 0; k < PCX_MAX_PROPS; k++)
              if (k < A.nprops) dps[k] = fma(JxW, dp[k], dps[k]);
          }
#pragma unroll
          for (int i = 0; i < NDOF; i++)
#pragma unroll
            for (int j = 0; j < ND; j++) S[i][j] = P[3 * i + j];
        } else {
          Dual F[9];
#pragma unroll
          for (int i = 0; i < 9; i++) F[i] = Dual((i % 4 == 0) ? 1.0 : 0.0, 0.0);
#pragma unroll
          for (int i = 0; i < NDOF; i++)
#pragma unroll
            for (int j = 0; j < ND; j++) { F[3 * i + j].v += H[i][j]; F[3 * i + j].d = dH[i][j]; }
          Dual W, P[9], dp[PCX_MAX_PROPS];
          if constexpr (VAR == 1) {
            double Fv[9];
#pragma unroll
            for (int i = 0; i < 9; i++) Fv[i] = F[i].v;
            if (A0.f33_in) Fv[8] = A0.f33_in[((long long)blockIdx.y * A0.ne + e) * tab.nq + q];
            else if (!plane_stress_solve<MODEL>(Fv, A.props)) bad = true;
            F[8].v = Fv[8];
            plane_stress_tangent<MODEL>(F, A.props, W, P, dp, want_dp, A.nprops);
          } else {
            model_eval<MODEL, Dual>(F, A.props, W, P, dp, want_dp);
          }
          if (want_dp) {
            // d(P:dH)/dp = tangent of dW/dp along dH
#pragma unroll
            for (int k = 0; k < PCX_MAX_PROPS; k++)
              if (k < A.nprops) dps[k] = fma(JxW, dp[k].d, dps[k]);
          }
#pragma unroll
          for (int i = 0; i < NDOF; i++)
#pragma unroll
            for (int j = 0; j < ND; j++) S[i][j] = P[3 * i + j].d;
        }
      }
      if (A.want_force) {
        // Q[i][k] = JxW sum_j S[i][j] Ji[j][k];  acc[a][i] += sum_k Q[i][k] dN[a][k]
        double Q[NDOF][ND];
#pragma unroll
        for (int i = 0; i < NDOF; i++)
#pragma unroll
          for (int k = 0; k < ND; k++) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < ND; j++) s = fma(S[i][j], Ji[j * ND + k], s);
            Q[i][k] = s * JxW;
          }
#pragma unroll
        for (int a = 0; a < NNPE; a++)
#pragma unroll
          for (int i = 0; i < NDOF; i++) {
            double s;
            if constexpr (ACC_SMEM) s = acc_s[a * NDOF + i][threadIdx.x];
            else s = acc[a][i];
#pragma unroll
            for (int k = 0; k < ND; k++) s = fma(Q[i][k], dN[a * 3 + k], s);
            if constexpr (MODEL == 0 && !TANGENT) s = fma(JxW * su, tab.N[q * TAB_MAX_NNPE + a], s);
            if constexpr (ACC_SMEM) acc_s[a * NDOF + i][threadIdx.x] = s;
            else acc[a][i] = s;
          }
      }
    }
    if (A.want_force) {
      if (A.fe) {
        double* o = A.fe + e * (NNPE * NDOF);
#pragma unroll
        for (int a = 0; a < NNPE; a++)
#pragma unroll
          for (int i = 0; i < NDOF; i++) o[a * NDOF + i] = ACC_SMEM ? acc_s[a * NDOF + i][threadIdx.x] : acc[a][i];
      } else {
#pragma unroll
        for (int a = 0; a < NNPE; a++)
#pragma unroll
          for (int i = 0; i < NDOF; i++)
            atomicAdd(A.f_atomic + (long long)cn[a] * NDOF + i, ACC_SMEM ? acc_s[a * NDOF + i][threadIdx.x] : acc[a][i]);
      }
    }
  }
#undef PCX_X
#undef PCX_U
  if (bad) atomicOr(A.flags, 1);
  if (A.epart) {
    const double t = block_sum<NTH>(Wsum, sh);
    if (threadIdx.x == 0) A.epart[blockIdx.x] = t;
  }
  if (A.ppart) {
    for (int k = 0; k < A.nprops; k++) {
      const double t = block_sum<NTH>(dps[k], sh);
      if (threadIdx.x == 0) A.ppart[(long long)blockIdx.x * PCX_MAX_PROPS + k] = t;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// SimpleFeFv sweeps, one thread per (element, quadrature point).  The state update and its reverse sweep are ~10x the
// arithmetic of a hyperelastic point and hold ~150 live doubles: one thread per ELEMENT (k_element, VAR = 2) keeps the
// element's X/U/acc (72 doubles for Hex8) live across them and runs only ne threads, so it spills and under-fills the
// machine.  Here a point's thread owns nothing but its point: geometry -> F -> branches -> (BWD) Q = JxW S J^-T in
// parent space [ne][nq][NDOF][ND], which k_visco_scatter contracts with the parent gradients into the element vector.
// MODE 0: z_old -> z_new, energy (and equilibrium property-gradient) partials per block.
// MODE 1: z_old, lam_in -> lam_out, Q  (reverse sweep of the path-dependent EnergyLoss).
// MODE 2: MODE 0 plus the flux of the step's internal force f = dPi/du at FIXED old state (lam = 0 reverse branch): the
//         forward sweep of the path-dependent residual / reaction losses (weak_form_loss_functions.py:151-192,242-276).
// MODE 3: reverse sweep of those losses.  With v = dloss/df of this step on the nodes, the flux carries
//         wpsi dpsi/dF + (adjoint of the new state pulled back) + d/deps dpsi/dF (u + eps v), and lam_out additionally
//         d/deps dpsi/dFv_old (u + eps v) -- the second-order terms from ONE forward-mode (dual) evaluation of the
//         equilibrium model and the closed-form tangent of the branch (fefv_branch); property-gradient partials are the
//         dual parts of JxW dW/dp.
// MODE 4 / 5 (2D, PlaneStress with state, path-dependent EnergyLoss): MODE 0 / 1 with the out-of-plane stretch from the
//         per-point root of the TOTAL P_33 at fixed old state (plane_stress_visco_solve; the roots are kept per step in
//         f33_out / f33_in).  The reverse sweep adds the implicit-function terms: with G_33 = dL/dF_33 of the point
//         (energy + adjoint of the new state) and the tangent (dP, dlam_old) of the lam = 0 branch along dF = E_33,
//         mu = -G_33 / dP_33, flux += mu dP_2D, lam_out += mu dlam_old, property partials mu d(dW/dp)
//         (dx/dF_ij = -dP_ij/dP_33 by the symmetry of the Hessian, dx/dz = -d(dpsi/dz)/dF_33 / dP_33).
// ----------------------------------------------------------------------------------------------
#ifndef PCX_VISCO_MINB_BWD
#define PCX_VISCO_MINB_BWD 3
#endif
#ifndef PCX_VISCO_MINB_FWD
#define PCX_VISCO_MINB_FWD 4
#endif
#ifndef PCX_VISCO_MINB_FWDF
#define PCX_VISCO_MINB_FWDF 3     // MODE 2 (196 registers unconstrained; 2 CTAs per SM: 5.59 ms per step of 10 sweeps at 2.1 M points, 3: 4.14)
#endif
#ifndef PCX_VISCO_MINB_BWDT
#define PCX_VISCO_MINB_BWDT 3     // MODE 3 (254 registers unconstrained)
#endif
template <int NNPE, int ND, int NDOF, int MODEL, int MODE>
__global__ void __launch_bounds__(128, MODE == 1 ? PCX_VISCO_MINB_BWD : MODE == 0 ? PCX_VISCO_MINB_FWD : MODE == 2 ? PCX_VISCO_MINB_FWDF : PCX_VISCO_MINB_BWDT) k_visco_qp(const __grid_constant__ ParentTab tab, const __grid_constant__ ElemArgs A0) {
  constexpr bool FLUX = MODE != 0 && MODE != 4;     // a point flux Q is written
  constexpr bool REV = MODE == 1 || MODE == 5;      // no energy partials
  constexpr int NTH = 128;
  __shared__ double sh[NTH / 32];
  const int nq = tab.nq;
  const long long npts = A0.ne * nq;
  const long long t = blockIdx.x * (long long)NTH + threadIdx.x;
  const bool live = t < npts;
  Props props;
#pragma unroll
  for (int k = 0; k < PCX_MAX_PROPS; k++) props.p[k] = (k < A0.nprops) ? __ldg(A0.props_dev + k) : 0.0;
  props_derive<MODEL>(props);
  const bool want_dp = MODE != 1 && A0.ppart != nullptr;   // MODE 5: the implicit-function part of the property gradient
  double Wq = 0.0;
  double dps[PCX_MAX_PROPS];
#pragma unroll
  for (int k = 0; k < PCX_MAX_PROPS; k++) dps[k] = 0.0;
  bool bad = false;
  if (live) {
    const long long e = t / nq;
    const int q = (int)(t - e * nq);
    const double* dN = tab.dN + q * (TAB_MAX_NNPE * 3);
    double J[ND * ND], Ji[ND * ND], Gu[NDOF][ND], Gv[MODE == 3 ? NDOF : 1][ND];
#pragma unroll
    for (int k = 0; k < ND * ND; k++) J[k] = 0.0;
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
      for (int k = 0; k < ND; k++) Gu[i][k] = 0.0;
    if constexpr (MODE == 3) {
#pragma unroll
      for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int k = 0; k < ND; k++) Gv[i][k] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < NNPE; a++) {
      const long long n = __ldg(A0.conns + e * NNPE + a);
#pragma unroll
      for (int j = 0; j < ND; j++) {
        const double x = __ldg(A0.coords + n * ND + j);
#pragma unroll
        for (int k = 0; k < ND; k++) J[k * ND + j] = fma(dN[a * 3 + k], x, J[k * ND + j]);
      }
#pragma unroll
      for (int i = 0; i < NDOF; i++) {
        const double u = __ldg(A0.us + n * NDOF + i);
#pragma unroll
        for (int k = 0; k < ND; k++) Gu[i][k] = fma(u, dN[a * 3 + k], Gu[i][k]);
        if constexpr (MODE == 3) {
          const double vv = __ldg(A0.v + n * NDOF + i);
#pragma unroll
          for (int k = 0; k < ND; k++) Gv[i][k] = fma(vv, dN[a * 3 + k], Gv[i][k]);
        }
      }
    }
    const double detJ = inv_det<ND>(J, Ji);
    const double JxW = detJ * tab.w[q];
    double F[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
      for (int j = 0; j < ND; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < ND; k++) s = fma(Gu[i][k], Ji[j * ND + k], s);
        F[3 * i + j] += s;
      }
    double W, P[9], dp[PCX_MAX_PROPS];
    const long long zq = t * (long long)A0.n_prony * 9;
    if constexpr (MODE == 3) {
      double dF[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int j = 0; j < ND; j++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < ND; k++) s = fma(Gv[i][k], Ji[j * ND + k], s);
          dF[3 * i + j] = s;
        }
      {
        Dual Fd[9], Wd, Pd[9], dpd[PCX_MAX_PROPS];
#pragma unroll
        for (int i = 0; i < 9; i++) Fd[i] = Dual(F[i], dF[i]);
        model_eval<MODEL, Dual>(Fd, props, Wd, Pd, dpd, want_dp);   // equilibrium branch, value + tangent along v
        W = Wd.v;
#pragma unroll
        for (int i = 0; i < 9; i++) P[i] = fma(A0.wpsi, Pd[i].v, Pd[i].d);
        if (want_dp) {
#pragma unroll
          for (int k = 0; k < PCX_MAX_PROPS; k++)
            if (k < A0.nprops) dp[k] = dpd[k].d;
        }
      }
      for (int b = 0; b < A0.n_prony; b++) {
        double zo[9], li[9], lo[9];
#pragma unroll
        for (int i = 0; i < 9; i++) { zo[i] = A0.z_old[zq + b * 9 + i]; li[i] = A0.lam_in[zq + b * 9 + i]; }
        W += fefv_branch(F, zo, A0.prony_kappa[b], A0.prony_c[b], nullptr, li, P, lo, A0.wpsi, dF, P, lo);
#pragma unroll
        for (int i = 0; i < 9; i++) A0.lam_out[zq + b * 9 + i] = lo[i];
      }
    } else if constexpr (MODE == 4 && ND == 2) {
      const double* zo = A0.z_old + zq;
      const double x_prev = A0.f33_in ? A0.f33_in[t] - 1.0 : __longlong_as_double(0x7ff8000000000000LL);   // previous step's root
      if (!plane_stress_visco_solve<MODEL>(F, props, A0.n_prony, zo, A0.prony_kappa, A0.prony_c, x_prev)) bad = true;
      if (A0.f33_out) A0.f33_out[t] = F[8];
      model_eval<MODEL, double>(F, props, W, P, dp, want_dp);
      for (int b = 0; b < A0.n_prony; b++) {
        double zn[9], lo[9];
        W += fefv_branch(F, zo + b * 9, A0.prony_kappa[b], A0.prony_c[b], zn, nullptr, nullptr, lo);
#pragma unroll
        for (int i = 0; i < 9; i++) A0.z_new[zq + b * 9 + i] = zn[i];
      }
    } else if constexpr (MODE == 5 && ND == 2) {
      const double* zo = A0.z_old + zq;
      F[8] = A0.f33_in[t];
      double dPt[9];
      {
        Dual Fd[9], Wd, Pd[9], dpd[PCX_MAX_PROPS];
#pragma unroll
        for (int i = 0; i < 9; i++) Fd[i] = Dual(F[i], i == 8 ? 1.0 : 0.0);
        model_eval<MODEL, Dual>(Fd, props, Wd, Pd, dpd, want_dp);
        W = Wd.v;
#pragma unroll
        for (int i = 0; i < 9; i++) { P[i] = Pd[i].v; dPt[i] = Pd[i].d; }
        if (want_dp) {
#pragma unroll
          for (int k = 0; k < PCX_MAX_PROPS; k++)
            if (k < A0.nprops) dp[k] = dpd[k].d;
        }
      }
      const double dF[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0};
      for (int b = 0; b < A0.n_prony; b++) {
        double li[9], lo[9], dlo[9];
#pragma unroll
        for (int i = 0; i < 9; i++) { li[i] = A0.lam_in[zq + b * 9 + i]; dlo[i] = 0.0; }
        W += fefv_branch(F, zo + b * 9, A0.prony_kappa[b], A0.prony_c[b], nullptr, li, P, lo, 1.0, dF, dPt, dlo);
#pragma unroll
        for (int i = 0; i < 9; i++) { A0.lam_out[zq + b * 9 + i] = lo[i]; A0.vscr[zq + b * 9 + i] = dlo[i]; }
      }
      const double mu = -P[8] / dPt[8];
#pragma unroll
      for (int i = 0; i < 9; i++) P[i] = fma(mu, dPt[i], P[i]);
      if (A0.pp_P) {          // pcx_state_step: PK1 stress at the plane-stress root (lam_in = 0)
#pragma unroll
        for (int i = 0; i < 9; i++) A0.pp_P[t * 9 + i] = P[i];
      }
      for (int b = 0; b < A0.n_prony; b++)
#pragma unroll
        for (int i = 0; i < 9; i++) A0.lam_out[zq + b * 9 + i] = fma(mu, A0.vscr[zq + b * 9 + i], A0.lam_out[zq + b * 9 + i]);
      if (want_dp) {
#pragma unroll
        for (int k = 0; k < PCX_MAX_PROPS; k++)
          if (k < A0.nprops) dp[k] *= mu;
      }
    } else {
      model_eval<MODEL, double>(F, props, W, P, dp, want_dp);   // equilibrium branch (simple_fefv.py:33-40)
      for (int b = 0; b < A0.n_prony; b++) {
        double zo[9], zn[9], li[9], lo[9];
#pragma unroll
        for (int i = 0; i < 9; i++) zo[i] = A0.z_old[zq + b * 9 + i];
        if constexpr (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 9; i++) li[i] = A0.lam_in[zq + b * 9 + i];
          W += fefv_branch(F, zo, A0.prony_kappa[b], A0.prony_c[b], nullptr, li, P, lo);
#pragma unroll
          for (int i = 0; i < 9; i++) A0.lam_out[zq + b * 9 + i] = lo[i];
        } else {
          W += fefv_branch(F, zo, A0.prony_kappa[b], A0.prony_c[b], zn, nullptr, MODE == 2 ? P : nullptr, lo);
#pragma unroll
          for (int i = 0; i < 9; i++) A0.z_new[zq + b * 9 + i] = zn[i];
        }
      }
      if constexpr (MODE == 2) {
        if (A0.pp_P) {
#pragma unroll
          for (int i = 0; i < 9; i++) A0.pp_P[t * 9 + i] = P[i];
        }
      }
    }
    if (!(W == W) || fabs(W) > 1.0e300) bad = true;
    if constexpr (FLUX) {
      double* Q = A0.qbuf + t * (NDOF * ND);
#pragma unroll
      for (int i = 0; i < NDOF; i++)
#pragma unroll
        for (int k = 0; k < ND; k++) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < ND; j++) s = fma(P[3 * i + j], Ji[j * ND + k], s);
          Q[i * ND + k] = s * JxW;
        }
    }
    if constexpr (MODE != 1) {
      if constexpr (!REV) Wq = JxW * W;
      if (want_dp) {
#pragma unroll
        for (int k = 0; k < PCX_MAX_PROPS; k++)
          if (k < A0.nprops) dps[k] = JxW * dp[k];
      }
    }
  }
  if (bad) atomicOr(A0.flags, 1);
  if constexpr (MODE != 1) {
    if (A0.epart) {
      const double s = block_sum<NTH>(Wq, sh);
      if (threadIdx.x == 0) A0.epart[blockIdx.x] = s;
    }
    if (want_dp) {
      for (int k = 0; k < A0.nprops; k++) {
        const double s = block_sum<NTH>(dps[k], sh);
        if (threadIdx.x == 0) A0.ppart[(long long)blockIdx.x * PCX_MAX_PROPS + k] = s;
      }
    }
  }
}

// element vector from the parent-space point fluxes: fe[a][i] = sum_q sum_k Q[q][i][k] dN[q][a][k].
// One thread per (element, node): the block's fluxes (EPB elements, contiguous) are staged through shared memory with
// coalesced loads, each thread contracts its node's NDOF components and writes them contiguously (HBM-bound:
// nq NDOF ND + NNPE NDOF doubles per element).
template <int NNPE>
__host__ __device__ constexpr int visco_scatter_epb() { return 128 / NNPE; }
template <int NNPE, int ND, int NDOF>
__global__ void __launch_bounds__(128) k_visco_scatter(const __grid_constant__ ParentTab tab, const __grid_constant__ ElemArgs A0) {
  constexpr int EPB = visco_scatter_epb<NNPE>();
  constexpr int QW = NDOF * ND;
  __shared__ double s_dN[TAB_MAX_NQ * TAB_MAX_NNPE * 3];
  __shared__ double s_Q[EPB * (TAB_MAX_NQ * QW + 1)];   // +1: elements of one warp on different banks
  const int nq = tab.nq;
  const int stride = nq * QW + 1;
  for (int i = threadIdx.x; i < TAB_MAX_NQ * TAB_MAX_NNPE * 3; i += 128) s_dN[i] = tab.dN[i];
  const long long e0 = blockIdx.x * (long long)EPB;
  const int nloc = (int)min((long long)EPB, A0.ne - e0);
  const double* Qg = A0.qbuf + e0 * nq * QW;
  for (int i = threadIdx.x; i < nloc * nq * QW; i += 128) {
    const int el = i / (nq * QW);
    s_Q[el * stride + (i - el * nq * QW)] = Qg[i];
  }
  __syncthreads();
  const int el = threadIdx.x / NNPE, a = threadIdx.x - el * NNPE;
  if (el >= nloc) return;
  double acc[NDOF];
#pragma unroll
  for (int i = 0; i < NDOF; i++) acc[i] = 0.0;
  for (int q = 0; q < nq; q++) {
    const double* Q = s_Q + el * stride + q * QW;
    const double* dN = s_dN + q * (TAB_MAX_NNPE * 3) + a * 3;
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
      for (int k = 0; k < ND; k++) acc[i] = fma(Q[i * ND + k], dN[k], acc[i]);
  }
  const long long e = e0 + el;
  if (A0.fe) {
#pragma unroll
    for (int i = 0; i < NDOF; i++) A0.fe[(e * NNPE + a) * NDOF + i] = acc[i];
  } else {
    const long long n = __ldg(A0.conns + e * NNPE + a);
#pragma unroll
    for (int i = 0; i < NDOF; i++) atomicAdd(A0.f_atomic + n * NDOF + i, acc[i]);
  }
}

// ----------------------------------------------------------------------------------------------
// Standalone geometry kernel (north_star item 1): gradN (ne,nq,nnpe,nd), JxW (ne,nq), x_q (ne,nq,nd).
// function_space.py:63-89 and base.py:322.  HBM-bound (writes ~ (nnpe*nd + 1 + nd) * 8 B per qp).
// One thread per (element, qp); the element's connectivity row is staged through shared memory
// with a coalesced block load.
// ----------------------------------------------------------------------------------------------
template <int NNPE, int ND>
__global__ void __launch_bounds__(256) k_geometry(const __grid_constant__ ParentTab tab, long long ne,
                                                  const double* __restrict__ coords, const int32_t* __restrict__ conns,
                                                  double* __restrict__ gradN, double* __restrict__ JxW,
                                                  double* __restrict__ xq, int epb /* elements per block */) {
  extern __shared__ int32_t s_conn[];  // [epb][NNPE]
  const long long e0 = (long long)blockIdx.x * epb;
  const int nloc = (int)min((long long)epb, ne - e0);
  for (int i = threadIdx.x; i < nloc * NNPE; i += blockDim.x) s_conn[i] = __ldg(conns + e0 * NNPE + i);
  __syncthreads();
  const int nq = tab.nq;
  for (int t = threadIdx.x; t < nloc * nq; t += blockDim.x) {
    const int el = t / nq, q = t - el * nq;
    const long long e = e0 + el;
    double X[NNPE][ND];
#pragma unroll
    for (int a = 0; a < NNPE; a++)
#pragma unroll
      for (int j = 0; j < ND; j++) X[a][j] = __ldg(coords + (long long)s_conn[el * NNPE + a] * ND + j);
    const double* dN = tab.dN + q * (TAB_MAX_NNPE * 3);
    double J[ND * ND], Ji[ND * ND];
#pragma unroll
    for (int k = 0; k < ND; k++)
#pragma unroll
      for (int j = 0; j < ND; j++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; a++) s = fma(dN[a * 3 + k], X[a][j], s);
        J[k * ND + j] = s;
      }
    const double det = inv_det<ND>(J, Ji);
    const long long qi = e * nq + q;
    if (JxW) JxW[qi] = det * tab.w[q];
    if (gradN) {
#pragma unroll
      for (int a = 0; a < NNPE; a++)
#pragma unroll
        for (int j = 0; j < ND; j++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < ND; k++) s = fma(Ji[j * ND + k], dN[a * 3 + k], s);
          gradN[(qi * NNPE + a) * ND + j] = s;
        }
    }
    if (xq) {
#pragma unroll
      for (int j = 0; j < ND; j++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; a++) s = fma(tab.N[q * TAB_MAX_NNPE + a], X[a][j], s);
        xq[qi * ND + j] = s;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Post-processing fields at the quadrature points: what pancax's element_pp wrappers
// evaluate for the Exodus output -- deformation_gradient (kinematic method), I1_bar (kinematic), pk1_stress
// (constitutive method = jacfwd(energy) w.r.t. the 3x3 gradient): pancax/physics_kernels/base.py:24-158,
// solid_mechanics.py:115-170, constitutive_models/mechanics/base.py:19-21,50-63,112-118.
// One thread per (element, qp), connectivity rows staged through shared memory; HBM-bound
// (19 doubles = 152 B written per point).  F, P: [ne][nq][3][3] row-major; I1bar: [ne][nq].
// ----------------------------------------------------------------------------------------------
template <int NNPE, int ND, int NDOF, int MODEL, bool PSTRESS = false>
__global__ void __launch_bounds__(256) k_element_pp(const __grid_constant__ ParentTab tab, long long ne,
                                                    const double* __restrict__ coords, const int32_t* __restrict__ conns,
                                                    const double* __restrict__ us, const double* __restrict__ props_dev,
                                                    int nprops, double* __restrict__ Fout, double* __restrict__ I1out,
                                                    double* __restrict__ Pout, int epb) {
  extern __shared__ int32_t s_conn[];  // [epb][NNPE]
  __shared__ double s_tile[8 * 32 * 9];
  __shared__ double s_dN[TAB_MAX_NQ * 25];   // parent gradients, row stride 25: lanes with different q hit different banks
  for (int i = threadIdx.x; i < TAB_MAX_NQ * TAB_MAX_NNPE * 3; i += blockDim.x) s_dN[(i / 24) * 25 + i % 24] = tab.dN[i];
  const long long e0 = (long long)blockIdx.x * epb;
  const int nloc = (int)min((long long)epb, ne - e0);
  for (int i = threadIdx.x; i < nloc * NNPE; i += blockDim.x) s_conn[i] = __ldg(conns + e0 * NNPE + i);
  Props props;
#pragma unroll
  for (int k = 0; k < PCX_MAX_PROPS; k++) props.p[k] = (k < nprops) ? __ldg(props_dev + k) : 0.0;
  props_derive<MODEL>(props);
  __syncthreads();
  const int nq = tab.nq;
  const int npts = nloc * nq;
  for (int tb = (threadIdx.x & ~31); tb < npts; tb += blockDim.x) {   // warp-uniform trip count
    const int t = tb + (threadIdx.x & 31);
    const bool livept = t < npts;
    const int tt = livept ? t : npts - 1;                              // dead lanes recompute the last point
    const int el = tt / nq, q = tt - el * nq;
    const double* dN = s_dN + q * 25;
    double J[ND * ND], Ji[ND * ND], Gu[NDOF][ND];
#pragma unroll
    for (int k = 0; k < ND * ND; k++) J[k] = 0.0;
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
      for (int k = 0; k < ND; k++) Gu[i][k] = 0.0;
#pragma unroll
    for (int a = 0; a < NNPE; a++) {
      const long long n = s_conn[el * NNPE + a];
      double X[ND], U[NDOF];
#pragma unroll
      for (int j = 0; j < ND; j++) X[j] = __ldg(coords + n * ND + j);
#pragma unroll
      for (int i = 0; i < NDOF; i++) U[i] = __ldg(us + n * NDOF + i);
#pragma unroll
      for (int k = 0; k < ND; k++) {
#pragma unroll
        for (int j = 0; j < ND; j++) J[k * ND + j] = fma(dN[a * 3 + k], X[j], J[k * ND + j]);
#pragma unroll
        for (int i = 0; i < NDOF; i++) Gu[i][k] = fma(U[i], dN[a * 3 + k], Gu[i][k]);
      }
    }
    inv_det<ND>(J, Ji);
    // F = I + [grad_u embedded in 3x3] (plane strain: tensor_math.py:64-65)
    double F[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
    for (int i = 0; i < NDOF; i++)
#pragma unroll
      for (int j = 0; j < ND; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < ND; k++) s = fma(Gu[i][k], Ji[j * ND + k], s);
        F[3 * i + j] += s;
      }
    if constexpr (PSTRESS) plane_stress_solve<MODEL>(F, props);   // solid_mechanics.py:49-71
    const long long qi = e0 * nq + t;
    if (I1out) {
      // mechanics/base.py:50-63: J^(-2/3) tr(F F^T)
      double cof[9];
      cofactor3(F, cof);
      const double Jd = F[0] * cof[0] + F[1] * cof[1] + F[2] * cof[2];
      double I1 = 0.0;
#pragma unroll
      for (int i = 0; i < 9; i++) I1 = fma(F[i], F[i], I1);
      if (livept) I1out[qi] = s_pow_m23(Jd) * I1;
    }
    double P[9];
    if (Pout) {
      double W, dp[PCX_MAX_PROPS];
      model_eval<MODEL, double>(F, props, W, P, dp, false);
    }
    // 3x3 outputs: a lane's 9 doubles are 72 B apart from its neighbour's, so a direct store touches 32
    // sectors per instruction.  Transpose through a per-warp shared tile (stride 9: conflict-free) and
    // write the warp's 32 x 9 doubles as 9 fully coalesced 256 B rows.  Points of a warp are consecutive
    // (t = threadIdx.x + k * blockDim.x), hence contiguous in the output.
    const int lane = threadIdx.x & 31;
    double* tile = s_tile + (threadIdx.x >> 5) * (32 * 9);
    const long long q0 = qi - lane;                       // first point of this warp's batch
    const int nvalid = min(32, npts - tb);                // live lanes of the batch
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
      double* dst = pass == 0 ? Fout : Pout;
      if (!dst) continue;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 9; i++) tile[lane * 9 + i] = pass == 0 ? F[i] : P[i];
      __syncwarp();
      for (int i = lane; i < nvalid * 9; i += 32) dst[q0 * 9 + i] = tile[i];
    }
  }
}

// node-centric fixed-order assembly for T time steps:
//   out[ts][n][i] = alpha * sum_{k in adj(n)} fe[ts][adj_idx[k]][i] (+ beta*add[ts][n][i])
__global__ void k_assemble(long long nn, int ndof, int T, long long nent, const long long* __restrict__ adj_ptr,
                           const int32_t* __restrict__ adj_idx, const double* __restrict__ fe,
                           double alpha, const double* __restrict__ add, double beta, double* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)T * nn * ndof) return;
  const long long ts = t / (nn * ndof);
  const long long rem = t - ts * nn * ndof;
  const long long n = rem / ndof;
  const int i = (int)(rem - n * ndof);
  fe += ts * nent * ndof;
  double s = 0.0;
  for (long long k = adj_ptr[n]; k < adj_ptr[n + 1]; k++) s += fe[(long long)adj_idx[k] * ndof + i];
  s *= alpha;
  if (add) s = fma(beta, add[t], s);
  out[t] = s;
}

// out = alpha*x + beta*y (y nullable)
__global__ void k_axpby(long long n, double alpha, const double* __restrict__ x, double beta,
                        const double* __restrict__ y, double* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  double s = alpha * x[t];
  if (y) s = fma(beta, y[t], s);
  out[t] = s;
}

// partial sums for the residual norm and the reaction force (fixed order per block)
__global__ void __launch_bounds__(256) k_resid_react_partials(const double* __restrict__ f, long long ndofs_per_t,
                                                              const long long* __restrict__ uidx,
                                                              long long nu, const int32_t* __restrict__ rnodes, long long nr,
                                                              int ndof, int rdof, double* __restrict__ part /* [T][grid][2] */) {
  __shared__ double sh[8];
  f += (long long)blockIdx.y * ndofs_per_t;
  part += (long long)blockIdx.y * gridDim.x * 2;
  double s2 = 0.0, sr = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nu; k += stride) {
    const double v = f[uidx[k]];
    s2 = fma(v, v, s2);
  }
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nr; k += stride)
    sr += f[(long long)rnodes[k] * ndof + rdof];
  const double a = block_sum<256>(s2, sh);
  const double b = block_sum<256>(sr, sh);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
  }
}

// fixed-order reduction of `n` partials with `ncomp` interleaved components, one block per row:
// out[row*out_stride + c] = sum_i part[(row*n + i)*stride + c]
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ part, long long n, int stride, int ncomp,
                                                         double* __restrict__ out, int out_stride) {
  __shared__ double sh[8];
  part += (long long)blockIdx.x * n * stride;
  out += (long long)blockIdx.x * out_stride;
  for (int c = 0; c < ncomp; c++) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += part[i * stride + c];
    const double t = block_sum<256>(s, sh);
    if (threadIdx.x == 0) out[c] = t;
  }
}

}  // namespace pcx
