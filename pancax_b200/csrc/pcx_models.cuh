// pcx_models.cuh -- per-quadrature-point constitutive math (device code).
//
// What pancax obtains by tracing `model.energy` through jax.jacfwd / jax.grad
// (pancax/constitutive_models/mechanics/base.py:112-118, physics_kernels/base.py:356-361) is written
// here in closed form: W(F), P = dW/dF and dW/d(property).  Every function is a template on the
// scalar type so the SAME source evaluated on dual numbers (value + one tangent) yields the
// material-tangent action dP = (d^2W/dF^2):dF and d(P:dF)/d(property) needed by the residual losses.
#pragma once
#include <math.h>
#include "../../include/pcx_abi.h"

namespace pcx {

// ------------------------------------------------------------------ dual numbers (forward mode)
struct Dual {
  double v, d;
  __host__ __device__ Dual() {}
  __host__ __device__ Dual(double v_) : v(v_), d(0.0) {}
  __host__ __device__ Dual(double v_, double d_) : v(v_), d(d_) {}
};
__host__ __device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__host__ __device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__host__ __device__ inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__host__ __device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
__host__ __device__ inline Dual operator/(Dual a, Dual b) {
  double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__host__ __device__ inline Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
__host__ __device__ inline Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
__host__ __device__ inline Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
__host__ __device__ inline Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
__host__ __device__ inline Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
__host__ __device__ inline Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
__host__ __device__ inline Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }
__host__ __device__ inline Dual operator/(double a, Dual b) {
  double q = a / b.v;
  return Dual(q, -q * b.d / b.v);
}
__host__ __device__ inline Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }

__host__ __device__ inline double value_of(double x) { return x; }
__host__ __device__ inline double value_of(Dual x) { return x.v; }
__host__ __device__ inline double tangent_of(double) { return 0.0; }
__host__ __device__ inline double tangent_of(Dual x) { return x.d; }

__host__ __device__ inline double s_log(double x) { return log(x); }
__host__ __device__ inline Dual s_log(Dual x) { return Dual(log(x.v), x.d / x.v); }
// x^p for a real exponent p (x > 0)
__host__ __device__ inline double s_pow(double x, double p) { return pow(x, p); }
__host__ __device__ inline Dual s_pow(Dual x, double p) {
  double r = pow(x.v, p);
  return Dual(r, p * r / x.v * x.d);
}

// x^p from a precomputed lx = log(x): exp(p lx).  CUDA's pow() carries an extended-precision log internally (~3x the
// FP64 instructions of log + exp); here |p lx| is O(1), so exp(p log x) is good to a few ulp -- and Swanson raises
// the same base to two different powers, which then share one log.
__host__ __device__ inline double s_pow_l(double, double lx, double p) { return exp(p * lx); }
__host__ __device__ inline Dual s_pow_l(Dual x, Dual lx, double p) {
  const double r = exp(p * lx.v);
  return Dual(r, p * r / x.v * x.d);
}

// x^(-2/3) (x > 0): rcbrt is ~4x cheaper than pow on the FP64 pipe and as accurate (1 ulp)
__device__ inline double s_pow_m23(double x) { const double r = rcbrt(x); return r * r; }
__device__ inline Dual s_pow_m23(Dual x) {
  const double r = rcbrt(x.v), r2 = r * r;
  return Dual(r2, (-2.0 / 3.0) * r2 / x.v * x.d);
}

// ------------------------------------------------------------------ 3x3 helpers (row-major F[3*i+j])
template <class T>
__host__ __device__ inline void cofactor3(const T* F, T* C) {
  C[0] = F[4] * F[8] - F[5] * F[7];
  C[1] = F[5] * F[6] - F[3] * F[8];
  C[2] = F[3] * F[7] - F[4] * F[6];
  C[3] = F[2] * F[7] - F[1] * F[8];
  C[4] = F[0] * F[8] - F[2] * F[6];
  C[5] = F[1] * F[6] - F[0] * F[7];
  C[6] = F[1] * F[5] - F[2] * F[4];
  C[7] = F[2] * F[3] - F[0] * F[5];
  C[8] = F[0] * F[4] - F[1] * F[3];
}

struct Props {
  double p[PCX_MAX_PROPS];
  double aux[6];   // per-model constants derived from p once per thread (props_derive), not once per quadrature point
};

// Volumetric part shared by NeoHookean and Gent: W_vol = 0.5 K (0.5 (J^2 - 1) - ln J)
// (neohookean.py:31, gent.py:37).

// ------------------------------------------------------------------ NeoHookean (neohookean.py:20-34)
//   W = 0.5 K (0.5 (J^2-1) - ln J) + 0.5 G (I1bar - 3),  I1bar = J^(-2/3) tr(F F^T)
//   P = [0.5 K (J - 1/J) - (G/3) I1bar / J] cof F + G J^(-2/3) F
template <class T>
__device__ inline void neohookean(const T* F, const Props& pr, T& W, T* P, T* dWdp, bool want_dp) {
  const double K = pr.p[0], G = pr.p[1];
  T cof[9];
  cofactor3(F, cof);
  T J = F[0] * cof[0] + F[1] * cof[1] + F[2] * cof[2];
  T I1 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2] + F[3] * F[3] + F[4] * F[4] + F[5] * F[5] +
         F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
  T Jm23 = s_pow_m23(J);
  T I1bar = Jm23 * I1;
  T wK = 0.5 * (0.5 * (J * J - 1.0) - s_log(J));
  T wG = 0.5 * (I1bar - 3.0);
  W = K * wK + G * wG;
  T Jinv = 1.0 / J;
  T a = 0.5 * K * (J - Jinv) - (G / 3.0) * I1bar * Jinv;
  T b = G * Jm23;
#pragma unroll
  for (int i = 0; i < 9; i++) P[i] = a * cof[i] + b * F[i];
  if (want_dp) {
    dWdp[0] = wK;
    dWdp[1] = wG;
  }
}

// ------------------------------------------------------------------ Gent (gent.py:22-39)
//   W_dev = -0.5 G Jm ln(1 - (I1bar-3)/Jm); I1bar is replaced by the constant Jm+3-0.001 when it
//   exceeds it (lax.cond, gent.py:30-34): zero F-gradient in that branch, Jm-gradient kept.
template <class T>
__device__ inline void gent(const T* F, const Props& pr, T& W, T* P, T* dWdp, bool want_dp) {
  const double K = pr.p[0], G = pr.p[1], Jm = pr.p[2];
  T cof[9];
  cofactor3(F, cof);
  T J = F[0] * cof[0] + F[1] * cof[1] + F[2] * cof[2];
  T I1 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2] + F[3] * F[3] + F[4] * F[4] + F[5] * F[5] +
         F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
  T Jm23 = s_pow_m23(J);
  T I1bar = Jm23 * I1;
  const bool clamped = value_of(I1bar) > Jm + 3.0 - 0.001;
  T wK = 0.5 * (0.5 * (J * J - 1.0) - s_log(J));
  T Jinv = 1.0 / J;
  T a = 0.5 * K * (J - Jinv);
  T Wdev, c1;
  T dWdJm;
  if (!clamped) {
    T x = (I1bar - 3.0) / Jm;
    T lg = s_log(1.0 - x);
    Wdev = -0.5 * G * Jm * lg;
    c1 = 0.5 * G / (1.0 - x);
    dWdJm = -0.5 * G * (lg + x / (1.0 - x));
  } else {
    const double I1c = Jm + 3.0 - 0.001;
    const double x = (I1c - 3.0) / Jm;
    const double lg = log(1.0 - x);
    Wdev = T(-0.5 * G * Jm * lg);
    c1 = T(0.0);
    // d/dJm of -0.5 G Jm ln(0.001/Jm) = -0.5 G (ln(0.001/Jm) - 1)
    dWdJm = T(-0.5 * G * (lg - 1.0));
  }
  W = K * wK + Wdev;
  T ca = a - c1 * (2.0 / 3.0) * I1bar * Jinv;
  T cb = 2.0 * c1 * Jm23;
#pragma unroll
  for (int i = 0; i < 9; i++) P[i] = ca * cof[i] + cb * F[i];
  if (want_dp) {
    dWdp[0] = wK;
    dWdp[1] = Wdev / G;
    dWdp[2] = dWdJm;
  }
}

// ------------------------------------------------------------------ Swanson (swanson.py:28-65)
//   props: K, A1, P1, B1, Q1, C1, R1, cutoff_strain
template <class T>
__device__ inline T swanson_term_dexp(T t, double tc, double A, double Pw) {
  // d/dPw of A/(Pw+1) (t^(Pw+1) - tc^(Pw+1))
  T lt = s_log(t);
  T tp = s_pow_l(t, lt, Pw + 1.0);
  double cp = pow(tc, Pw + 1.0);
  double inv = 1.0 / (Pw + 1.0);
  return A * ((tp * (lt * inv - inv * inv)) - cp * (log(tc) * inv - inv * inv));
}

template <class T>
__device__ inline void swanson(const T* F, const Props& pr, T& W, T* P, T* dWdp, bool want_dp) {
  const double K = pr.p[0], A1 = pr.p[1], P1 = pr.p[2], B1 = pr.p[3], Q1 = pr.p[4], C1 = pr.p[5],
               R1 = pr.p[6], cs = pr.p[7];
  const double tc = (1.0 / 3.0) * (3.0 + cs * cs) - 1.0;
  T cof[9];
  cofactor3(F, cof);
  T J = F[0] * cof[0] + F[1] * cof[1] + F[2] * cof[2];
  // C = F^T F
  T C[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
  T I1 = C[0] + C[4] + C[8];
  T trC2 = C[0] * C[0] + C[4] * C[4] + C[8] * C[8] + 2.0 * (C[1] * C[1] + C[2] * C[2] + C[5] * C[5]);
  T I2 = 0.5 * (I1 * I1 - trC2);
  T Jm23 = s_pow_m23(J);
  T Jm43 = Jm23 * Jm23;
  T I1bar = Jm23 * I1;
  T I2bar = Jm43 * I2;
  T t1 = (1.0 / 3.0) * I1bar - 1.0 + tc;
  T t2 = (1.0 / 3.0) * I2bar - 1.0 + tc;
  T lJ = s_log(J);
  T wK = J * lJ - J + 1.0;
  T lt1 = s_log(t1), lt2 = s_log(t2);
  T t1P = s_pow_l(t1, lt1, P1), t1R = s_pow_l(t1, lt1, R1), t2Q = s_pow_l(t2, lt2, Q1);
  T wA = pr.aux[3] * (t1P * t1 - pr.aux[0]);   // 1.5 / (P1 + 1) (t1^(P1+1) - tc^(P1+1)), constants from props_derive
  T wB = pr.aux[4] * (t2Q * t2 - pr.aux[1]);
  T wC = pr.aux[5] * (t1R * t1 - pr.aux[2]);
  W = K * wK + A1 * wA + B1 * wB + C1 * wC;
  T dJ = K * lJ;
  T d1 = 0.5 * (A1 * t1P + C1 * t1R);
  T d2 = 0.5 * B1 * t2Q;
  // P = dJ cof + d1 (2 Jm23 F - (2/3) I1bar/J cof) + d2 (Jm43 (2 I1 F - 2 F C) - (4/3) I2bar/J cof)
  T Jinv = 1.0 / J;
  T ccof = dJ - d1 * (2.0 / 3.0) * I1bar * Jinv - d2 * (4.0 / 3.0) * I2bar * Jinv;
  T cF = 2.0 * d1 * Jm23 + 2.0 * d2 * Jm43 * I1;
  T cFC = -2.0 * d2 * Jm43;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      T FC = F[3 * i] * C[j] + F[3 * i + 1] * C[3 + j] + F[3 * i + 2] * C[6 + j];
      P[3 * i + j] = ccof * cof[3 * i + j] + cF * F[3 * i + j] + cFC * FC;
    }
  if (want_dp) {
    dWdp[0] = wK;
    dWdp[1] = wA;
    dWdp[2] = 1.5 * swanson_term_dexp(t1, tc, A1, P1);
    dWdp[3] = wB;
    dWdp[4] = 1.5 * swanson_term_dexp(t2, tc, B1, Q1);
    dWdp[5] = wC;
    dWdp[6] = 1.5 * swanson_term_dexp(t1, tc, C1, R1);
    dWdp[7] = T(0.0);
  }
}

// constants of a model that depend on the properties only: evaluated once per thread, before the quadrature loop
template <int MODEL>
__device__ inline void props_derive(Props& pr) {
#pragma unroll
  for (int k = 0; k < 6; k++) pr.aux[k] = 0.0;
  if constexpr (MODEL == PCX_MODEL_SWANSON) {
    const double cs = pr.p[7], tc = (1.0 / 3.0) * (3.0 + cs * cs) - 1.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double e1 = pr.p[2 + 2 * k] + 1.0;   // P1 + 1, Q1 + 1, R1 + 1
      pr.aux[k] = pow(tc, e1);
      pr.aux[3 + k] = 1.5 / e1;
    }
  }
}

// ------------------------------------------------------------------ Hencky (hencky.py:10-18)
// Closed-form eigen decomposition of a symmetric 3x3 tensor, following the algorithm of
// pancax/math/tensor_math.py:76-295 step by step (trace shift, trigonometric root with a Pade
// approximant of cos(acos(x)/3), pivoted QR for the first eigenvector, 2x2 Wilkinson step,
// degenerate-case fallbacks, ascending sort) so that degenerate inputs (C = I at t = 0) take the
// same branches as the reference.
__device__ inline double cos_acos_third(double x) {  // tensor_math.py:307-334
  double x2 = x * x, x4 = x2 * x2;
  double numer = 0.866025403784438713 + 2.12714890259493060 * x +
                 ((1.89202064815951569 + 0.739603278343401613 * x) * x2 +
                  (0.121973926953064794 + x * (0.00655637626263929360 + 0.0000390884982780803443 * x)) * x4);
  double denom = 1.0 + 2.26376989330935617 * x +
                 ((1.80461009751278976 + 0.603976798217196003 * x) * x2 +
                  (0.0783255761115461708 + 0.00268525944538021629 * x) * x4);
  return numer / denom;
}
__device__ inline double sgn(double x) { return (x > 0.0) - (x < 0.0); }

// A: symmetric 3x3 (row-major). lam[3] ascending, V column k = eigenvector k (V[3*i+k]), NOT normalised.
__device__ inline void eigen_sym33_non_unit(const double* A, double* lam, double* V) {
  double cxx = A[0], cyy = A[4], czz = A[8];
  const double cxy = 0.5 * (A[1] + A[3]), cyz = 0.5 * (A[5] + A[7]), czx = 0.5 * (A[6] + A[2]);
  const double c1 = (cxx + cyy + czz) / 3.0;
  cxx -= c1; cyy -= c1; czz -= c1;
  const double cxy2 = cxy * cxy, cyz2 = cyz * cyz, czx2 = czx * czx, cxxcyy = cxx * cyy;
  const double c2 = cxxcyy + cyy * czz + czz * cxx - cxy2 - cyz2 - czx2;
  const bool neg = c2 < 0.0;
  const double threeOverA = neg ? -3.0 / c2 : 1.0;
  const double sq = neg ? sqrt(threeOverA) : 1.0;
  const double c3 = cxx * cyz2 + cyy * czx2 - 2.0 * cxy * cyz * czx + czz * (cxy2 - cxxcyy);
  const double rr = -0.5 * c3 * threeOverA * sq;
  const double arg = fmin(fabs(rr), 1.0);
  double eval2 = neg ? (2.0 * cos_acos_third(arg) * sgn(rr)) / sq : 1.0;

  double r0[3] = {cxx - eval2, cxy, czx};
  double r1[3] = {cxy, cyy - eval2, cyz};
  double r2[3] = {czx, cyz, czz - eval2};
  const double k0 = r0[0] * r0[0] + cxy2 + czx2;
  const double k1 = cxy2 + r1[1] * r1[1] + cyz2;
  const double k2 = czx2 + cyz2 + r2[2] * r2[2];
  const bool k0gk1 = k1 <= k0, k0gk2 = k2 <= k0, k1gk2 = k2 <= k1;
  const bool k0L = k0gk1 && k0gk2;
  const bool k1L = k1gk2 && !k0gk1;
  const bool k2L = !(k0L || k1L);
  double kr[3], row2[3], row3[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    kr[i] = (k0L ? r0[i] : 0.0) + (k1L ? r1[i] : 0.0) + (k2L ? r2[i] : 0.0);
    row2[i] = k0L ? r1[i] : r0[i];
    row3[i] = k2L ? r1[i] : r2[i];
  }
  const double kiki = 1.0 / ((k0L ? k0 : 0.0) + (k1L ? k1 : 0.0) + (k2L ? k2 : 0.0));
  const double d1 = kiki * (kr[0] * row2[0] + kr[1] * row2[1] + kr[2] * row2[2]);
  const double d2 = kiki * (kr[0] * row3[0] + kr[1] * row3[1] + kr[2] * row3[2]);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    row2[i] -= d1 * kr[i];
    row3[i] -= d2 * kr[i];
  }
  const double a0 = row2[0] * row2[0] + row2[1] * row2[1] + row2[2] * row2[2];
  const double a1 = row3[0] * row3[0] + row3[1] * row3[1] + row3[2] * row3[2];
  const bool a0le = a0 <= a1;
  double ar[3];
#pragma unroll
  for (int i = 0; i < 3; i++) ar[i] = a0le ? row3[i] : row2[i];
  const double aiai = 1.0 / (a0le ? a1 : a0);

  double e2[3] = {kr[1] * ar[2] - kr[2] * ar[1], kr[2] * ar[0] - kr[0] * ar[2], kr[0] * ar[1] - kr[1] * ar[0]};

  const double k11 = cxx * kr[0] + cxy * kr[1] + czx * kr[2];
  const double k21 = cxy * kr[0] + cyy * kr[1] + cyz * kr[2];
  const double k31 = czx * kr[0] + cyz * kr[1] + czz * kr[2];
  const double a12 = cxx * ar[0] + cxy * ar[1] + czx * ar[2];
  const double a22 = cxy * ar[0] + cyy * ar[1] + cyz * ar[2];
  const double a32 = czx * ar[0] + cyz * ar[1] + czz * ar[2];
  double rxx = (kr[0] * k11 + kr[1] * k21 + kr[2] * k31) * kiki;
  const double kaxy = kr[0] * a12 + kr[1] * a22 + kr[2] * a32;
  double ryy = (ar[0] * a12 + ar[1] * a22 + ar[2] * a32) * aiai;
  const double rxy2 = kaxy * kaxy * aiai * kiki;

  const double b = 0.5 * (rxx - ryy);
  const double sqrtTerm = sqrt(b * b + rxy2) * sgn(b);
  double eval0 = ryy + b - sqrtTerm;
  double eval1 = rxx + ryy - eval0;
  rxx -= eval0;
  ryy -= eval0;
  const double rxx2 = rxx * rxx, ryy2 = ryy * ryy;
  const bool lt = rxx2 < ryy2;
  const double fac1 = lt ? kaxy * aiai : rxx;
  const double fac2 = lt ? ryy : kiki * kaxy;
  double e0[3];
  const bool both_zero = (rxx2 == 0.0) && (rxy2 == 0.0);
#pragma unroll
  for (int i = 0; i < 3; i++) e0[i] = both_zero ? ar[i] : fac1 * ar[i] - fac2 * kr[i];
  double e1[3] = {e2[1] * e0[2] - e2[2] * e0[1], e2[2] * e0[0] - e2[0] * e0[2], e2[0] * e0[1] - e2[1] * e0[0]};

  eval0 += c1; eval1 += c1; eval2 += c1;
  const bool ok = c2 < (c1 * c1) * (-1.0e-30);
  if (!ok) {
    eval0 = eval1 = eval2 = c1;
    e0[0] = 1.0; e0[1] = 0.0; e0[2] = 0.0;
    e1[0] = 0.0; e1[1] = 1.0; e1[2] = 0.0;
    e2[0] = 0.0; e2[1] = 0.0; e2[2] = 1.0;
  }
  // stable ascending sort of (eval0, eval1, eval2) carrying the vectors (argsort, tensor_math.py:276)
  // (compare-exchanges on named registers: indexing the vectors through a sorted index array would put them in local memory)
  auto cswap = [](double& la, double* va, double& lb, double* vb) {
    const bool sw = lb < la;   // strict: equal eigenvalues keep their order (stable)
    const double tl = sw ? lb : la; lb = sw ? la : lb; la = tl;
#pragma unroll
    for (int i = 0; i < 3; i++) { const double tv = sw ? vb[i] : va[i]; vb[i] = sw ? va[i] : vb[i]; va[i] = tv; }
  };
  cswap(eval0, e0, eval1, e1);
  cswap(eval1, e1, eval2, e2);
  cswap(eval0, e0, eval1, e1);
  lam[0] = eval0; lam[1] = eval1; lam[2] = eval2;
#pragma unroll
  for (int i = 0; i < 3; i++) { V[3 * i] = e0[i]; V[3 * i + 1] = e1[i]; V[3 * i + 2] = e2[i]; }
}

__device__ inline void eigen_sym33_unit(const double* A, double* lam, double* V) {  // tensor_math.py:281-295
  double cmax = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) cmax = fmax(cmax, fabs(A[3 * i]) + fabs(A[3 * i + 1]) + fabs(A[3 * i + 2]));
  const double inv = cmax > 0.0 ? 1.0 / cmax : 1.0;
  double S[9];
#pragma unroll
  for (int i = 0; i < 9; i++) S[i] = inv * A[i];
  eigen_sym33_non_unit(S, lam, V);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double rn = 1.0 / sqrt(V[k] * V[k] + V[3 + k] * V[3 + k] + V[6 + k] * V[6 + k]);
    V[k] *= rn; V[3 + k] *= rn; V[6 + k] *= rn;
    lam[k] *= cmax;
  }
}

// W = 0.5 K tr(E)^2 + G E:E with E = 0.5 log(F^T F) (mtk_log_sqrt, tensor_math.py:337-340).
// P = 2 F S, S = L(Ebar), Ebar = K tr(E) I + 2 G E, L = the (self-adjoint) derivative rule of
// mtk_log_sqrt (tensor_math.py:343-424), evaluated in the eigenbasis.  The tangent action is the Dual overload below.
__device__ inline void hencky(const double* F, const Props& pr, double& W, double* P, double* dWdp, bool want_dp) {
  const double K = pr.p[0], G = pr.p[1];
  double C[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
  double lam[3], V[9];
  eigen_sym33_unit(C, lam, V);
  // E = V diag(e) V^T, e_i = 0.5 log(lam_i): tr E = sum e_i, E:E = sum e_i^2 (V orthonormal)
  const double e[3] = {0.5 * log(lam[0]), 0.5 * log(lam[1]), 0.5 * log(lam[2])};
  const double trE = e[0] + e[1] + e[2];
  const double EE = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  W = 0.5 * K * trE * trE + G * EE;
  // S = L(Ebar) with Ebar = K tr(E) I + 2 G E coaxial with C: in the eigenbasis hHat = 0.5 V^T Ebar V is diagonal,
  // so of the rule's six coefficients (tensor_math.py:369-383) only l_iiii = hHat_ii / lam_i survive -- the
  // off-diagonal ones multiply entries that vanish identically -- and S = V diag(w) V^T, w_i = (K trE + 2 G e_i) / (2 lam_i)
  double FV[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++)
      FV[3 * r + k] = (F[3 * r] * V[k] + F[3 * r + 1] * V[3 + k] + F[3 * r + 2] * V[6 + k]) * ((K * trE + 2.0 * G * e[k]) / lam[k]);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int j = 0; j < 3; j++) P[3 * r + j] = FV[3 * r] * V[3 * j] + FV[3 * r + 1] * V[3 * j + 1] + FV[3 * r + 2] * V[3 * j + 2];
  if (want_dp) {
    dWdp[0] = 0.5 * trE * trE;
    dWdp[1] = EE;
  }
}
// Tangent of the above: dP = (d^2 W / dF^2) : dF, what the reference obtains by differentiating the custom rule a
// second time (JAX differentiates through eigen_sym33_unit and relative_log_difference; base.py:363-385).
// Differentiating the eigenvector algorithm with dual numbers would be ill-conditioned wherever two stretches
// approach each other, and is NaN at F = I (the reference's own result there: every shipped load ramp at t = 0).
// S = dW/dC is an isotropic tensor function of C, S = sum_i w_i(mu) n_i n_i with mu_i the eigenvalues of C,
//   w_i = (K trE + 2 G e_i) / (2 mu_i),  e_i = 0.5 ln mu_i,
// so its derivative has the closed form (in the eigenbasis, D = V^T dC V):
//   dS_ii = sum_k (dw_i/dmu_k) D_kk,   dw_i/dmu_k = K / (4 mu_i mu_k) + delta_ik (G - K trE - 2 G e_i) / (2 mu_i^2)
//   dS_ij = [(w_i - w_j) / (mu_i - mu_j)] D_ij
//         = [-K trE + G (g(x) - 2 e_j)] / (2 mu_i mu_j) D_ij,  x = (mu_i - mu_j) / mu_j,  g(x) = log1p(x) / x
// -- the divided difference written so that nothing cancels as mu_i -> mu_j (g -> 1 reproduces the diagonal
// limit), hence finite and exact at repeated stretches and at F = I.  dP = 2 (dF S + F dS).
__device__ inline void hencky(const Dual* F, const Props& pr, Dual& W, Dual* P, Dual* dWdp, bool want_dp) {
  const double K = pr.p[0], G = pr.p[1];
  double Fv[9], dF[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { Fv[i] = F[i].v; dF[i] = F[i].d; }
  double C[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = Fv[i] * Fv[j] + Fv[3 + i] * Fv[3 + j] + Fv[6 + i] * Fv[6 + j];
  double mu[3], V[9];
  eigen_sym33_unit(C, mu, V);
  const double e[3] = {0.5 * log(mu[0]), 0.5 * log(mu[1]), 0.5 * log(mu[2])};
  const double trE = e[0] + e[1] + e[2];
  const double rmu[3] = {1.0 / mu[0], 1.0 / mu[1], 1.0 / mu[2]};   // one reciprocal per stretch instead of a division per use
  double w[3];
#pragma unroll
  for (int i = 0; i < 3; i++) w[i] = (K * trE + 2.0 * G * e[i]) * (0.5 * rmu[i]);
  // FV = F V, dFV = dF V (columns = eigen directions)
  double FV[9], dFV[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      FV[3 * r + k] = Fv[3 * r] * V[k] + Fv[3 * r + 1] * V[3 + k] + Fv[3 * r + 2] * V[6 + k];
      dFV[3 * r + k] = dF[3 * r] * V[k] + dF[3 * r + 1] * V[3 + k] + dF[3 * r + 2] * V[6 + k];
    }
  // D = V^T (dF^T F + F^T dF) V
  double D[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double sacc = 0.0;
#pragma unroll
      for (int r = 0; r < 3; r++) sacc += dFV[3 * r + i] * FV[3 * r + j] + FV[3 * r + i] * dFV[3 * r + j];
      D[3 * i + j] = sacc;
    }
  // M = dS in the eigenbasis
  double M[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double sacc = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) sacc += K * (0.25 * rmu[i] * rmu[k]) * D[4 * k];
    M[4 * i] = sacc + (G - K * trE - 2.0 * G * e[i]) * (0.5 * rmu[i] * rmu[i]) * D[4 * i];
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i + 1; j < 3; j++) {
      const double x = (mu[i] - mu[j]) * rmu[j];
      const double gx = (x == 0.0) ? 1.0 : log1p(x) / x;
      const double dd = (-K * trE + G * (gx - 2.0 * e[j])) * (0.5 * rmu[i] * rmu[j]);
      M[3 * i + j] = M[3 * j + i] = dd * 0.5 * (D[3 * i + j] + D[3 * j + i]);
    }
  // dP = 2 (dF S + F dS) = 2 (dFV diag(w) + FV M) V^T ;  P = 2 FV diag(w) V^T
  double A1[9], A0[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      A0[3 * r + k] = FV[3 * r + k] * w[k];
      A1[3 * r + k] = dFV[3 * r + k] * w[k] + FV[3 * r] * M[k] + FV[3 * r + 1] * M[3 + k] + FV[3 * r + 2] * M[6 + k];
    }
  double dW = 0.0;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double pv = 2.0 * (A0[3 * r] * V[3 * j] + A0[3 * r + 1] * V[3 * j + 1] + A0[3 * r + 2] * V[3 * j + 2]);
      const double pd = 2.0 * (A1[3 * r] * V[3 * j] + A1[3 * r + 1] * V[3 * j + 1] + A1[3 * r + 2] * V[3 * j + 2]);
      P[3 * r + j] = Dual(pv, pd);
      dW += pv * dF[3 * r + j];
    }
  const double EE = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
  W = Dual(0.5 * K * trE * trE + G * EE, dW);
  if (want_dp) {
    // d(P:dF)/dK and d(P:dF)/dG: P = K P_K + G P_G with S_K = V diag(trE / (2 mu)) V^T, S_G = V diag(e / mu) V^T
    double tK = 0.0, tG = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double q = 0.5 * D[4 * i];   // (F V)_i . (dF V)_i
      tK += trE * rmu[i] * q;
      tG += 2.0 * e[i] * rmu[i] * q;
    }
    dWdp[0] = Dual(0.5 * trE * trE, tK);
    dWdp[1] = Dual(EE, tG);
  }
}

// ------------------------------------------------------------------ Blatz-Ko (blatz_ko.py:6-24)
//   W = (G/2) (I2 / I3 + 2 sqrt(I3) - 5),  I2 = (tr(C)^2 - tr(C^2)) / 2,  I3 = J^2  (mechanics/base.py:65-86)
//   dI2/dF = 2 (I1 F - F C),  d sqrt(I3)/dF = cof F,  dI3/dF = 2 J cof F
//   P = G [ (I1 F - F C) / J^2 + (1 - I2 / J^3) cof F ]
template <class T>
__device__ inline void blatz_ko(const T* F, const Props& pr, T& W, T* P, T* dWdp, bool want_dp) {
  const double G = pr.p[0];
  T cof[9];
  cofactor3(F, cof);
  T J = F[0] * cof[0] + F[1] * cof[1] + F[2] * cof[2];
  T C[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
  T I1 = C[0] + C[4] + C[8];
  T trC2 = C[0] * C[0] + C[4] * C[4] + C[8] * C[8] + 2.0 * (C[1] * C[1] + C[2] * C[2] + C[5] * C[5]);
  T I2 = 0.5 * (I1 * I1 - trC2);
  T Jinv = 1.0 / J;
  T Jinv2 = Jinv * Jinv;
  T w = 0.5 * (I2 * Jinv2 + 2.0 * J - 5.0);
  W = G * w;
  T cF = G * Jinv2 * I1;
  T cFC = -(G * Jinv2);
  T ccof = G * (1.0 - I2 * Jinv2 * Jinv);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      T FC = F[3 * i] * C[j] + F[3 * i + 1] * C[3 + j] + F[3 * i + 2] * C[6 + j];
      P[3 * i + j] = ccof * cof[3 * i + j] + cF * F[3 * i + j] + cFC * FC;
    }
  if (want_dp) dWdp[0] = w;
}

template <int MODEL, class T>
__device__ inline void model_eval(const T* F, const Props& pr, T& W, T* P, T* dWdp, bool want_dp) {
  if constexpr (MODEL == PCX_MODEL_NEOHOOKEAN) neohookean(F, pr, W, P, dWdp, want_dp);
  else if constexpr (MODEL == PCX_MODEL_GENT) gent(F, pr, W, P, dWdp, want_dp);
  else if constexpr (MODEL == PCX_MODEL_SWANSON) swanson(F, pr, W, P, dWdp, want_dp);
  else if constexpr (MODEL == PCX_MODEL_BLATZ_KO) blatz_ko(F, pr, W, P, dWdp, want_dp);
  else hencky(F, pr, W, P, dWdp, want_dp);
}

// ------------------------------------------------------------------ SimpleFeFv (hyperviscoelasticity/simple_fefv.py)
// One Prony branch (G, tau) of the finite-strain viscoelastic model, state Fv (3x3) per quadrature point:
//   Fe = F Fv_old^-1,  E = log_sqrt(Fe^T Fe) (mtk_log_sqrt, tensor_math.py:337-424),
//   dEv = dt / (1 + dt/tau) dev(E) / tau = c dev(E),  Ee = E - dEv,  Dv = dEv / dt       (simple_fefv.py:85-108)
//   psi = G ||dev Ee||^2 + G tau ||dev Dv||^2 = kappa ||dev E||^2,  kappa = G [(1 - c)^2 + tau c^2 / dt^2]
//   Fv_new = expm(dEv) Fv_old
// (c and kappa are computed on the host per step).  Everything is a spectral function of Ce = Fe^T Fe, so ONE
// closed-form eigen decomposition (the reference's own algorithm) serves the energy, the state update and -- for the
// reverse sweep over time that jax derives by differentiating the lax.scan -- the adjoints:
//   psi_bar = 1 (the caller scales by JxW) and lam = adjoint of Fv_new  ->  S += d/dF,  lam_old = adjoint of Fv_old
//   X = c dev E:   Xbar = V (Gamma o sym(V^T (lam Fv_old^T) V)) V^T, Gamma_ij = (e^xi - e^xj) / (xi - xj)   [d expm]
//   Ebar = 2 kappa dev E + c dev(Xbar);   Cbar = V (Lambda o (V^T Ebar V)) V^T,  Lambda_ii = 1 / (2 mu_i),
//          Lambda_ij = (ln mu_i - ln mu_j) / (2 (mu_i - mu_j))                                        [d log_sqrt, self-adjoint]
//   Febar = 2 Fe Cbar;  S += Febar A^T;  Abar = F^T Febar;  lam_old = expm(X) lam - A^T Abar A^T,  A = Fv_old^-1
// with both divided differences written through expm1 / log1p so that nothing cancels at repeated stretches (F = I).

__device__ inline void mat33_mul(const double* A, const double* B, double* C) {      // C = A B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ inline void mat33_mul_tn(const double* A, const double* B, double* C) {   // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ inline void mat33_mul_nt(const double* A, const double* B, double* C) {   // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}

// F, Fv_old: 3x3 row-major.  Fv_new / lam / S / lam_old may be null (forward sweep: Fv_new only; reverse sweep: the
// last three).  Returns psi of the branch.
//
// Second order (residual / reaction losses of a model with state, weak_form_loss_functions.py:151-192,242-276): the
// loss holds v . f with f = dPi/du at FIXED old state, so its gradient needs the derivative of (dpsi/dF, dpsi/dFv_old)
// along dF = sum_a v_a (x) grad N_a -- the dual part of the lam = 0 reverse sweep.  With Cbar = kappa V diag(de_i / mu_i) V^T
// = kappa [g(Ce) - eb Ce^-1], g(mu) = ln mu / (2 mu), eb = ln det Ce / 6, its directional derivative along
// dCe = dFe^T Fe + Fe^T dFe is, in the eigenbasis (H = V^T dCe V),
//   dCbar^_ij = kappa { [log1p(xr)/xr - 2 de_j] / (2 mu_i mu_j) H_ij - delta_ij (sum_k H_kk / (6 mu_k)) / mu_i },  xr = (mu_i - mu_j)/mu_j
// (first divided differences of g merged with the eb Ce^-1 dCe Ce^-1 term; symmetric in i, j; the diagonal is the limit
// xr -> 0) -- no eigenvector derivatives, nothing singular at repeated stretches.  Then dFebar = 2 (dFe Cbar + Fe dCbar),
// dS += dFebar A^T, dlam_old -= A^T (dF^T Febar + F^T dFebar) A^T.  wpsi = cotangent of psi (the real part's scale).
__device__ inline double fefv_branch(const double* F, const double* Fv_old, double kappa, double c, double* Fv_new,
                                     const double* lam, double* S, double* lam_old, double wpsi = 1.0,
                                     const double* dF = nullptr, double* dS = nullptr, double* dlam_old = nullptr) {
  double cofv[9], A[9];
  cofactor3(Fv_old, cofv);
  const double detv = Fv_old[0] * cofv[0] + Fv_old[1] * cofv[1] + Fv_old[2] * cofv[2];
  const double rdetv = 1.0 / detv;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) A[3 * i + j] = cofv[3 * j + i] * rdetv;   // inverse = cof^T / det
  double Fe[9], Ce[9];
  mat33_mul(F, A, Fe);
  mat33_mul_tn(Fe, Fe, Ce);
  double mu[3], V[9];
  eigen_sym33_unit(Ce, mu, V);
  const double e[3] = {0.5 * log(mu[0]), 0.5 * log(mu[1]), 0.5 * log(mu[2])};
  const double eb = (e[0] + e[1] + e[2]) * (1.0 / 3.0);
  const double de[3] = {e[0] - eb, e[1] - eb, e[2] - eb};
  const double psi = kappa * (de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
  const double x[3] = {c * de[0], c * de[1], c * de[2]};
  const double ex[3] = {exp(x[0]), exp(x[1]), exp(x[2])};
  double expX[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) expX[3 * i + j] = V[3 * i] * ex[0] * V[3 * j] + V[3 * i + 1] * ex[1] * V[3 * j + 1] + V[3 * i + 2] * ex[2] * V[3 * j + 2];
  if (Fv_new) mat33_mul(expX, Fv_old, Fv_new);
  if (S) {
    double Eh[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Eh[i] = 0.0;
    Eh[0] = 2.0 * kappa * wpsi * de[0]; Eh[4] = 2.0 * kappa * wpsi * de[1]; Eh[8] = 2.0 * kappa * wpsi * de[2];
    if (lam) {
      double M[9], T[9], Mh[9];
      mat33_mul_nt(lam, Fv_old, M);     // lam Fv_old^T
      mat33_mul_tn(V, M, T);            // V^T M
      mat33_mul(T, V, Mh);              // V^T M V
      // divided differences of exp at the eigenvalues: gam_ij = (ex_i - ex_j) / (x_i - x_j) = ex_j expm1(d) / d, symmetric
      double Xh[9];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Xh[4 * i] = ex[i] * Mh[4 * i];
#pragma unroll
        for (int j = 0; j < i; j++) {
          const double d = x[i] - x[j];
          const double gam = (d == 0.0) ? ex[j] : ex[j] * expm1(d) / d;
          Xh[3 * i + j] = Xh[3 * j + i] = gam * 0.5 * (Mh[3 * i + j] + Mh[3 * j + i]);
        }
      }
      const double trX = (Xh[0] + Xh[4] + Xh[8]) * (1.0 / 3.0);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Eh[3 * i + j] += c * (Xh[3 * i + j] - (i == j ? trX : 0.0));
      mat33_mul(expX, lam, lam_old);    // expm(X)^T lam, expm(X) symmetric
    } else {
#pragma unroll
      for (int i = 0; i < 9; i++) lam_old[i] = 0.0;
    }
    // divided differences of 0.5 log at the eigenvalues (symmetric): L_ij = 0.5 (log1p(xr) / xr) / mu_j, xr = (mu_i - mu_j) / mu_j
    double Ch[9];
    const double rmu[3] = {1.0 / mu[0], 1.0 / mu[1], 1.0 / mu[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Ch[4 * i] = 0.5 * rmu[i] * Eh[4 * i];
#pragma unroll
      for (int j = 0; j < i; j++) {
        const double xr = (mu[i] - mu[j]) * rmu[j];
        const double L = 0.5 * ((xr == 0.0) ? 1.0 : log1p(xr) / xr) * rmu[j];
        Ch[3 * i + j] = L * Eh[3 * i + j];
        Ch[3 * j + i] = L * Eh[3 * j + i];
      }
    }
    double T1[9], Cb[9], Feb[9], T2[9], Ab[9];
    mat33_mul(V, Ch, T1);
    mat33_mul_nt(T1, V, Cb);            // V Ch V^T
    mat33_mul(Fe, Cb, Feb);
#pragma unroll
    for (int i = 0; i < 9; i++) Feb[i] *= 2.0;
    mat33_mul_nt(Feb, A, T2);           // Febar A^T
#pragma unroll
    for (int i = 0; i < 9; i++) S[i] += T2[i];
    mat33_mul_tn(F, Feb, Ab);           // F^T Febar
    mat33_mul_tn(A, Ab, T1);            // A^T Abar
    mat33_mul_nt(T1, A, T2);            // A^T Abar A^T
#pragma unroll
    for (int i = 0; i < 9; i++) lam_old[i] -= T2[i];
    if (dF) {
      double dFe[9], dC[9], H[9], dCh[9], dCb[9], Cb0[9], Feb0[9], dFeb[9];
      mat33_mul(dF, A, dFe);
      mat33_mul_tn(Fe, dFe, T1);        // Fe^T dFe
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) dC[3 * i + j] = T1[3 * i + j] + T1[3 * j + i];
      mat33_mul_tn(V, dC, T1);
      mat33_mul(T1, V, H);              // V^T dCe V
      const double deb = (H[0] * rmu[0] + H[4] * rmu[1] + H[8] * rmu[2]) * (1.0 / 6.0);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        dCh[4 * i] = kappa * rmu[i] * ((0.5 - de[i]) * rmu[i] * H[4 * i] - deb);
#pragma unroll
        for (int j = 0; j < i; j++) {
          const double xr = (mu[i] - mu[j]) * rmu[j];
          const double cf = kappa * (((xr == 0.0) ? 1.0 : log1p(xr) / xr) - 2.0 * de[j]) * 0.5 * rmu[i] * rmu[j];
          dCh[3 * i + j] = cf * H[3 * i + j];
          dCh[3 * j + i] = cf * H[3 * j + i];
        }
      }
      mat33_mul(V, dCh, T1);
      mat33_mul_nt(T1, V, dCb);         // V dCh V^T
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          Cb0[3 * i + j] = kappa * (V[3 * i] * de[0] * rmu[0] * V[3 * j] + V[3 * i + 1] * de[1] * rmu[1] * V[3 * j + 1] +
                                    V[3 * i + 2] * de[2] * rmu[2] * V[3 * j + 2]);
      mat33_mul(Fe, Cb0, Feb0);
      mat33_mul(dFe, Cb0, T1);
      mat33_mul(Fe, dCb, T2);
#pragma unroll
      for (int i = 0; i < 9; i++) { Feb0[i] *= 2.0; dFeb[i] = 2.0 * (T1[i] + T2[i]); }
      mat33_mul_nt(dFeb, A, T1);        // dFebar A^T
#pragma unroll
      for (int i = 0; i < 9; i++) dS[i] += T1[i];
      mat33_mul_tn(dF, Feb0, T1);       // dF^T Febar
      mat33_mul_tn(F, dFeb, T2);        // F^T dFebar
#pragma unroll
      for (int i = 0; i < 9; i++) T1[i] += T2[i];
      mat33_mul_tn(A, T1, T2);
      mat33_mul_nt(T2, A, T1);          // A^T dAbar A^T
#pragma unroll
      for (int i = 0; i < 9; i++) dlam_old[i] -= T1[i];
    }
  }
  return psi;
}

// ------------------------------------------------------------------ PlaneStress (solid_mechanics.py:32-71)
// The out-of-plane displacement gradient x = grad_u_33 is the root of sigma_33(x) = 0, which the reference finds
// per quadrature point with scalar_root_find.find_root (Newton safeguarded by bisection, guess 0.05, bracket
// [-0.99, 10], x_tol 1e-13; math/scalar_root_find.py:18-160).  With F = [[F_2D, 0], [0, 1 + x]],
// sigma_33 = P_33 F_33 / J, so the root is that of P_33 = dW/dx.  Here: the same safeguarded Newton iteration on
// P_33, its derivative from one dual-number evaluation seeded on F_33, started from the incompressible estimate
// F_33 = 1 / det F_2D (2 - 4 iterations instead of the reference's ~8 from 0.05; the root in the bracket is unique,
// P_33 increases with x), iterated to |dx| < 1e-15 or to the rounding floor (tighter than the reference's 1e-13, so both agree to 1e-13).
// F: in-plane entries set, F[8] ignored on entry; on exit F[8] = 1 + x.  Returns false if 60 iterations did not converge.
template <int MODEL>
__device__ inline bool plane_stress_solve(double* F, const Props& pr) {
  const double j2 = F[0] * F[4] - F[1] * F[3];
  double lo = -0.99, hi = 10.0;
  double x = 1.0 / j2 - 1.0;
  if constexpr (MODEL == PCX_MODEL_HENCKY) {
    // closed form: F = diag(F_2D, l3) has tr E = ln det F_2D + ln l3, and P_33 = (K tr E + 2 G ln l3) / l3 = 0 gives
    // l3 = (det F_2D)^(-K / (K + 2 G)); the iteration below then only confirms it
    x = exp(-pr.p[0] / (pr.p[0] + 2.0 * pr.p[1]) * log(j2)) - 1.0;
  }
  x = fmin(fmax(x, lo), hi);
  if (!(x == x)) x = 0.05;
  bool ok = false;
  double dx_prev = 1.0e300;
#pragma unroll 1
  for (int it = 0; it < 60; it++) {
    Dual Fd[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Fd[i] = Dual(F[i], 0.0);
    Fd[8] = Dual(1.0 + x, 1.0);
    Dual W, P[9], dp[PCX_MAX_PROPS];
    model_eval<MODEL, Dual>(Fd, pr, W, P, dp, false);
    const double g = P[8].v, dg = P[8].d;
    if (g < 0.0) lo = x; else hi = x;
    double xn = x - g / dg;
    // a Newton step that lands ON a bracket end (an earlier iterate sat on the root to rounding and became that end) is
    // the root, not an escape: taking it ends the iteration, bisecting instead would halve the bracket ~50 more times
    const double slack = 8.0 * 2.220446049250313e-16 * fmax(1.0, fabs(x));
    if (dg > 0.0 && xn <= lo && xn >= lo - slack) xn = lo;
    else if (dg > 0.0 && xn >= hi && xn <= hi + slack) xn = hi;
    else if (!(xn > lo) || !(xn < hi) || !(dg > 0.0)) xn = 0.5 * (lo + hi);   // bisection step (rtsafe)
    const double dx = fabs(xn - x);
    const bool newton = xn != 0.5 * (lo + hi);
    x = xn;
    // converged: step below 1e-15, or below 1e-13 and no longer contracting (a bisection step halves it) -- the quadratic phase has hit the rounding
    // floor of P_33 (a few 1e-16 through the eigen decomposition of Hencky), where steps only jitter; or (r2) an accepted
    // Newton step below 3e-8, whose remaining error (g''/2g') dx^2 is below 1e-15: the confirming evaluation is not run
    if ((newton && dx < 3.0e-8) || dx < 1.0e-15 || (dx < 1.0e-13 && dx > 0.75 * dx_prev)) { ok = true; break; }
    dx_prev = dx;
  }
  F[8] = 1.0 + x;
  return ok;
}

// Tangent of the plane-stress response along an in-plane direction dF_2D (Fd[i].d set in-plane, Fd[8].d ignored):
// the implicit function theorem on P_33(F_2D, x) = 0 gives dx = -(dP_33 | dx = 0) / (dP_33/dx), and every output is
// linear in the direction, so two dual evaluations (in-plane seed, unit seed on F_33) combine to the total tangent
//   dP = dP|_(dF_2D, 0) + dx dP|_(0, 1),   d(dW/dp) likewise
// (what jax.lax.custom_root's tangent rule produces, math/scalar_root_find.py:45-50).  Fd[8].v must hold the root.
template <int MODEL>
__device__ inline void plane_stress_tangent(Dual* Fd, const Props& pr, Dual& W, Dual* P, Dual* dWdp, bool want_dp, int nprops) {
  Fd[8].d = 0.0;
  model_eval<MODEL, Dual>(Fd, pr, W, P, dWdp, want_dp);
  Dual Fe[9], We, Pe[9], dpe[PCX_MAX_PROPS];
#pragma unroll
  for (int i = 0; i < 9; i++) Fe[i] = Dual(Fd[i].v, 0.0);
  Fe[8].d = 1.0;
  model_eval<MODEL, Dual>(Fe, pr, We, Pe, dpe, want_dp);
  const double dx = -P[8].d / Pe[8].d;
#pragma unroll
  for (int i = 0; i < 9; i++) P[i].d = fma(dx, Pe[i].d, P[i].d);
  W.d = fma(dx, We.d, W.d);
  if (want_dp) {
#pragma unroll
    for (int k = 0; k < PCX_MAX_PROPS; k++)
      if (k < nprops) dWdp[k].d = fma(dx, dpe[k].d, dWdp[k].d);
  }
  Fd[8].d = dx;
}

// ------------------------------------------------------------------ PlaneStress with a model with state (SimpleFeFv)
// The reference's root find runs on the Cauchy stress of the WHOLE model at fixed old state (solid_mechanics.py:49-71 with
// mechanics/base.py:10-17,112-118: cauchy_stress = jacfwd(energy)[0] F^T / J, energy = psi_eq + sum_b psi_b), so here
//   g(x)  = P_33 = dpsi_eq/dF_33 + sum_b S_b,33          (S_b: the lam = 0 reverse branch of fefv_branch)
//   g'(x) = its tangent along dF = E_33                    (dual evaluation of the equilibrium model + fefv_branch's closed-form tangent)
// and the same safeguarded Newton iteration as plane_stress_solve.  F: in-plane entries set; on exit F[8] = 1 + x.
// zo: the old viscous deformation gradients [n_prony][9].
template <int MODEL>
__device__ inline void plane_stress_visco_g(const double* F, const Props& pr, int n_prony, const double* zo, const double* kappa,
                                            const double* cc, double& g, double& dg) {
  Dual Fd[9], W, P[9], dp[PCX_MAX_PROPS];
#pragma unroll
  for (int i = 0; i < 9; i++) Fd[i] = Dual(F[i], i == 8 ? 1.0 : 0.0);
  model_eval<MODEL, Dual>(Fd, pr, W, P, dp, false);
  g = P[8].v;
  dg = P[8].d;
  const double dF[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0};
  for (int b = 0; b < n_prony; b++) {
    double S[9], lo[9], dS[9], dlo[9];
#pragma unroll
    for (int i = 0; i < 9; i++) S[i] = dS[i] = dlo[i] = 0.0;
    fefv_branch(F, zo + 9 * b, kappa[b], cc[b], nullptr, nullptr, S, lo, 1.0, dF, dS, dlo);
    g += S[8];
    dg += dS[8];
  }
}

template <int MODEL>
__device__ inline bool plane_stress_visco_solve(double* F, const Props& pr, int n_prony, const double* zo, const double* kappa,
                                                const double* cc, double x_start /* previous step's root, or NaN */) {
  const double j2 = F[0] * F[4] - F[1] * F[3];
  double lo = -0.99, hi = 10.0;
  // start: the root of the previous time step when there is one (the out-of-plane stretch moves little between load
  // steps: 2 iterations instead of 4-5), else the incompressible estimate
  double x = fmin(fmax((x_start == x_start) ? x_start : 1.0 / j2 - 1.0, lo), hi);
  if (!(x == x)) x = 0.05;
  bool ok = false;
  double dx_prev = 1.0e300;
#pragma unroll 1
  for (int it = 0; it < 60; it++) {
    F[8] = 1.0 + x;
    double g, dg;
    plane_stress_visco_g<MODEL>(F, pr, n_prony, zo, kappa, cc, g, dg);
    if (g < 0.0) lo = x; else hi = x;
    double xn = x - g / dg;
    const double slack = 8.0 * 2.220446049250313e-16 * fmax(1.0, fabs(x));
    if (dg > 0.0 && xn <= lo && xn >= lo - slack) xn = lo;
    else if (dg > 0.0 && xn >= hi && xn <= hi + slack) xn = hi;
    else if (!(xn > lo) || !(xn < hi) || !(dg > 0.0)) xn = 0.5 * (lo + hi);
    const double dx = fabs(xn - x);
    const bool newton = xn != 0.5 * (lo + hi);
    x = xn;
    // an accepted Newton step below 3e-8 leaves an error ~ (g''/2g') dx^2 < 1e-15: every evaluation costs a dual pass through
    // the equilibrium model and the eigen decomposition + tangent of every branch, so the confirming iteration is not run
    if ((newton && dx < 3.0e-8) || dx < 1.0e-15 || (dx < 1.0e-13 && dx > 0.75 * dx_prev)) { ok = true; break; }
    dx_prev = dx;
  }
  F[8] = 1.0 + x;
  return ok;
}

}  // namespace pcx
