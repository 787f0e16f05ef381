"""Oracle (test infrastructure): constitutive energies, torch float64 on the CPU.

All functions are batched over leading dimensions: ``grad_u`` has shape (..., 3, 3).
torch.autograd stands in for jax.jacfwd / jax.grad (the reference obtains the PK1
stress as ``jax.jacfwd(energy)``: pancax/constitutive_models/mechanics/base.py:112-118).
"""
from __future__ import annotations

import torch

F64 = torch.float64


# ----------------------------------------------------------------------------- kinematics
def deformation_gradient(grad_u):
    """F = grad_u + I.  mechanics/base.py:19-21."""
    return grad_u + torch.eye(3, dtype=grad_u.dtype)


def jacobian(grad_u):
    """J = det F (guard rails are commented out in the reference).  base.py:86-105."""
    return torch.linalg.det(deformation_gradient(grad_u))


def I1_bar(grad_u):
    """J^(-2/3) tr(F F^T).  base.py:50-63."""
    F = deformation_gradient(grad_u)
    I1 = torch.einsum("...ij,...ij->...", F, F)  # trace(F @ F.T)
    J = jacobian(grad_u)
    return torch.pow(J, -2.0 / 3.0) * I1


def I2_bar(grad_u):
    """J^(-4/3) * 0.5 (I1^2 - tr C^2).  base.py:73-80."""
    F = deformation_gradient(grad_u)
    C = F.transpose(-1, -2) @ F
    C2 = C @ C
    I1 = torch.einsum("...ii->...", C)
    I2 = 0.5 * (I1 ** 2 - torch.einsum("...ii->...", C2))
    J = jacobian(grad_u)
    return torch.pow(J, -4.0 / 3.0) * I2


def tensor_2D_to_3D(H):
    """Embed (..., r, c) in zeros (..., 3, 3).  pancax/math/tensor_math.py:64-65."""
    out = torch.zeros(H.shape[:-2] + (3, 3), dtype=H.dtype)
    out[..., : H.shape[-2], : H.shape[-1]] = H
    return out


# ----------------------------------------------------------------------------- properties
def bounded_property(prop_min, prop_max, prop_val):
    """p = p_min + (p_max - p_min) * sigmoid(prop_val[0]).
    pancax/constitutive_models/properties.py:33-38."""
    return prop_min + (prop_max - prop_min) * torch.sigmoid(prop_val)


# ----------------------------------------------------------------------------- models
def neohookean_energy(grad_u, K, G):
    """hyperelasticity/neohookean.py:20-34."""
    J = jacobian(grad_u)
    I_1_bar = I1_bar(grad_u)
    W_vol = 0.5 * K * (0.5 * (J ** 2 - 1) - torch.log(J))
    W_dev = 0.5 * G * (I_1_bar - 3.0)
    return W_vol + W_dev


def gent_energy(grad_u, K, G, Jm):
    """hyperelasticity/gent.py:22-39.  lax.cond replaces I1_bar by the constant
    Jm + 3 - 0.001 in the guarded branch (zero F-gradient there, Jm-gradient kept)."""
    J = jacobian(grad_u)
    I_1_bar = I1_bar(grad_u)
    check_value = I_1_bar > Jm + 3.0 - 0.001
    I_1_bar = torch.where(check_value, Jm + 3.0 - 0.001 + 0.0 * I_1_bar.detach(), I_1_bar)
    W_vol = 0.5 * K * (0.5 * (J ** 2 - 1) - torch.log(J))
    W_dev = -0.5 * G * Jm * torch.log(1.0 - ((I_1_bar - 3.0) / Jm))
    return W_vol + W_dev


def swanson_energy(grad_u, K, A1, P1, B1, Q1, C1, R1, cutoff_strain):
    """hyperelasticity/swanson.py:28-65."""
    tau_cutoff = (1.0 / 3.0) * (3.0 + cutoff_strain ** 2) - 1.0
    J = jacobian(grad_u)
    I_1_bar = I1_bar(grad_u)
    I_2_bar = I2_bar(grad_u)
    tau_1 = (1.0 / 3.0) * I_1_bar - 1.0
    tau_2 = (1.0 / 3.0) * I_2_bar - 1.0
    tau_tilde_1 = tau_1 + tau_cutoff
    tau_tilde_2 = tau_2 + tau_cutoff
    W_vol = K * (J * torch.log(J) - J + 1.0)
    W_dev_tau = (3.0 / 2.0 * (
        A1 / (P1 + 1.0) * (tau_tilde_1 ** (P1 + 1.0))
        + B1 / (Q1 + 1.0) * (tau_tilde_2 ** (Q1 + 1.0))
        + C1 / (R1 + 1.0) * (tau_tilde_1 ** (R1 + 1.0))))
    W_dev_cutoff = (3.0 / 2.0 * (
        A1 / (P1 + 1.0) * (tau_cutoff ** (P1 + 1.0))
        + B1 / (Q1 + 1.0) * (tau_cutoff ** (Q1 + 1.0))
        + C1 / (R1 + 1.0) * (tau_cutoff ** (R1 + 1.0))))
    return W_vol + (W_dev_tau - W_dev_cutoff)


# ---- Hencky: closed-form symmetric 3x3 eigen solver + log-sqrt with the reference's JVP
def cos_of_acos_divided_by_3(x):
    """Pade approximant of cos(acos(x)/3).  tensor_math.py:307-334."""
    x2 = x * x
    x4 = x2 * x2
    numer = (0.866025403784438713 + 2.12714890259493060 * x
             + ((1.89202064815951569 + 0.739603278343401613 * x) * x2
                + (0.121973926953064794
                   + x * (0.00655637626263929360 + 0.0000390884982780803443 * x)) * x4))
    denom = (1.0 + 2.26376989330935617 * x
             + ((1.80461009751278976 + 0.603976798217196003 * x) * x2
                + (0.0783255761115461708 + 0.00268525944538021629 * x) * x4))
    return numer / denom


def _w(c, a, b):
    return torch.where(c, a, b)


def eigen_sym33_non_unit(T):
    """tensor_math.py:76-278, batched.  Returns (evals (...,3) ascending, evecs (...,3,3)
    with eigenvectors in columns, not normalised)."""
    cxx = T[..., 0, 0]
    cyy = T[..., 1, 1]
    czz = T[..., 2, 2]
    cxy = 0.5 * (T[..., 0, 1] + T[..., 1, 0])
    cyz = 0.5 * (T[..., 1, 2] + T[..., 2, 1])
    czx = 0.5 * (T[..., 2, 0] + T[..., 0, 2])

    c1 = (cxx + cyy + czz) / 3.0
    cxx = cxx - c1
    cyy = cyy - c1
    czz = czz - c1

    cxy_cxy = cxy * cxy
    cyz_cyz = cyz * cyz
    czx_czx = czx * czx
    cxx_cyy = cxx * cyy

    c2 = cxx_cyy + cyy * czz + czz * cxx - cxy_cxy - cyz_cyz - czx_czx
    one = torch.ones_like(c2)
    zero = torch.zeros_like(c2)

    c2Negative = c2 < 0
    denom = _w(c2Negative, c2, one)
    ThreeOverA = _w(c2Negative, -3.0 / denom, one)
    sqrtThreeOverA = _w(c2Negative, torch.sqrt(ThreeOverA), one)

    c3 = (cxx * cyz_cyz + cyy * czx_czx - 2.0 * cxy * cyz * czx
          + czz * (cxy_cxy - cxx_cyy))
    rr = -0.5 * c3 * ThreeOverA * sqrtThreeOverA
    arg = torch.minimum(torch.abs(rr), one)
    cos_thd3 = cos_of_acos_divided_by_3(arg)
    two_cos_thd3 = 2.0 * cos_thd3 * torch.sign(rr)
    eval2 = _w(c2Negative, two_cos_thd3 / sqrtThreeOverA, one)

    crow0 = torch.stack([cxx - eval2, cxy, czx], -1)
    crow1 = torch.stack([cxy, cyy - eval2, cyz], -1)
    crow2 = torch.stack([czx, cyz, czz - eval2], -1)

    k0 = crow0[..., 0] * crow0[..., 0] + cxy_cxy + czx_czx
    k1 = cxy_cxy + crow1[..., 1] * crow1[..., 1] + cyz_cyz
    k2 = czx_czx + cyz_cyz + crow2[..., 2] * crow2[..., 2]

    k0gk1 = k1 <= k0
    k0gk2 = k2 <= k0
    k1gk2 = k2 <= k1
    k0_largest = k0gk1 & k0gk2
    k1_largest = k1gk2 & (~k0gk1)
    k2_largest = ~(k0_largest | k1_largest)

    def sel3(c0, c1_, c2_, v0, v1, v2):
        return (_w(c0, v0, torch.zeros_like(v0)) + _w(c1_, v1, torch.zeros_like(v1))
                + _w(c2_, v2, torch.zeros_like(v2)))

    u = lambda c: c.unsqueeze(-1)  # noqa: E731
    k_row1 = sel3(u(k0_largest), u(k1_largest), u(k2_largest), crow0, crow1, crow2)
    row2 = _w(u(k0_largest), crow1, crow0)
    row3 = _w(u(k2_largest), crow1, crow2)

    ki_ki = 1.0 / sel3(k0_largest, k1_largest, k2_largest, k0, k1, k2)
    ki_dpr1 = ki_ki * (k_row1 * row2).sum(-1)
    ki_dpr2 = ki_ki * (k_row1 * row3).sum(-1)
    row2 = row2 - u(ki_dpr1) * k_row1
    row3 = row3 - u(ki_dpr2) * k_row1

    a0 = (row2 * row2).sum(-1)
    a1 = (row3 * row3).sum(-1)
    a0lea1 = a0 <= a1
    a_row2 = _w(u(a0lea1), row3, row2)
    ai_ai = 1.0 / _w(a0lea1, a1, a0)

    evec2 = torch.stack([
        k_row1[..., 1] * a_row2[..., 2] - k_row1[..., 2] * a_row2[..., 1],
        k_row1[..., 2] * a_row2[..., 0] - k_row1[..., 0] * a_row2[..., 2],
        k_row1[..., 0] * a_row2[..., 1] - k_row1[..., 1] * a_row2[..., 0]], -1)

    k_atr11 = cxx * k_row1[..., 0] + cxy * k_row1[..., 1] + czx * k_row1[..., 2]
    k_atr21 = cxy * k_row1[..., 0] + cyy * k_row1[..., 1] + cyz * k_row1[..., 2]
    k_atr31 = czx * k_row1[..., 0] + cyz * k_row1[..., 1] + czz * k_row1[..., 2]
    a_atr12 = cxx * a_row2[..., 0] + cxy * a_row2[..., 1] + czx * a_row2[..., 2]
    a_atr22 = cxy * a_row2[..., 0] + cyy * a_row2[..., 1] + cyz * a_row2[..., 2]
    a_atr32 = czx * a_row2[..., 0] + cyz * a_row2[..., 1] + czz * a_row2[..., 2]

    rm2xx = (k_row1[..., 0] * k_atr11 + k_row1[..., 1] * k_atr21
             + k_row1[..., 2] * k_atr31) * ki_ki
    k_a_rm2xy = (k_row1[..., 0] * a_atr12 + k_row1[..., 1] * a_atr22
                 + k_row1[..., 2] * a_atr32)
    rm2yy = (a_row2[..., 0] * a_atr12 + a_row2[..., 1] * a_atr22
             + a_row2[..., 2] * a_atr32) * ai_ai
    rm2xy_rm2xy = k_a_rm2xy * k_a_rm2xy * ai_ai * ki_ki

    # Wilkinson shift (safe_sqrt == sqrt in value: pancax/math/math.py:14-25)
    b = 0.5 * (rm2xx - rm2yy)
    sqrtTerm = torch.sqrt(b * b + rm2xy_rm2xy) * torch.sign(b)
    eval0 = rm2yy + b - sqrtTerm
    eval1 = rm2xx + rm2yy - eval0
    rm2xx = rm2xx - eval0
    rm2yy = rm2yy - eval0
    rm2xx2 = rm2xx * rm2xx
    rm2yy2 = rm2yy * rm2yy
    lt = rm2xx2 < rm2yy2
    fac1 = _w(lt, k_a_rm2xy * ai_ai, rm2xx)
    fac2 = _w(lt, rm2yy, ki_ki * k_a_rm2xy)
    evec0 = u(fac1) * a_row2 - u(fac2) * k_row1
    both_zero = (rm2xx2 == 0.0) & (rm2xy_rm2xy == 0.0)
    evec0 = _w(u(both_zero), a_row2, evec0)

    evec1 = torch.stack([
        evec2[..., 1] * evec0[..., 2] - evec2[..., 2] * evec0[..., 1],
        evec2[..., 2] * evec0[..., 0] - evec2[..., 0] * evec0[..., 2],
        evec2[..., 0] * evec0[..., 1] - evec2[..., 1] * evec0[..., 0]], -1)

    eval0 = eval0 + c1
    eval1 = eval1 + c1
    eval2 = eval2 + c1

    c2tol = (c1 * c1) * (-1.0e-30)
    ok = c2 < c2tol
    eval0 = _w(ok, eval0, c1)
    eval1 = _w(ok, eval1, c1)
    eval2 = _w(ok, eval2, c1)
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=T.dtype).expand_as(evec0)
    ey = torch.tensor([0.0, 1.0, 0.0], dtype=T.dtype).expand_as(evec0)
    ez = torch.tensor([0.0, 0.0, 1.0], dtype=T.dtype).expand_as(evec0)
    evec0 = _w(u(ok), evec0, ex)
    evec1 = _w(u(ok), evec1, ey)
    evec2 = _w(u(ok), evec2, ez)

    evals = torch.stack([eval0, eval1, eval2], -1)
    evecs = torch.stack([evec0, evec1, evec2], -1)  # columns
    idx = torch.argsort(evals, dim=-1, stable=True)
    evals = torch.gather(evals, -1, idx)
    evecs = torch.gather(evecs, -1, idx.unsqueeze(-2).expand_as(evecs))
    del zero
    return evals, evecs


def eigen_sym33_unit(T):
    """tensor_math.py:281-295 (inf-norm scaling, unit eigenvectors)."""
    cmax = T.abs().sum(-1).max(-1).values  # np.linalg.norm(ord=inf) = max row sum
    cmaxInv = torch.where(cmax > 0.0, 1.0 / cmax, torch.ones_like(cmax))
    evals, evecs = eigen_sym33_non_unit(cmaxInv[..., None, None] * T)
    evecs = evecs / torch.linalg.norm(evecs, dim=-2, keepdim=True)
    return cmax.unsqueeze(-1) * evals, evecs


def relative_log_difference(lam1, lam2):
    """tensor_math.py:517-551 (5 % switch to a 9th-order Taylor series)."""
    haveLargeDiff = torch.abs(lam1 - lam2) > 0.05 * torch.minimum(lam1, lam2)
    lamFake = torch.where(haveLargeDiff, lam2, 2.0 * lam2)
    no_tol = torch.log(lam1 / lamFake) / (lam1 - lamFake)
    frac = (lam1 - lam2) / (lam1 + lam2)
    frac2 = frac * frac
    frac4 = frac2 * frac2
    taylor = (2.0 + (2.0 / 3.0) * frac2 + (2.0 / 5.0) * frac4
              + (2.0 / 7.0) * frac4 * frac2 + (2.0 / 9.0) * frac4 * frac4) / (lam1 + lam2)
    return torch.where(haveLargeDiff, no_tol, taylor)


def mtk_log_sqrt_value(A):
    """V diag(0.5 log lam) V^T.  tensor_math.py:337-340."""
    lam, V = eigen_sym33_unit(A)
    return (V * (0.5 * torch.log(lam)).unsqueeze(-2)) @ V.transpose(-1, -2)


def mtk_log_sqrt_jvp(C, H):
    """Tangent of mtk_log_sqrt at C in direction H, the reference's custom JVP rule.
    tensor_math.py:343-424 (written with V instead of the expanded t00..t22 sums)."""
    lam, V = eigen_sym33_unit(C)
    hHat = 0.5 * (V.transpose(-1, -2) @ H @ V)
    l1, l2, l3 = lam[..., 0], lam[..., 1], lam[..., 2]
    L = torch.zeros_like(hHat)
    L[..., 0, 0] = hHat[..., 0, 0] / l1
    L[..., 1, 1] = hHat[..., 1, 1] / l2
    L[..., 2, 2] = hHat[..., 2, 2] / l3
    l1212 = 0.5 * (hHat[..., 0, 1] + hHat[..., 1, 0]) * relative_log_difference(l1, l2)
    l2323 = 0.5 * (hHat[..., 1, 2] + hHat[..., 2, 1]) * relative_log_difference(l2, l3)
    l3131 = 0.5 * (hHat[..., 2, 0] + hHat[..., 0, 2]) * relative_log_difference(l3, l1)
    L[..., 0, 1] = l1212
    L[..., 1, 0] = l1212
    L[..., 1, 2] = l2323
    L[..., 2, 1] = l2323
    L[..., 2, 0] = l3131
    L[..., 0, 2] = l3131
    return V @ L @ V.transpose(-1, -2)


class _MtkLogSqrt(torch.autograd.Function):
    """mtk_log_sqrt with the reference's derivative rule.  The JVP map H -> sol is
    self-adjoint for symmetric arguments (it is V (G o (V^T H V)) V^T with a symmetric
    coefficient table G), so the VJP applied by reverse-mode AD to a cotangent Ebar is
    the same map applied to sym(Ebar) -- this is what JAX derives by transposing the
    custom JVP rule (linear in H)."""

    @staticmethod
    def forward(ctx, C):
        ctx.save_for_backward(C)
        return mtk_log_sqrt_value(C)

    @staticmethod
    def backward(ctx, Ebar):
        (C,) = ctx.saved_tensors
        # transpose of H -> V (G o 0.5 V^T H V) V^T (note hHat is NOT symmetrised for the
        # diagonal entries and IS symmetrised for the off-diagonal ones; both are
        # self-adjoint operations)
        return mtk_log_sqrt_jvp(C, Ebar)


def hencky_energy(grad_u, K, G):
    """hyperelasticity/hencky.py:10-18 with log_strain base.py:107-110."""
    F = deformation_gradient(grad_u)
    C = F.transpose(-1, -2) @ F
    E = _MtkLogSqrt.apply(C)
    trE = torch.einsum("...ii->...", E)
    return 0.5 * K * trE ** 2 + G * torch.einsum("...ij,...ij->...", E, E)


def I2(grad_u):
    """mechanics/base.py:65-71"""
    F = deformation_gradient(grad_u)
    C = F.transpose(-1, -2) @ F
    I1 = torch.einsum("...ii->...", C)
    return 0.5 * (I1 ** 2 - torch.einsum("...ii->...", C @ C))


def blatz_ko_energy(grad_u, G):
    """hyperelasticity/blatz_ko.py:9-24: W = G/2 (I2/I3 + 2 sqrt(I3) - 5), I3 = J^2 (base.py:82-84)."""
    J = jacobian(grad_u)
    I3 = J * J
    return (G / 2.0) * (I2(grad_u) / I3 + 2.0 * torch.sqrt(I3) - 5.0)


def simple_fefv_energy(model, grad_u, props, Z_old, dt, moduli, relaxation_times):
    """hyperviscoelasticity/simple_fefv.py:20-40,85-108 (SimpleFeFv.energy, a_T = 1 as in the reference :27):
    psi = psi_eq(grad_u) + sum_b [G_b ||dev Ee_b||^2 + G_b tau_b ||dev Dv_b||^2], per Prony branch b
      Fe_trial = F Fv_old^-1,  Ee_trial = log_sqrt(Fe_trial^T Fe_trial),
      dEv = dt / (1 + dt / tau) dev(Ee_trial) / tau,  Ee = Ee_trial - dEv,  Dv = dEv / dt,
      Fv_new = expm(dEv) Fv_old.
    grad_u (..., 3, 3), Z_old (..., nb, 3, 3) -> (psi (...), Z_new (..., nb, 3, 3))."""
    psi = energy(model, grad_u, props)
    F = deformation_gradient(grad_u)
    eye = torch.eye(3, dtype=F.dtype)
    news = []
    for b, (G, tau) in enumerate(zip(moduli, relaxation_times)):
        Fv_old = Z_old[..., b, :, :]
        Fe = F @ torch.linalg.inv(Fv_old)
        Ee_trial = _MtkLogSqrt.apply(Fe.transpose(-1, -2) @ Fe)
        dev_tr = Ee_trial - torch.einsum("...ii->...", Ee_trial)[..., None, None] / 3.0 * eye
        dEv = dt * (1.0 / (1.0 + dt / tau)) * dev_tr / tau                    # state_increment :104-108
        Ee = Ee_trial - dEv
        Dv = dEv / dt
        dev_e = Ee - torch.einsum("...ii->...", Ee)[..., None, None] / 3.0 * eye
        dev_d = Dv - torch.einsum("...ii->...", Dv)[..., None, None] / 3.0 * eye
        psi = psi + G * (dev_e * dev_e).sum((-1, -2)) + (G * tau) * (dev_d * dev_d).sum((-1, -2))
        news.append(torch.linalg.matrix_exp(dEv) @ Fv_old)
    return psi, torch.stack(news, dim=-3)


MODELS = {
    "blatz_ko": (blatz_ko_energy, ("shear_modulus",)),
    "neohookean": (neohookean_energy, ("bulk_modulus", "shear_modulus")),
    "gent": (gent_energy, ("bulk_modulus", "shear_modulus", "Jm_parameter")),
    "swanson": (swanson_energy, ("bulk_modulus", "A1", "P1", "B1", "Q1", "C1", "R1",
                                 "cutoff_strain")),
    "hencky": (hencky_energy, ("bulk_modulus", "shear_modulus")),
}


def energy(model: str, grad_u, props):
    fn, names = MODELS[model]
    return fn(grad_u, *[props[n] for n in names])


def pk1_stress(model: str, grad_u, props):
    """P = dW/d(grad_u) by AD, as base.py:112-118 does with jacfwd."""
    g = grad_u.detach().clone().requires_grad_(True)
    W = energy(model, g, props)
    (P,) = torch.autograd.grad(W.sum(), g)
    return P
