"""Oracle (test infrastructure): element interpolants, potential energy, internal force,
residual norm and reaction force -- torch float64 on the CPU, batched over elements.

Restates pancax/physics_kernels/base.py:297-385,418-427 and
pancax/fem/function_space.py:63-89.
"""
from __future__ import annotations

import torch

from . import models
from .fem import ParentElement

F64 = torch.float64


class Physics:
    """kind in {"poisson", "solid"}; for "solid": model name, props dict (floats or torch
    scalars) and formulation in {"plane_strain", "three_dimensional"}.
    pancax/physics_kernels/solid_mechanics.py:86-109; poisson.py:8-19."""

    def __init__(self, kind, ndof, model=None, props=None, formulation=None, source=None, prony=None):
        self.kind, self.ndof = kind, ndof
        self.model, self.props, self.formulation = model, props, formulation
        self.source = source  # Poisson f(x): callable on (..., nd) tensors
        self.prony = prony    # (moduli, relaxation_times): SimpleFeFv around `model` (simple_fefv.py:11-14)

    def energy_density_with_state(self, grad_u, Z_old, dt):
        """(psi (ne, nq), Z_new (ne, nq, nb, 3, 3)) -- SolidMechanics.energy with a SimpleFeFv model.  PlaneStress
        (solid_mechanics.py:49-71): the root find runs on the stress of the WHOLE model at fixed old state --
        cauchy_stress = jacfwd(energy)[0] F^T / J with energy = (psi_eq + sum psi_neq, Z_new), mechanics/base.py:10-17,
        112-118 -- so the out-of-plane gradient depends on grad_u AND on the old state."""
        if self.formulation == "plane_stress":
            def psi_of(H3):
                return models.simple_fefv_energy(self.model, H3, self.props, Z_old, dt, self.prony[0], self.prony[1])[0]
            H3 = plane_stress_gradient(self.model, grad_u, self.props, energy_fn=psi_of)
        else:
            H3 = self.modify_field_gradient(grad_u)
        return models.simple_fefv_energy(self.model, H3, self.props, Z_old, dt, self.prony[0], self.prony[1])

    def energy_density(self, x_q, u_q, grad_u):
        """(ne, nq) energy density.  solid_mechanics.py:102-109 / poisson.py:16-19."""
        if self.kind == "poisson":
            f = self.source(x_q)                                   # (ne, nq)
            # 0.5 * dot(grad_u, grad_u.T) - f*u, summed (1x1)    poisson.py:18-19
            return 0.5 * (grad_u * grad_u).sum((-1, -2)) - f * u_q[..., 0]
        return models.energy(self.model, self.modify_field_gradient(grad_u), self.props)

    def modify_field_gradient(self, grad_u):
        """formulation.modify_field_gradient, solid_mechanics.py:26-29 (PlaneStrain), :49-71 (PlaneStress),
        :80-83 (ThreeDimensional)."""
        if self.formulation == "plane_strain":
            return models.tensor_2D_to_3D(grad_u)
        if self.formulation == "three_dimensional":
            return grad_u
        if self.formulation == "plane_stress":
            return plane_stress_gradient(self.model, grad_u, self.props)
        raise ValueError(self.formulation)


def plane_stress_gradient(model, grad_u, props, energy_fn=None):
    """PlaneStress.modify_field_gradient (solid_mechanics.py:49-71): the out-of-plane entry x = grad_u_33 is the
    root of sigma_33(x) = 0, found by scalar_root_find.find_root (Newton safeguarded by bisection, guess 0.05,
    bracket [-0.99, 10], x_tol 1e-13; math/scalar_root_find.py:18-160) and differentiated by the implicit function
    theorem (jax.lax.custom_root).  sigma_33 = P_33 F_33 / J for this block structure, so the root is that of
    P_33 = dW/dx.  Here: a batched safeguarded Newton iteration without a tape, then two differentiable Newton
    steps from the converged root, which reproduce the derivatives of the implicit function up to third order
    (Newton's map is a second-order contraction at the root) -- what the outer gradients of the residual losses need."""
    E33 = torch.zeros(3, 3, dtype=F64)
    E33[2, 2] = 1.0

    # energy_fn(H3) -> psi: the energy the stress is taken of (a model with state: psi at FIXED old state, differentiable
    # w.r.t. that state through its closure); default: the hyperelastic model
    def g_and_dg(H2, x, create_graph):
        H3 = models.tensor_2D_to_3D(H2) + x[..., None, None] * E33
        W = models.energy(model, H3, props) if energy_fn is None else energy_fn(H3)
        (g,) = torch.autograd.grad(W.sum(), x, create_graph=True)
        (dg,) = torch.autograd.grad(g.sum(), x, create_graph=create_graph)
        return g, dg

    gu = grad_u.detach()
    lo = torch.full(gu.shape[:-2], -0.99, dtype=F64)
    hi = torch.full(gu.shape[:-2], 10.0, dtype=F64)
    # incompressible estimate F33 = 1 / det(F_2D) as the start (the reference starts from 0.05; the root is unique
    # in the bracket for these models, P_33 is increasing in x)
    x = (1.0 / torch.linalg.det(gu + torch.eye(2, dtype=F64)) - 1.0).clamp(-0.99, 10.0)
    for _ in range(60):
        g, dg = g_and_dg(gu, x.detach().clone().requires_grad_(True), False)
        g, dg = g.detach(), dg.detach()
        lo = torch.where(g < 0, x, lo)
        hi = torch.where(g < 0, hi, x)
        xn = x - g / dg
        bad = ~(xn > lo) | ~(xn < hi) | ~(dg > 0)
        xn = torch.where(bad, 0.5 * (lo + hi), xn)
        done = bool((xn - x).abs().max() < 1e-15)
        x = xn
        if done:
            break
    x = x.detach().clone().requires_grad_(True)
    for _ in range(2):
        g, dg = g_and_dg(grad_u, x, True)
        x = x - g / dg
    return models.tensor_2D_to_3D(grad_u) + x[..., None, None] * E33


def element_geometry(pe: ParentElement, X_el):
    """X_el (ne, nnpe, nd) -> grad_N (ne, nq, nnpe, nd), JxW (ne, nq), detJ (ne, nq).

    function_space.py:63-69 (JxWs), :80-89 (shape_function_gradients):
      J = (X^T dN)^T   ->  J[k, j] = sum_a dN[q, a, k] X[a, j]
      Jinv = inv(J);  grad_N = (Jinv dN^T)^T  -> grad_N[a, j] = sum_k Jinv[j, k] dN[q, a, k]
      JxW = det(J) * w
    """
    dN = torch.as_tensor(pe.dN, dtype=F64)                          # (nq, nnpe, nd)
    w = torch.as_tensor(pe.w, dtype=F64)
    J = torch.einsum("qak,eaj->eqkj", dN, X_el)
    detJ = torch.linalg.det(J)
    Jinv = torch.linalg.inv(J)
    grad_N = torch.einsum("eqjk,qak->eqaj", Jinv, dN)
    return grad_N, detJ * w, detJ


def element_interpolants(pe: ParentElement, X_el, U_el):
    """base.py:310-325.  Returns x_q (ne,nq,nd), u_q (ne,nq,ndof), grad_u (ne,nq,ndof,nd),
    JxW (ne,nq).  grad_u[i, j] = sum_a U[a, i] grad_N[a, j]."""
    N = torch.as_tensor(pe.N, dtype=F64)
    grad_N, JxW, _ = element_geometry(pe, X_el)
    x_q = torch.einsum("qa,eaj->eqj", N, X_el)
    u_q = torch.einsum("qa,eai->eqi", N, U_el)
    grad_u = torch.einsum("eai,eqaj->eqij", U_el, grad_N)
    return x_q, u_q, grad_u, JxW


def potential_energy(physics: Physics, pe: ParentElement, coords, conns, us):
    """Pi = sum_e sum_q JxW * W.  base.py:297-307 (element_integral), :335-354."""
    X_el = coords[conns]                                            # base.py:338-339
    U_el = us[conns]
    x_q, u_q, grad_u, JxW = element_interpolants(pe, X_el, U_el)
    W = physics.energy_density(x_q, u_q, grad_u)
    return (JxW * W).sum(dim=1).sum()


def potential_energy_with_state(physics: Physics, pe: ParentElement, coords, conns, us, Z_old, dt):
    """(Pi, Z_new): base.py:297-307,335-354 with state variables threaded through the element integral."""
    X_el, U_el = coords[conns], us[conns]
    _, _, grad_u, JxW = element_interpolants(pe, X_el, U_el)
    W, Z_new = physics.energy_density_with_state(grad_u, Z_old, dt)
    return (JxW * W).sum(dim=1).sum(), Z_new


def potential_energy_and_internal_force(physics, pe, coords, conns, us, create_graph=False):
    """(Pi, f = dPi/dus).  base.py:356-361."""
    if not us.requires_grad:
        us = us.detach().clone().requires_grad_(True)
    pi = potential_energy(physics, pe, coords, conns, us)
    (f,) = torch.autograd.grad(pi, us, create_graph=create_graph)
    return pi, f


R_FLOOR = 1.0e-10


def residual_norm(f, unknown_idx):
    """R = || f.ravel()[unknownIndices] ||_2.  base.py:369-371, :381.
    Convention (shared with the CUDA path): when R <= R_FLOOR (f is zero up to
    rounding noise, e.g. t = 0 with a ramped Dirichlet wrapper) dR/df := 0; the reference's
    jnp.linalg.norm gives NaN or a noise direction there."""
    fu = f.reshape(-1)[unknown_idx]
    s = (fu * fu).sum()
    ok = s > R_FLOOR * R_FLOOR
    safe = torch.where(ok, s, torch.ones_like(s))
    return torch.where(ok, torch.sqrt(safe), torch.sqrt(s.detach()))


def reaction_force(f, reaction_nodes, reaction_dof):
    """sum f[reaction_nodes, reaction_dof].  base.py:382-384."""
    return f[reaction_nodes, reaction_dof].sum()


def element_fields(physics: Physics, pe: ParentElement, coords, conns, us):
    """Post-processing fields at the quadrature points: (F (ne,nq,3,3), I1_bar (ne,nq), P (ne,nq,3,3)).

    Restates what the element_pp wrappers evaluate (pancax/physics_kernels/base.py:24-158 with the
    methods registered in solid_mechanics.py:115-170): the formulation embeds grad_u in 3x3
    (modify_field_gradient), then deformation_gradient / I1_bar (kinematic methods,
    mechanics/base.py:19-21,50-63) and pk1_stress = jacfwd(energy) w.r.t. the 3x3 gradient
    (constitutive method, mechanics/base.py:112-118)."""
    X_el, U_el = coords[conns], us[conns]
    _, _, grad_u, _ = element_interpolants(pe, X_el, U_el)
    H = physics.modify_field_gradient(grad_u)
    if physics.formulation == "plane_stress":
        H = H.detach()
    F = models.deformation_gradient(H)
    I1b = models.I1_bar(H)
    P = models.pk1_stress(physics.model, H, physics.props)
    return F, I1b, P
