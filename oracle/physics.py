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

    def __init__(self, kind, ndof, model=None, props=None, formulation=None, source=None):
        self.kind, self.ndof = kind, ndof
        self.model, self.props, self.formulation = model, props, formulation
        self.source = source  # Poisson f(x): callable on (..., nd) tensors

    def energy_density(self, x_q, u_q, grad_u):
        """(ne, nq) energy density.  solid_mechanics.py:102-109 / poisson.py:16-19."""
        if self.kind == "poisson":
            f = self.source(x_q)                                   # (ne, nq)
            # 0.5 * dot(grad_u, grad_u.T) - f*u, summed (1x1)    poisson.py:18-19
            return 0.5 * (grad_u * grad_u).sum((-1, -2)) - f * u_q[..., 0]
        if self.formulation == "plane_strain":
            H = models.tensor_2D_to_3D(grad_u)                     # solid_mechanics.py:26-29
        elif self.formulation == "three_dimensional":
            H = grad_u                                             # solid_mechanics.py:80-83
        else:
            raise ValueError(self.formulation)
        return models.energy(self.model, H, self.props)


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
    H = models.tensor_2D_to_3D(grad_u) if physics.formulation == "plane_strain" else grad_u
    F = models.deformation_gradient(H)
    I1b = models.I1_bar(H)
    P = models.pk1_stress(physics.model, H, physics.props)
    return F, I1b, P
