"""Oracle (test infrastructure) for the strong-form / collocation path: the jet
(u, grad u, laplacian u, du/dt) of the Field at points and the strong-form residual losses.

Restates pancax/physics_kernels/base.py:209-234 (field_gradients = jacfwd wrt x, field_hessians,
field_laplacians = trace of the hessian, field_time_derivatives = jacfwd wrt t),
poisson.py:26-29 (residual = -lap(u) - f(x)), heat_equation.py:11-17 (rho c du/dt - k lap(u)) and
loss_functions/strong_form_loss_functions.py:8-27 (mean over points of residual^2, mean over times),
with torch.func in the role of jax.jacfwd / jax.hessian.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.func import hessian, jacfwd, vmap

F64 = torch.float64


def field_jet(fld, x, t):
    """x (N, nd), t (N,) -> u (N, ndof), grad_u (N, ndof, nd), lap_u (N, ndof), u_t (N, ndof)."""
    x = torch.as_tensor(x, dtype=F64)
    t = torch.as_tensor(t, dtype=F64).reshape(-1)

    def f(xi, ti):                              # base.py:205-207 field_values on one point
        return fld(xi[None, :], ti[None])[0]
    u = vmap(f)(x, t)
    gu = vmap(jacfwd(f, argnums=0))(x, t)                     # base.py:209-211
    H = vmap(hessian(f, argnums=0))(x, t)                     # base.py:213-215  (N, ndof, nd, nd)
    lap = torch.diagonal(H, dim1=-2, dim2=-1).sum(-1)         # base.py:217-226
    ut = vmap(jacfwd(f, argnums=1))(x, t)                     # base.py:228-229
    return u, gu, lap, ut


def residual(kind, coeffs, fld, x, t, source=None):
    """Strong-form residual (N, ndof).  kind "poisson": -lap(u) - f(x) (poisson.py:26-29);
    "heat": rho c u_t - k lap(u) with coeffs = (rho, c, k) (heat_equation.py:11-17)."""
    _, _, lap, ut = field_jet(fld, x, t)
    if kind == "poisson":
        f = source(torch.as_tensor(x, dtype=F64)) if source is not None else 0.0
        f = f.reshape(-1, 1) if isinstance(f, torch.Tensor) else f
        return -lap - f
    if kind == "heat":
        rho, c, k = coeffs
        return rho * c * ut - k * lap
    raise ValueError(kind)


def residual_loss(kind, coeffs, fld, x, t, weight=1.0, source=None, target=None):
    """weight * mean((residual - target)^2) over a batch of points (strong_form_loss_functions.py:15-27 when
    every time step carries the same points; the collocation example's user loss with ``target``)."""
    r = residual(kind, coeffs, fld, x, t, source)
    if target is not None:
        r = r - torch.as_tensor(target, dtype=F64).reshape(r.shape)
    return weight * torch.square(r).mean(), r


def loss_and_grad(problem, flat_w, kind, coeffs, x, t, weight=1.0, source=None, target=None):
    """(loss, residuals, d loss / d flat_w) with the restated Field of ``problem`` (oracle.losses.OracleProblem)."""
    from .losses import _build, _leaves
    w, pr = _leaves(problem, flat_w, None)
    fld, _, _ = _build(problem, w, pr)
    val, r = residual_loss(kind, coeffs, fld, x, t, weight, source, target)
    (gw,) = torch.autograd.grad(val, [w])
    return float(val.detach()), r.detach().numpy(), gw.numpy()


def jet_numpy(problem, flat_w, x, t):
    from .losses import _build, _leaves
    w, pr = _leaves(problem, flat_w, None)
    fld, _, _ = _build(problem, w, pr)
    with torch.no_grad():
        pass
    out = field_jet(fld, x, t)
    return tuple(o.detach().numpy() for o in out)
