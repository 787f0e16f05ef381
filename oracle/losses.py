"""Oracle (test infrastructure): loss functions and d(loss)/d(params).

Restates pancax/loss_functions/weak_form_loss_functions.py (EnergyLoss :8-88,
EnergyAndResidualLoss :115-212 via path_independent_call :194-205,
EnergyResidualAndReactionLoss :215-308) and data_loss_functions.py:7-24
(FullFieldDataLoss).  The time ``vmap`` of the reference is a Python loop here (same
math, bounded memory).  torch.autograd is the reference's ``eqx.filter_value_and_grad``
(pancax/optimizers.py:77-88).
"""
from __future__ import annotations

from dataclasses import dataclass, field as dc_field
from typing import Callable, Optional

import numpy as np
import torch

from . import models
from .fem import ParentElement
from .field import Field, unflatten_mlp
from .physics import (Physics, potential_energy, potential_energy_and_internal_force,
                      reaction_force, residual_norm)

F64 = torch.float64


@dataclass
class MLPSpec:
    n_in: int
    n_out: int
    width: int
    depth: int
    use_final_bias: bool = False


@dataclass
class PropSpec:
    """A constitutive property: fixed float, or bounded+trainable
    (pancax/constitutive_models/properties.py:10-38)."""
    name: str
    value: float = 0.0                 # fixed value, or prop_val[0] (raw) when bounded
    bounded: bool = False
    prop_min: float = 0.0
    prop_max: float = 0.0


@dataclass
class OracleProblem:
    coords: np.ndarray                 # (nn, nd) f64
    conns: np.ndarray                  # (ne, nnpe) int
    elem: str
    q_order: int
    times: np.ndarray                  # (nt,)
    kind: str                          # "poisson" | "solid"
    ndof: int
    mlp: MLPSpec
    model: Optional[str] = None
    formulation: Optional[str] = None
    props: list = dc_field(default_factory=list)          # list[PropSpec]
    bc_func: Optional[Callable] = None                     # g(x, t, z) on torch tensors
    source: Optional[Callable] = None                      # Poisson f(x)
    unknown_idx: Optional[np.ndarray] = None
    reaction_nodes: Optional[np.ndarray] = None
    reaction_dof: int = 0
    global_outputs: Optional[np.ndarray] = None            # (nt,)
    prony: Optional[tuple] = None                          # (moduli, relaxation_times): SimpleFeFv around `model`


def _build(problem: OracleProblem, flat_w, prop_raw):
    layers_np = unflatten_mlp(flat_w, problem.mlp.n_in, problem.mlp.n_out,
                              problem.mlp.width, problem.mlp.depth, problem.mlp.use_final_bias)
    layers = [(W, b) for W, b in layers_np]
    fld = Field(problem.coords, problem.times, layers, problem.bc_func)
    props = {}
    k = 0
    for p in problem.props:
        if p.bounded:
            props[p.name] = models.bounded_property(p.prop_min, p.prop_max, prop_raw[k])
            k += 1
        else:
            props[p.name] = p.value
    phys = Physics(problem.kind, problem.ndof, problem.model, props, problem.formulation,
                   problem.source, prony=problem.prony)
    pe = ParentElement(problem.elem, problem.q_order)
    return fld, phys, pe


def _leaves(problem, flat_w, prop_raw):
    w = torch.as_tensor(np.asarray(flat_w), dtype=F64).clone().requires_grad_(True)
    raw0 = [p.value for p in problem.props if p.bounded] if prop_raw is None else prop_raw
    pr = torch.as_tensor(np.asarray(raw0, dtype=np.float64), dtype=F64).clone()
    pr.requires_grad_(True)
    return w, pr


def field_values(problem, flat_w, t, prop_raw=None):
    """us (nn, ndof) at time t -- physics.vmap_field_values, base.py:248-250."""
    w, pr = _leaves(problem, flat_w, prop_raw)
    fld, _, _ = _build(problem, w, pr)
    with torch.no_grad():
        return fld(torch.as_tensor(problem.coords, dtype=F64), float(t)).numpy()


def energy_and_force(problem, us, prop_raw=None):
    """(Pi, f (nn, ndof)) for given nodal values -- base.py:356-361."""
    w, pr = _leaves(problem, np.zeros(_nparams(problem.mlp)), prop_raw)
    _, phys, pe = _build(problem, w, pr)
    coords = torch.as_tensor(problem.coords, dtype=F64)
    conns = torch.as_tensor(problem.conns, dtype=torch.long)
    pi, f = potential_energy_and_internal_force(phys, pe, coords, conns,
                                                torch.as_tensor(us, dtype=F64))
    return float(pi), f.detach().numpy()


def element_fields(problem, us, prop_raw=None):
    """(F, I1_bar, P) at the quadrature points as numpy arrays -- oracle/physics.element_fields."""
    from .physics import element_fields as _ef
    w, pr = _leaves(problem, np.zeros(_nparams(problem.mlp)), prop_raw)
    _, phys, pe = _build(problem, w, pr)
    coords = torch.as_tensor(problem.coords, dtype=F64)
    conns = torch.as_tensor(problem.conns, dtype=torch.long)
    F, I1b, P = _ef(phys, pe, coords, conns, torch.as_tensor(us, dtype=F64))
    return F.detach().numpy(), I1b.detach().numpy(), P.detach().numpy()


def stiffness_action(problem, us, v, prop_raw=None):
    """K v = d/d(eps) f(us + eps v) by double backward (K symmetric => grad of <f, v>)."""
    w, pr = _leaves(problem, np.zeros(_nparams(problem.mlp)), prop_raw)
    _, phys, pe = _build(problem, w, pr)
    coords = torch.as_tensor(problem.coords, dtype=F64)
    conns = torch.as_tensor(problem.conns, dtype=torch.long)
    u = torch.as_tensor(us, dtype=F64).clone().requires_grad_(True)
    _, f = potential_energy_and_internal_force(phys, pe, coords, conns, u, create_graph=True)
    (Kv,) = torch.autograd.grad((f * torch.as_tensor(v, dtype=F64)).sum(), u)
    return Kv.numpy()


def _nparams(m: MLPSpec):
    sizes = [m.n_in] + [m.width] * m.depth + [m.n_out]
    n = 0
    for i in range(len(sizes) - 1):
        n += sizes[i] * sizes[i + 1]
        if i < len(sizes) - 2 or m.use_final_bias:
            n += sizes[i + 1]
    return n


def loss_value(problem: OracleProblem, loss: dict, w, pr):
    """Differentiable loss (torch scalar) and aux dict.

    ``loss`` = {"kind": "energy"|"energy_residual"|"energy_residual_reaction",
                "energy_weight", "residual_weight", "reaction_weight"}"""
    fld, phys, pe = _build(problem, w, pr)
    coords = torch.as_tensor(problem.coords, dtype=F64)
    conns = torch.as_tensor(problem.conns, dtype=torch.long)
    kind = loss["kind"]
    we = loss.get("energy_weight", 1.0)
    wr = loss.get("residual_weight", 1.0)
    wg = loss.get("reaction_weight", 1.0)
    nt = len(problem.times)
    pis, Rs, reacts = [], [], []
    # loss["first_step"] = 1: the path-dependent call of a stateless model (weak_form:48-72,242-276): the scan
    # starts at time step 1; residual / reaction terms are still divided by len(times)
    first = int(loss.get("first_step", 0))
    if problem.prony is not None:
        # EnergyLoss.path_dependent_call with state variables (weak_form_loss_functions.py:48-72): the state of
        # every quadrature point starts at initial_state() (Fv = I per Prony branch, simple_fefv.py:72-75) and is
        # threaded through the steps n = 1 .. nt-1.  dt: times[1]-times[0] for step 1, and -- the update at the END
        # of the body, :63 -- times[n-1]-times[n-2] for step n >= 2.
        from .physics import potential_energy_with_state
        assert first == 1, "state variables: path-dependent calls only"
        ne, nq, nb = conns.shape[0], len(pe.w), len(problem.prony[0])
        Z = torch.eye(3, dtype=F64).expand(ne, nq, nb, 3, 3).clone()
        ts = problem.times
        dt = float(ts[1] - ts[0])
        energy, R, reaction = 0.0, 0.0, 0.0
        for n in range(1, nt):
            us = fld(coords, float(ts[n]))
            pi, Z_new = potential_energy_with_state(phys, pe, coords, conns, us, Z, dt)
            energy = energy + pi
            if kind != "energy":
                # :146-192, :242-276 with base.py:356-385: f = dPi/dus at FIXED old state (value_and_grad, argnums=3,
                # has_aux); the residual norm / reaction still depend on the weights through the old state as well
                (f,) = torch.autograd.grad(pi, us, create_graph=True)
                R = R + residual_norm(f, torch.as_tensor(problem.unknown_idx, dtype=torch.long))
                if kind == "energy_residual_reaction":
                    rn = torch.as_tensor(problem.reaction_nodes, dtype=torch.long)
                    r = reaction_force(f, rn, problem.reaction_dof)
                    reaction = reaction + torch.square(r - float(problem.global_outputs[n]))
            Z = Z_new
            dt = float(ts[n] - ts[n - 1])
        if kind == "energy":
            return we * energy, dict(energy=energy.detach())
        if kind == "energy_residual":
            return we * energy + wr * (R / nt), dict(energy=energy.detach(), residual=(R / nt).detach())
        return (we * energy + wr * (R / nt) + wg * (reaction / nt),
                dict(energy=energy.detach(), residual=(R / nt).detach(), global_data_loss=(reaction / nt).detach()))
    for t in problem.times[first:]:
        us = fld(coords, float(t))                           # weak_form:42,208,304
        if kind == "energy":
            pis.append(potential_energy(phys, pe, coords, conns, us))   # :43-45
        else:
            pi, f = potential_energy_and_internal_force(phys, pe, coords, conns, us,
                                                        create_graph=True)
            pis.append(pi)
            uidx = torch.as_tensor(problem.unknown_idx, dtype=torch.long)
            Rs.append(residual_norm(f, uidx))
            if kind == "energy_residual_reaction":
                rn = torch.as_tensor(problem.reaction_nodes, dtype=torch.long)
                reacts.append(reaction_force(f, rn, problem.reaction_dof))
    energy = torch.stack(pis).sum()
    if kind == "energy":
        # weak_form_loss_functions.py:86-88
        return we * energy, dict(energy=energy.detach())
    if kind == "energy_residual":
        # :203-205   pi, R = sum(pis), mean(Rs)
        R = torch.stack(Rs).mean()
        return we * energy + wr * R, dict(energy=energy.detach(), residual=R.detach())
    if kind == "energy_residual_reaction":
        # :289-300
        R = torch.stack(Rs).sum() / nt
        reactions = torch.stack(reacts)
        outs = torch.as_tensor(problem.global_outputs, dtype=F64)[first:]
        reaction_loss = torch.square(reactions - outs).sum() / nt
        val = we * energy + wr * R + wg * reaction_loss
        return val, dict(energy=energy.detach(), residual=R.detach(),
                         global_data_loss=reaction_loss.detach(),
                         reactions=reactions.detach())
    raise ValueError(kind)


def loss_and_grad(problem: OracleProblem, loss: dict, flat_w, prop_raw=None):
    """(loss, aux, dloss/dflat_w, dloss/dprop_raw) -- optimizers.py:77-88."""
    w, pr = _leaves(problem, flat_w, prop_raw)
    val, aux = loss_value(problem, loss, w, pr)
    gw, gp = torch.autograd.grad(val, [w, pr], allow_unused=True)
    gw = torch.zeros_like(w) if gw is None else gw
    gp = torch.zeros_like(pr) if gp is None else gp
    aux = {k: v.numpy() for k, v in aux.items()}
    return float(val.detach()), aux, gw.numpy(), gp.numpy()


def full_field_data_loss_and_grad(problem: OracleProblem, flat_w, inputs, outputs,
                                  weight=1.0):
    """FullFieldDataLoss: w * mean((u(x_i, t_i) - u_i)^2) over a batch.
    pancax/loss_functions/data_loss_functions.py:13-24 (inputs = [x..., t])."""
    w, pr = _leaves(problem, flat_w, None)
    fld, _, _ = _build(problem, w, pr)
    inp = torch.as_tensor(inputs, dtype=F64)
    nd = problem.coords.shape[1]
    xs, ts = inp[:, :nd], inp[:, nd]
    u_pred = fld(xs, ts)
    n_out = u_pred.shape[1]
    val = weight * torch.square(u_pred - torch.as_tensor(outputs, dtype=F64)[:, :n_out]).mean()
    (gw,) = torch.autograd.grad(val, [w])
    return float(val.detach()), gw.numpy()
