"""pancax/loss_functions: the weak-form losses and the full-field data loss, evaluated by the
CUDA library.  ``loss(params, problem, *args) -> (loss, aux)`` as in the reference;
``loss.value_and_grad(params, problem, *args) -> ((loss, aux), grads)`` is what
``eqx.filter_value_and_grad(loss.filtered_loss, has_aux=True)`` returns in pancax/optimizers.py:79-88.
Values are device tensors (no host synchronisation inside)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .constitutive_models import BoundedProperty, SimpleFeFv
from .engine import Engine, tabulate_bc
from .physics_kernels import BaseStrongFormPhysics, HeatEquation, Poisson, SolidMechanics


class Grads:
    """Gradient 'pytree': flat MLP weight gradient + trainable-property gradient."""

    def __init__(self, fields, props):
        self.fields, self.props = fields, props

    def __add__(self, other):
        return Grads(self.fields + other.fields, self.props + other.props)

    def scaled(self, c):
        return Grads(self.fields * c, self.props * c)


class TracedLoss:
    """A loss value that carries d(value)/d(weights, properties): what the built-in losses return while a
    UserDefinedLossFunction is being differentiated.  The reference lets jax trace the user's Python function
    (loss_functions/utils.py:30-34, optimizers.py:77-88); here the function is EXECUTED once and the arithmetic it does on
    the built-in losses' values (+, -, *, /, unary -, with numbers or with each other) is differentiated by the chain
    rule on this object -- the built-in losses and their gradients themselves are the library's fused CUDA calls."""

    def __init__(self, value, grads: Grads):
        self.value, self.grads = value, grads

    @staticmethod
    def _split(other):
        return (other.value, other.grads) if isinstance(other, TracedLoss) else (other, None)

    def __add__(self, other):
        v, g = self._split(other)
        return TracedLoss(self.value + v, self.grads if g is None else self.grads + g)

    __radd__ = __add__

    def __neg__(self):
        return TracedLoss(-self.value, self.grads.scaled(-1.0))

    def __sub__(self, other):
        v, g = self._split(other)
        return TracedLoss(self.value - v, self.grads if g is None else self.grads + g.scaled(-1.0))

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        v, g = self._split(other)
        if g is None:
            return TracedLoss(self.value * v, self.grads.scaled(v))
        return TracedLoss(self.value * v, self.grads.scaled(v) + g.scaled(self.value))      # product rule

    __rmul__ = __mul__

    def __truediv__(self, other):
        v, g = self._split(other)
        if g is None:
            return TracedLoss(self.value / v, self.grads.scaled(1.0 / v))
        return TracedLoss(self.value / v, self.grads.scaled(1.0 / v) + g.scaled(-self.value / (v * v)))

    def __float__(self):
        return float(self.value)


_TRACING = [False]      # set while UserDefinedLossFunction.value_and_grad runs the user's function


def _wrap_bc(bc_func, ndof):
    """Accept batched wrappers (x (N,nd), t, z (N,ndof)) -- pancax_b200.bvps -- and, for small
    meshes, per-point wrappers written for pancax (x (nd,), t, z (ndof,))."""
    def batched(x, t, z):
        try:
            out = np.asarray(bc_func(x, t, z), dtype=np.float64)
            if out.shape == z.shape:
                return out
        except Exception:
            pass
        out = np.empty_like(z)
        tt = np.broadcast_to(np.asarray(t, dtype=np.float64), (x.shape[0],))
        for n in range(x.shape[0]):
            out[n] = np.asarray(bc_func(_AtArray(x[n]), float(tt[n]), _AtArray(z[n])), dtype=np.float64)
        return out
    return batched


class _AtArray(np.ndarray):
    """numpy array with jax's functional ``.at[idx].set(v)`` so per-point wrappers written for
    pancax (bvps/*.py use ``u_out.at[0].set(...)``) can be tabulated on the host."""
    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)

    class _At:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            a = self.a

            class _S:
                def set(self_inner, v):
                    b = a.copy().view(_AtArray)
                    b[idx] = v
                    return b
            return _S()

    @property
    def at(self):
        return _AtArray._At(self)


def get_engine(problem, params, deterministic: bool = True) -> Engine:
    """One Engine (pcx_handle) per (problem, field shape); built on first use, like the first
    traced call of a jitted step."""
    fld = params.fields
    net = fld.networks
    shard = getattr(problem, "shard", None)
    if shard is not None and getattr(shard, "x_mins", None) is not None and getattr(fld, "_shard_bounds_seen", None) is not shard:
        # fields.py:70-71 normalises with the bounds of the WHOLE mesh.  A Field built on a shard took the shard's
        # own bounds from problem.coords: every rank would then evaluate a different u(x, t) from the same weights.
        # A Field whose bounds were set by hand to something else is left alone.  Checked once per (Field, shard): the
        # reduction over the coordinates costs milliseconds and this function runs on every loss call.
        if getattr(shard, "local_bounds", None) is None:
            shard.local_bounds = (problem.coords.min(axis=0), problem.coords.max(axis=0))
        loc_lo, loc_hi = shard.local_bounds
        if np.array_equal(fld.x_mins, loc_lo) and np.array_equal(fld.x_maxs, loc_hi):
            fld.x_mins, fld.x_maxs = shard.x_mins.copy(), shard.x_maxs.copy()
        fld._shard_bounds_seen = shard
    # everything baked into the pcx_handle at set-up is part of the key; the wrapper itself (not its id(), which can
    # be reused after garbage collection) is held by the key
    key = (net.n_inputs, net.n_outputs, net.n_neurons, net.n_layers, net.use_final_bias,
           fld.dirichlet_bc_func, deterministic,
           tuple(np.asarray(fld.x_mins, dtype=np.float64).tolist()), tuple(np.asarray(fld.x_maxs, dtype=np.float64).tolist()),
           float(fld.t_min), float(fld.t_max), bool(fld.normalize_time))
    eng = problem._engines.get(key)
    if eng is not None:
        return eng
    dom = problem.domain
    gd = getattr(problem, "global_data", None)
    unk = None if dom.dof_manager is None else dom.dof_manager.unknownIndices
    rn = None if gd is None else gd.reaction_nodes
    if shard is not None:     # entries owned by this shard first (pcx_set_ownership)
        unk, n_uo, rn, n_ro = shard.plan.order_owned_first(unk, rn, problem.n_dofs)
    eng = Engine(dom.coords, dom.conns, dom.mesh.elem_type, dom.q_order, problem.n_dofs, dom.times,
                 unknown_idx=unk, reaction_nodes=rn,
                 reaction_dof=0 if gd is None else gd.reaction_dof, deterministic=deterministic,
                 device=net.weights.device)
    if shard is not None:
        eng.set_ownership(n_uo, n_ro)
    phys = problem.physics
    if isinstance(phys, HeatEquation):
        eng.set_physics("poisson")          # strong form only: no energy functional, the engine just needs a physics
    elif isinstance(phys, Poisson):
        _, _, xq = eng.element_geometry(want_grad_N=False, want_JxW=False)
        eng.set_physics("poisson", source_q=np.asarray(phys.f(xq.cpu().numpy()), dtype=np.float64))
    elif isinstance(phys, SolidMechanics):
        model = phys.constitutive_model
        prony = None
        if isinstance(model, SimpleFeFv):      # equilibrium model + Prony branches (state variables)
            prony = (model.prony_series.moduli, model.prony_series.relaxation_times)
            model = model.eq_model
        props = []
        for name in model.properties():
            p = getattr(model, name)
            props.append((p.prop_min, p.prop_max) if isinstance(p, BoundedProperty) else float(p))
        eng.set_physics("solid", model.name, phys.formulation.name, props, prony=prony)
    else:
        raise NotImplementedError(f"physics {type(phys)} is not on the B200 hot path")
    a = b = None
    if fld.dirichlet_bc_func is not None:
        a, b = tabulate_bc(_wrap_bc(fld.dirichlet_bc_func, problem.n_dofs), dom.coords, dom.times,
                           problem.n_dofs)
    eng.set_field(net.n_inputs, net.n_outputs, net.n_neurons, net.n_layers, net.use_final_bias,
                  bc_a=a, bc_b=b, x_min=fld.x_mins, x_max=fld.x_maxs, t_min=fld.t_min, t_max=fld.t_max,
                  normalize_time=fld.normalize_time)
    problem._engines[key] = eng
    return eng


class BaseLossFunction:
    def filtered_loss(self, diff_params, static_params, *args, **kwargs):
        # base_loss_function.py:23-25; params are not partitioned here: the trainable subset is
        # selected by the optimizer's filter
        return self.__call__(diff_params, *args, **kwargs)

    def value_and_grad(self, params, problem, *args):
        raise NotImplementedError


class PhysicsLossFunction(BaseLossFunction):
    kind: str = None
    deterministic: bool = True

    def _weights(self):
        return dict(energy_weight=1.0, residual_weight=1.0, reaction_weight=1.0)

    def _run(self, params, problem, want_grad, path_dependent=None):
        first_step = 0
        model = getattr(problem.physics, "constitutive_model", None)
        has_state = model is not None and model.num_state_variables != 0
        if getattr(self, "is_path_dependent", False) if path_dependent is None else path_dependent:
            # weak_form_loss_functions.py:48-72,151-192,242-276: the scan starts at time step 1.  Without state variables
            # that is the only difference from the path-independent call; with them (SimpleFeFv) the library threads the
            # state through the steps and -- residual / reaction losses -- adds the second-order terms through the history.
            if self.kind == "energy_residual" and path_dependent is None:
                raise AssertionError("EnergyAndResidualLoss.__call__: the path-dependent case asserts in the reference too "
                                     "(weak_form_loss_functions.py:145-147); call path_dependent_call")
            first_step = 1
        elif has_state:
            raise NotImplementedError("models with state variables need the path-dependent call (is_path_dependent=True)")
        eng = get_engine(problem, params, self.deterministic)
        gd = getattr(problem, "global_data", None)
        gout = None
        if self.kind == "energy_residual_reaction":
            if gd is None:
                raise ValueError("EnergyResidualAndReactionLoss needs problem.global_data")
            gout = gd.outputs
        shard = getattr(problem, "shard", None)
        if shard is not None and shard.plan.n_shards > 1 and self.kind != "energy":
            # one element shard per rank: phased ABI, two small exchanges in between; the returned
            # values are PARTIAL -- sum them over ranks (parallel.PackedResult.allreduce) like EnergyLoss
            if first_step:
                raise NotImplementedError("path-dependent residual losses on shards")
            return eng.sharded_loss_and_grad(
                self.kind, params.fields.networks.weights, params.prop_raw if eng.n_train else None,
                shard.exchange, global_outputs=gout, **self._weights())
        loss, aux, gw, gp = eng.loss_and_grad(
            self.kind, params.fields.networks.weights,
            params.prop_raw if eng.n_train else None, global_outputs=gout, want_grad=want_grad,
            first_step=first_step, **self._weights())
        return loss, aux, gw, gp

    def _aux(self, aux):
        raise NotImplementedError

    def __call__(self, params, problem, *args):
        if _TRACING[0]:
            (loss, aux), grads = self.value_and_grad(params, problem, *args)
            return TracedLoss(loss, grads), aux
        loss, aux, _, _ = self._run(params, problem, False)
        return loss[0], self._aux(aux)

    def value_and_grad(self, params, problem, *args):
        loss, aux, gw, gp = self._run(params, problem, True)
        return (loss[0], self._aux(aux)), Grads(gw, gp)


class EnergyLoss(PhysicsLossFunction):
    """weak_form_loss_functions.py:8-88: L = w * sum_t Pi_t, aux energy."""
    kind = "energy"

    def __init__(self, is_path_dependent: Optional[bool] = False, use_fori_loop: Optional[bool] = False,
                 weight: Optional[float] = 1.0):
        self.is_path_dependent, self.use_fori_loop, self.weight = is_path_dependent, use_fori_loop, weight

    def _weights(self):
        return dict(energy_weight=self.weight)

    def _aux(self, aux):
        return dict(energy=aux[0])


class EnergyAndResidualLoss(PhysicsLossFunction):
    """weak_form_loss_functions.py:115-212.  The reference's ``__call__`` drops its return value
    (:145-149); here it returns what ``path_independent_call`` (:194-205) computes:
    L = w_e sum_t Pi_t + w_r mean_t R_t."""
    kind = "energy_residual"

    def __init__(self, energy_weight: Optional[float] = 1.0, is_path_dependent: Optional[bool] = False,
                 residual_weight: Optional[float] = 1.0, use_fori_loop: Optional[bool] = False):
        self.energy_weight, self.residual_weight = energy_weight, residual_weight
        self.is_path_dependent, self.use_fori_loop = is_path_dependent, use_fori_loop

    def _weights(self):
        return dict(energy_weight=self.energy_weight, residual_weight=self.residual_weight)

    def _aux(self, aux):
        return dict(energy=aux[0], residual=aux[1])

    def path_independent_call(self, params, problem):
        loss, aux, _, _ = self._run(params, problem, False, path_dependent=False)
        return loss[0], self._aux(aux)

    def path_dependent_call(self, params, problem):
        """weak_form_loss_functions.py:151-192: L = w_e sum_{n>=1} Pi_n + w_r sum_{n>=1} R_n / len(times), the state
        threaded through the steps."""
        loss, aux, _, _ = self._run(params, problem, False, path_dependent=True)
        return loss[0], self._aux(aux)

    def path_dependent_value_and_grad(self, params, problem):
        loss, aux, gw, gp = self._run(params, problem, True, path_dependent=True)
        return (loss[0], self._aux(aux)), Grads(gw, gp)


class EnergyResidualAndReactionLoss(PhysicsLossFunction):
    """weak_form_loss_functions.py:215-308: L = w_e sum Pi_t + w_r sum R_t / nt
    + w_g mean_t (reaction_t - data_t)^2; aux energy, residual, global_data_loss, reactions(nt)."""
    kind = "energy_residual_reaction"

    def __init__(self, energy_weight: Optional[float] = 1.0, is_path_dependent: Optional[bool] = False,
                 residual_weight: Optional[float] = 1.0, reaction_weight: Optional[float] = 1.0,
                 use_fori_loop: Optional[bool] = False):
        self.energy_weight, self.residual_weight = energy_weight, residual_weight
        self.reaction_weight = reaction_weight
        self.is_path_dependent, self.use_fori_loop = is_path_dependent, use_fori_loop

    def _weights(self):
        return dict(energy_weight=self.energy_weight, residual_weight=self.residual_weight,
                    reaction_weight=self.reaction_weight)

    def _aux(self, aux):
        return dict(energy=aux[0], residual=aux[1], global_data_loss=aux[2], reactions=aux[4:])


class FullFieldDataLoss(BaseLossFunction):
    """data_loss_functions.py:7-24: w * mean((u(x_i, t_i) - u_i)^2) on a batch (inputs = [x..., t])."""

    def __init__(self, weight: Optional[float] = 1.0):
        self.weight = weight

    def _run(self, params, problem, inputs, outputs):
        eng = get_engine(problem, params)
        inputs = np.asarray(inputs, dtype=np.float64)
        frac = 1.0
        shard = getattr(problem, "shard", None)
        if shard is not None and shard.plan.n_shards > 1:
            # the batch rows are sharded over the ranks (rows rank, rank + R, ...); w * mean over the WHOLE batch
            # = sum over ranks of (w n_r / N) * mean over the rank's rows, so the returned values are PARTIAL and add up in
            # the packed allreduce like the weak-form losses (the field is a function of (x, t, theta): any rank can
            # evaluate any row)
            n_all = inputs.shape[0]
            inputs = inputs[shard.plan.index::shard.plan.n_shards]
            outputs = np.asarray(outputs, dtype=np.float64)[shard.plan.index::shard.plan.n_shards]
            frac = inputs.shape[0] / n_all
            if inputs.shape[0] == 0:
                z = torch.zeros(1, dtype=torch.float64, device=params.fields.networks.weights.device)
                return z, torch.zeros_like(params.fields.networks.weights), 0.0
        nd = problem.n_dims
        a = b = None
        bc = params.fields.dirichlet_bc_func
        if bc is not None:
            g = _wrap_bc(bc, problem.n_dofs)
            xs, ts = inputs[:, :nd], inputs[:, nd]
            z0 = np.zeros((inputs.shape[0], problem.n_dofs))
            a = g(xs, ts, z0)
            b = g(xs, ts, z0 + 1.0) - a
        loss, gw = eng.field_data_loss_and_grad(params.fields.networks.weights, inputs, outputs, a, b,
                                                weight=self.weight * frac)
        return loss, gw, frac

    def __call__(self, params, problem, inputs, outputs):
        if _TRACING[0]:
            (loss, aux), grads = self.value_and_grad(params, problem, inputs, outputs)
            return TracedLoss(loss, grads), aux
        loss, _, _ = self._run(params, problem, inputs, outputs)
        return loss[0], {"field_data_loss": loss[0] / self.weight if self.weight else loss[0]}

    def value_and_grad(self, params, problem, inputs, outputs):
        loss, gw, _ = self._run(params, problem, inputs, outputs)
        aux = {"field_data_loss": loss[0] / self.weight if self.weight else loss[0]}
        return (loss[0], aux), Grads(gw, torch.zeros_like(params.prop_raw))


class StrongFormResidualLoss(BaseLossFunction):
    """strong_form_loss_functions.py:8-27: weight * mean_t mean_nodes residual^2 with the physics' strong-form
    residual at the mesh nodes (collocation points); aux ``residual``.  The jet of the field (value, coordinate
    tangents, Laplacian, time derivative) and its adjoint are fused CUDA kernels (pcx_strong_form_loss_and_grad).
    ``__call__(params, problem, inputs, outputs)`` evaluates mean((residual - outputs)^2) on a batch of (x..., t)
    rows instead: the user loss of examples/forward_problems/poisson/collocation_example.py:36-41."""

    def __init__(self, weight: Optional[float] = 1.0):
        self.weight = weight

    def _run(self, params, problem, inputs, outputs, want_grad):
        from .jets import tabulate_bc_jet
        phys = problem.physics
        if not isinstance(phys, BaseStrongFormPhysics):
            raise TypeError(f"{type(phys)} has no strong form on the B200 path")
        eng = get_engine(problem, params)

        def tables(pts):
            tab = tabulate_bc_jet(params.fields.dirichlet_bc_func, pts, problem.n_dofs)
            src = phys.strong_form_source(pts[:, :problem.n_dims])
            if src is not None:
                src = np.asarray(src, dtype=np.float64).reshape(pts.shape[0], problem.n_dofs)
            return eng._dev(pts), eng._dev(tab), eng._dev(src)
        if inputs is None:          # every (time step, node): static, tabulated once per problem
            cache = problem._engines.setdefault("_strong_form_static", {})
            # keyed by id() but holding the wrapper itself: an id() reused after garbage collection cannot pair another
            # wrapper with these tables
            bc = params.fields.dirichlet_bc_func
            hit = cache.get(id(bc))
            if hit is None or hit[0] is not bc:
                hit = cache[id(bc)] = (bc, tables(problem.domain.collocation_points()))
            pts, tab, src = hit[1]
        else:
            pts, tab, src = tables(np.ascontiguousarray(inputs, dtype=np.float64))
        c_t, c_lap = phys.strong_form_coefficients()
        loss, _, gw = eng.strong_form_loss_and_grad(params.fields.networks.weights, pts, c_t, c_lap, self.weight,
                                                    tab=tab, src=src, target=outputs, want_grad=want_grad)
        return loss, gw

    def __call__(self, params, problem, inputs=None, outputs=None):
        loss, _ = self._run(params, problem, inputs, outputs, False)
        return loss[0], dict(residual=loss[0] / self.weight if self.weight else loss[0])

    def value_and_grad(self, params, problem, inputs=None, outputs=None):
        loss, gw = self._run(params, problem, inputs, outputs, True)
        aux = dict(residual=loss[0] / self.weight if self.weight else loss[0])
        return (loss[0], aux), Grads(gw, torch.zeros_like(params.prop_raw))


class CombineLossFunctions(BaseLossFunction):
    """loss_functions/utils.py:4-27.  Data losses inside the combination receive the extra
    ``(inputs, outputs)`` arguments; physics losses do not (example_v2.py:106-109)."""

    def __init__(self, *funcs, with_props=False):
        self.funcs = funcs
        self.with_props = with_props

    def _args_for(self, f, args):
        return args if isinstance(f, FullFieldDataLoss) else ()

    def __call__(self, params, domain, *args):
        loss, aux = 0.0, dict()
        for f in self.funcs:
            tl, ta = f(params, domain, *self._args_for(f, args))
            loss = loss + tl
            aux.update(ta)
        if self.with_props:
            aux.update({"props": params.properties()})
        return loss, aux

    def value_and_grad(self, params, domain, *args):
        loss, aux, grads = 0.0, dict(), None
        for f in self.funcs:
            (tl, ta), g = f.value_and_grad(params, domain, *self._args_for(f, args))
            loss = loss + tl
            aux.update(ta)
            grads = g if grads is None else grads + g
        if self.with_props:
            aux.update({"props": params.properties()})
        return (loss, aux), grads


class UserDefinedLossFunction(BaseLossFunction):
    """loss_functions/utils.py:30-34: ``func(params, problem, *args) -> (loss, aux)`` written by the user.  Every shipped
    example builds it from the library's own loss functions (examples/inverse_problems/mechanics/path-dependent/
    script.py:121-127: ``loss_1 + loss_2`` of an EnergyResidualAndReactionLoss and a FullFieldDataLoss), and that is what
    runs here: the function is executed, the built-in losses inside it run as the fused CUDA calls and hand back values
    that carry their gradients (TracedLoss), and the arithmetic between them is differentiated by the chain rule.  A
    function that computes its loss some other way (from params / problem directly) has no gradient on this path and
    is refused -- there is no tracing compiler and no CPU fallback behind it."""

    def __init__(self, func):
        self.func = func

    def __call__(self, params, problem, *args):
        return self.func(params, problem, *args)

    def value_and_grad(self, params, problem, *args):
        if _TRACING[0]:
            raise RuntimeError("UserDefinedLossFunction.value_and_grad is not re-entrant")
        _TRACING[0] = True
        try:
            out = self.func(params, problem, *args)
        finally:
            _TRACING[0] = False
        loss, aux = out if isinstance(out, tuple) else (out, {})
        if not isinstance(loss, TracedLoss):
            raise NotImplementedError("the user-defined loss must be an arithmetic combination (+ - * /) of the library's loss "
                                      "functions evaluated on (params, problem): anything else has no gradient on the B200 path")
        return (loss.value, aux), loss.grads
