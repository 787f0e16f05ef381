"""pancax/optimizers.py: Adam with init/step (:44-68,142-180) and LBFGS (:189-417).  The update itself is one CUDA
kernel per leaf (pcx_adam_step) reproducing optax.chain(scale_by_adam(), scale_by_schedule(
exponential_decay(lr, transition_steps, decay_rate)), scale(-1)) as the reference composes it."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Optional

import torch

from . import _lib
from ._lib import check
from .engine import _ptr, _stream


@dataclass
class OptimizerState:
    opt_state: Any
    epoch: int
    filter_spec: Any
    is_ensemble: bool


class Adam:
    def __init__(self, loss_function, learning_rate: float, clip_gradients: Optional[bool] = False,
                 decay_rate: Optional[float] = 0.99, has_aux: Optional[bool] = False,
                 jit: Optional[bool] = True, transition_steps: Optional[int] = 500) -> None:
        self.loss_function = loss_function
        self.learning_rate, self.decay_rate = float(learning_rate), float(decay_rate)
        self.transition_steps = int(transition_steps)
        self.clip_gradients, self.has_aux, self.jit = bool(clip_gradients), has_aux, jit
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8          # optax.scale_by_adam defaults

    def init(self, params, filter_spec: Optional[Any] = None):
        if filter_spec is None:
            filter_spec = params.freeze_physics_normalization_filter()
        if getattr(params, "is_ensemble", False):
            # optimizers.py:44-62: one optimiser state over the stacked leaves == one per member with a shared count
            return OptimizerState([self.init(m, filter_spec) for m in params.members], 0, filter_spec, True)
        w, p = params.fields.networks.weights, params.prop_raw
        opt_state = dict(mu_w=torch.zeros_like(w), nu_w=torch.zeros_like(w),
                         mu_p=torch.zeros_like(p), nu_p=torch.zeros_like(p), count=0,
                         count_dev=torch.zeros(1, dtype=torch.float64, device=w.device))
        return OptimizerState(opt_state, 0, filter_spec, False)

    def _schedule(self, count):
        # optax.exponential_decay (non-staircase): init * decay_rate ** (count / transition_steps)
        return self.learning_rate * self.decay_rate ** (count / self.transition_steps)

    def step(self, params, state: OptimizerState, problem, *args):
        if state.is_ensemble:
            return self.ensemble_step(params, state, problem, *args)
        return self.single_step(params, state, problem, *args)

    def ensemble_step(self, params, state: OptimizerState, problem, *args):
        """optimizers.py:109-139: eqx.filter_vmap(filter_value_and_grad(loss)) over the ensemble axis, then ONE
        optax update over the stacked leaves -- element-wise, so it equals the members' own updates with the
        shared step count.  The members run back to back on the same Engine (each launch already fills the GPU);
        losses / aux come back stacked along axis 0 like the vmapped reference returns them."""
        losses, auxs = [], []
        for m, st in zip(params.members, state.opt_state):
            _, _, out = self.single_step(m, st, problem, *args)
            loss, aux = out if self.has_aux else (out, {})
            losses.append(loss)
            auxs.append(aux)
        state.epoch += 1
        loss = torch.stack([torch.as_tensor(v) for v in losses])
        if not self.has_aux:
            return params, state, loss
        aux = {k: torch.stack([torch.as_tensor(a[k]) for a in auxs]) for k in auxs[0]
               if all(isinstance(a[k], torch.Tensor) for a in auxs)}
        return params, state, (loss, aux)

    def single_step(self, params, state: OptimizerState, problem, *args):
        (loss, aux), grads = self.loss_function.value_and_grad(params, problem, *args)
        st = state.opt_state
        gw, gp = grads.fields, grads.props
        if self.clip_gradients:   # optax.clip_by_global_norm(1.0)
            parts = [gw.flatten()] + ([gp.flatten()] if gp.numel() else [])
            gn = torch.linalg.vector_norm(torch.cat(parts))
            scale = torch.clamp(1.0 / gn, max=1.0)
            gw, gp = gw * scale, gp * scale
        lr = self._schedule(st["count"])
        st["count"] += 1
        lib = _lib.load()
        spec = state.filter_spec
        if spec.get("fields", True):
            w = params.fields.networks.weights
            check(lib.pcx_adam_step(_ptr(w), _ptr(gw), _ptr(st["mu_w"]), _ptr(st["nu_w"]), w.numel(),
                                    st["count"], lr, self.b1, self.b2, self.eps, _stream()))
        if spec.get("props", True) and params.prop_raw.numel():
            p = params.prop_raw
            check(lib.pcx_adam_step(_ptr(p), _ptr(gp), _ptr(st["mu_p"]), _ptr(st["nu_p"]), p.numel(),
                                    st["count"], lr, self.b1, self.b2, self.eps, _stream()))
        state.epoch += 1
        return params, state, ((loss, aux) if self.has_aux else loss)


    # ------------------------------------------------------------------ device-resident training step
    def device_step(self, params, state: OptimizerState, problem, *args):
        """Same update as ``step`` with the update counter and the learning-rate schedule on the device
        (pcx_adam_step_sched): nothing depends on host state, so the call can be captured in a CUDA graph.
        Gradient clipping is not available on this path."""
        if self.clip_gradients:
            raise NotImplementedError("clip_gradients is not available on the device-resident step")
        (loss, aux), grads = self.loss_function.value_and_grad(params, problem, *args)
        st = state.opt_state
        lib = _lib.load()
        spec = state.filter_spec
        w, p = params.fields.networks.weights, params.prop_raw
        do_w = spec.get("fields", True)
        do_p = spec.get("props", True) and p.numel() > 0
        sched = (self.learning_rate, self.decay_rate, float(self.transition_steps), self.b1, self.b2, self.eps)
        if do_w:
            check(lib.pcx_adam_step_sched(_ptr(w), _ptr(grads.fields), _ptr(st["mu_w"]), _ptr(st["nu_w"]), w.numel(),
                                          _ptr(st["count_dev"]), 0 if do_p else 1, *sched, _stream()))
        if do_p:
            check(lib.pcx_adam_step_sched(_ptr(p), _ptr(grads.props), _ptr(st["mu_p"]), _ptr(st["nu_p"]), p.numel(),
                                          _ptr(st["count_dev"]), 1, *sched, _stream()))
        if not (do_w or do_p):
            check(lib.pcx_adam_step_sched(None, None, None, None, 0, _ptr(st["count_dev"]), 1, *sched, _stream()))
        return (loss, aux)

    def capture(self, params, state: OptimizerState, problem, *args):
        """Capture ONE training step (loss + gradient + Adam update) in a CUDA graph.  Returns
        (graph, (loss, aux)): every ``graph.replay()`` advances the parameters by one step and refreshes the
        static ``loss`` / ``aux`` device tensors.  For launch-bound (small) problems."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        snap = [t.clone() for t in (params.fields.networks.weights, params.prop_raw, state.opt_state["mu_w"],
                                    state.opt_state["nu_w"], state.opt_state["mu_p"], state.opt_state["nu_p"],
                                    state.opt_state["count_dev"])]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            self.device_step(params, state, problem, *args)          # warm-up (builds the engine, allocates)
            s.synchronize()
            for t, v in zip((params.fields.networks.weights, params.prop_raw, state.opt_state["mu_w"],
                             state.opt_state["nu_w"], state.opt_state["mu_p"], state.opt_state["nu_p"],
                             state.opt_state["count_dev"]), snap):
                t.copy_(v)                                            # undo the warm-up step
            with torch.cuda.graph(g, stream=s):
                out = self.device_step(params, state, problem, *args)
        torch.cuda.current_stream().wait_stream(s)
        return g, out


class LBFGS:
    """pancax/optimizers.py:189-417: same constructor / init / step interface.  The reference delegates the update to
    ``optimistix.LBFGS(rtol, atol)`` (third-party, unpinned, not vendored: PARITY UNPINNED); what is restated here is the
    published algorithm it implements -- limited-memory BFGS direction by the two-loop recursion over the last
    ``history`` (s, y) pairs, backtracking Armijo line search (slope 0.1, step halving from 1), Cauchy termination
    ``|df| < atol + rtol |f|`` and ``||dx|| < atol + rtol ||x||`` -- with every vector operation on the device and the
    loss + gradient coming from the CUDA library.  One ``step`` = one accepted iterate.  Ensembles are refused like in
    the reference (:409-417)."""

    def __init__(self, loss_function, has_aux: Optional[bool] = False, jit: Optional[bool] = True,
                 rtol: Optional[float] = 1.0e-3, atol: Optional[float] = 1.0e-3, options=None, tags=None,
                 history: int = 10, max_backtracks: int = 20) -> None:
        self.loss_function, self.has_aux, self.jit = loss_function, has_aux, jit
        self.rtol, self.atol = float(rtol), float(atol)
        self.options = {} if options is None else options
        self.tags = frozenset() if tags is None else tags
        self.history, self.max_backtracks = int(history), int(max_backtracks)

    def init(self, params, filter_spec: Optional[Any] = None):
        if filter_spec is None:
            filter_spec = params.freeze_physics_normalization_filter()
        if getattr(params, "is_ensemble", False):
            raise NotImplementedError("Per-ensemble LBFGS is not implemented (optimizers.py:409-417)")
        return OptimizerState(dict(s=[], y=[], f=None, g=None, done=False), 0, filter_spec, False)

    # flat view of the trainable leaves
    def _leaves(self, params, spec):
        out = []
        if spec.get("fields", True):
            out.append(params.fields.networks.weights)
        if spec.get("props", True) and params.prop_raw.numel():
            out.append(params.prop_raw)
        return out

    def _eval(self, params, spec, problem, args):
        (loss, aux), grads = self.loss_function.value_and_grad(params, problem, *args)
        gs = []
        if spec.get("fields", True):
            gs.append(grads.fields.reshape(-1))
        if spec.get("props", True) and params.prop_raw.numel():
            gs.append(grads.props.reshape(-1)[:params.prop_raw.numel()])
        return loss, aux, torch.cat(gs).clone()

    def step(self, params, state: OptimizerState, problem, *args):
        st, spec = state.opt_state, state.filter_spec
        leaves = self._leaves(params, spec)
        x0 = torch.cat([v.reshape(-1) for v in leaves]).clone()

        def set_x(x):
            o = 0
            for v in leaves:
                v.copy_(x[o:o + v.numel()].reshape(v.shape))
                o += v.numel()
        if st["f"] is None:
            st["f"], st["aux"], st["g"] = self._eval(params, spec, problem, args)
        f0, g0 = st["f"], st["g"]
        # two-loop recursion: d = -H g
        q = g0.clone()
        alphas = []
        for s_k, y_k in zip(reversed(st["s"]), reversed(st["y"])):
            rho = 1.0 / torch.dot(y_k, s_k)
            a = rho * torch.dot(s_k, q)
            alphas.append((a, rho, s_k, y_k))
            q -= a * y_k
        if st["s"]:
            s_k, y_k = st["s"][-1], st["y"][-1]
            q *= torch.dot(s_k, y_k) / torch.dot(y_k, y_k)
        for a, rho, s_k, y_k in reversed(alphas):
            b = rho * torch.dot(y_k, q)
            q += (a - b) * s_k
        d = -q
        slope = torch.dot(g0, d)
        if not bool(slope < 0):          # not a descent direction (curvature pairs gone bad): restart from steepest descent
            st["s"], st["y"] = [], []
            d, slope = -g0, -torch.dot(g0, g0)
        t, accepted = 1.0 if st["s"] else float(1.0 / max(1.0, float(g0.norm()))), False
        for _ in range(self.max_backtracks):
            set_x(x0 + t * d)
            f1, aux1, g1 = self._eval(params, spec, problem, args)
            if bool(torch.isfinite(f1)) and bool(f1 <= f0 + 0.1 * t * slope):
                accepted = True
                break
            t *= 0.5
        if not accepted:
            set_x(x0)
            st["done"] = True
        else:
            s_new, y_new = t * d, g1 - g0
            if bool(torch.dot(s_new, y_new) > 1e-12 * s_new.norm() * y_new.norm()):
                st["s"].append(s_new)
                st["y"].append(y_new)
                if len(st["s"]) > self.history:
                    st["s"].pop(0)
                    st["y"].pop(0)
            df, dx = abs(float(f1 - f0)), float(s_new.norm())
            st["done"] = df < self.atol + self.rtol * abs(float(f1)) and dx < self.atol + self.rtol * float(x0.norm())
            st["f"], st["aux"], st["g"] = f1, aux1, g1
        state.epoch += 1
        loss = st["f"]
        return params, state, ((loss, st["aux"]) if self.has_aux else loss)
