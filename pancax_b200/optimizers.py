"""pancax/optimizers.py: Adam with init/step (:44-68,142-180).  The update itself is one CUDA
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
        w, p = params.fields.networks.weights, params.prop_raw
        opt_state = dict(mu_w=torch.zeros_like(w), nu_w=torch.zeros_like(w),
                         mu_p=torch.zeros_like(p), nu_p=torch.zeros_like(p), count=0)
        return OptimizerState(opt_state, 0, filter_spec, False)

    def _schedule(self, count):
        # optax.exponential_decay (non-staircase): init * decay_rate ** (count / transition_steps)
        return self.learning_rate * self.decay_rate ** (count / self.transition_steps)

    def step(self, params, state: OptimizerState, problem, *args):
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
