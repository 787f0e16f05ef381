"""Oracle (test infrastructure): Field + MLP forward, torch float64 on the CPU.

Restates pancax/networks/fields.py:79-101 (input normalisation, MLP, Dirichlet wrapper)
and pancax/networks/mlp.py:49-85, which wraps ``equinox.nn.MLP`` (equinox ~=0.13,
pyproject.toml:18; NOT vendored in /root/reference, so its published algorithm is
restated): ``depth`` hidden Linear layers ``h = act(W h + b)`` with ``W (out, in)``,
``b (out,)``, followed by a final Linear (bias only if ``use_final_bias``) and the
identity final activation.  equinox.nn.Linear initialises W and b ~ U(-1/sqrt(in), 1/sqrt(in)).
"""
from __future__ import annotations

import math

import numpy as np
import torch

F64 = torch.float64


def init_mlp(n_in, n_out, width, depth, seed=0, use_final_bias=False):
    """Weights as a list of (W, b|None), drawn from numpy.random.default_rng(seed) so the
    same arrays can be injected into the oracle, the kernels and (elsewhere) pancax."""
    rng = np.random.default_rng(seed)
    sizes = [n_in] + [width] * depth + [n_out]
    layers = []
    for i in range(len(sizes) - 1):
        fan_in, fan_out = sizes[i], sizes[i + 1]
        lim = 1.0 / math.sqrt(fan_in)
        W = rng.uniform(-lim, lim, size=(fan_out, fan_in))
        last = i == len(sizes) - 2
        b = None if (last and not use_final_bias) else rng.uniform(-lim, lim, size=(fan_out,))
        layers.append((W, b))
    return layers


def flatten_mlp(layers) -> np.ndarray:
    """Flat f64 vector in equinox leaf order [W0, b0, W1, b1, ..., Wout(, bout)]."""
    out = []
    for W, b in layers:
        out.append(np.asarray(W, dtype=np.float64).ravel())
        if b is not None:
            out.append(np.asarray(b, dtype=np.float64).ravel())
    return np.concatenate(out)


def unflatten_mlp(flat, n_in, n_out, width, depth, use_final_bias=False):
    sizes = [n_in] + [width] * depth + [n_out]
    layers, o = [], 0
    for i in range(len(sizes) - 1):
        fi, fo = sizes[i], sizes[i + 1]
        W = flat[o:o + fi * fo].reshape(fo, fi)
        o += fi * fo
        last = i == len(sizes) - 2
        if last and not use_final_bias:
            b = None
        else:
            b = flat[o:o + fo]
            o += fo
        layers.append((W, b))
    assert o == flat.shape[0]
    return layers


def mlp_forward(layers, h, activation=torch.tanh):
    """equinox.nn.MLP.__call__ batched over rows of ``h`` (N, n_in)."""
    n = len(layers)
    for i, (W, b) in enumerate(layers):
        h = h @ W.T
        if b is not None:
            h = h + b
        if i < n - 1:
            h = activation(h)
    return h


class Field:
    """pancax/networks/fields.py:10-101 for ``seperate_networks=False``."""

    def __init__(self, coords, times, layers, dirichlet_bc_func=None):
        coords = torch.as_tensor(coords, dtype=F64)
        times = torch.as_tensor(times, dtype=F64)
        self.layers = layers  # list of (torch W, torch b|None)
        self.x_mins = coords.min(dim=0).values           # fields.py:70-71
        self.x_maxs = coords.max(dim=0).values
        self.t_min = times.min()                         # fields.py:68-69
        self.t_max = times.max()
        # fields.py:73-76  (jnp.allclose default rtol=1e-5, atol=1e-8)
        self.normalize_time = not bool(torch.isclose(self.t_min, self.t_max,
                                                     rtol=1e-5, atol=1e-8))
        self.dirichlet_bc_func = dirichlet_bc_func or (lambda x, t, z: z)

    def inputs(self, x, t):
        x_norm = (x - self.x_mins) / (self.x_maxs - self.x_mins)      # fields.py:80
        t = torch.as_tensor(t, dtype=F64)
        t_norm = (t - self.t_min) / (self.t_max - self.t_min) if self.normalize_time else t
        t_col = t_norm.expand(x.shape[0]).unsqueeze(-1) if t_norm.ndim == 0 \
            else t_norm.reshape(-1, 1)
        return torch.cat([x_norm, t_col], dim=1)                       # fields.py:88

    def __call__(self, x, t):
        """x (N, nd), t scalar or (N,) -> u (N, ndof).  (vmap of fields.py:79-101, i.e.
        pancax/physics_kernels/base.py:248-250.)"""
        z = mlp_forward(self.layers, self.inputs(x, t))
        return self.dirichlet_bc_func(x, t, z)
