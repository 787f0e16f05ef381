"""Oracle (test infrastructure): the shipped Dirichlet wrappers g(x, t, z), batched torch.

x (N, nd), t scalar or (N,), z (N, ndof) -> u (N, ndof).  Every shipped wrapper is
diagonal-affine in z.
"""
from __future__ import annotations

import torch


def _t(t, x):
    t = torch.as_tensor(t, dtype=x.dtype)
    return t if t.ndim == 0 else t.reshape(-1)


_AXIS = {"x": 0, "y": 1, "z": 2}


def UniaxialTensionLinearRamp(final_displacement, length, direction, n_dimensions):
    """pancax/bvps/uniaxial_tension.py:3-101."""
    if n_dimensions == 2 and direction not in ("x", "y"):
        raise ValueError("Direction must be x or y for this BVP.")
    if n_dimensions == 3 and direction not in ("x", "y", "z"):
        raise ValueError("Direction must be x, y, or z for this BVP.")
    if n_dimensions not in (2, 3):
        raise ValueError("Dimensions can be either 2 or 3.")
    d = _AXIS[direction]

    def bc_func(xs, t, nn):
        t = _t(t, xs)
        c = xs[:, d]
        cols = []
        for i in range(n_dimensions):
            ui = c * (c - length) * t * nn[:, i] / length ** 2
            if i == d:
                ui = c * t * final_displacement / length + ui
            cols.append(ui)
        return torch.stack(cols, dim=1)

    return bc_func


def SimpleShearLinearRamp(final_displacement, length, direction):
    """pancax/bvps/simple_shear.py:1-22."""
    if direction != "xy":
        raise ValueError("Direction must be x or y for this BVP.")

    def bc_func(xs, t, z):
        t = _t(t, xs)
        x = xs[:, 0]
        u0 = x * (x - length) * t * z[:, 0] / length ** 2
        u1 = x * t * final_displacement / length + x * (x - length) * t * z[:, 1] / length ** 2
        return torch.stack([u0, u1], dim=1)

    return bc_func


def BiaxialLinearRamp(final_displacement_x, final_displacement_y, length_x, length_y):
    """pancax/bvps/biaxial_tension.py:1-45 (note: final_displacement_y is unused there)."""

    def bc_func(xs, t, z):
        t = _t(t, xs)
        x, y = xs[:, 0], xs[:, 1]
        s = (x * (x - length_x) * t / length_x ** 2 * y * (y - length_y) * t / length_y ** 2)
        u0 = s * z[:, 0] + x * t * final_displacement_x / length_x
        u1 = s * z[:, 1]
        return torch.stack([u0, u1], dim=1)

    return bc_func


def poisson_example_bc(xs, t, z):
    """examples/forward_problems/poisson/variational_example.py:30-32."""
    x, y = xs[:, 0:1], xs[:, 1:2]
    return x * (1.0 - x) * y * (1.0 - y) * z


def poisson_example_source(x):
    """examples/forward_problems/poisson/variational_example.py:25."""
    import math
    return 2 * math.pi ** 2 * torch.sin(2.0 * math.pi * x[..., 0]) * torch.sin(2.0 * math.pi * x[..., 1])
