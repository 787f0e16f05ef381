"""Dirichlet wrappers g(x, t, z) shipped by pancax (pancax/bvps/*.py), as batched array
functions: x (N, nd), t float, z (N, ndof) -> (N, ndof).  They are evaluated once on the host at
the mesh nodes (Engine / tabulate_bc); all of them are diagonal-affine in z."""
from __future__ import annotations

import numpy as np

_AXIS = {"x": 0, "y": 1, "z": 2}


def UniaxialTensionLinearRamp(final_displacement: float, length: float, direction: str,
                              n_dimensions: int):
    """pancax/bvps/uniaxial_tension.py:3-101 (same argument names and errors)."""
    if n_dimensions == 2:
        if direction not in ("x", "y"):
            raise ValueError("Direction must be x or y for this BVP.")
    elif n_dimensions == 3:
        if direction not in ("x", "y", "z"):
            raise ValueError("Direction must be x, y, or z for this BVP.")
    else:
        raise ValueError("Dimensions can be either 2 or 3.")
    d = _AXIS[direction]

    def bc_func(xs, t, nn):
        c = xs[:, d]
        out = np.empty_like(nn)
        for i in range(n_dimensions):
            out[:, i] = c * (c - length) * t * nn[:, i] / length ** 2
        out[:, d] += c * t * final_displacement / length
        return out

    return bc_func


def SimpleShearLinearRamp(final_displacement: float, length: float, direction: str):
    """pancax/bvps/simple_shear.py:1-22."""
    if direction != "xy":
        raise ValueError("Direction must be x or y for this BVP.")

    def bc_func(xs, t, z):
        x = xs[:, 0]
        out = np.empty_like(z)
        out[:, 0] = x * (x - length) * t * z[:, 0] / length ** 2
        out[:, 1] = x * t * final_displacement / length + x * (x - length) * t * z[:, 1] / length ** 2
        return out

    return bc_func


def BiaxialLinearRamp(final_displacement_x: float, final_displacement_y: float, length_x: float,
                      length_y: float):
    """pancax/bvps/biaxial_tension.py:1-45 (final_displacement_y is unused in the reference too)."""

    def bc_func(xs, t, z):
        x, y = xs[:, 0], xs[:, 1]
        s = x * (x - length_x) * t / length_x ** 2 * y * (y - length_y) * t / length_y ** 2
        out = np.empty_like(z)
        out[:, 0] = s * z[:, 0] + x * t * final_displacement_x / length_x
        out[:, 1] = s * z[:, 1]
        return out

    return bc_func
