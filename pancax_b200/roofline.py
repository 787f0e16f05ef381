"""Algorithmic work of the hot path per unit, as formulas of the problem shape -- what
``bench.py`` divides by the measured kernel times.  Nothing here is measured; peaks come from
MEASURED_PEAKS.json / pcx_measure_dmma_peak at run time.

Notation: P_w = sum_l in_l * out_l of the MLP, rho = nodes per quadrature point."""
from __future__ import annotations

NNPE = {"hex8": 8, "quad4": 4, "tri3": 3}
NDIM = {"hex8": 3, "quad4": 2, "tri3": 2}


def mlp_pw(n_in: int, n_out: int, width: int, depth: int) -> int:
    return n_in * width + (depth - 1) * width * width + width * n_out


def mlp_forward_flops_per_row(n_in, n_out, width, depth) -> int:
    """2 P_w multiply-adds per (time step, node) row; the depth*width tanh are counted separately."""
    return 2 * mlp_pw(n_in, n_out, width, depth)


def mlp_backward_flops_per_row(n_in, n_out, width, depth) -> int:
    """Delta chain (2 P_w minus the input gradient of layer 0, which is not needed) + weight gradients
    (2 P_w): 4 P_w - 2 in_0 width."""
    return 4 * mlp_pw(n_in, n_out, width, depth) - 2 * n_in * width


def mlp_tanh_per_row(width, depth) -> int:
    return width * depth


def mlp_padded_dmma_per_tile(width, depth, n_in=4, n_out=3):
    """DMMA.8x8x4 instructions the kernels actually issue per 8-row forward tile / per 64-row adjoint tile:
    width pads to NT 8-wide slots tiles (incl. the bias slot) and KS 4-deep k-steps."""
    need_nt = (width + 1 + 7) // 8
    nt = 2 if need_nt <= 2 else 4 if need_nt <= 4 else 7 if need_nt <= 7 else 8
    ks = (width + 3) // 4
    fwd = nt + (depth - 1) * ks * nt + ks
    chain = nt + (depth - 1) * ks * nt                     # per warp (8 rows)
    dw = (depth - 1) * nt * nt * 16 + nt * 16 + nt * 16    # per 64-row tile: hidden + output + layer-0 blocks
    return dict(NT=nt, KS=ks, forward_per_8_rows=fwd, adjoint_per_64_rows=8 * chain + dw)


def activation_bytes_per_row(width, depth) -> int:
    """Stored hidden activations (fragment layout, padded to NT tiles of 8 doubles)."""
    need_nt = (width + 1 + 7) // 8
    nt = 2 if need_nt <= 2 else 4 if need_nt <= 4 else 7 if need_nt <= 7 else 8
    return depth * nt * 8 * 8


def element_flops_per_qp(elem: str, ndof: int, tangent: bool = False) -> int:
    """J (nnpe nd^2 FMA) + grad u (nnpe ndof nd) + force scatter (nnpe ndof nd) + inverse / H / Q products +
    constitutive update (~150 flop NeoHookean incl. log / rcbrt at their FMA-equivalent cost).  ~800 flop for
    Hex8 energy+force, ~1300 with the dual-number tangent."""
    nnpe, nd = NNPE[elem], NDIM[elem]
    fma = nnpe * nd * nd + 2 * nnpe * ndof * nd + 2 * ndof * nd * nd + 3 * nd * nd
    base = 2 * fma + 200
    return int(base * (1.65 if tangent else 1.0))


def element_bytes_per_qp(elem: str, ndof: int, nq: int, nodes_per_element: float = 1.0) -> float:
    """conn 4 B/node-slot + (coords + u + 2 x f) amortised over the element's quadrature points, geometry
    recomputed.  ~16 B (Hex8), ~28 B (Quad4)."""
    nnpe, nd = NNPE[elem], NDIM[elem]
    per_element = 4 * nnpe + nodes_per_element * 8 * (nd + ndof + 2 * ndof)
    return per_element / nq


def step_flops_per_qp_eval(elem, n_in, n_out, width, depth, nodes_per_qp, residual: bool = False) -> float:
    """Total algorithmic flop per quadrature-point evaluation of one loss+gradient step."""
    mlp = (mlp_forward_flops_per_row(n_in, n_out, width, depth) +
           mlp_backward_flops_per_row(n_in, n_out, width, depth)) * nodes_per_qp
    el = element_flops_per_qp(elem, n_out) + (element_flops_per_qp(elem, n_out, True) if residual else 0)
    return mlp + el
