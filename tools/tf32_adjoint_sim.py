"""CPU emulation of the TF32 adjoint's arithmetic (no GPU needed): which TF32 scheme holds the tolerance?

Runs the hand-written reverse pass of the Field MLP on the oracle's seeded cases with the products emulated in
  * single-pass TF32 (operands rounded to 10 mantissa bits, cvt.rna), with and without a hi/lo split of the weights,
  * 3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 intermediate results, fp32 activations)
and prints the relative error of every gradient leaf against torch.autograd in fp64.  Result recorded in
pancax_b200/csrc/pcx_mlp_tf32.cuh / DESIGN.md: single pass 1.5e-4 .. 1.6e-3 (fails 1e-4), 3xTF32 ~1e-7.

    python tools/tf32_adjoint_sim.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import losses as ol                      # noqa: E402
from oracle.field import mlp_forward, unflatten_mlp   # noqa: E402
from tests import cases                               # noqa: E402

F64 = torch.float64


def tf32(x):
    """fp32 -> tf32 with round-to-nearest, ties away (cvt.rna.tf32.f32), returned as float64 values"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32).astype(np.float64)


def f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


def mm_fp64(a, b):
    return a @ b


def mm_tf32(a, b):
    return tf32(a) @ tf32(b)


def mm_tf32_splitw(a, b):
    """chain products only: weights b split hi/lo (removes the systematic part), rows a single-pass"""
    bh = tf32(b)
    return tf32(a) @ bh + tf32(a) @ tf32(b - bh)


def mm_3xtf32(a, b):
    a, b = f32(a), f32(b)
    ah, bh = tf32(a), tf32(b)
    al, bl = tf32(a - ah), tf32(b - bh)
    return f32(f32(al @ bh) + f32(ah @ bl) + f32(ah @ bh))


def cotangent_and_inputs(case):
    """rows (time step, node): MLP inputs X and the cotangent ZB on the network output of EnergyLoss"""
    prob = cases.oracle_problem(case)
    w, pr = ol._leaves(prob, case.weights, case.prop_raw)
    fld, phys, pe = ol._build(prob, w, pr)
    coords = torch.as_tensor(prob.coords, dtype=F64)
    conns = torch.as_tensor(prob.conns, dtype=torch.long)
    layers = unflatten_mlp(case.weights, *case.mlp)
    tl = [(torch.as_tensor(W), None if b is None else torch.as_tensor(b)) for W, b in layers]
    X, ZB = [], []
    for t in prob.times:
        inp = fld.inputs(coords, float(t)).detach()
        z = mlp_forward(tl, inp).requires_grad_(True)
        us = fld.dirichlet_bc_func(coords, float(t), z)
        pi = ol.potential_energy(phys, pe, coords, conns, us)
        (zb,) = torch.autograd.grad(cases.LOSS_WEIGHTS["energy_weight"] * pi, z)
        X.append(inp.numpy())
        ZB.append(zb.numpy())
    return layers, np.concatenate(X), np.concatenate(ZB)


def adjoint(layers, X, ZB, mm_chain, mm_contract, rh):
    """flat gradient in equinox leaf order; rh = storage precision of the hidden activations"""
    hs = [X]
    for W, b in layers[:-1]:
        hs.append(np.tanh(hs[-1] @ W.T + b))
    g = [None] * len(layers)
    d = ZB
    g[-1] = (mm_contract(d.T, rh(hs[-1])), None)
    d = mm_chain(d, layers[-1][0]) * (1 - rh(hs[-1]) ** 2)
    for L in range(len(layers) - 2, -1, -1):
        g[L] = (mm_contract(d.T, rh(hs[L]) if L > 0 else hs[L]),
                mm_contract(d.T, np.ones((d.shape[0], 1))).ravel())
        if L > 0:
            d = mm_chain(d, layers[L][0]) * (1 - rh(hs[L]) ** 2)
    out = []
    for a, b in g:
        out.append(a.ravel())
        if b is not None:
            out.append(b.ravel())
    return np.concatenate(out)


def run(name, n):
    case = cases.make_case(name, n=n, nt=2)
    _, _, gref, _ = cases.run_oracle_loss(case, "energy")
    layers, X, ZB = cotangent_and_inputs(case)
    ident = lambda x: x   # noqa: E731
    schemes = {"fp64 (hand adjoint vs autograd)": (mm_fp64, mm_fp64, ident),
               "1xTF32": (mm_tf32, mm_tf32, f32),
               "1xTF32 + hi/lo weights": (mm_tf32_splitw, mm_tf32, f32),
               "3xTF32": (mm_3xtf32, mm_3xtf32, f32)}
    sizes = [case.mlp[0]] + [case.mlp[2]] * case.mlp[3] + [case.mlp[1]]
    print(f"{name} n={n}: {X.shape[0]} rows")
    for label, (mc, mk, rh) in schemes.items():
        e = adjoint(layers, X, ZB, mc, mk, rh) - gref
        o, worst = 0, 0.0
        for i in range(len(sizes) - 1):
            for cnt in ([sizes[i] * sizes[i + 1]] + ([sizes[i + 1]] if i < len(sizes) - 2 else [])):
                sl = slice(o, o + cnt)
                o += cnt
                worst = max(worst, np.linalg.norm(e[sl]) / np.linalg.norm(gref[sl]))
        print(f"   {label:34s} whole vector {np.linalg.norm(e) / np.linalg.norm(gref):.2e}   worst leaf {worst:.2e}")


if __name__ == "__main__":
    run("hex8_neohookean", 10)
    run("hex8_gent", 14)
    run("quad4_neohookean", 40)
