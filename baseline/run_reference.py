#!/usr/bin/env python
"""baseline/run_reference.py -- run the REAL pancax (jax + equinox + optax) on the seeded parity cases of this repository
and dump what its own eqx.filter_value_and_grad(loss.filtered_loss) returns (last row).

STATUS: source only in this repository's build image -- jax, equinox, optax and netCDF4 are not installed and there is no
network, so this script has never run here (it is syntax-checked by tests/test_integration_sources.py).  On a box with
`pip install pancax` (or PYTHONPATH=/path/to/pancax):

    JAX_ENABLE_X64=1 JAX_PLATFORMS=cpu python baseline/run_reference.py --out tests/golden/loss_grad_jax.npz
    JAX_ENABLE_X64=1 JAX_PLATFORMS=cpu python baseline/run_reference.py --time --n 32      # CPU arm of bench.py's metric

The dump uses the SAME keys as tests/golden/loss_grad.npz (produced by executing the reference's source on a torch
stand-in, oracle/gen_golden_ad.py); tests/test_jax_dump.py compares the two files, the oracle and the CUDA path against it
whenever tests/golden/loss_grad_jax.npz exists.  That closes the one gap the stand-in leaves: equinox.nn.MLP, optax and
XLA's own det / inv / reduction order are then the real ones.

Weights are INJECTED (tests/cases.py: numpy seeds, equinox leaf order), never re-drawn: jax PRNG streams cannot be
reproduced without jax.  Meshes are written as Exodus II files with pancax_b200.fem.write_exodus_mesh (NetCDF-3, the
subset pancax/fem/read_exodus_mesh.py reads) into a temporary directory.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_pancax_problem(case, workdir):
    """The reference objects of one tests/cases.py Case: VariationalDomain on the written mesh, physics, Dirichlet
    wrapper, ForwardProblem / InverseProblem, Parameters with the case's weights injected."""
    import equinox as eqx
    import jax
    import jax.numpy as jnp
    import pancax as px
    from pancax_b200.fem import Mesh, write_exodus_mesh
    mesh = Mesh(case.mesh["coords"], case.mesh["conns"].astype(np.int32), case.mesh["elem"],
                {k: np.asarray(v, dtype=np.int32) for k, v in case.mesh["nodesets"].items()})
    path = os.path.join(workdir, f"{case.name}.g")
    write_exodus_mesh(path, mesh)
    times = jnp.asarray(case.times)
    domain = px.VariationalDomain(path, times, q_order=case.q_order)
    if case.kind == "poisson":
        physics = px.Poisson(lambda x: 2 * jnp.pi ** 2 * jnp.sin(2.0 * jnp.pi * x[0]) * jnp.sin(2.0 * jnp.pi * x[1]))

        def bc(x, t, z):
            return x[0] * (1.0 - x[0]) * x[1] * (1.0 - x[1]) * z
        dbcs = []
    else:
        names = {"neohookean": ("NeoHookean", ("bulk_modulus", "shear_modulus")),
                 "gent": ("Gent", ("bulk_modulus", "shear_modulus", "Jm_parameter")),
                 "swanson": ("Swanson", ("bulk_modulus", "A1", "P1", "B1", "Q1", "C1", "R1", "cutoff_strain")),
                 "hencky": ("Hencky", ("bulk_modulus", "shear_modulus")), "blatz_ko": ("BlatzKo", ("shear_modulus",))}
        cls, fields = names[case.model]
        kw = {}
        for name, p in zip(fields, case.props):
            if isinstance(p, tuple):          # trainable: BoundedProperty with the raw value injected below
                kw[name] = px.BoundedProperty(p[0], p[1], key=jax.random.PRNGKey(0))
            else:
                kw[name] = float(p)
        model = getattr(px, cls)(**kw)
        for name, p in zip(fields, case.props):
            if isinstance(p, tuple):
                model = eqx.tree_at(lambda m, n=name: getattr(m, n).prop_val, model, jnp.asarray([p[2]]))
        if case.prony is not None:
            model = px.SimpleFeFv(model, px.PronySeries(moduli=jnp.asarray(case.prony[0]), relaxation_times=jnp.asarray(case.prony[1])),
                                  px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
        form = {"three_dimensional": px.ThreeDimensional, "plane_strain": px.PlaneStrain, "plane_stress": px.PlaneStress}[case.formulation]()
        physics = px.SolidMechanics(model, form)
        nd = case.mesh["coords"].shape[1]
        bc = px.UniaxialTensionLinearRamp(final_displacement=case.bc[1], length=case.bc[2], direction=case.bc[3], n_dimensions=nd)
        sets = ("nset_3", "nset_5") if nd == 3 else ("nset_1", "nset_3")
        dbcs = [px.DirichletBC(s, c) for s in sets for c in range(case.ndof)]
    problem = px.ForwardProblem(domain, physics, [], dbcs, [])
    if case.reaction_nodes is not None:
        gd = type("GD", (), dict(outputs=jnp.asarray(case.global_outputs), reaction_nodes=jnp.asarray(case.reaction_nodes),
                                 reaction_dof=int(case.reaction_dof)))()
        problem = eqx.tree_at(lambda p: p.global_data, problem, gd, is_leaf=lambda x: x is None) \
            if hasattr(problem, "global_data") else problem
    n_in, n_out, width, depth = case.mlp
    params = px.Parameters(problem, jax.random.PRNGKey(0), dirichlet_bc_func=bc, n_layers=depth, n_neurons=width)
    net = params.fields.networks
    leaves, treedef = jax.tree_util.tree_flatten(net)
    out, off = [], 0
    for leaf in leaves:                      # equinox leaf order: W0 (out, in), b0, W1, b1, ..., Wout
        if hasattr(leaf, "dtype") and jnp.issubdtype(leaf.dtype, jnp.inexact):
            out.append(jnp.asarray(case.weights[off:off + leaf.size]).reshape(leaf.shape))
            off += leaf.size
        else:
            out.append(leaf)
    assert off == case.weights.size, (off, case.weights.size)
    params = eqx.tree_at(lambda p: p.fields.networks, params, jax.tree_util.tree_unflatten(treedef, out))
    return problem, params


def loss_for(kind, lw):
    import pancax as px
    if kind == "energy":
        return px.EnergyLoss(weight=lw["energy_weight"])
    if kind == "energy_residual":
        return px.EnergyAndResidualLoss(energy_weight=lw["energy_weight"], residual_weight=lw["residual_weight"])
    return px.EnergyResidualAndReactionLoss(energy_weight=lw["energy_weight"], residual_weight=lw["residual_weight"],
                                            reaction_weight=lw["reaction_weight"])


def value_and_grad(loss_fn, kind, params, problem):
    """What AbstractOptimizer._single_step differentiates (pancax/optimizers.py:77-88)."""
    import equinox as eqx
    import jax
    import jax.numpy as jnp
    call = loss_fn.path_independent_call if kind == "energy_residual" else loss_fn      # __call__ drops the value (:145-149)
    (loss, aux), grads = eqx.filter_jit(eqx.filter_value_and_grad(lambda p: call(p, problem), has_aux=True))(params)
    gw = jnp.concatenate([g.ravel() for g in jax.tree_util.tree_leaves(grads.fields.networks)])
    gp = [g.ravel() for g in jax.tree_util.tree_leaves(grads.physics)]
    return loss, aux, np.asarray(gw), np.concatenate([np.asarray(g) for g in gp]) if gp else np.zeros(0)


def dump(out_path):
    from oracle.gen_golden_ad import CASES, T0
    from tests.cases import LOSS_WEIGHTS, make_case
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        for name, n, nt, width, depth, q, kinds in CASES:
            case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
            tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
            problem, params = build_pancax_problem(case, wd)
            for kind in kinds:
                loss, aux, gw, gp = value_and_grad(loss_for(kind, LOSS_WEIGHTS), kind, params, problem)
                out[f"{tag}__{kind}__loss"], out[f"{tag}__{kind}__gw"], out[f"{tag}__{kind}__gp"] = np.asarray(loss), gw, gp
                for k, v in aux.items():
                    out[f"{tag}__{kind}__aux_{k}"] = np.asarray(v)
                print(f"{tag:48s} {kind:26s} loss {float(loss): .12e} |gw| {np.linalg.norm(gw):.6e}", flush=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


def timeit(n, nt, steps):
    """CPU arm of BASELINE.json's metric on the real pancax: loss+grad evals/s of config C3's physics on Hex8 n^3."""
    import jax
    from tests.cases import make_case
    case = make_case("hex8_neohookean", n=n, nt=nt, width=50, depth=4, permute=False)
    with tempfile.TemporaryDirectory() as wd:
        problem, params = build_pancax_problem(case, wd)
        fn = loss_for("energy", dict(energy_weight=1.0))
        value_and_grad(fn, "energy", params, problem)          # compile
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            loss, _, gw, _ = value_and_grad(fn, "energy", params, problem)
            jax.block_until_ready(loss)
            ts.append(time.perf_counter() - t0)
    evals = n ** 3 * 8 * nt
    print(json.dumps({"impl": "pancax/jax", "jax": jax.__version__, "backend": jax.default_backend(), "n": n, "nt": nt,
                      "evals_per_step": evals, "median_s": float(np.median(ts)), "evals_per_s": evals / float(np.median(ts)),
                      "host_cpus": os.cpu_count()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "loss_grad_jax.npz"))
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--nt", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    os.environ.setdefault("JAX_ENABLE_X64", "1")
    import jax
    jax.config.update("jax_enable_x64", True)
    if args.time:
        timeit(args.n, args.nt, args.steps)
    else:
        dump(args.out)


if __name__ == "__main__":
    main()
