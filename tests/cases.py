"""Shared problem factory for the parity tests, smoke() and bench.py: builds the same seeded
inputs for the CUDA engine (pancax_b200) and for the CPU oracle (oracle/)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

import numpy as np

from oracle import bvps as obvps
from oracle import fem as ofem
from oracle import losses as olosses
from oracle.field import flatten_mlp, init_mlp

SWANSON_PROPS = dict(bulk_modulus=10.0, A1=0.93074321, P1=-0.07673672, B1=0.0, Q1=0.5,
                     C1=0.10448312, R1=1.71691036, cutoff_strain=0.01)


@dataclass
class Case:
    name: str
    mesh: dict
    q_order: int
    times: np.ndarray
    kind: str
    ndof: int
    mlp: tuple                      # (n_in, n_out, width, depth)
    weights: np.ndarray
    model: str = None
    formulation: str = None
    props: list = dc_field(default_factory=list)     # floats or (min, max, raw)
    bc: tuple = None                # ("uniaxial", final_disp, length, direction) | ("poisson",) | None
    unknown_idx: np.ndarray = None
    reaction_nodes: np.ndarray = None
    reaction_dof: int = 0
    global_outputs: np.ndarray = None
    prony: tuple = None             # (moduli, relaxation_times): SimpleFeFv around `model` (names "..._visco_<model>")

    @property
    def prop_raw(self):
        return np.array([p[2] for p in self.props if isinstance(p, tuple)], dtype=np.float64)


def _prop_names(model):
    return olosses.models.MODELS[model][1]


PRONY = ([1.0, 0.5], [10.0, 1.0])   # Prony series of the visco cases: moduli, relaxation times


def make_case(name, n=4, nt=2, width=50, depth=4, seed=0, permute=True, q_order=2, t0=0.0) -> Case:
    if "_visco_" in name:            # hyperviscoelasticity/simple_fefv.py around the named equilibrium model
        c = make_case(name.replace("_visco_", "_"), n, nt, width, depth, seed, permute, q_order, t0)
        c.name, c.prony = name, PRONY
        if c.model == "neohookean" and not any(isinstance(p, tuple) for p in c.props):
            c.props = [10.0, 0.855]  # examples/forward_problems/mechanics/example_hyper_visco_3d.py:59-63
        return c
    times = np.linspace(t0, 1.0, nt)
    if name.startswith("hex8_"):
        model = name.split("_", 1)[1]
        mesh = ofem.structured_mesh("hex8", n, permute_seed=7 if permute else None)
        props = {"neohookean": [1000.0, 1.0], "gent": [0.833, 0.3846, 3.0], "blatz_ko": [0.3846],
                 "swanson": list(SWANSON_PROPS.values()), "hencky": [0.833, 0.3846]}[model]
        mlp = (4, 3, width, depth)
        c = Case(name, mesh, q_order, times, "solid", 3, mlp,
                 flatten_mlp(init_mlp(*mlp, seed=seed)), model, "three_dimensional", props,
                 bc=("uniaxial", 0.5 if model != "neohookean" else 1.0, 1.0, "y"))
        bcs = [("nset_3", 0), ("nset_3", 1), ("nset_3", 2), ("nset_5", 0), ("nset_5", 1), ("nset_5", 2)]
        c.unknown_idx = ofem.unknown_indices(mesh["coords"].shape[0], 3, mesh["nodesets"], bcs)
        c.reaction_nodes = mesh["nodesets"]["nset_5"]
        c.reaction_dof = 1
        c.global_outputs = np.linspace(0.0, 0.3, nt)
        return c
    if name.startswith("quad4_") or name.startswith("tri3_"):
        elem, model = name.split("_", 1)
        mesh = ofem.structured_mesh(elem, n, permute_seed=11 if permute else None)
        trainable = model.endswith("_bounded")
        model = model.replace("_bounded", "")
        formulation = "plane_stress" if model.endswith("_pstress") else "plane_strain"   # solid_mechanics.py:20-71
        model = model.replace("_pstress", "")
        props = {"neohookean": [1000.0, 1.0], "gent": [0.833, 0.3846, 3.0], "blatz_ko": [0.3846],
                 "swanson": list(SWANSON_PROPS.values()), "hencky": [0.833, 0.3846]}[model]
        if trainable:
            props = [0.833, (0.01, 5.0, 0.3)] if model == "neohookean" else [(0.01, 5.0, 0.3)] if model == "blatz_ko" \
                else [(0.5, 2.0, -0.2), (0.01, 5.0, 0.3)] + props[2:]
        mlp = (3, 2, width, depth)
        c = Case(name, mesh, q_order, times, "solid", 2, mlp,
                 flatten_mlp(init_mlp(*mlp, seed=seed)), model, formulation, props,
                 bc=("uniaxial", 0.5, 1.0, "y"))
        bcs = [("nset_1", 0), ("nset_1", 1), ("nset_3", 0), ("nset_3", 1)]
        c.unknown_idx = ofem.unknown_indices(mesh["coords"].shape[0], 2, mesh["nodesets"], bcs)
        c.reaction_nodes = mesh["nodesets"]["nset_1"]
        c.reaction_dof = 1
        c.global_outputs = np.linspace(0.0, 0.2, nt)
        return c
    if name.startswith("poisson_"):
        elem = name.split("_", 1)[1]
        mesh = ofem.structured_mesh(elem, n, permute_seed=13 if permute else None)
        nd = mesh["coords"].shape[1]
        mlp = (nd + 1, 1, width, depth)
        c = Case(name, mesh, q_order, times, "poisson", 1, mlp,
                 flatten_mlp(init_mlp(*mlp, seed=seed)), bc=("poisson",))
        return c
    raise ValueError(name)


# ------------------------------------------------------------------ oracle side
def oracle_bc(case):
    if case.bc is None:
        return None
    if case.bc[0] == "uniaxial":
        nd = case.mesh["coords"].shape[1]
        return obvps.UniaxialTensionLinearRamp(case.bc[1], case.bc[2], case.bc[3], nd)
    if case.bc[0] == "poisson":
        if case.mesh["coords"].shape[1] == 2:
            return obvps.poisson_example_bc
        return lambda x, t, z: (x[:, 0:1] * (1 - x[:, 0:1]) * x[:, 1:2] * (1 - x[:, 1:2])
                                * x[:, 2:3] * (1 - x[:, 2:3])) * z
    raise ValueError(case.bc)


def oracle_problem(case) -> olosses.OracleProblem:
    props = []
    if case.kind == "solid":
        for nme, p in zip(_prop_names(case.model), case.props):
            if isinstance(p, tuple):
                props.append(olosses.PropSpec(nme, p[2], True, p[0], p[1]))
            else:
                props.append(olosses.PropSpec(nme, float(p)))
    m = case.mlp
    return olosses.OracleProblem(
        coords=case.mesh["coords"], conns=case.mesh["conns"], elem=case.mesh["elem"],
        q_order=case.q_order, times=case.times, kind=case.kind, ndof=case.ndof,
        mlp=olosses.MLPSpec(m[0], m[1], m[2], m[3]), model=case.model,
        formulation=case.formulation, props=props, bc_func=oracle_bc(case),
        source=obvps.poisson_example_source if case.kind == "poisson" else None,
        unknown_idx=case.unknown_idx, reaction_nodes=case.reaction_nodes,
        reaction_dof=case.reaction_dof, global_outputs=case.global_outputs, prony=case.prony)


LOSS_WEIGHTS = dict(energy_weight=1.3, residual_weight=0.7, reaction_weight=2.1)


def run_oracle_loss(case, kind, weights=None, **kw):
    lw = dict(LOSS_WEIGHTS)
    lw.update(kw)
    w = case.weights if weights is None else weights
    return olosses.loss_and_grad(oracle_problem(case), dict(kind=kind, **lw), w, case.prop_raw)


# ------------------------------------------------------------------ engine side
def engine_bc(case):
    from pancax_b200 import bvps
    if case.bc is None:
        return None
    if case.bc[0] == "uniaxial":
        nd = case.mesh["coords"].shape[1]
        return bvps.UniaxialTensionLinearRamp(case.bc[1], case.bc[2], case.bc[3], nd)
    if case.bc[0] == "poisson":
        def f(x, t, z):
            s = np.ones(x.shape[0])
            for j in range(x.shape[1]):
                s = s * x[:, j] * (1.0 - x[:, j])
            return s[:, None] * z
        return f
    raise ValueError(case.bc)


def poisson_source_np(xq):
    return 2 * math.pi ** 2 * np.sin(2.0 * math.pi * xq[..., 0]) * np.sin(2.0 * math.pi * xq[..., 1])


def make_engine(case, deterministic=True):
    from pancax_b200 import Engine, tabulate_bc
    eng = Engine(case.mesh["coords"], case.mesh["conns"], case.mesh["elem"], case.q_order, case.ndof,
                 case.times, unknown_idx=case.unknown_idx, reaction_nodes=case.reaction_nodes,
                 reaction_dof=case.reaction_dof, deterministic=deterministic)
    if case.kind == "poisson":
        _, _, xq = eng.element_geometry(want_grad_N=False, want_JxW=False)
        eng.set_physics("poisson", source_q=poisson_source_np(xq.cpu().numpy()))
    else:
        props = [(p[0], p[1]) if isinstance(p, tuple) else p for p in case.props]
        eng.set_physics("solid", case.model, case.formulation, props, prony=case.prony)
    a = b = None
    bc = engine_bc(case)
    if bc is not None:
        a, b = tabulate_bc(bc, case.mesh["coords"], case.times, case.ndof)
    eng.set_field(*case.mlp, bc_a=a, bc_b=b)
    return eng


def run_engine_loss(case, kind, weights=None, deterministic=True, eng=None, **kw):
    lw = dict(LOSS_WEIGHTS)
    lw.update(kw)
    own = eng is None
    if own:
        eng = make_engine(case, deterministic)
    w = case.weights if weights is None else weights
    loss, aux, gw, gp = eng.loss_and_grad(kind, w, case.prop_raw if eng.n_train else None,
                                          global_outputs=case.global_outputs, **lw)
    out = (float(loss.cpu()[0]), aux.cpu().numpy(), gw.cpu().numpy(), gp.cpu().numpy())
    if own:
        eng.close()
    return out


def make_shard_engines(case, n_shards, deterministic=True, scramble=False):
    """The same problem cut into ``n_shards`` contiguous element ranges (each with the closure nodes of
    its elements, renumbered locally) -> (engines, plans).  ``scramble`` permutes the element order first
    so that shards interleave in space (many shared nodes, nodes shared by more than two shards)."""
    from pancax_b200 import Engine, parallel, tabulate_bc
    coords, conns = case.mesh["coords"], case.mesh["conns"]
    if scramble:
        conns = conns[np.random.default_rng(5).permutation(conns.shape[0])]
    ne = conns.shape[0]
    subs = []
    for r in range(n_shards):
        lo, hi = (ne * r) // n_shards, (ne * (r + 1)) // n_shards
        nodes, inv = np.unique(conns[lo:hi], return_inverse=True)
        subs.append((nodes, inv.reshape(-1, conns.shape[1]).astype(np.int32)))
    plans = parallel.make_shard_plans([s[0] for s in subs])
    engines = []
    bc = engine_bc(case)
    for (nodes, lconns), plan in zip(subs, plans):
        g2l = -np.ones(coords.shape[0], dtype=np.int64)
        g2l[nodes] = np.arange(nodes.size)
        unk = rn = None
        if case.unknown_idx is not None:
            gn, comp = case.unknown_idx // case.ndof, case.unknown_idx % case.ndof
            keep = g2l[gn] >= 0
            unk = g2l[gn[keep]] * case.ndof + comp[keep]
        if case.reaction_nodes is not None:
            loc = g2l[case.reaction_nodes]
            rn = loc[loc >= 0]
        unk, n_uo, rn, n_ro = plan.order_owned_first(unk, rn, case.ndof)
        eng = Engine(coords[nodes], lconns, case.mesh["elem"], case.q_order, case.ndof, case.times,
                     unknown_idx=unk, reaction_nodes=rn, reaction_dof=case.reaction_dof,
                     deterministic=deterministic)
        eng.set_ownership(n_uo, n_ro)
        props = [(p[0], p[1]) if isinstance(p, tuple) else p for p in case.props]
        eng.set_physics("solid", case.model, case.formulation, props, prony=case.prony)
        a, b = tabulate_bc(bc, coords[nodes], case.times, case.ndof)
        eng.set_field(*case.mlp, bc_a=a, bc_b=b, x_min=coords.min(0), x_max=coords.max(0))   # GLOBAL bounds
        engines.append(eng)
    return engines, plans
