"""Generates tests/golden/c4_vanilla.npz: config C4 of BASELINE.json LITERALLY -- the reference's own inverse-problem
example (examples/inverse_problems/mechanics/vanilla/example_v2.py:20-109): its TRI3 mesh `mesh/mesh.g` (902 elements,
492 nodes, q_order 2 -> 2 706 quadrature points), its DIC file `data/data_full_field.csv` (10 332 rows) and its global
force-displacement file `data/data_global_data.csv` (21 load steps), NeoHookean(K = 0.833, G = BoundedProperty(0.01, 5)),
PlaneStrain, UniaxialTensionLinearRamp(max u_x of the DIC data, 1.0, 'x'), Dirichlet BCs on nodeset_2 / nodeset_4,
loss = EnergyResidualAndReactionLoss(residual_weight=250e9, reaction_weight=250e9) + FullFieldDataLoss(weight=10e9) on the
first batch of 1 024 rows of FullFieldDataLoader.dataloader (pancax/data.py:97-105: np.random.permutation, seeded here).

The reference's OWN loss functions / physics kernels / Field / function space / constitutive model / Dirichlet wrapper
are executed on the torch stand-in for jax/equinox (oracle/refshim_torch.py); the mesh, the data and the injected
weights travel inside the .npz so the tests need nothing from /root/reference.  Not reproduced from example_v2.py: the
user-defined L1 regulariser of its Python loss (off the hot path) and the jax.random initialisation (weights and the raw
shear-modulus value are injected from numpy seeds).

Time step 0 of the load ramp has u = 0, hence f = 0 exactly: jnp.linalg.norm's gradient is NaN there under jax (SURVEY
F9); the torch stand-in returns the zero subgradient, which is the documented convention of the CUDA path and the oracle.

    python -m oracle.gen_golden_c4          # build container only (needs /root/reference)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EX = os.path.join(REF, "examples", "inverse_problems", "mechanics", "vanilla")
OUT = os.path.join(ROOT, "tests", "golden", "c4_vanilla.npz")
BATCH = 1024
G_RAW = 0.3          # injected raw value of the BoundedProperty (sigmoid argument)


def load_inputs():
    """Everything example_v2.py:20-66 builds from files, with numpy / pandas only."""
    import pandas
    sys.path.insert(0, ROOT)
    from oracle import fem as ofem
    from oracle.field import flatten_mlp, init_mlp
    mesh = ofem.read_exodus_mesh(os.path.join(EX, "mesh", "mesh.g"))
    ff = pandas.read_csv(os.path.join(EX, "data", "data_full_field.csv"))
    ff.columns = ff.columns.str.strip()
    ff_in, ff_out = ff[["x", "y", "t"]].values, ff[["u_x", "u_y"]].values                 # data.py:31-34
    gd = pandas.read_csv(os.path.join(EX, "data", "data_global_data.csv"))
    gd.columns = gd.columns.str.strip()
    times_in, forces_in = gd["t"].values, gd["f_x"].values
    g_out = np.interp(times_in, times_in, forces_in)                                      # data.py:181-183 (interpolate=False)
    times = np.linspace(0.0, 1.0, len(g_out))                                             # example_v2.py:36
    reaction_nodes = mesh["nodesets"]["nodeset_4"]                                        # data.py:222-225, nset_id = 4
    bcs = [("nodeset_2", 0), ("nodeset_2", 1), ("nodeset_4", 0), ("nodeset_4", 1)]        # example_v2.py:56-61
    unknown = ofem.unknown_indices(mesh["coords"].shape[0], 2, mesh["nodesets"], bcs)
    np.random.seed(0)
    perm = np.random.permutation(np.arange(ff_in.shape[0]))                               # data.py:98
    batch = perm[:BATCH]
    weights = flatten_mlp(init_mlp(3, 2, 50, 3, seed=0))                                  # Parameters defaults: 3 x 50
    return dict(coords=mesh["coords"], conns=mesh["conns"], times=times, global_outputs=g_out,
                reaction_nodes=reaction_nodes.astype(np.int32), reaction_dof=np.array(0), unknown_idx=unknown,
                nodeset_2=mesh["nodesets"]["nodeset_2"], nodeset_4=mesh["nodesets"]["nodeset_4"],
                batch_inputs=ff_in[batch], batch_outputs=ff_out[batch], final_displacement=np.array(np.max(ff_out[:, 0])),
                weights=weights, g_raw=np.array([G_RAW]), n_full_field_rows=np.array(ff_in.shape[0]))


def main():
    d = load_inputs()
    from oracle import refshim_torch as shim
    shim.import_reference(REF)
    from pancax.fem.elements import SimplexTriElement
    from pancax.fem.quadrature_rules import QuadratureRule
    from pancax.fem.function_space import NonAllocatedFunctionSpace
    from pancax.constitutive_models import NeoHookean, BoundedProperty
    from pancax.physics_kernels import SolidMechanics, PlaneStrain
    from pancax.bvps.uniaxial_tension import UniaxialTensionLinearRamp
    from pancax.loss_functions.weak_form_loss_functions import EnergyResidualAndReactionLoss
    from pancax.loss_functions.data_loss_functions import FullFieldDataLoss
    from pancax.networks.parameters import Parameters, _Parameters
    import pancax.networks.fields as ref_fields

    el = SimplexTriElement(1)
    space = NonAllocatedFunctionSpace(types.SimpleNamespace(parentElement=el), QuadratureRule(el, 2))
    coords = torch.as_tensor(d["coords"], dtype=torch.float64)
    conns = torch.as_tensor(d["conns"], dtype=torch.long)
    times = torch.as_tensor(d["times"], dtype=torch.float64)
    bp = BoundedProperty.__new__(BoundedProperty)                  # skip the jax.random draw: inject the raw value
    object.__setattr__(bp, "prop_min", 0.01)
    object.__setattr__(bp, "prop_max", 5.0)
    raw = torch.tensor([G_RAW], dtype=torch.float64, requires_grad=True)
    object.__setattr__(bp, "prop_val", raw)
    physics = SolidMechanics(NeoHookean(bulk_modulus=0.833, shear_modulus=bp), PlaneStrain())
    bc = UniaxialTensionLinearRamp(final_displacement=float(d["final_displacement"]), length=1.0, direction="x", n_dimensions=2)
    dofm = types.SimpleNamespace(unknownIndices=torch.as_tensor(d["unknown_idx"], dtype=torch.long))
    domain = types.SimpleNamespace(coords=coords, conns=conns, fspace=space, dof_manager=dofm, times=times)
    gd = types.SimpleNamespace(outputs=torch.as_tensor(d["global_outputs"], dtype=torch.float64),
                               reaction_nodes=torch.as_tensor(d["reaction_nodes"], dtype=torch.long), reaction_dof=0)
    problem = types.SimpleNamespace(coords=coords, times=times, domain=domain, physics=physics, global_data=gd,
                                    n_dims=2, n_dofs=2, conns=conns, fspace=space)
    fld = ref_fields.Field(problem, np.array([0, 0], dtype=np.uint32), dirichlet_bc_func=bc)      # defaults: 3 x 50
    params = Parameters.__new__(Parameters)
    inner = _Parameters.__new__(_Parameters)
    for k, v in (("fields", fld), ("physics", physics), ("state", None)):
        object.__setattr__(inner, k, v)
    for k, v in (("is_ensemble", False), ("n_ensemble", 1), ("parameters", inner)):
        object.__setattr__(params, k, v)
    leaves = shim.mlp_leaves(fld.networks.mlp)
    off = 0
    with torch.no_grad():
        for leaf in leaves:
            n = leaf.numel()
            leaf.copy_(torch.as_tensor(d["weights"][off:off + n]).reshape(leaf.shape))
            off += n
    assert off == d["weights"].size

    phys_loss = EnergyResidualAndReactionLoss(residual_weight=250.e9, reaction_weight=250.e9)     # example_v2.py:85-87
    data_loss = FullFieldDataLoss(weight=10.e9)                                                    # :88
    l1, a1 = phys_loss(params, problem)
    l2, a2 = data_loss(params, problem, torch.as_tensor(d["batch_inputs"]), torch.as_tensor(d["batch_outputs"]))
    out = dict(d)
    for tag, loss in (("physics", l1), ("data", l2), ("total", l1 + l2)):
        gs = torch.autograd.grad(loss, leaves + [raw], allow_unused=True, retain_graph=True)
        gw = torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(gs[:-1], leaves)])
        gp = gs[-1] if gs[-1] is not None else torch.zeros(1, dtype=torch.float64)
        out[f"{tag}__loss"], out[f"{tag}__gw"], out[f"{tag}__gp"] = np.array(float(loss)), gw.numpy(), gp.numpy()
        print(f"{tag:8s} loss {float(loss): .12e} |gw| {float(gw.norm()):.6e} gp {gp.numpy()}")
    for k, v in {**a1, **a2}.items():
        out[f"aux_{k}"] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
