"""Generates tests/golden/hex8_fixture.npz: a mid-size END-TO-END golden on the reference's own Cubit-numbered Hex8
fixture test/fem/mesh_hex8.g (4 096 elements, 4 913 nodes, q_order 2 -> 32 768 quadrature points): EnergyLoss and its
weight gradient for 3D NeoHookean (K = 1000, G = 1: config C3's physics) and Gent, nt = 2, MLP 4->50x4->3, obtained by
EXECUTING the reference's own EnergyLoss / SolidMechanics / Field / NonAllocatedFunctionSpace on the torch stand-in for
jax (oracle/refshim_torch.py).  The mesh travels inside the .npz (compressed), so the tests need nothing from
/root/reference.

    python -m oracle.gen_golden_hex8_fixture          # build container only (needs /root/reference)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "hex8_fixture.npz")
MODELS = {"neohookean": [1000.0, 1.0], "gent": [0.833, 0.3846, 3.0]}
ENERGY_WEIGHT = 1.3


def main():
    sys.path.insert(0, ROOT)
    from oracle import fem as ofem
    from oracle.field import flatten_mlp, init_mlp
    mesh = ofem.read_exodus_mesh(os.path.join(REF, "test", "fem", "mesh_hex8.g"))
    weights = flatten_mlp(init_mlp(4, 3, 50, 4, seed=0))
    times_np = np.linspace(0.0, 1.0, 2)
    out = dict(coords=mesh["coords"], conns=mesh["conns"], times=times_np, weights=weights,
               energy_weight=np.array(ENERGY_WEIGHT))
    for k, v in mesh["nodesets"].items():
        out[f"nodeset__{k}"] = v
    from oracle import refshim_torch as shim
    shim.import_reference(REF)
    from pancax.fem.elements import Hex8Element
    from pancax.fem.quadrature_rules import QuadratureRule
    from pancax.fem.function_space import NonAllocatedFunctionSpace
    from pancax.constitutive_models import Gent, NeoHookean
    from pancax.physics_kernels import SolidMechanics, ThreeDimensional
    from pancax.bvps.uniaxial_tension import UniaxialTensionLinearRamp
    from pancax.loss_functions.weak_form_loss_functions import EnergyLoss
    from pancax.networks.parameters import Parameters, _Parameters
    import pancax.networks.fields as ref_fields
    el = Hex8Element()
    space = NonAllocatedFunctionSpace(types.SimpleNamespace(parentElement=el), QuadratureRule(el, 2))
    coords = torch.as_tensor(mesh["coords"], dtype=torch.float64)
    conns = torch.as_tensor(mesh["conns"], dtype=torch.long)
    times = torch.as_tensor(times_np, dtype=torch.float64)
    for name, props in MODELS.items():
        model = NeoHookean(bulk_modulus=props[0], shear_modulus=props[1]) if name == "neohookean" else \
            Gent(bulk_modulus=props[0], shear_modulus=props[1], Jm_parameter=props[2])
        physics = SolidMechanics(model, ThreeDimensional())
        bc = UniaxialTensionLinearRamp(1.0 if name == "neohookean" else 0.5, 1.0, "y", 3)
        domain = types.SimpleNamespace(coords=coords, conns=conns, fspace=space, dof_manager=types.SimpleNamespace(unknownIndices=None),
                                       times=times)
        problem = types.SimpleNamespace(coords=coords, times=times, domain=domain, physics=physics, global_data=None,
                                        n_dims=3, n_dofs=3, conns=conns, fspace=space)
        fld = ref_fields.Field(problem, np.array([0, 0], dtype=np.uint32), n_layers=4, n_neurons=50, dirichlet_bc_func=bc)
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
                leaf.copy_(torch.as_tensor(weights[off:off + n]).reshape(leaf.shape))
                off += n
        loss, aux = EnergyLoss(weight=ENERGY_WEIGHT)(params, problem)
        gs = torch.autograd.grad(loss, leaves)
        gw = torch.cat([g.reshape(-1) for g in gs]).numpy()
        out[f"{name}__loss"], out[f"{name}__gw"] = np.array(float(loss.detach())), gw
        out[f"{name}__aux_energy"] = np.asarray(aux["energy"].detach().numpy())
        print(f"{name:12s} loss {float(loss.detach()): .12e} |gw| {np.linalg.norm(gw):.6e}", flush=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
