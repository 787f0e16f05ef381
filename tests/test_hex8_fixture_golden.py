"""Mid-size END-TO-END golden on the reference's own Cubit-numbered Hex8 fixture test/fem/mesh_hex8.g (4 096 elements,
32 768 quadrature points, nt = 2, MLP 4->50x4->3): loss and weight gradient produced by EXECUTING the reference's own
EnergyLoss / SolidMechanics / Field / NonAllocatedFunctionSpace on the torch stand-in for jax
(oracle/gen_golden_hex8_fixture.py -> tests/golden/hex8_fixture.npz, mesh included).

* not gpu: the CPU oracle against those vectors;  * gpu: the CUDA path through the C ABI.  Tolerance 1e-10 relative."""
import os

import numpy as np
import pytest

from oracle.gen_golden_hex8_fixture import MODELS
from tests.cases import Case, run_engine_loss, run_oracle_loss

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hex8_fixture.npz"))
TOL = 1e-10


def fixture_case(model):
    nodesets = {k.split("__", 1)[1]: G[k] for k in G.files if k.startswith("nodeset__")}
    mesh = dict(coords=G["coords"], conns=G["conns"], elem="hex8", nodesets=nodesets)
    return Case(f"hex8_fixture_{model}", mesh, 2, G["times"], "solid", 3, (4, 3, 50, 4), G["weights"], model,
                "three_dimensional", list(MODELS[model]), bc=("uniaxial", 1.0 if model == "neohookean" else 0.5, 1.0, "y"))


def _check(out, model):
    L, aux, gw, _ = out
    e = aux["energy"] if isinstance(aux, dict) else aux[0]
    assert G["conns"].shape == (4096, 8)
    assert abs(L - float(G[f"{model}__loss"])) <= TOL * abs(float(G[f"{model}__loss"]))
    assert abs(e - float(G[f"{model}__aux_energy"])) <= TOL * abs(float(G[f"{model}__aux_energy"]))
    d = np.linalg.norm(gw - G[f"{model}__gw"]) / np.linalg.norm(G[f"{model}__gw"])
    assert d < TOL, d


@pytest.mark.parametrize("model", sorted(MODELS))
def test_oracle_matches_reference_on_hex8_fixture(model):
    _check(run_oracle_loss(fixture_case(model), "energy", energy_weight=float(G["energy_weight"])), model)


@pytest.mark.gpu
@pytest.mark.parametrize("model", sorted(MODELS))
@pytest.mark.parametrize("deterministic", [True, False])
def test_cuda_matches_reference_on_hex8_fixture(cuda_device, model, deterministic):
    _check(run_engine_loss(fixture_case(model), "energy", deterministic=deterministic, energy_weight=float(G["energy_weight"])), model)
