"""Config C4 of BASELINE.json, literally: the reference's own inverse-problem example
(examples/inverse_problems/mechanics/vanilla/example_v2.py:20-109) -- its TRI3 mesh (902 elements), its DIC file (first
dataloader batch of 1 024 of the 10 332 rows), its 21-step global force-displacement file, NeoHookean with a trainable
shear modulus, loss = EnergyResidualAndReactionLoss(250e9, 250e9) + FullFieldDataLoss(10e9).

tests/golden/c4_vanilla.npz (oracle/gen_golden_c4.py) holds the inputs and the (loss, aux, d loss / d theta, d loss / d G)
obtained by EXECUTING the reference's own loss functions on the torch stand-in for jax.

* not gpu: the CPU oracle against those vectors;
* gpu:     the CUDA path through the pancax-named host API (InverseProblem / Parameters / CombineLossFunctions, i.e. the
           call example_v2.py makes) AND through the C ABI engine, against the same vectors.
Tolerance 1e-10 relative."""
import os

import numpy as np
import pytest

from tests.cases import Case, make_engine, run_engine_loss, run_oracle_loss

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "c4_vanilla.npz"))
TOL = 1e-10
LW = dict(energy_weight=1.0, residual_weight=250.e9, reaction_weight=250.e9)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def c4_case():
    mesh = dict(coords=G["coords"], conns=G["conns"], elem="tri3",
                nodesets={"nodeset_2": G["nodeset_2"], "nodeset_4": G["nodeset_4"]})
    return Case("c4_vanilla", mesh, 2, G["times"], "solid", 2, (3, 2, 50, 3), G["weights"], "neohookean", "plane_strain",
                [0.833, (0.01, 5.0, float(G["g_raw"][0]))], bc=("uniaxial", float(G["final_displacement"]), 1.0, "x"),
                unknown_idx=G["unknown_idx"], reaction_nodes=G["reaction_nodes"], reaction_dof=int(G["reaction_dof"]),
                global_outputs=G["global_outputs"])


def _check_physics(L, aux, gw, gp):
    assert G["times"].shape == (21,) and G["conns"].shape == (902, 3) and int(G["n_full_field_rows"]) == 10332
    assert abs(L - float(G["physics__loss"])) <= TOL * abs(float(G["physics__loss"]))
    assert rel(gw, G["physics__gw"]) < TOL
    assert rel(gp, G["physics__gp"]) < TOL
    if isinstance(aux, dict):
        e, r, g, re = aux["energy"], aux["residual"], aux["global_data_loss"], aux["reactions"]
    else:
        e, r, g, re = aux[0], aux[1], aux[2], aux[4:]
    assert abs(e - float(G["aux_energy"])) <= TOL * abs(float(G["aux_energy"]))
    assert abs(r - float(G["aux_residual"])) <= TOL * abs(float(G["aux_residual"]))
    assert abs(g - float(G["aux_global_data_loss"])) <= TOL * abs(float(G["aux_global_data_loss"]))
    assert rel(re, G["aux_reactions"]) < TOL


def test_oracle_matches_reference_on_c4():
    from oracle import losses as ol
    from tests.cases import oracle_problem
    case = c4_case()
    L, aux, gw, gp = run_oracle_loss(case, "energy_residual_reaction", **LW)
    _check_physics(L, aux, gw, gp)
    Ld, gwd = ol.full_field_data_loss_and_grad(oracle_problem(case), case.weights, G["batch_inputs"], G["batch_outputs"], weight=10.e9)
    assert abs(Ld - float(G["data__loss"])) <= TOL * abs(float(G["data__loss"]))
    assert rel(gwd, G["data__gw"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("deterministic", [True, False])
def test_cuda_engine_matches_reference_on_c4(cuda_device, deterministic):
    case = c4_case()
    _check_physics(*run_engine_loss(case, "energy_residual_reaction", deterministic=deterministic, **LW))


@pytest.mark.gpu
def test_host_api_runs_example_v2_and_matches_reference(cuda_device):
    """The calls of example_v2.py:20-109 on the host mirror: InverseProblem, Parameters, the two losses combined."""
    import torch
    import pancax_b200 as px
    from pancax_b200.fem import Mesh
    mesh = Mesh(G["coords"], G["conns"].astype(np.int32), "tri3",
                {"nodeset_2": G["nodeset_2"].astype(np.int32), "nodeset_4": G["nodeset_4"].astype(np.int32)})
    times = G["times"]
    domain = px.VariationalDomain(mesh, times)
    model = px.NeoHookean(bulk_modulus=0.833, shear_modulus=px.BoundedProperty(0.01, 5.0, key=0))
    physics = px.SolidMechanics(model, px.PlaneStrain())
    gd = px.GlobalData.from_arrays(times, G["global_outputs"], G["reaction_nodes"], 0)
    bcs = [px.DirichletBC(nset_name="nodeset_2", component=0), px.DirichletBC(nset_name="nodeset_2", component=1),
           px.DirichletBC(nset_name="nodeset_4", component=0), px.DirichletBC(nset_name="nodeset_4", component=1)]
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    bc = px.UniaxialTensionLinearRamp(final_displacement=float(G["final_displacement"]), length=1.0, direction="x", n_dimensions=2)
    params = px.Parameters(problem, key=0, dirichlet_bc_func=bc)                 # defaults: 3 layers x 50 neurons
    params.fields.networks.set_weights(G["weights"])
    params.prop_raw.copy_(torch.as_tensor(G["g_raw"]))
    loss_fn = px.CombineLossFunctions(px.EnergyResidualAndReactionLoss(residual_weight=250.e9, reaction_weight=250.e9),
                                      px.FullFieldDataLoss(weight=10.e9))
    (loss, aux), grads = loss_fn.value_and_grad(params, problem, G["batch_inputs"], G["batch_outputs"])
    assert abs(float(loss) - float(G["total__loss"])) <= TOL * abs(float(G["total__loss"]))
    assert rel(grads.fields.cpu().numpy(), G["total__gw"]) < TOL
    assert rel(grads.props.cpu().numpy(), G["total__gp"]) < TOL
    assert abs(float(aux["field_data_loss"]) - float(G["aux_field_data_loss"])) <= TOL * abs(float(G["aux_field_data_loss"]))
    assert rel(aux["reactions"].cpu().numpy(), G["aux_reactions"]) < TOL
