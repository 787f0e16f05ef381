"""Parity at mid sizes DIRECTLY against the oracle (the golden meshes have 2..6 elements per side, the
full-size C2 / C3 checks are property tests).  The oracle finishes these in seconds on the host:

  * Hex8 32^3 (262 144 qp), nt = 2, EnergyLoss, MLP 4->50x4->3  -- the CPU arm's own sample of config C3;
  * Hex8 64^3 (2.1 M qp), nt = 2, EnergyLoss -- one eighth of C3;
  * Quad4 128^2 (65 536 qp), nt = 11, EnergyResidualAndReactionLoss + EnergyAndResidualLoss, MLP 3->50x3->2 -- config
    C2's loss and time grid on a mesh the second-order AD of the oracle still handles;
  * the same with a trainable shear modulus (config C4's property gradient).

Tolerance: 1e-10 relative on loss, aux and every gradient (BASELINE.json north_star)."""
import numpy as np
import pytest

from tests.cases import make_case, run_engine_loss, run_oracle_loss

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize("n,deterministic", [(32, True), (32, False), (64, True)])
def test_hex8_energy_loss_vs_oracle(cuda_device, n, deterministic):
    case = make_case("hex8_neohookean", n=n, nt=2, permute=(n == 32))
    L, aux, gw, _ = run_engine_loss(case, "energy", deterministic=deterministic)
    Lo, auxo, gwo, _ = run_oracle_loss(case, "energy")
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert abs(aux[0] - auxo["energy"]) <= TOL * abs(auxo["energy"])
    assert aux[3] == 0.0
    assert rel(gw, gwo) < TOL


@pytest.mark.parametrize("model", ["gent", "swanson", "hencky"])
def test_hex8_32_other_models_vs_oracle(cuda_device, model):
    case = make_case(f"hex8_{model}", n=32, nt=2, permute=False)
    L, aux, gw, _ = run_engine_loss(case, "energy")
    Lo, auxo, gwo, _ = run_oracle_loss(case, "energy")
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL


@pytest.mark.parametrize("name", ["quad4_neohookean", "quad4_neohookean_bounded"])
@pytest.mark.parametrize("kind", ["energy_residual_reaction", "energy_residual"])
def test_quad4_128_nt11_residual_losses_vs_oracle(cuda_device, name, kind):
    case = make_case(name, n=128, nt=11, width=50, depth=3, permute=True)
    case.times = np.linspace(0.1, 1.0, 11)              # F9: ||f|| > 0 at every step
    case.global_outputs = np.linspace(0.0, 0.2, 11)
    L, aux, gw, gp = run_engine_loss(case, kind)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert abs(aux[0] - auxo["energy"]) <= TOL * abs(auxo["energy"])
    assert abs(aux[1] - auxo["residual"]) <= TOL * abs(auxo["residual"])
    if kind == "energy_residual_reaction":
        assert rel(aux[4:], auxo["reactions"]) < TOL
        assert abs(aux[2] - auxo["global_data_loss"]) <= TOL * abs(auxo["global_data_loss"])
    assert rel(gw, gwo) < TOL
    if gpo.size:
        assert rel(gp, gpo) < TOL
