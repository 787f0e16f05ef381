"""GPU parity of the PlaneStress formulation (pancax/physics_kernels/solid_mechanics.py:32-71: grad_u_33 = root of
sigma_33 = 0 per quadrature point, math/scalar_root_find.py, differentiated through jax.lax.custom_root) against the
CPU oracle, which is itself pinned to vectors produced by executing the reference's own PlaneStress on the stand-in
(tests/test_loss_grad_golden.py, cases *_pstress).  Tolerance 1e-10 relative (fp64)."""
import numpy as np
import pytest
import torch

from tests.cases import make_case, make_engine, oracle_problem, run_engine_loss, run_oracle_loss

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


@pytest.mark.parametrize("name", ["quad4_neohookean_pstress", "quad4_gent_pstress", "quad4_swanson_pstress",
                                  "quad4_hencky_pstress", "tri3_neohookean_pstress"])
@pytest.mark.parametrize("deterministic", [True, False])
def test_energy_force_plane_stress(cuda_device, name, deterministic):
    from oracle import losses as ol
    case = make_case(name, n=5, nt=2)
    eng = make_engine(case, deterministic=deterministic)
    prob = oracle_problem(case)
    us = ol.field_values(prob, case.weights, 1.0)
    Pi, f, _ = eng.energy_force(us)
    Pi_o, f_o = ol.energy_and_force(prob, us)
    assert abs(float(Pi.cpu()[0]) - Pi_o) <= TOL * abs(Pi_o), name
    assert rel(f.cpu().numpy(), f_o) < TOL, name
    eng.close()


@pytest.mark.parametrize("name", ["quad4_neohookean_pstress", "quad4_gent_pstress", "quad4_swanson_pstress",
                                  "quad4_hencky_pstress", "tri3_neohookean_pstress"])
def test_stiffness_action_plane_stress(cuda_device, name):
    """K v with the implicit derivative of the root (dx = -dP_33 / (dP_33/dx)); K stays symmetric."""
    from oracle import losses as ol
    case = make_case(name, n=4, nt=2)
    eng = make_engine(case)
    prob = oracle_problem(case)
    us = ol.field_values(prob, case.weights, 1.0)
    rng = np.random.default_rng(2)
    v = rng.standard_normal(us.shape)
    Kv, _ = eng.stiffness_action(us, v)
    Kv_o = ol.stiffness_action(prob, us, v)
    assert rel(Kv.cpu().numpy(), Kv_o) < TOL, name
    w2 = rng.standard_normal(us.shape)
    Kw, _ = eng.stiffness_action(us, w2)
    a = float((torch.as_tensor(w2) * Kv.cpu()).sum())
    b = float((torch.as_tensor(v) * Kw.cpu()).sum())
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    eng.close()


@pytest.mark.parametrize("name", ["quad4_neohookean_pstress", "quad4_gent_pstress_bounded", "tri3_neohookean_pstress_bounded",
                                  "quad4_hencky_pstress"])
@pytest.mark.parametrize("kind", ["energy", "energy_residual", "energy_residual_reaction"])
def test_losses_plane_stress(cuda_device, name, kind):
    case = make_case(name, n=5, nt=3)
    case.times = np.linspace(0.25, 1.0, 3)   # parity on steps with ||f|| > 0
    L, aux, gw, gp = run_engine_loss(case, kind)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= TOL * abs(Lo), (name, kind)
    assert aux[3] == 0.0, "root find did not converge somewhere"
    if kind != "energy":
        assert abs(aux[1] - auxo["residual"]) <= TOL * abs(auxo["residual"])
    if kind == "energy_residual_reaction":
        assert rel(aux[4:], auxo["reactions"]) < TOL
    assert rel(gw, gwo) < TOL, (name, kind)
    if gpo.size:
        assert rel(gp, gpo) < TOL, (name, kind, gp, gpo)


def test_post_processing_fields_plane_stress(cuda_device):
    """element_pp outputs (F, I1_bar, P) with F_33 from the root; sigma_33 ~ P_33 = 0 at every point."""
    from oracle import losses as ol
    case = make_case("quad4_neohookean_pstress", n=5, nt=2)
    eng = make_engine(case)
    prob = oracle_problem(case)
    us = ol.field_values(prob, case.weights, 1.0)
    F, I1, P = eng.element_fields(us)
    F_o, I1_o, P_o = ol.element_fields(prob, us)
    assert rel(F.cpu().numpy().reshape(F_o.shape), F_o) < 1e-12
    assert rel(I1.cpu().numpy().reshape(I1_o.shape), I1_o) < 1e-12
    Pn = P.cpu().numpy().reshape(P_o.shape)
    assert rel(Pn, P_o) < TOL
    # x converged to 1e-15 and dP_33/dx ~ K = 1000 here
    assert np.abs(Pn[..., 2, 2]).max() <= 1e-10 * np.abs(Pn).max()
    assert np.all(F.cpu().numpy().reshape(F_o.shape)[..., 2, 2] != 1.0)
    eng.close()


def test_plane_stress_needs_a_2d_mesh(cuda_device):
    from pancax_b200 import Engine
    from pancax_b200._lib import PcxError
    case = make_case("hex8_neohookean", n=2)
    eng = Engine(case.mesh["coords"], case.mesh["conns"], "hex8", 2, 3, case.times)
    with pytest.raises(PcxError, match="PlaneStress"):
        eng.set_physics("solid", "neohookean", "plane_stress", [1000.0, 1.0])
    eng.close()


@pytest.mark.parametrize("name", ["quad4_neohookean_pstress", "quad4_gent_pstress", "quad4_swanson_pstress",
                                  "quad4_hencky_pstress"])
def test_root_find_converges_everywhere(cuda_device, name):
    """4 096 quadrature points x 4 load levels: every per-point root find must report convergence (aux[3] == 0) --
    a Newton iterate that sits on the root to rounding becomes a bracket end, and the later Newton steps that land on
    that end must be taken, not bisected away (Hencky ran out of its iteration budget there before)."""
    case = make_case(name, n=32, nt=4)
    case.times = np.linspace(0.25, 1.0, 4)
    L, aux, gw, gp = run_engine_loss(case, "energy_residual")
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy_residual")
    assert aux[3] == 0.0, "root find did not converge somewhere"
    assert abs(L - Lo) <= TOL * abs(Lo), name
    assert rel(gw, gwo) < 1e-9, name
