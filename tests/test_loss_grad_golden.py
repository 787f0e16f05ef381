"""End-to-end parity against tests/golden/loss_grad.npz: loss, aux and d(loss)/d(theta, props) produced by
EXECUTING the reference's own loss functions / physics kernels / Field / constitutive models
(pancax/loss_functions/weak_form_loss_functions.py:74-88,194-205,278-300, data_loss_functions.py:13-24,
physics_kernels/base.py:248-385) on the torch-backed jax stand-in (oracle/gen_golden_ad.py).

* not gpu: the CPU oracle against the golden vectors (this is what pins the oracle's gradients);
* gpu:     the CUDA path, through the C ABI, against the same vectors.
Tolerance 1e-10 relative (fp64, BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from oracle.gen_golden_ad import CASES, T0, VISCO_CASES
from tests.cases import make_case, run_engine_loss, run_oracle_loss

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_grad.npz"))
TOL = 1e-10
PARAMS = [(c[0], c[1], c[2], c[3], c[4], c[5], k) for c in CASES for k in c[6]]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


def _check(out, tag, kind, nt, tol=TOL):
    L, aux, gw, gp = out
    if isinstance(aux, dict):      # oracle: dict with the reference's aux keys; engine: PCX_AUX_* layout
        a = np.zeros(4 + nt)
        a[0] = aux["energy"]
        a[1] = aux.get("residual", 0.0)
        a[2] = aux.get("global_data_loss", 0.0)
        if "reactions" in aux:
            a[4:] = aux["reactions"]
        aux = a
    key = f"{tag}__{kind}__"
    assert abs(L - float(G[key + "loss"])) <= tol * abs(float(G[key + "loss"])), (tag, kind, L, float(G[key + "loss"]))
    assert rel(gw, G[key + "gw"]) < tol, (tag, kind, "gw", rel(gw, G[key + "gw"]))
    if G[key + "gp"].size:
        assert rel(gp, G[key + "gp"]) < tol, (tag, kind, "gp", gp, G[key + "gp"])
    assert abs(aux[0] - float(G[key + "aux_energy"])) <= tol * abs(float(G[key + "aux_energy"]))
    if kind != "energy":
        assert abs(aux[1] - float(G[key + "aux_residual"])) <= tol * abs(float(G[key + "aux_residual"]))
    if kind == "energy_residual_reaction":
        assert abs(aux[2] - float(G[key + "aux_global_data_loss"])) <= tol * abs(float(G[key + "aux_global_data_loss"]))
        assert rel(aux[4:4 + nt], G[key + "aux_reactions"]) < tol


@pytest.mark.parametrize("name,n,nt,width,depth,q,kind", PARAMS)
def test_oracle_matches_reference_loss_and_grad(name, n, nt, width, depth, q, kind):
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    _check(run_oracle_loss(case, kind), tag, kind, nt)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,nt,width,depth,q,kind", PARAMS)
def test_cuda_matches_reference_loss_and_grad(cuda_device, name, n, nt, width, depth, q, kind):
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    for det in (True, False):
        _check(run_engine_loss(case, kind, deterministic=det), tag, kind, nt)


@pytest.mark.parametrize("name,n,nt,width,depth,q", [c[:6] for c in CASES if not c[0].startswith("poisson")])
def test_oracle_matches_reference_internal_force(name, n, nt, width, depth, q):
    import torch
    from oracle import losses as ol
    from tests.cases import oracle_problem
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    prob = oracle_problem(case)
    us = ol.field_values(prob, case.weights, case.times[-1])
    assert rel(us, G[f"{tag}__us_last"]) < 1e-13
    pi, f = ol.energy_and_force(prob, G[f"{tag}__us_last"], case.prop_raw)
    assert abs(pi - float(G[f"{tag}__Pi_last"])) <= TOL * abs(float(G[f"{tag}__Pi_last"]))
    assert rel(f, G[f"{tag}__f_last"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,nt,width,depth,q", [c[:6] for c in CASES if not c[0].startswith("poisson")])
def test_cuda_matches_reference_internal_force(cuda_device, name, n, nt, width, depth, q):
    from tests.cases import make_engine
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    eng = make_engine(case)
    us = eng.field_forward(case.weights, nt - 1).cpu().numpy()
    assert rel(us, G[f"{tag}__us_last"]) < 1e-13
    pi, f, _ = eng.energy_force(G[f"{tag}__us_last"], case.prop_raw if eng.n_train else None)
    assert abs(float(pi.cpu()[0]) - float(G[f"{tag}__Pi_last"])) <= TOL * abs(float(G[f"{tag}__Pi_last"]))
    assert rel(f.cpu().numpy(), G[f"{tag}__f_last"]) < TOL
    eng.close()


def _ffd_inputs():
    case = make_case("quad4_neohookean", n=5, nt=3, width=50, depth=3)
    return case, G["ffd_quad4_n5_nt3_w50_d3__inputs"], G["ffd_quad4_n5_nt3_w50_d3__outputs"]


def test_oracle_matches_reference_full_field_data_loss():
    from oracle import losses as ol
    from tests.cases import oracle_problem
    case, inputs, outputs = _ffd_inputs()
    L, gw = ol.full_field_data_loss_and_grad(oracle_problem(case), case.weights, inputs, outputs, 3.0)
    assert abs(L - float(G["ffd_quad4_n5_nt3_w50_d3__loss"])) <= TOL * abs(L)
    assert rel(gw, G["ffd_quad4_n5_nt3_w50_d3__gw"]) < TOL


@pytest.mark.gpu
def test_cuda_matches_reference_full_field_data_loss(cuda_device):
    from tests.cases import engine_bc, make_engine
    case, inputs, outputs = _ffd_inputs()
    eng = make_engine(case)
    g = engine_bc(case)
    xs, ts = inputs[:, :2], inputs[:, 2]
    a = np.stack([g(xs[i:i + 1], float(ts[i]), np.zeros((1, 2)))[0] for i in range(len(ts))])
    b = np.stack([g(xs[i:i + 1], float(ts[i]), np.ones((1, 2)))[0] for i in range(len(ts))]) - a
    L, gw = eng.field_data_loss_and_grad(case.weights, inputs, outputs, a, b, weight=3.0)
    assert abs(float(L.cpu()[0]) - float(G["ffd_quad4_n5_nt3_w50_d3__loss"])) <= TOL * abs(float(L.cpu()[0]))
    assert rel(gw.cpu().numpy(), G["ffd_quad4_n5_nt3_w50_d3__gw"]) < TOL
    eng.close()


# ---- post-processing element variables: deformation_gradient, I1_bar, pk1_stress as the
# reference's element_pp wrappers return them (physics_kernels/base.py:24-158, solid_mechanics.py:115-170)
PP_CASES = [c[:6] for c in CASES if not c[0].startswith("poisson")]


@pytest.mark.parametrize("name,n,nt,width,depth,q", PP_CASES)
def test_oracle_matches_reference_element_fields(name, n, nt, width, depth, q):
    from oracle import losses as ol
    from tests.cases import oracle_problem
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    F, I1b, P = ol.element_fields(oracle_problem(case), G[f"{tag}__us_last"], case.prop_raw)
    assert rel(F, G[f"{tag}__pp_F"]) < 1e-13
    assert rel(I1b, G[f"{tag}__pp_I1bar"]) < 1e-12
    assert rel(P, G[f"{tag}__pp_P"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,nt,width,depth,q", PP_CASES)
def test_cuda_matches_reference_element_fields(cuda_device, name, n, nt, width, depth, q):
    from tests.cases import make_engine
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
    tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
    eng = make_engine(case)
    F, I1b, P = eng.element_fields(G[f"{tag}__us_last"], case.prop_raw if eng.n_train else None)
    assert rel(F.cpu().numpy(), G[f"{tag}__pp_F"]) < 1e-13
    assert rel(I1b.cpu().numpy(), G[f"{tag}__pp_I1bar"]) < 1e-12
    assert rel(P.cpu().numpy(), G[f"{tag}__pp_P"]) < TOL
    # outputs are optional
    F2, I2, P2 = eng.element_fields(G[f"{tag}__us_last"], case.prop_raw if eng.n_train else None,
                                    want_F=False, want_P=False)
    assert F2 is None and P2 is None and np.array_equal(I2.cpu().numpy(), I1b.cpu().numpy())
    eng.close()


# ---- path-dependent calls of stateless models: the time loop starts at step 1
PD = "pathdep_quad4_neohookean_n5_nt4_w20_d2__"


def _pd_case():
    return make_case("quad4_neohookean", n=5, nt=4, width=20, depth=2, t0=0.1)


@pytest.mark.parametrize("kind", ["energy", "energy_residual_reaction"])
def test_oracle_matches_reference_path_dependent_call(kind):
    L, aux, gw, _ = run_oracle_loss(_pd_case(), kind, first_step=1)
    assert abs(L - float(G[PD + kind + "__loss"])) <= TOL * abs(L)
    assert rel(gw, G[PD + kind + "__gw"]) < TOL
    assert abs(float(aux["energy"]) - float(G[PD + kind + "__aux_energy"])) <= TOL * abs(float(aux["energy"]))
    if kind != "energy":
        assert abs(float(aux["residual"]) - float(G[PD + kind + "__aux_residual"])) <= TOL * float(aux["residual"])
        assert abs(float(aux["global_data_loss"]) - float(G[PD + kind + "__aux_global_data_loss"])) <= \
            TOL * float(aux["global_data_loss"])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["energy", "energy_residual_reaction"])
def test_cuda_matches_reference_path_dependent_call(cuda_device, kind):
    from tests.cases import LOSS_WEIGHTS, make_engine
    case = _pd_case()
    eng = make_engine(case)
    loss, aux, gw, _ = eng.loss_and_grad(kind, case.weights, global_outputs=case.global_outputs, first_step=1, **LOSS_WEIGHTS)
    L, aux = float(loss.cpu()[0]), aux.cpu().numpy()
    assert abs(L - float(G[PD + kind + "__loss"])) <= TOL * abs(L)
    assert rel(gw.cpu().numpy(), G[PD + kind + "__gw"]) < TOL
    assert abs(aux[0] - float(G[PD + kind + "__aux_energy"])) <= TOL * abs(aux[0])
    if kind != "energy":
        assert abs(aux[1] - float(G[PD + kind + "__aux_residual"])) <= TOL * aux[1]
        assert abs(aux[2] - float(G[PD + kind + "__aux_global_data_loss"])) <= TOL * aux[2]
    eng.close()


# ---------------------------------------------------------------- state variables (SimpleFeFv, path-dependent EnergyLoss)
def _check_visco(out, tag):
    L, aux, gw, _ = out
    e = aux["energy"] if isinstance(aux, dict) else aux[0]
    key = f"{tag}__energy__"
    assert abs(L - float(G[key + "loss"])) <= TOL * abs(float(G[key + "loss"])), (tag, L, float(G[key + "loss"]))
    assert abs(float(e) - float(G[key + "aux_energy"])) <= TOL * abs(float(G[key + "aux_energy"]))
    assert rel(gw, G[key + "gw"]) < TOL, (tag, rel(gw, G[key + "gw"]))


@pytest.mark.parametrize("name,n,nt,width,depth,q", VISCO_CASES)
def test_oracle_matches_reference_visco_path_dependent(name, n, nt, width, depth, q):
    """The reference's own SimpleFeFv.energy / EnergyLoss.path_dependent_call (lax.scan over the steps, state threaded
    through, jax.scipy.linalg.expm, mtk_log_sqrt with its registered rule) executed on the stand-in."""
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
    _check_visco(run_oracle_loss(case, "energy", first_step=1), f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}")


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,nt,width,depth,q", VISCO_CASES)
def test_cuda_matches_reference_visco_path_dependent(cuda_device, name, n, nt, width, depth, q):
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
    for det in (True, False):
        _check_visco(run_engine_loss(case, "energy", deterministic=det, first_step=1), f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}")


# ---------------------------------------------------------------- state variables under the residual / reaction losses
# EnergyAndResidualLoss.path_dependent_call (weak_form_loss_functions.py:151-192) and
# EnergyResidualAndReactionLoss.path_dependent_call (:242-276) with SimpleFeFv: goldens from the reference's own calls on
# the stand-in (oracle/gen_golden_visco_resid.py) -- what examples/inverse_problems/mechanics/path-dependent runs.
GV = np.load(os.path.join(os.path.dirname(__file__), "golden", "visco_resid.npz"))
VISCO_RESID_CASES = [("hex8_visco_neohookean", 2, 4, 12, 2, 2), ("quad4_visco_gent", 3, 3, 12, 2, 2),
                     ("hex8_visco_hencky", 2, 3, 8, 2, 2), ("tri3_visco_neohookean_bounded", 3, 3, 10, 2, 2)]
RESID_KINDS = ["energy_residual", "energy_residual_reaction"]


def _check_visco_resid(out, tag, kind):
    L, aux, gw, gp = out
    key = f"{tag}__{kind}__"
    names = ["energy", "residual", "global_data_loss"][:2 if kind == "energy_residual" else 3]
    assert abs(L - float(GV[key + "loss"])) <= TOL * abs(float(GV[key + "loss"])), (tag, kind, L, float(GV[key + "loss"]))
    for i, nme in enumerate(names):
        a = float(aux[nme]) if isinstance(aux, dict) else float(aux[i])
        assert abs(a - float(GV[key + "aux_" + nme])) <= TOL * abs(float(GV[key + "aux_" + nme])), (tag, kind, nme)
    assert rel(gw, GV[key + "gw"]) < TOL, (tag, kind, rel(gw, GV[key + "gw"]))
    if GV[key + "gp"].size:
        assert rel(gp, GV[key + "gp"]) < TOL, (tag, kind, gp, GV[key + "gp"])


@pytest.mark.parametrize("kind", RESID_KINDS)
@pytest.mark.parametrize("name,n,nt,width,depth,q", VISCO_RESID_CASES)
def test_oracle_matches_reference_visco_residual_losses(name, n, nt, width, depth, q, kind):
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
    _check_visco_resid(run_oracle_loss(case, kind, first_step=1), f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}", kind)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", RESID_KINDS)
@pytest.mark.parametrize("name,n,nt,width,depth,q", VISCO_RESID_CASES)
def test_cuda_matches_reference_visco_residual_losses(cuda_device, name, n, nt, width, depth, q, kind):
    case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
    for det in (True, False):
        _check_visco_resid(run_engine_loss(case, kind, deterministic=det, first_step=1),
                           f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}", kind)


# ---------------------------------------------------------------- PlaneStress with state: no golden (the reference's own
# call does not finish on the stand-in) -- the oracle's gradient is checked against central differences of
# its own loss, and the CUDA path against the oracle (tests/test_gpu_visco.py::test_plane_stress_with_state)
@pytest.mark.parametrize("name", ["tri3_visco_gent_pstress_bounded", "quad4_visco_neohookean_pstress"])
def test_oracle_plane_stress_with_state_gradient_is_consistent(name):
    case = make_case(name, n=2, nt=3, width=8, depth=2)
    L, aux, gw, gp = run_oracle_loss(case, "energy", first_step=1)
    rng = np.random.default_rng(3)
    d = rng.standard_normal(gw.shape)
    d /= np.linalg.norm(d)
    h = 1e-6
    Lp = run_oracle_loss(case, "energy", weights=case.weights + h * d, first_step=1)[0]
    Lm = run_oracle_loss(case, "energy", weights=case.weights - h * d, first_step=1)[0]
    fd = (Lp - Lm) / (2 * h)
    assert abs(fd - gw @ d) <= 1e-6 * max(abs(fd), 1e-3), (fd, gw @ d)
