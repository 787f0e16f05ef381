"""GPU parity of the state-variable path: SimpleFeFv (pancax/constitutive_models/mechanics/
hyperviscoelasticity/simple_fefv.py) under the path-dependent EnergyLoss (weak_form_loss_functions.py:48-72) -- the
forward sweep over time with the states kept and the reverse sweep jax derives from the lax.scan -- against the CPU
oracle (torch autograd through the same time loop), which is pinned to vectors produced by executing the reference's
own SimpleFeFv / path_dependent_call on the stand-in (tests/test_loss_grad_golden.py, cases *_visco_*).
Tolerance 1e-10 relative (fp64)."""
import numpy as np
import pytest

from tests.cases import LOSS_WEIGHTS, make_case, make_engine, run_engine_loss, run_oracle_loss

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


@pytest.mark.parametrize("name,n,nt", [("hex8_visco_neohookean", 3, 4), ("hex8_visco_gent", 3, 3), ("hex8_visco_hencky", 2, 3),
                                       ("hex8_visco_swanson", 2, 3), ("hex8_visco_blatz_ko", 2, 4),
                                       ("quad4_visco_neohookean", 5, 5), ("tri3_visco_gent", 4, 3)])
@pytest.mark.parametrize("deterministic", [True, False])
def test_path_dependent_energy_loss_with_state(cuda_device, name, n, nt, deterministic):
    case = make_case(name, n=n, nt=nt, width=20, depth=3)
    L, aux, gw, _ = run_engine_loss(case, "energy", deterministic=deterministic, first_step=1)
    Lo, auxo, gwo, _ = run_oracle_loss(case, "energy", first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo), (name, L, Lo)
    assert abs(aux[0] - auxo["energy"]) <= TOL * abs(auxo["energy"])
    assert aux[3] == 0.0
    assert rel(gw, gwo) < TOL, (name, rel(gw, gwo))


def test_non_uniform_time_steps(cuda_device):
    """dt of step n is times[1]-times[0] for n = 1 and times[n-1]-times[n-2] afterwards (the update at the end of
    the scan body, weak_form_loss_functions.py:63) -- visible only with non-uniform times."""
    case = make_case("hex8_visco_neohookean", n=2, nt=5, width=16, depth=2)
    case.times = np.array([0.0, 0.1, 0.35, 0.5, 1.0])
    L, aux, gw, _ = run_engine_loss(case, "energy", first_step=1)
    Lo, auxo, gwo, _ = run_oracle_loss(case, "energy", first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL


def test_relaxation_limits(cuda_device):
    """tau -> infinity: no viscous flow within the steps, every branch adds G ||dev E||^2 of the TOTAL log strain (an
    elastic Hencky-type shear term); tau -> 0: the branch relaxes completely and only the dissipation term
    G tau ||dev Dv||^2 -> 0 remains, i.e. the equilibrium model alone."""
    base = make_case("hex8_neohookean", n=3, nt=3, width=16, depth=2)
    base.props = [10.0, 0.855]
    Le, _, ge, _ = run_engine_loss(base, "energy", first_step=1)
    fast = make_case("hex8_visco_neohookean", n=3, nt=3, width=16, depth=2)
    fast.prony = ([1.0, 0.5], [1e-9, 1e-10])
    Lf, _, gf, _ = run_engine_loss(fast, "energy", first_step=1)
    assert abs(Lf - Le) <= 1e-7 * abs(Le) and rel(gf, ge) < 1e-7
    slow = make_case("hex8_visco_neohookean", n=3, nt=3, width=16, depth=2)
    slow.prony = ([1.0], [1e12])
    Ls, _, _, _ = run_engine_loss(slow, "energy", first_step=1)
    Lso, _, _, _ = run_oracle_loss(slow, "energy", first_step=1)
    assert Ls > Le and abs(Ls - Lso) <= TOL * abs(Lso)


def test_value_only_and_tf32_adjoint(cuda_device):
    case = make_case("hex8_visco_neohookean", n=3, nt=4, width=20, depth=3)
    eng = make_engine(case)
    L, aux, gw, _ = run_engine_loss(case, "energy", eng=eng, first_step=1)
    loss, aux2, _, _ = eng.loss_and_grad("energy", case.weights, None, want_grad=False, first_step=1, energy_weight=1.3)
    assert float(loss.cpu()[0]) == L
    eng.set_option("tf32_adjoint", 1)
    L32, _, gw32, _ = run_engine_loss(case, "energy", eng=eng, first_step=1)
    eng.close()
    assert L32 == L and rel(gw32, gw) < 1e-5 and np.any(gw32 != gw)


def test_state_models_are_refused_elsewhere(cuda_device):
    from pancax_b200._lib import PcxError
    case = make_case("hex8_visco_neohookean", n=2, nt=3, width=8, depth=2)
    eng = make_engine(case)
    us = np.zeros((eng.nn, 3))
    with pytest.raises(PcxError, match="state variables"):
        eng.energy_force(us)
    with pytest.raises(PcxError, match="state variables"):
        eng.loss_and_grad("energy", case.weights, None)                       # path-independent call
    with pytest.raises(PcxError, match="state variables"):
        eng.loss_and_grad("energy_residual", case.weights, None)              # residual loss, path-independent call
    eng.close()
    ps = make_case("quad4_visco_neohookean_pstress", n=3, nt=3, width=8, depth=2)
    eng = make_engine(ps)
    with pytest.raises(PcxError, match="PlaneStress"):      # third derivatives of the energy: EnergyLoss only
        eng.loss_and_grad("energy_residual", ps.weights, None, first_step=1)
    eng.close()


def test_public_api_example_hyper_visco(cuda_device):
    """examples/forward_problems/mechanics/example_hyper_visco_3d.py: SimpleFeFv(NeoHookean, PronySeries, WLF) with
    the path-dependent energy loss and Adam."""
    import pancax_b200 as px
    import torch
    mesh = px.create_structured_mesh("hex8", 3)
    times = np.linspace(0.0, 1.0, 4)
    domain = px.VariationalDomain(mesh, times)
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=10.0, shear_modulus=0.855),
                          px.PronySeries(moduli=[1.0], relaxation_times=[10.0]), px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
    physics = px.SolidMechanics(model, px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    problem = px.ForwardProblem(domain, physics, [], bcs, [])
    params = px.Parameters(problem, key=0, dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
    loss_fn = px.EnergyLoss(is_path_dependent=True)
    opt = px.Adam(loss_fn, learning_rate=1e-3, has_aux=True)
    st = opt.init(params)
    l0 = None
    for _ in range(20):
        params, st, (loss, aux) = opt.step(params, st, problem)
        l0 = float(loss) if l0 is None else l0
    assert float(loss) < l0 and torch.isfinite(loss)
    with pytest.raises(Exception):
        px.EnergyLoss()(params, problem)      # the path-independent call is undefined for a model with state


def test_state_variables_output(cuda_device, tmp_path):
    """pcx_get_state + the post-processor's `state_variables` (solid_mechanics.py:157-168): the kept state history
    equals the oracle's threaded state, step 0 is initial_state(), and the Exodus file carries
    state_variable_{k}_{q} element variables like the reference's example_hyper_visco_3d.py asks for."""
    import torch
    import pancax_b200 as px
    from oracle import losses as ol
    from oracle.physics import potential_energy_with_state
    from pancax_b200.post_processor import read_exodus_results
    from tests.cases import oracle_problem
    case = make_case("hex8_visco_neohookean", n=2, nt=3, width=12, depth=2)
    eng = make_engine(case)
    with pytest.raises(Exception, match="no state history"):
        eng.get_state(1)
    run_engine_loss(case, "energy", eng=eng, first_step=1)
    z0, z2 = eng.get_state(0).cpu().numpy(), eng.get_state(2).cpu().numpy()
    assert z0.shape == (eng.ne, eng.nq, 18)
    assert np.array_equal(z0, np.tile(np.eye(3).ravel(), (eng.ne, eng.nq, 2)))
    prob = oracle_problem(case)
    w, pr = ol._leaves(prob, case.weights, case.prop_raw)
    fld, phys, pe = ol._build(prob, w, pr)
    coords, conns = torch.as_tensor(prob.coords), torch.as_tensor(prob.conns, dtype=torch.long)
    Z = torch.eye(3, dtype=torch.float64).expand(eng.ne, eng.nq, 2, 3, 3).clone()
    dt = float(case.times[1] - case.times[0])
    with torch.no_grad():
        for n in (1, 2):
            _, Z = potential_energy_with_state(phys, pe, coords, conns, fld(coords, float(case.times[n])), Z, dt)
    assert rel(z2, Z.reshape(eng.ne, eng.nq, 18).numpy()) < 1e-12
    eng.close()
    # through the public post-processor
    mesh = px.create_structured_mesh("hex8", 2)
    times = np.linspace(0.0, 1.0, 3)
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=10.0, shear_modulus=0.855), px.PronySeries([1.0], [10.0]), px.WLF(17.44, 51.6, 60.0))
    problem = px.ForwardProblem(px.VariationalDomain(mesh, times), px.SolidMechanics(model, px.ThreeDimensional()), [],
                                [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)], [])
    params = px.Parameters(problem, 0, n_layers=2, n_neurons=12, dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
    pp = px.PostProcessor(mesh, "exodus")
    out = str(tmp_path / "visco.e")
    pp.init(params, problem, out, node_variables=["field_values", "internal_force"],
            element_variables=["deformation_gradient", "state_variables", "pk1_stress"])
    pp.write_outputs(params, problem)
    pp.close()
    _, nod, el = read_exodus_results(out)
    # internal_force / pk1_stress of a state model: evaluated step by step from the step's old state (pcx_state_step),
    # as the reference's _write_step_outputs does (post_processor.py:244-354)
    from pancax_b200.loss_functions import get_engine
    e2 = get_engine(problem, params)
    wts = params.fields.networks.weights
    z = e2.initial_state()
    for n in range(3):
        _, z, f_n, P_n = e2.state_step(e2.field_forward(wts, n), z, 0.0 if n == 0 else float(times[n] - times[n - 1]))
        assert np.allclose(el["P_yy_3"][n], P_n[:, 2, 1, 1].cpu().numpy(), rtol=1e-12, atol=1e-14)
        assert np.allclose(nod["internal_force_y"][n], f_n[:, 1].cpu().numpy(), rtol=1e-12, atol=1e-14)
        assert np.allclose(el["state_variable_5_2"][n], z[:, 1, 4].cpu().numpy(), rtol=1e-12, atol=1e-14)
    assert np.abs(el["P_yy_3"][2]).max() > 1e-3
    assert "displ_x" in nod and "state_variable_1_1" in el and "state_variable_9_8" in el and "F_xx_1" in el
    assert np.allclose(el["state_variable_1_1"][0], 1.0) and np.allclose(el["state_variable_2_1"][0], 0.0)
    assert not np.allclose(el["state_variable_1_1"][2], 1.0)          # the viscous stretch has evolved by the last step


def test_recompute_mode_when_the_workspace_cannot_hold_every_step(cuda_device, monkeypatch):
    """With fewer resident time steps than nt (here forced with PCX_TCHUNK = 1; in production: PCX_WORKSPACE_GB too
    small for nt x nodes x 1.8 KB of activations) the reverse sweep recomputes each step's activations right before
    its MLP adjoint instead of keeping them: same loss, same gradient to rounding."""
    case = make_case("hex8_visco_gent", n=3, nt=4, width=20, depth=3)
    ref = run_engine_loss(case, "energy", first_step=1)
    monkeypatch.setenv("PCX_TCHUNK", "1")
    got = run_engine_loss(case, "energy", first_step=1)
    assert got[0] == ref[0]
    assert rel(got[2], ref[2]) < 1e-13


@pytest.mark.parametrize("name", ["quad4_visco_neohookean_bounded", "quad4_visco_gent_bounded"])
def test_trainable_equilibrium_properties(cuda_device, name):
    """BoundedProperty on the equilibrium model of a SimpleFeFv (examples/inverse_problems/mechanics/path-dependent/
    script.py:38-48): the Prony branches and the state evolution do not depend on it, d loss / d prop_raw is the sum
    of JxW dpsi_eq/dp over the steps, chained through the sigmoid."""
    case = make_case(name, n=5, nt=4, width=16, depth=2)
    L, aux, gw, gp = run_engine_loss(case, "energy", first_step=1)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy", first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL
    assert gpo.size and rel(gp, gpo) < TOL, (gp, gpo)


# ---------------------------------------------------------------- residual / reaction losses with state (second order)
@pytest.mark.parametrize("deterministic", [True, False])
@pytest.mark.parametrize("kind", ["energy_residual", "energy_residual_reaction"])
@pytest.mark.parametrize("name,n,nt", [("hex8_visco_neohookean", 3, 4), ("hex8_visco_gent", 2, 3), ("hex8_visco_hencky", 2, 3),
                                       ("hex8_visco_swanson", 2, 3), ("hex8_visco_blatz_ko", 2, 3),
                                       ("quad4_visco_neohookean", 5, 5), ("tri3_visco_gent", 4, 3),
                                       ("quad4_visco_gent_bounded", 4, 4)])
def test_path_dependent_residual_losses_with_state(cuda_device, name, n, nt, kind, deterministic):
    """EnergyAndResidualLoss / EnergyResidualAndReactionLoss.path_dependent_call (weak_form_loss_functions.py:151-192,
    242-276) with SimpleFeFv: f_n = dPi_n/du_n at fixed old state, R_n and reaction_n from it, and the gradient with the
    second-order terms through the state history (k_visco_qp MODE 2 / 3) against the oracle's double backward."""
    case = make_case(name, n=n, nt=nt, width=16, depth=2)
    L, aux, gw, gp = run_engine_loss(case, kind, deterministic=deterministic, first_step=1)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind, first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo), (L, Lo)
    assert abs(aux[0] - float(auxo["energy"])) <= TOL * abs(float(auxo["energy"]))
    assert abs(aux[1] - float(auxo["residual"])) <= TOL * abs(float(auxo["residual"]))
    if kind == "energy_residual_reaction":
        assert abs(aux[2] - float(auxo["global_data_loss"])) <= TOL * abs(float(auxo["global_data_loss"]))
    assert rel(gw, gwo) < TOL, rel(gw, gwo)
    if gpo.size:
        assert rel(gp, gpo) < TOL, (gp, gpo)


def test_residual_losses_with_state_value_only_recompute_and_zero_energy_weight(cuda_device, monkeypatch):
    case = make_case("hex8_visco_neohookean", n=3, nt=4, width=12, depth=2)
    eng = make_engine(case)
    ref = run_engine_loss(case, "energy_residual_reaction", eng=eng, first_step=1)
    loss, aux, _, _ = eng.loss_and_grad("energy_residual_reaction", case.weights, None, global_outputs=case.global_outputs,
                                        want_grad=False, first_step=1, **LOSS_WEIGHTS)
    assert float(loss.cpu()[0]) == ref[0]
    # pure residual + reaction loss (energy weight 0): the flux then holds second-order terms only
    got = run_engine_loss(case, "energy_residual_reaction", eng=eng, first_step=1, energy_weight=0.0)
    exp = run_oracle_loss(case, "energy_residual_reaction", first_step=1, energy_weight=0.0)
    assert abs(got[0] - exp[0]) <= TOL * abs(exp[0]) and rel(got[2], exp[2]) < TOL
    eng.close()
    monkeypatch.setenv("PCX_TCHUNK", "1")        # activations of one step resident: recompute mode
    got = run_engine_loss(case, "energy_residual_reaction", first_step=1)
    assert got[0] == ref[0] and rel(got[2], ref[2]) < 1e-13


def test_public_api_inverse_path_dependent_visco(cuda_device):
    """examples/inverse_problems/mechanics/path-dependent/script.py: SimpleFeFv with a trainable equilibrium shear
    modulus, EnergyResidualAndReactionLoss(is_path_dependent=True) + Adam on a small mesh."""
    import pancax_b200 as px
    import torch
    mesh = px.create_structured_mesh("hex8", 3)
    times = np.linspace(0.0, 1.0, 4)
    domain = px.VariationalDomain(mesh, times)
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=10.0, shear_modulus=px.BoundedProperty(0.01, 5.0, key=1)),
                          px.PronySeries(moduli=[1.0, 0.5], relaxation_times=[10.0, 1.0]), px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
    physics = px.SolidMechanics(model, px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    gd = px.GlobalData.from_arrays(times, 0.3 * times, mesh.nodeSets["nset_5"], 1)
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
    loss_fn = px.EnergyResidualAndReactionLoss(is_path_dependent=True, residual_weight=10.0, reaction_weight=10.0)
    opt = px.Adam(loss_fn, learning_rate=1e-3, has_aux=True)
    st = opt.init(params)
    l0 = None
    for _ in range(20):
        params, st, (loss, aux) = opt.step(params, st, problem)
        l0 = float(loss) if l0 is None else l0
    assert torch.isfinite(loss) and float(loss) < l0


# ---------------------------------------------------------------- PlaneStress with state (path-dependent EnergyLoss)
@pytest.mark.parametrize("deterministic", [True, False])
@pytest.mark.parametrize("name,n,nt", [("quad4_visco_neohookean_pstress", 4, 4), ("tri3_visco_gent_pstress_bounded", 4, 3),
                                       ("quad4_visco_hencky_pstress", 3, 3), ("quad4_visco_swanson_pstress", 3, 3),
                                       ("tri3_visco_neohookean_pstress_bounded", 3, 4)])
def test_plane_stress_with_state(cuda_device, name, n, nt, deterministic):
    """PlaneStress.modify_field_gradient (solid_mechanics.py:49-71) with SimpleFeFv: the per-point root find runs on the
    Cauchy stress of the whole model at fixed old state, so the out-of-plane stretch depends on the displacement gradient
    AND on the old viscous state (k_visco_qp MODE 4); the reverse sweep (MODE 5) adds the implicit-function terms to the
    point flux, to the adjoint of the old state and to the property gradient.  Against the oracle (differentiable Newton
    steps on the same root)."""
    case = make_case(name, n=n, nt=nt, width=12, depth=2)
    L, aux, gw, gp = run_engine_loss(case, "energy", deterministic=deterministic, first_step=1)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy", first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo), (L, Lo)
    assert rel(gw, gwo) < TOL, rel(gw, gwo)
    if gpo.size:
        assert rel(gp, gpo) < TOL, (gp, gpo)
    assert aux[3] == 0.0          # every root find converged


@pytest.mark.parametrize("name,n,nt,shards", [("hex8_visco_neohookean", 4, 4, 2), ("quad4_visco_gent_bounded", 6, 3, 3),
                                              ("quad4_visco_neohookean_pstress", 5, 3, 2)])
def test_state_models_on_element_shards_sum_to_the_whole_mesh(cuda_device, name, n, nt, shards):
    """SURVEY 8e for models with state: the viscous states belong to the quadrature points of a shard's own elements and
    the field at a node is a pure function of (x, t, theta), so the path-dependent EnergyLoss needs no exchange -- the
    shards' (loss, gradient, property gradient) add up to the whole-mesh values (what the packed allreduce does)."""
    from tests.cases import make_shard_engines
    case = make_case(name, n=n, nt=nt, width=12, depth=2)
    L, aux, gw, gp = run_engine_loss(case, "energy", first_step=1)
    engines, _ = make_shard_engines(case, shards, scramble=True)
    Ls, gws, gps = 0.0, 0.0, 0.0
    for eng in engines:
        l, a, g, p = run_engine_loss(case, "energy", eng=eng, first_step=1)
        Ls, gws, gps = Ls + l, gws + g, gps + p
        eng.close()
    assert abs(Ls - L) <= 1e-13 * abs(L)
    assert rel(gws, gw) < 1e-12
    if np.size(gp):
        assert rel(gps, gp) < 1e-12


@pytest.mark.parametrize("nb,kind", [(1, "energy"), (3, "energy_residual_reaction"), (20, "energy"), (20, "energy_residual")])
def test_prony_series_lengths_and_two_time_steps(cuda_device, nb, kind):
    """1, 3 (the reference's test_hyperviscoelastic.py fixture) and PCX_MAX_PRONY = 20 branches, nt = 2 (ONE active step of
    the scan): forward / reverse sweeps loop over the branches, the state arrays scale with them."""
    case = make_case("hex8_visco_neohookean", n=2, nt=2, width=10, depth=2)
    rng = np.random.default_rng(nb)
    case.prony = (list(rng.uniform(0.2, 2.0, nb)), list(10.0 ** rng.uniform(-1.0, 2.0, nb)))
    L, aux, gw, gp = run_engine_loss(case, kind, first_step=1)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind, first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo), (L, Lo)
    assert rel(gw, gwo) < TOL, rel(gw, gwo)


def test_more_than_max_prony_branches_is_refused(cuda_device):
    from pancax_b200._lib import PcxError
    case = make_case("hex8_visco_neohookean", n=2, nt=2, width=10, depth=2)
    case.prony = ([1.0] * 21, [1.0] * 21)
    with pytest.raises((PcxError, ValueError, AssertionError, IndexError)):
        make_engine(case)


@pytest.mark.parametrize("name,q", [("hex8_visco_neohookean", 1), ("quad4_visco_gent", 1), ("quad4_visco_gent", 3), ("tri3_visco_neohookean", 1),
                                    ("quad4_visco_neohookean_pstress", 3)])
@pytest.mark.parametrize("kind", ["energy", "energy_residual_reaction"])
def test_state_models_on_other_quadrature_orders(cuda_device, name, q, kind):
    """Hex8 / Quad4 / Tri3 with 1 point and Quad4 with 9 points per element (quadrature_rules.py:109-245): the per-point
    sweeps and the flux scatter take nq from the parent table."""
    if "pstress" in name and kind != "energy":
        pytest.skip("PlaneStress with state: EnergyLoss only")
    case = make_case(name, n=3, nt=3, width=10, depth=2, q_order=q)
    L, aux, gw, gp = run_engine_loss(case, kind, first_step=1)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind, first_step=1)
    assert abs(L - Lo) <= TOL * abs(Lo), (L, Lo)
    assert rel(gw, gwo) < TOL, rel(gw, gwo)


@pytest.mark.parametrize("name,kind", [("hex8_visco_neohookean", "energy_residual_reaction"), ("quad4_visco_gent_bounded", "energy_residual"),
                                       ("quad4_visco_neohookean_pstress", "energy")])
def test_state_model_losses_repeat_bitwise(cuda_device, name, kind):
    """Deterministic assembly: two evaluations of the same step (fresh engines) agree bit for bit -- fixed-order block
    partials, CSR assembly of the point fluxes, fixed-order weight-gradient reduction."""
    case = make_case(name, n=4, nt=4, width=16, depth=2)
    a = run_engine_loss(case, kind, first_step=1)
    b = run_engine_loss(case, kind, first_step=1)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("mode", [1, 2])
def test_state_model_residual_loss_with_the_tf32_adjoint(cuda_device, mode):
    """PCX_OPT_TF32_ADJOINT (1: mma.sync, 2: tcgen05 contraction) only changes the adjoint through the MLP: the loss and aux
    of the state-model residual + reaction loss stay bit-identical, the weight gradient moves by <= 1e-5 per the stated
    tolerance."""
    case = make_case("hex8_visco_neohookean", n=3, nt=4, width=50, depth=4)
    eng = make_engine(case)
    L, aux, gw, _ = run_engine_loss(case, "energy_residual_reaction", eng=eng, first_step=1)
    eng.set_option("tf32_adjoint", mode)
    L32, aux32, gw32, _ = run_engine_loss(case, "energy_residual_reaction", eng=eng, first_step=1)
    eng.close()
    assert L32 == L and np.array_equal(aux32, aux)
    assert rel(gw32, gw) < 1e-5 and np.any(gw32 != gw)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("hex8_visco_neohookean", 3), ("hex8_visco_hencky", 2), ("quad4_visco_gent_bounded", 4),
                                    ("tri3_visco_swanson", 4), ("quad4_visco_neohookean_pstress", 4),
                                    ("tri3_visco_gent_pstress", 4)])
@pytest.mark.parametrize("deterministic", [True, False])
def test_state_step_matches_oracle(cuda_device, name, n, deterministic):
    """pcx_state_step = the reference's op-level calls for a state-variable model at a GIVEN old state:
    physics.potential_energy(params, domain, t, us, state_old, dt) -> (Pi, state_new) (base.py:349-354), the internal force
    at fixed old state (base.py:356-361) and element_pp(pk1_stress) = jacfwd(energy)(grad_u, theta, state_old, dt)
    (base.py:24-74, mechanics/base.py:112-118), against the oracle's autograd from a non-trivial old state (two steps)."""
    import torch
    from oracle import losses as ol
    from oracle.physics import element_interpolants, potential_energy_with_state
    from tests.cases import oracle_problem
    case = make_case(name, n=n, nt=3, width=12, depth=2)
    eng = make_engine(case, deterministic=deterministic)
    prob = oracle_problem(case)
    w, pr = ol._leaves(prob, case.weights, case.prop_raw)
    fld, phys, pe = ol._build(prob, w, pr)
    coords, conns = torch.as_tensor(prob.coords), torch.as_tensor(prob.conns, dtype=torch.long)
    nb = len(case.prony[0])
    Z = torch.eye(3, dtype=torch.float64).expand(eng.ne, eng.nq, nb, 3, 3).clone()
    z = eng.initial_state()
    assert np.array_equal(z.cpu().numpy(), Z.reshape(eng.ne, eng.nq, nb * 9).numpy())
    prop_raw = case.prop_raw if eng.n_train else None
    for step, dt in ((1, 0.3), (2, 0.7)):
        us = fld(coords, float(case.times[step])).detach().clone().requires_grad_(True)
        X_el, U_el = coords[conns], us[conns]
        _, _, grad_u, JxW = element_interpolants(pe, X_el, U_el)
        grad_u.retain_grad()
        W, Z_new = phys.energy_density_with_state(grad_u, Z, dt)
        Pi = (JxW * W).sum()
        (f,) = torch.autograd.grad(Pi, us, retain_graph=True)
        from oracle import models
        if phys.formulation == "plane_stress":      # PK1 stress at the root of the total sigma_33 at fixed old state
            from oracle.physics import plane_stress_gradient
            H3 = plane_stress_gradient(phys.model, grad_u.detach(), phys.props, energy_fn=lambda H: models.simple_fefv_energy(
                phys.model, H, phys.props, Z, dt, phys.prony[0], phys.prony[1])[0]).detach().clone().requires_grad_(True)
        else:
            H3 = phys.modify_field_gradient(grad_u.detach()).clone().requires_grad_(True)
        psi = models.simple_fefv_energy(phys.model, H3, phys.props, Z, dt, phys.prony[0], phys.prony[1])[0]
        (P,) = torch.autograd.grad(psi.sum(), H3)
        Pi_o, Z_chk = potential_energy_with_state(phys, pe, coords, conns, us.detach(), Z, dt)
        Pi = Pi.detach()
        assert abs(float(Pi_o.detach()) - float(Pi)) <= 1e-14 * abs(float(Pi))
        pi_g, zn, f_g, P_g = eng.state_step(us.detach().numpy(), z, dt, prop_raw)
        assert abs(float(pi_g) - float(Pi)) <= 1e-12 * abs(float(Pi))
        assert rel(zn.cpu().numpy(), Z_new.detach().reshape(eng.ne, eng.nq, nb * 9).numpy()) < 1e-12
        assert rel(f_g.cpu().numpy(), f.numpy()) < 1e-10
        assert rel(P_g.cpu().numpy(), P.numpy()) < 1e-10
        # outputs are optional; the state is advanced the same way
        _, zn2, f2, P2 = eng.state_step(us.detach().numpy(), z, dt, prop_raw, want_force=False, want_P=False)
        assert f2 is None and P2 is None and torch.equal(zn2, zn)
        Z, z = Z_new.detach(), zn
    # dt = 0 (the post-processor's step 0): finite, state unchanged
    pi0, z0, _, P0 = eng.state_step(us.detach().numpy(), z, 0.0, prop_raw)
    assert np.isfinite(float(pi0)) and bool(torch.isfinite(P0).all()) and rel(z0.cpu().numpy(), z.cpu().numpy()) < 1e-14
    with pytest.raises(Exception, match="dt must be"):
        eng.state_step(us.detach().numpy(), z, -1.0, prop_raw)
    eng.close()


@pytest.mark.gpu
def test_state_step_refused_without_state(cuda_device):
    case = make_case("hex8_neohookean", n=2, nt=2, width=12, depth=2)
    eng = make_engine(case)
    with pytest.raises(Exception, match="no state variables"):
        eng.state_step(np.zeros((eng.nn, 3)), np.zeros((eng.ne, eng.nq, 9)), 0.1)
    eng.close()


@pytest.mark.gpu
def test_state_step_is_cuda_graph_capturable(cuda_device):
    """pcx_state_step keeps the ABI's contract (no allocation, synchronisation or host staging inside a compute call):
    captured on a handle that has never run, replayed with new displacements / old state in the caller's buffers."""
    import torch
    case = make_case("hex8_visco_neohookean", n=3, nt=3, width=12, depth=2)
    warm, eng = make_engine(case), make_engine(case)
    w = torch.as_tensor(case.weights, device=eng.device)
    us = warm.field_forward(w, 2).clone()
    z = warm.initial_state()
    _, z1, _, _ = warm.state_step(us * 0.5, z, 0.4)
    ref = [o.clone() for o in warm.state_step(us, z1, 0.3)]
    zbuf = z1.clone()
    torch.cuda.synchronize()
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out = eng.state_step(us, zbuf, 0.3)
    torch.cuda.current_stream().wait_stream(s)
    for o in out:
        o.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(out, ref))
    us.mul_(1.1)
    zbuf.copy_(z)
    g.replay()
    torch.cuda.synchronize()
    ref2 = warm.state_step(us, z, 0.3)
    assert all(torch.equal(a, b) for a, b in zip(out, ref2))
    eng.close()
    warm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["neohookean", "gent", "swanson", "hencky"])
def test_state_step_matches_reference_points(cuda_device, model):
    """pcx_state_step against the reference's OWN SimpleFeFv.energy / pk1_stress at a given old state
    (tests/golden/state_step.npz, produced by oracle/gen_golden_state_step.py from simple_fefv.py:20-40 and
    mechanics/base.py:112-118 on the torch stand-in): one Hex8 element with the affine displacement u = H X, so every
    quadrature point sees grad_u = H; the same old state at every point."""
    import os
    import pancax_b200 as px
    from pancax_b200 import Engine
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "state_step.npz"))
    mesh = px.create_structured_mesh("hex8", 1)
    X = np.asarray(mesh.coords, dtype=np.float64)
    eng = Engine(X, mesh.conns, "hex8", 2, 3, np.linspace(0.0, 1.0, 3))
    moduli, taus = g[f"{model}__prony"]
    eng.set_physics("solid", model, "three_dimensional", [float(p) for p in g[f"{model}__props"]],
                    prony=(list(moduli), list(taus)))
    eng.set_field(4, 3, 8, 2)
    _, JxW, _ = eng.element_geometry(want_grad_N=False, want_xq=False)
    vol = float(JxW.sum())
    for k in range(g[f"{model}__grad_u"].shape[0]):
        H, Z, dt = g[f"{model}__grad_u"][k], g[f"{model}__Z_old"][k], float(g[f"{model}__dt"][k])
        us = X @ H.T
        z_old = np.broadcast_to(Z.reshape(1, 1, -1), (eng.ne, eng.nq, Z.size)).copy()
        Pi, zn, f, P = eng.state_step(us, z_old, dt)
        psi, Zn, Pg = g[f"{model}__psi"][k], g[f"{model}__Z_new"][k], g[f"{model}__P"][k]
        assert abs(float(Pi) - psi * vol) <= 1e-12 * abs(psi * vol)
        assert np.abs(zn.cpu().numpy() - Zn.reshape(1, 1, -1)).max() <= 1e-10        # expm: Pade there, spectral here
        assert np.abs(P.cpu().numpy() - Pg[None, None]).max() <= 1e-10 * max(1.0, np.abs(Pg).max())
        # patch test: a homogeneous stress field is in equilibrium at the interior -- here every node is on the boundary,
        # so check the resultant instead: the nodal forces sum to zero
        assert np.abs(f.cpu().numpy().sum(0)).max() <= 1e-12 * max(1.0, np.abs(Pg).max())
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["neohookean", "gent", "swanson", "hencky"])
def test_plane_stress_state_step_matches_reference_points(cuda_device, model):
    """PlaneStress with a state-variable model against the REFERENCE's own formulation, point by point: the `ps_` arrays
    of tests/golden/state_step.npz hold PlaneStress.modify_field_gradient (solid_mechanics.py:49-71: find_root on
    sigma_33 of the whole model at fixed old state) followed by SimpleFeFv.energy / pk1_stress at that root.  One Quad4
    element with the affine displacement u = H X and the same old state at every quadrature point (k_visco_qp MODE 4 / 5
    through pcx_state_step)."""
    import os
    import pancax_b200 as px
    from pancax_b200 import Engine
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "state_step.npz"))
    mesh = px.create_structured_mesh("quad4", 1)
    X = np.asarray(mesh.coords, dtype=np.float64)
    eng = Engine(X, mesh.conns, "quad4", 2, 2, np.linspace(0.0, 1.0, 3))
    moduli, taus = g[f"{model}__prony"]
    eng.set_physics("solid", model, "plane_stress", [float(p) for p in g[f"{model}__ps_props"]],
                    prony=(list(moduli), list(taus)))
    eng.set_field(3, 2, 8, 2)
    _, JxW, _ = eng.element_geometry(want_grad_N=False, want_xq=False)
    area = float(JxW.sum())
    for k in range(g[f"{model}__ps_grad_u"].shape[0]):
        H, Z, dt = g[f"{model}__ps_grad_u"][k], g[f"{model}__ps_Z_old"][k], float(g[f"{model}__ps_dt"][k])
        z_old = np.broadcast_to(Z.reshape(1, 1, -1), (eng.ne, eng.nq, Z.size)).copy()
        Pi, zn, f, P = eng.state_step(X @ H.T, z_old, dt)
        psi, Zn, Pg = g[f"{model}__ps_psi"][k], g[f"{model}__ps_Z_new"][k], g[f"{model}__ps_P"][k]
        assert abs(float(Pi) - psi * area) <= 1e-10 * abs(psi * area)
        assert np.abs(zn.cpu().numpy() - Zn.reshape(1, 1, -1)).max() <= 1e-10
        assert np.abs(P.cpu().numpy() - Pg[None, None]).max() <= 1e-10 * max(1.0, np.abs(Pg).max())
        assert np.abs(f.cpu().numpy().sum(0)).max() <= 1e-12 * max(1.0, np.abs(Pg).max())
    eng.close()
