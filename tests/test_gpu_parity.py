"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded
inputs.  Tolerance: 1e-10 relative in fp64 (BASELINE.json north_star); indices bit-exact."""
import numpy as np
import pytest
import torch

from tests.cases import (make_case, make_engine, oracle_problem, run_engine_loss, run_oracle_loss)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


@pytest.mark.parametrize("elem,n,q", [("hex8", 5, 2), ("hex8", 4, 1), ("quad4", 9, 2), ("quad4", 6, 3),
                                      ("quad4", 7, 1), ("tri3", 8, 2), ("tri3", 8, 1)])
def test_element_geometry(cuda_device, elem, n, q):
    from oracle import fem as ofem
    from oracle.physics import element_geometry
    from pancax_b200 import Engine
    mesh = ofem.structured_mesh(elem, n, permute_seed=3)
    rng = np.random.default_rng(0)
    mesh["coords"] = mesh["coords"] + 0.02 * rng.standard_normal(mesh["coords"].shape)  # distort
    eng = Engine(mesh["coords"], mesh["conns"], elem, q, 1, np.array([0.0, 1.0]))
    gN, JxW, xq = eng.element_geometry()
    pe = ofem.ParentElement(elem, q)
    X = torch.as_tensor(mesh["coords"])[torch.as_tensor(mesh["conns"], dtype=torch.long)]
    gN_o, JxW_o, _ = element_geometry(pe, X)
    xq_o = torch.einsum("qa,eaj->eqj", torch.as_tensor(pe.N), X)
    assert rel(gN.cpu().numpy(), gN_o.numpy()) < 1e-13
    assert rel(JxW.cpu().numpy(), JxW_o.numpy()) < 1e-13
    assert rel(xq.cpu().numpy(), xq_o.numpy()) < 1e-14
    eng.close()


@pytest.mark.parametrize("name", ["hex8_neohookean", "quad4_neohookean", "poisson_quad4", "tri3_gent"])
def test_field_forward_backward(cuda_device, name):
    from oracle import losses as ol
    case = make_case(name, n=5, nt=3)
    eng = make_engine(case)
    prob = oracle_problem(case)
    rng = np.random.default_rng(1)
    for t_idx in (0, 2):
        us = eng.field_forward(case.weights, t_idx).cpu().numpy()
        us_o = ol.field_values(prob, case.weights, case.times[t_idx])
        assert rel(us, us_o) < 1e-13, (name, t_idx)
        # VJP against torch autograd
        g = rng.standard_normal(us.shape)
        gw = eng.field_backward(case.weights, t_idx, g).cpu().numpy()
        w = torch.as_tensor(case.weights).clone().requires_grad_(True)
        pr = torch.as_tensor(case.prop_raw)
        fld, _, _ = ol._build(prob, w, pr)
        u = fld(torch.as_tensor(case.mesh["coords"]), float(case.times[t_idx]))
        (gw_o,) = torch.autograd.grad((u * torch.as_tensor(g)).sum(), w)
        assert rel(gw, gw_o.numpy()) < 1e-12, (name, t_idx)
    eng.close()


@pytest.mark.parametrize("name", ["hex8_neohookean", "hex8_gent", "hex8_swanson", "hex8_hencky", "hex8_blatz_ko",
                                  "quad4_neohookean", "quad4_gent", "quad4_swanson", "quad4_hencky", "quad4_blatz_ko",
                                  "quad4_blatz_ko_pstress",
                                  "tri3_neohookean", "poisson_quad4", "poisson_tri3", "poisson_hex8"])
@pytest.mark.parametrize("deterministic", [True, False])
def test_energy_force(cuda_device, name, deterministic):
    from oracle import losses as ol
    case = make_case(name, n=4, nt=2)
    eng = make_engine(case, deterministic=deterministic)
    prob = oracle_problem(case)
    us = ol.field_values(prob, case.weights, 1.0)
    Pi, f, _ = eng.energy_force(us, case.prop_raw if eng.n_train else None)
    Pi_o, f_o = ol.energy_and_force(prob, us)
    assert abs(float(Pi.cpu()[0]) - Pi_o) <= TOL * abs(Pi_o), name
    assert rel(f.cpu().numpy(), f_o) < TOL, name
    eng.close()


@pytest.mark.parametrize("name", ["hex8_neohookean", "hex8_gent", "hex8_swanson", "hex8_hencky", "hex8_blatz_ko",
                                  "quad4_neohookean", "quad4_gent", "quad4_swanson", "quad4_hencky", "quad4_blatz_ko",
                                  "quad4_blatz_ko_pstress", "tri3_hencky", "tri3_neohookean", "poisson_quad4"])
def test_stiffness_action(cuda_device, name):
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
    # symmetry of K: <w, K v> == <v, K w>
    w2 = rng.standard_normal(us.shape)
    Kw, _ = eng.stiffness_action(us, w2)
    a = float((torch.as_tensor(w2) * Kv.cpu()).sum())
    b = float((torch.as_tensor(v) * Kw.cpu()).sum())
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    eng.close()


def test_hencky_tangent_at_repeated_stretches(cuda_device):
    """Hencky's tangent action is the closed-form derivative of the isotropic tensor function (pcx_models.cuh), not
    AD through the eigenvector algorithm: it stays finite and exact where stretches coincide.  At F = I (every shipped
    load ramp at t = 0) the reference's second derivative is NaN; the exact limit is linear elasticity,
    dP = K tr(eps) I + 2 G eps.  Checked against that closed form assembled by the oracle's own B-matrices
    (K v = d/d eps f(eps v) of the small-strain energy), and for symmetry of K."""
    import torch as th
    from oracle import fem as ofem
    from oracle.physics import element_geometry
    case = make_case("hex8_hencky", n=3)
    eng = make_engine(case)
    Kb, G = case.props
    rng = np.random.default_rng(4)
    us = np.zeros((eng.nn, 3))
    v = rng.standard_normal(us.shape)
    Kv, _ = eng.stiffness_action(us, v)
    Kv = Kv.cpu().numpy()
    assert np.all(np.isfinite(Kv))
    # linear elasticity: f = sum_e sum_q JxW B^T sigma(grad v), sigma = K tr(eps) I + 2 G eps
    pe = ofem.ParentElement("hex8", 2)
    conns = th.as_tensor(case.mesh["conns"], dtype=th.long)
    X = th.as_tensor(case.mesh["coords"])[conns]
    gN, JxW, _ = element_geometry(pe, X)                      # (ne,nq,nnpe,nd), (ne,nq)
    V = th.as_tensor(v)[conns]                                 # (ne,nnpe,3)
    H = th.einsum("eai,eqaj->eqij", V, gN)
    eps = 0.5 * (H + H.transpose(-1, -2))
    tr = th.einsum("eqii->eq", eps)
    sig = Kb * tr[..., None, None] * th.eye(3, dtype=th.float64) + 2.0 * G * eps
    fe = th.einsum("eq,eqij,eqaj->eai", JxW, sig, gN)
    f = th.zeros_like(th.as_tensor(v))
    f.index_add_(0, conns.reshape(-1), fe.reshape(-1, 3))
    assert rel(Kv, f.numpy()) < 1e-12
    w2 = rng.standard_normal(us.shape)
    Kw, _ = eng.stiffness_action(us, w2)
    a, b = float((w2 * Kv).sum()), float((v * Kw.cpu().numpy()).sum())
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # two equal stretches exactly (uniaxial state): finite, symmetric, and equal to a central difference of f
    us = np.zeros((eng.nn, 3))
    us[:, 1] = 0.2 * case.mesh["coords"][:, 1]
    Kv, _ = eng.stiffness_action(us, v)
    h = 1e-6
    _, fp, _ = eng.energy_force(us + h * v)
    _, fm, _ = eng.energy_force(us - h * v)
    fd = ((fp - fm) / (2 * h)).cpu().numpy()
    assert rel(Kv.cpu().numpy(), fd) < 1e-7
    eng.close()


@pytest.mark.parametrize("name", ["hex8_neohookean", "hex8_gent", "hex8_swanson", "hex8_hencky",
                                  "quad4_neohookean", "tri3_neohookean", "poisson_quad4"])
def test_energy_loss_and_grad(cuda_device, name):
    case = make_case(name, n=5, nt=3)
    L, aux, gw, gp = run_engine_loss(case, "energy")
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy")
    assert abs(L - Lo) <= TOL * abs(Lo), name
    assert abs(aux[0] - auxo["energy"]) <= TOL * abs(auxo["energy"])
    assert aux[3] == 0.0
    assert rel(gw, gwo) < TOL, name


@pytest.mark.parametrize("name", ["quad4_neohookean", "quad4_gent", "quad4_swanson", "hex8_neohookean",
                                  "tri3_neohookean_bounded", "quad4_gent_bounded", "hex8_hencky", "quad4_hencky_bounded",
                                  "hex8_blatz_ko", "quad4_blatz_ko_bounded"])
@pytest.mark.parametrize("kind", ["energy_residual", "energy_residual_reaction"])
def test_residual_losses(cuda_device, name, kind):
    case = make_case(name, n=5, nt=3)
    case.times = np.linspace(0.25, 1.0, 3)   # F9: parity on steps with ||f|| > 0
    L, aux, gw, gp = run_engine_loss(case, kind)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= TOL * abs(Lo), (name, kind)
    assert abs(aux[1] - auxo["residual"]) <= TOL * abs(auxo["residual"])
    if kind == "energy_residual_reaction":
        assert rel(aux[4:], auxo["reactions"]) < TOL
        assert abs(aux[2] - auxo["global_data_loss"]) <= TOL * abs(auxo["global_data_loss"])
    assert rel(gw, gwo) < TOL, (name, kind)
    if gpo.size:
        assert rel(gp, gpo) < TOL, (name, kind)


def test_residual_loss_at_t0_is_finite(cuda_device):
    """F9: at t = 0 the ramped wrapper gives u == 0, f == 0 exactly; the gradient must stay finite
    (convention d||f||/df := 0) and agree with the oracle using the same convention."""
    case = make_case("quad4_neohookean", n=5, nt=3)
    L, aux, gw, gp = run_engine_loss(case, "energy_residual_reaction")
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy_residual_reaction")
    assert np.all(np.isfinite(gw)) and np.isfinite(L)
    assert abs(aux[4]) < 1e-12
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL


@pytest.mark.parametrize("kind", ["energy_residual", "energy_residual_reaction"])
def test_residual_losses_with_more_than_64_time_steps(cuda_device, kind):
    """The per-step residual coefficients are computed by one thread per time step of the batch: a batch of more
    than 64 steps must fill all of them (a <<<1, 64>>> launch left steps >= 64 at zero)."""
    case = make_case("quad4_neohookean", n=3, nt=70, width=12, depth=2)
    case.times = np.linspace(0.2, 1.0, 70)
    case.global_outputs = np.linspace(0.0, 0.2, 70)
    L, aux, gw, gp = run_engine_loss(case, kind)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert abs(aux[1] - auxo["residual"]) <= TOL * abs(auxo["residual"])
    if kind == "energy_residual_reaction":
        assert rel(aux[4:], auxo["reactions"]) < TOL
        assert np.all(aux[4 + 64:] != 0.0)
    assert rel(gw, gwo) < TOL


def test_time_chunk_not_dividing_nt(cuda_device, monkeypatch):
    case = make_case("quad4_neohookean", n=4, nt=7, width=16, depth=2)
    case.times = np.linspace(0.2, 1.0, 7)
    case.global_outputs = np.linspace(0.0, 0.2, 7)
    ref = run_oracle_loss(case, "energy_residual_reaction")
    monkeypatch.setenv("PCX_TCHUNK", "3")
    out = run_engine_loss(case, "energy_residual_reaction")
    assert abs(ref[0] - out[0]) <= TOL * abs(ref[0])
    assert rel(out[1][4:], ref[1]["reactions"]) < TOL
    assert rel(out[2], ref[2]) < TOL


def test_time_chunking_matches(cuda_device, monkeypatch):
    case = make_case("quad4_neohookean", n=5, nt=5)
    case.times = np.linspace(0.2, 1.0, 5)
    ref = run_engine_loss(case, "energy_residual_reaction")
    monkeypatch.setenv("PCX_TCHUNK", "2")
    out = run_engine_loss(case, "energy_residual_reaction")
    assert abs(ref[0] - out[0]) <= 1e-13 * abs(ref[0])
    assert rel(out[2], ref[2]) < 1e-12


def test_full_field_data_loss(cuda_device):
    from oracle import losses as ol
    case = make_case("tri3_neohookean", n=6, nt=4)
    eng = make_engine(case)
    prob = oracle_problem(case)
    rng = np.random.default_rng(5)
    nb = 1024
    pts = np.column_stack([rng.uniform(0, 1, (nb, 2)), rng.uniform(0, 1, nb)])
    outs = 0.1 * rng.standard_normal((nb, 2))
    from pancax_b200 import bvps
    bc = bvps.UniaxialTensionLinearRamp(case.bc[1], case.bc[2], case.bc[3], 2)
    a = np.empty((nb, 2)); b = np.empty((nb, 2))
    for i in range(nb):   # per-row time
        a[i] = bc(pts[i:i + 1, :2], pts[i, 2], np.zeros((1, 2)))[0]
        b[i] = bc(pts[i:i + 1, :2], pts[i, 2], np.ones((1, 2)))[0] - a[i]
    loss, gw = eng.field_data_loss_and_grad(case.weights, pts, outs, a, b, weight=10.0)
    Lo, gwo = ol.full_field_data_loss_and_grad(prob, case.weights, pts, outs, weight=10.0)
    assert abs(float(loss.cpu()[0]) - Lo) <= TOL * abs(Lo)
    assert rel(gw.cpu().numpy(), gwo) < TOL
    eng.close()


def test_deterministic_bitwise_repeat(cuda_device):
    case = make_case("hex8_neohookean", n=6, nt=2)
    eng = make_engine(case, deterministic=True)
    r1 = run_engine_loss(case, "energy", eng=eng)
    r2 = run_engine_loss(case, "energy", eng=eng)
    assert r1[0] == r2[0]
    assert np.array_equal(r1[2], r2[2])
    eng.close()


@pytest.mark.parametrize("width,depth", [(10, 2), (20, 1), (30, 3), (49, 3), (50, 3), (50, 5), (51, 2), (55, 2), (60, 2)])
def test_mlp_shapes(cuda_device, width, depth):
    case = make_case("quad4_neohookean", n=4, nt=2, width=width, depth=depth)
    L, aux, gw, gp = run_engine_loss(case, "energy")
    Lo, auxo, gwo, gpo = run_oracle_loss(case, "energy")
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL


@pytest.mark.parametrize("name,n,shards,scramble", [("quad4_neohookean", 8, 2, False), ("quad4_gent_bounded", 7, 3, True),
                                                     ("hex8_neohookean", 4, 2, False), ("hex8_swanson", 3, 4, True),
                                                     ("tri3_neohookean_bounded", 6, 3, True)])
@pytest.mark.parametrize("kind", ["energy_residual", "energy_residual_reaction"])
def test_sharded_residual_losses_sum_to_whole_mesh(cuda_device, name, n, shards, scramble, kind):
    """residual / reaction losses on element shards (phased ABI + two exchanges) == the fused
    whole-mesh call, including nodes shared by more than two shards and trainable properties."""
    from pancax_b200 import parallel
    from tests.cases import LOSS_WEIGHTS, make_shard_engines
    case = make_case(name, n=n, nt=3, width=20, depth=2, t0=0.2)
    L, aux, gw, gp = run_engine_loss(case, kind)
    engines, plans = make_shard_engines(case, shards, scramble=scramble)
    assert any(p.n_shared > 0 for p in plans)
    Ls, auxs, gws, gps = parallel.run_shards_in_process(
        engines, plans, kind, case.weights, case.prop_raw if engines[0].n_train else None,
        global_outputs=case.global_outputs, **LOSS_WEIGHTS)
    assert abs(float(Ls.cpu()[0]) - L) <= 1e-11 * abs(L)
    a2 = auxs.cpu().numpy()
    assert rel(a2[:2], aux[:2]) < 1e-11
    if kind == "energy_residual_reaction":
        assert rel(a2[2], aux[2]) < 1e-10 and rel(a2[4:], aux[4:]) < 1e-11
    assert rel(gws.cpu().numpy(), gw) < 1e-10
    if gp.size:
        assert rel(gps.cpu().numpy(), gp) < 1e-10
    for e in engines:
        e.close()


def test_phased_api_call_order_is_enforced(cuda_device):
    from pancax_b200._lib import PcxError
    case = make_case("quad4_neohookean", n=4, nt=2, width=12, depth=2)
    eng = make_engine(case)
    eng.loss_begin("energy_residual", case.weights)
    with pytest.raises(PcxError):
        eng.loss_finish(1)            # mid missing
    with pytest.raises(PcxError):
        eng.loss_begin("energy", case.weights)   # the phased API is for the residual losses only
    eng.close()


@pytest.mark.parametrize("name,kind", [("hex8_neohookean", "energy"), ("quad4_neohookean_bounded", "energy_residual_reaction"),
                                       ("tri3_gent", "energy_residual")])
def test_null_step_culling_is_exact(cuda_device, name, kind):
    """t = 0 of a load ramp has b == 0: u = a and the cotangent on the network output vanishes, so the MLP is
    skipped there (PCX_OPT_CULL_NULL_STEPS).  Same loss / aux / gradients as evaluating every row."""
    case = make_case(name, n=5, nt=4, width=20, depth=2)     # times 0, 1/3, 2/3, 1: step 0 is null
    eng = make_engine(case)
    assert list(eng.null_steps()) == [True, False, False, False]
    outs = []
    for cull in (1, 0):
        eng.set_option("cull_null_steps", cull)
        outs.append(run_engine_loss(case, kind, eng=eng))
        us0 = eng.field_forward(case.weights, 0).cpu().numpy()
        assert np.array_equal(us0, np.zeros_like(us0))
        g0 = eng.field_backward(case.weights, 0, np.ones_like(us0)).cpu().numpy()
        assert np.abs(g0).max() == 0.0
    (L1, a1, g1, p1), (L0, a0, g0, p0) = outs
    assert abs(L1 - L0) <= 1e-13 * abs(L0) and rel(a1, a0) < 1e-13
    assert rel(g1, g0) < 1e-12
    if p0.size:
        assert rel(p1, p0) < 1e-12
    eng.close()


@pytest.mark.parametrize("name", ["hex8_neohookean", "quad4_neohookean", "tri3_neohookean_bounded"])
@pytest.mark.parametrize("kind", ["energy", "energy_residual", "energy_residual_reaction"])
def test_single_element_single_time_step_all_dofs_constrained(cuda_device, name, kind):
    """Smallest inputs: ONE element, ONE time step.  Every node of a one-element mesh sits on a Dirichlet node set, so
    unknownIndices is EMPTY (not missing): R = ||f[[]]|| = 0 with a zero gradient, as jnp.linalg.norm of an empty slice
    gives; energy and reaction terms remain."""
    case = make_case(name, n=1, nt=1, width=9, depth=1, t0=0.5)
    assert case.unknown_idx.size == 0
    L, aux, gw, gp = run_engine_loss(case, kind)
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= TOL * abs(Lo)
    assert rel(gw, gwo) < TOL
    if kind != "energy":
        assert aux[1] == 0.0 and float(auxo["residual"]) == 0.0
    if gpo.size:
        assert rel(gp, gpo) < TOL


def test_full_field_data_loss_single_row(cuda_device):
    """A batch of ONE DIC row (ragged tail batches of a data loader)."""
    import torch
    from oracle import losses as ol
    from tests.cases import oracle_problem
    case = make_case("quad4_neohookean", n=3, nt=2, width=10, depth=2)
    eng = make_engine(case)
    inp, out = np.array([[0.3, 0.6, 0.5]]), np.array([[0.01, -0.02]])
    from pancax_b200 import tabulate_bc
    from tests.cases import engine_bc
    bc = engine_bc(case)
    a = bc(inp[:, :2], inp[:, 2], np.zeros((1, 2)))
    b = bc(inp[:, :2], inp[:, 2], np.ones((1, 2))) - a
    loss, gw = eng.field_data_loss_and_grad(torch.as_tensor(case.weights, device="cuda"), inp, out, a, b, weight=2.0)
    Lo, gwo = ol.full_field_data_loss_and_grad(oracle_problem(case), case.weights, inp, out, weight=2.0)
    assert abs(float(loss[0]) - Lo) <= TOL * abs(Lo) and rel(gw.cpu().numpy(), gwo) < TOL
    eng.close()


@pytest.mark.parametrize("name,kind", [("quad4_neohookean", "energy_residual_reaction"), ("hex8_gent", "energy"),
                                       ("hex8_visco_neohookean", "energy_residual")])
def test_mlp_with_final_bias(cuda_device, name, kind):
    """networks/mlp.py:61-79: ``use_final_bias=True`` (MLPBasis and user networks) -- one more leaf [.., Wout, bout] in the
    equinox order; forward, adjoint and (TF32 mode) the bias gradient of the output layer."""
    from oracle import losses as ol
    from oracle.field import flatten_mlp, init_mlp
    from tests.cases import LOSS_WEIGHTS
    case = make_case(name, n=3, nt=3 if "visco" in name else 2, width=12, depth=2, t0=0.0 if "visco" in name else 0.3)
    n_in, n_out, width, depth = case.mlp
    w = flatten_mlp(init_mlp(n_in, n_out, width, depth, seed=3, use_final_bias=True))
    assert w.size == case.weights.size + n_out
    case.weights, case.mlp = w, (n_in, n_out, width, depth, True)
    first = 1 if "visco" in name else 0
    eng = make_engine(case)
    op = oracle_problem(case)
    op.mlp.use_final_bias = True
    Lo, auxo, gwo, gpo = ol.loss_and_grad(op, dict(kind=kind, first_step=first, **LOSS_WEIGHTS), w, None)
    for mode, tol in ((0, TOL), (1, 1e-5)):
        eng.set_option("tf32_adjoint", mode)
        loss, aux, gw, gp = eng.loss_and_grad(kind, w, None, global_outputs=case.global_outputs, first_step=first, **LOSS_WEIGHTS)
        assert abs(float(loss[0]) - Lo) <= TOL * abs(Lo)
        assert rel(gw.cpu().numpy(), gwo) < tol, (mode, rel(gw.cpu().numpy(), gwo))
        assert rel(gw.cpu().numpy()[-n_out:], gwo[-n_out:]) < max(tol, 1e-5 if mode else TOL)      # the final-bias leaf
    eng.close()
