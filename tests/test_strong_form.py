"""Strong-form / collocation path.  tests/golden/strong_form.npz was produced by EXECUTING the
reference's own StrongFormResidualLoss / Poisson.strong_form_residual / HeatEquation.strong_form_residual /
BasePhysics.field_{gradients,laplacians,time_derivatives} on the torch-backed jax stand-in
(oracle/gen_golden_ad.py).  not gpu: the oracle (torch.func restatement) against those vectors;
gpu: the CUDA jet kernels through the C ABI against the same vectors."""
import os

import numpy as np
import pytest

from tests.cases import make_case

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "strong_form.npz"))
TOL = 1e-10
RCK = (1.7, 0.9, 0.35)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a - b).ravel())
    return d / n if n > 0 else d


def _case():
    return make_case("poisson_quad4", n=4, nt=2, width=20, depth=3)


def _points(case):
    """All (time step, node) collocation points, time-major: what CollocationDataLoader builds (domains.py)."""
    x = np.concatenate([case.mesh["coords"]] * len(case.times))
    t = np.repeat(case.times, case.mesh["coords"].shape[0])
    return x, t


def test_oracle_jet_matches_reference():
    from oracle import strong_form as sf
    from tests.cases import oracle_problem
    case = _case()
    x = case.mesh["coords"]
    t = np.full(x.shape[0], case.times[-1])
    u, gu, lap, ut = sf.jet_numpy(oracle_problem(case), case.weights, x, t)
    k = "poisson_quad4_n4_nt2_w20_d3__"
    assert rel(u, G[k + "u"]) < 1e-13
    assert rel(gu, G[k + "grad_u"]) < 1e-12
    assert rel(lap, G[k + "lap_u"]) < 1e-11
    assert rel(ut, G[k + "u_t"]) < 1e-12


@pytest.mark.parametrize("kind", ["poisson", "heat"])
def test_oracle_strong_form_loss_matches_reference(kind):
    from oracle import bvps, strong_form as sf
    from tests.cases import oracle_problem
    case = _case()
    x, t = _points(case)
    k = f"{kind}_quad4_n4_nt2_w20_d3__"
    w = 2.5 if kind == "poisson" else 1.5
    L, r, gw = sf.loss_and_grad(oracle_problem(case), case.weights, kind, RCK if kind == "heat" else None, x, t, w,
                                source=bvps.poisson_example_source if kind == "poisson" else None)
    assert rel(r.reshape(len(case.times), -1, 1), G[k + "residuals"]) < 1e-11
    assert abs(L - float(G[k + "loss"])) <= TOL * abs(L)
    assert rel(gw, G[k + "gw"]) < TOL


# ------------------------------------------------------------------------------------------ CUDA path
def poisson_wrapper_tables(x, ndof=1):
    """(n, ndof, 2*(nd+3)) tables of u = x(1-x)y(1-y) z (collocation_example.py:11-13):
    (a, d_j a, lap a, a_t | b, d_j b, lap b, b_t)."""
    X, Y = x[:, 0], x[:, 1]
    tab = np.zeros((x.shape[0], ndof, 2 * (2 + 3)))
    b = X * (1 - X) * Y * (1 - Y)
    tab[:, 0, 5] = b
    tab[:, 0, 6] = (1 - 2 * X) * Y * (1 - Y)
    tab[:, 0, 7] = X * (1 - X) * (1 - 2 * Y)
    tab[:, 0, 8] = -2 * Y * (1 - Y) - 2 * X * (1 - X)
    return tab


def _engine(case):
    from tests.cases import make_engine
    return make_engine(case)


@pytest.mark.gpu
def test_cuda_jet_matches_reference(cuda_device):
    case = _case()
    eng = _engine(case)
    x = case.mesh["coords"]
    pts = np.column_stack([x, np.full(x.shape[0], case.times[-1])])
    u, gu, lap, ut = eng.points_jet(case.weights, pts, poisson_wrapper_tables(x))
    k = "poisson_quad4_n4_nt2_w20_d3__"
    assert rel(u.cpu().numpy(), G[k + "u"]) < 1e-13
    assert rel(gu.cpu().numpy(), G[k + "grad_u"]) < 1e-12
    assert rel(lap.cpu().numpy(), G[k + "lap_u"]) < 1e-11
    assert rel(ut.cpu().numpy(), G[k + "u_t"]) < 1e-12
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["poisson", "heat"])
def test_cuda_strong_form_loss_matches_reference(cuda_device, kind):
    from tests.cases import poisson_source_np
    case = _case()
    eng = _engine(case)
    x, t = _points(case)
    pts = np.column_stack([x, t])
    tab = poisson_wrapper_tables(x)
    k = f"{kind}_quad4_n4_nt2_w20_d3__"
    if kind == "poisson":
        L, r, gw = eng.strong_form_loss_and_grad(case.weights, pts, 0.0, 1.0, 2.5, tab=tab,
                                                 src=poisson_source_np(x)[:, None], want_resid=True)
    else:
        L, r, gw = eng.strong_form_loss_and_grad(case.weights, pts, RCK[0] * RCK[1], RCK[2], 1.5, tab=tab, want_resid=True)
    assert rel(r.cpu().numpy().reshape(len(case.times), -1, 1), G[k + "residuals"]) < 1e-11
    assert abs(float(L.cpu()[0]) - float(G[k + "loss"])) <= TOL * abs(float(G[k + "loss"]))
    assert rel(gw.cpu().numpy(), G[k + "gw"]) < TOL
    # value only
    L2, _, g2 = eng.strong_form_loss_and_grad(case.weights, pts, 0.0, 1.0, 2.5, tab=tab, want_grad=False)
    assert g2 is None
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,width,depth,npts", [("hex8_neohookean", 3, 50, 4, 333), ("quad4_neohookean", 5, 30, 2, 64),
                                                     ("hex8_neohookean", 2, 12, 1, 7), ("quad4_neohookean", 4, 55, 3, 1000)])
def test_cuda_jet_and_heat_loss_match_oracle_on_other_shapes(cuda_device, name, n, width, depth, npts, monkeypatch):
    """nd = 2 / 3, several dofs, ragged batches, identity wrapper, target term; chunked when the jet workspace is small."""
    from oracle import strong_form as sf
    from oracle.losses import _build, _leaves
    from pancax_b200 import Engine
    case = make_case(name, n=n, nt=3, width=width, depth=depth)
    nd, ndof = case.mesh["coords"].shape[1], case.ndof
    if npts == 1000:
        monkeypatch.setenv("PCX_JET_WORKSPACE_GB", "0.004")        # forces several chunks
    eng = Engine(case.mesh["coords"], case.mesh["conns"], case.mesh["elem"], 2, ndof, case.times)
    eng.set_physics("solid", "neohookean", case.formulation, [1000.0, 1.0])
    eng.set_field(*case.mlp)
    rng = np.random.default_rng(2)
    x = rng.uniform(0.0, 1.0, (npts, nd))
    t = rng.uniform(0.0, 1.0, npts)
    pts = np.column_stack([x, t])
    prob = __import__("tests.cases", fromlist=["oracle_problem"]).oracle_problem(case)
    prob.bc_func = None
    u, gu, lap, ut = eng.points_jet(case.weights, pts)
    uo, guo, lapo, uto = sf.jet_numpy(prob, case.weights, x, t)
    assert rel(u.cpu().numpy(), uo) < 1e-13 and rel(gu.cpu().numpy(), guo) < 1e-12
    assert rel(lap.cpu().numpy(), lapo) < 1e-11 and rel(ut.cpu().numpy(), uto) < 1e-12
    target = 0.1 * rng.standard_normal((npts, ndof))
    L, r, gw = eng.strong_form_loss_and_grad(case.weights, pts, 0.7, 0.3, 1.9, target=target, want_resid=True)
    Lo, ro, gwo = sf.loss_and_grad(prob, case.weights, "heat", (0.7, 1.0, 0.3), x, t, 1.9, target=target)
    assert rel(r.cpu().numpy() - target, ro) < 1e-11
    assert abs(float(L.cpu()[0]) - Lo) <= TOL * abs(Lo)
    assert rel(gw.cpu().numpy(), gwo) < TOL
    eng.close()


# ------------------------------------------------------------------------------------------ host mirror
def test_jet_tabulation_of_dirichlet_wrappers():
    """pancax_b200/jets.py: second-order forward mode on the host gives (a, grad a, lap a, a_t, b, ...) of a
    Python wrapper written per point (x[0]) like the reference's, incl. .at[].set and numpy ufuncs."""
    from pancax_b200.jets import tabulate_bc_jet
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 1, (11, 3))
    X, Y, T = pts[:, 0], pts[:, 1], pts[:, 2]
    tab = tabulate_bc_jet(lambda x, t, z: x[0] * (1. - x[0]) * x[1] * (1. - x[1]) * z, pts, 1)
    assert np.abs(tab - poisson_wrapper_tables(pts[:, :2])).max() < 1e-15

    def uniaxial(x, t, z):              # pancax/bvps/uniaxial_tension.py, 2D, direction y
        u = z
        u = u.at[0].set(x[1] * (x[1] - 1.0) * t * z[0] / 1.0 ** 2)
        u = u.at[1].set(x[1] * t * 0.5 / 1.0 + x[1] * (x[1] - 1.0) * t * z[1] / 1.0 ** 2)
        return u
    tab = tabulate_bc_jet(uniaxial, pts, 2)
    assert np.array_equal(tab[:, 1, 0], Y * T * 0.5) and np.array_equal(tab[:, 1, 4], Y * 0.5)      # a, a_t
    assert np.array_equal(tab[:, 0, 5], Y * (Y - 1.0) * T) and np.array_equal(tab[:, 0, 8], 2.0 * T)  # b, lap b
    tab = tabulate_bc_jet(lambda x, t, z: np.sin(x[0]) * np.exp(-t) * z + np.cos(x[1]), pts, 1)
    assert np.abs(tab[:, 0, 3] + np.cos(Y)).max() < 1e-15 and np.abs(tab[:, 0, 9] + np.sin(X) * np.exp(-T)).max() < 1e-15
    with pytest.raises(NotImplementedError):
        tabulate_bc_jet(lambda x, t, z: z * z, pts, 1)
    assert tabulate_bc_jet(None, pts, 1) is None


def test_collocation_domain_points_and_problem_checks():
    import pancax_b200 as px
    mesh = px.create_structured_mesh("quad4", 3)
    dom = px.CollocationDomain(mesh, np.linspace(0.0, 1.0, 3))
    pts = dom.collocation_points()
    assert pts.shape == (3 * 16, 3) and np.array_equal(pts[16:32, 2], np.full(16, 0.5))
    assert np.array_equal(pts[32:, :2], mesh.coords)
    px.ForwardProblem(dom, px.HeatEquation(1.0, 2.0, 3.0))
    px.ForwardProblem(dom, px.Poisson(lambda x: x[..., 0]))
    with pytest.raises(px.problems.DomainPhysicsCompatabilityError):
        px.ForwardProblem(px.VariationalDomain(mesh, np.linspace(0, 1, 2)), px.HeatEquation())
    assert px.HeatEquation(2.0, 3.0, 5.0).strong_form_coefficients() == (6.0, 5.0)


@pytest.mark.gpu
def test_strong_form_loss_through_the_public_api(cuda_device):
    """CollocationDomain + Poisson + StrongFormResidualLoss (strong_form_loss_functions.py:8-27) == the vectors the
    reference's own classes produced; then Adam drives the residual down (heat equation)."""
    import pancax_b200 as px
    from tests.cases import poisson_source_np
    case = _case()
    mesh = px.Mesh(case.mesh["coords"], case.mesh["conns"], "quad4", case.mesh["nodesets"])
    dom = px.CollocationDomain(mesh, case.times)
    problem = px.ForwardProblem(dom, px.Poisson(poisson_source_np))
    params = px.Parameters(problem, key=0, n_layers=3, n_neurons=20,
                           dirichlet_bc_func=lambda x, t, z: x[0] * (1. - x[0]) * x[1] * (1. - x[1]) * z)
    params.fields.networks.set_weights(case.weights)
    lf = px.StrongFormResidualLoss(weight=2.5)
    (L, aux), g = lf.value_and_grad(params, problem)
    k = "poisson_quad4_n4_nt2_w20_d3__"
    assert abs(float(L) - float(G[k + "loss"])) <= TOL * abs(float(G[k + "loss"]))
    assert rel(g.fields.cpu().numpy(), G[k + "gw"]) < TOL
    L2, aux2 = lf(params, problem)
    assert abs(float(L2) - float(L)) <= 1e-14 * abs(float(L)) and abs(float(aux2["residual"]) * 2.5 - float(L)) < 1e-12 * abs(float(L))
    # batch form of the collocation example: mean((residual - outputs)^2)
    pts = dom.collocation_points()[5:25]
    (Lb, _), _ = lf.value_and_grad(params, problem, pts, np.zeros((20, 1)))
    assert np.isfinite(float(Lb))
    # training on the heat equation lowers the residual
    problem = px.ForwardProblem(px.CollocationDomain(mesh, np.linspace(0.0, 1.0, 4)), px.HeatEquation(1.0, 1.0, 0.1))
    params = px.Parameters(problem, key=1, n_layers=2, n_neurons=16,
                           dirichlet_bc_func=lambda x, t, z: np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]) * np.exp(-t) + t * z)
    opt = px.Adam(px.StrongFormResidualLoss(), learning_rate=2e-3)
    st = opt.init(params)
    l0 = None
    for it in range(200):
        params, st, loss = opt.step(params, st, problem)
        l0 = float(loss) if l0 is None else l0
    assert float(loss) < 0.5 * l0


@pytest.mark.gpu
def test_jet_without_the_time_stream_is_bit_identical(cuda_device):
    """The d/dt stream is only carried when somebody asks for it (pcx_points_jet with u_t = NULL, strong-form losses with
    c_t = 0: Poisson): u, grad u and the Laplacian -- and a Poisson-type loss with its gradient -- are bit-identical to the
    six-stream evaluation with the time terms switched off by a zero coefficient downstream."""
    import torch
    from tests.cases import make_case, make_engine
    case = make_case("poisson_hex8", n=3, nt=2, width=20, depth=3)
    eng = make_engine(case)
    rng = np.random.default_rng(4)
    pts = rng.uniform(0.05, 0.95, (777, 4))
    w = torch.as_tensor(case.weights, device=eng.device)
    u6, g6, l6, _ = eng.points_jet(w, pts)
    u5, g5, l5, none = eng.points_jet(w, pts, want=("u", "grad_u", "lap_u"))
    assert none is None and torch.equal(u5, u6) and torch.equal(g5, g6) and torch.equal(l5, l6)
    # loss: c_t = 0 drops the stream; a denormal-small c_t keeps it with no measurable contribution
    L5, _, gw5 = eng.strong_form_loss_and_grad(w, pts, 0.0, 0.7, 2.0)
    L6, _, gw6 = eng.strong_form_loss_and_grad(w, pts, 1e-300, 0.7, 2.0)
    assert float(L5[0]) == float(L6[0]) and float((gw5 - gw6).norm()) <= 1e-15 * float(gw6.norm())
    eng.close()


def test_collocation_data_loader():
    """pancax/domains.py:88-130: rows (x..., t) of every node at every time with zero targets, shuffled FULL batches."""
    import pancax_b200 as px
    mesh = px.create_structured_mesh("quad4", 3)
    dom = px.CollocationDomain(mesh, np.linspace(0.0, 1.0, 3))
    dl = px.CollocationDataLoader(dom, num_fields=2)
    assert len(dl) == 48 and dl.inputs.shape == (48, 3) and dl.outputs.shape == (48, 2) and not dl.outputs.any()
    assert np.array_equal(dl.inputs, dom.collocation_points())
    batches = list(dl.dataloader(10))
    assert len(batches) == 4 and all(b[0].shape == (10, 3) and b[1].shape == (10, 2) for b in batches)
    assert len({tuple(r) for b in batches for r in b[0]}) == 40
