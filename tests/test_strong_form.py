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
