"""GPU tests of the TF32 tensor-core adjoint (PCX_OPT_TF32_ADJOINT; BASELINE.json north_star item 4 and
config C5 "fp64 vs TF32 tensor-core adjoint").

Stated tolerance of the mode: loss and aux are bit-identical to the fp64 mode (the forward pass, the element
sweeps and the reductions are untouched); every leaf of d loss / d weights is within TOL_TF32 = 1e-5 relative
(2-norm) of the fp64 path / the oracle / the reference-executed golden vectors -- inside the 1e-4 north_star
allows for TF32 mode.  Property gradients do not pass through the MLP adjoint and stay at 1e-10.

Every test runs for both kernels of the mode: tf32_adjoint = 1 (mma.sync, pcx_mlp_tf32.cuh) and tf32_adjoint = 2
(weight-gradient contractions on tcgen05.mma kind::tf32 with TMEM accumulators, pcx_mlp_umma.cuh; widths 49..55 with
2..4 hidden layers, anything else in mode 2 runs the mma.sync kernel)."""
import numpy as np
import pytest

from tests.cases import make_case, make_engine, run_engine_loss, run_oracle_loss

pytestmark = pytest.mark.gpu
TOL_TF32 = 1e-5


def leaves(case):
    """(name, slice) of every equinox leaf in the flat weight vector."""
    n_in, n_out, width, depth = case.mlp
    sizes = [n_in] + [width] * depth + [n_out]
    out, o = [], 0
    for i in range(len(sizes) - 1):
        nW = sizes[i] * sizes[i + 1]
        out.append((f"W{i}", slice(o, o + nW)))
        o += nW
        if i < len(sizes) - 2:
            out.append((f"b{i}", slice(o, o + sizes[i + 1])))
            o += sizes[i + 1]
    assert o == case.weights.size
    return out


def leaf_errors(case, g, g_ref):
    return {nme: np.linalg.norm(g[sl] - g_ref[sl]) / max(np.linalg.norm(g_ref[sl]), 1e-300) for nme, sl in leaves(case)}


@pytest.fixture(params=[1, 2], ids=["mma_sync", "tcgen05"])
def tf32_mode(request):
    return request.param


def both_modes(case, kind, deterministic=True, mode=1):
    eng = make_engine(case, deterministic)
    out = {}
    for m in (0, mode):
        eng.set_option("tf32_adjoint", m)
        out[m] = run_engine_loss(case, kind, eng=eng)
    eng.close()
    return out[0], out[mode]


@pytest.mark.parametrize("name,n,nt,kind", [
    ("hex8_neohookean", 6, 2, "energy"),
    ("hex8_gent", 7, 3, "energy"),                      # 3 time steps, 512 nodes: several partial 128-row tiles
    ("hex8_swanson", 5, 2, "energy_residual"),
    ("hex8_hencky", 5, 2, "energy"),
    ("quad4_neohookean", 19, 11, "energy_residual_reaction"),   # C2-like: nt = 11, 400 nodes per step
    ("quad4_gent_bounded", 9, 2, "energy_residual_reaction"),
    ("tri3_neohookean_bounded", 8, 2, "energy_residual"),
    ("poisson_quad4", 23, 2, "energy"),
    ("poisson_hex8", 5, 2, "energy"),
])
def test_tf32_adjoint_matches_fp64_mode(cuda_device, tf32_mode, name, n, nt, kind):
    case = make_case(name, n=n, nt=nt)
    (L0, aux0, gw0, gp0), (L1, aux1, gw1, gp1) = both_modes(case, kind, mode=tf32_mode)
    assert L1 == L0 and np.array_equal(aux1, aux0), "loss / aux must not depend on the adjoint mode"
    if gp0.size:
        assert np.allclose(gp1, gp0, rtol=1e-10, atol=0), "property gradients do not pass through the MLP adjoint"
    errs = leaf_errors(case, gw1, gw0)
    assert max(errs.values()) < TOL_TF32, errs
    assert np.any(gw1 != gw0), "TF32 mode returned the fp64 bits: the option did not take effect"


@pytest.mark.parametrize("name,kind", [("hex8_neohookean", "energy"), ("quad4_neohookean", "energy_residual_reaction")])
def test_tf32_adjoint_matches_oracle(cuda_device, tf32_mode, name, kind):
    case = make_case(name, n=5, nt=2)
    eng = make_engine(case)
    eng.set_option("tf32_adjoint", tf32_mode)
    L, aux, gw, gp = run_engine_loss(case, kind, eng=eng)
    eng.close()
    Lo, auxo, gwo, gpo = run_oracle_loss(case, kind)
    assert abs(L - Lo) <= 1e-10 * abs(Lo)
    errs = leaf_errors(case, gw, gwo)
    assert max(errs.values()) < TOL_TF32, errs


@pytest.mark.parametrize("width,depth", [(7, 1), (15, 2), (24, 3), (31, 5), (49, 2), (50, 4), (55, 3), (50, 3), (52, 5)])
def test_tf32_adjoint_network_shapes(cuda_device, tf32_mode, width, depth):
    """Every tile count the kernel is instantiated for (NT = 2, 4, 7) and depths 1..5."""
    case = make_case("quad4_neohookean", n=12, nt=2, width=width, depth=depth)
    (L0, aux0, gw0, _), (L1, aux1, gw1, _) = both_modes(case, "energy", mode=tf32_mode)
    assert L1 == L0
    errs = leaf_errors(case, gw1, gw0)
    assert max(errs.values()) < TOL_TF32, (width, depth, errs)


def test_tf32_adjoint_refuses_wide_networks(cuda_device):
    from pancax_b200._lib import PcxError
    case = make_case("quad4_neohookean", n=4, nt=2, width=60, depth=2)
    eng = make_engine(case)
    with pytest.raises(PcxError, match="TF32 adjoint"):
        eng.set_option("tf32_adjoint", 1)
    eng.close()


def test_tf32_adjoint_with_and_without_null_step_culling(cuda_device, tf32_mode):
    case = make_case("hex8_neohookean", n=6, nt=3)
    eng = make_engine(case)
    eng.set_option("tf32_adjoint", tf32_mode)
    res = {}
    for cull in (1, 0):
        eng.set_option("cull_null_steps", cull)
        res[cull] = run_engine_loss(case, "energy", eng=eng)
    eng.close()
    assert res[1][0] == res[0][0]
    errs = leaf_errors(case, res[1][2], res[0][2])
    assert max(errs.values()) < TOL_TF32, errs


def test_tf32_adjoint_atomic_assembly_and_repeatability(cuda_device, tf32_mode):
    case = make_case("hex8_gent", n=6, nt=2)
    eng = make_engine(case, deterministic=True)
    eng.set_option("tf32_adjoint", tf32_mode)
    a = run_engine_loss(case, "energy", eng=eng)
    b = run_engine_loss(case, "energy", eng=eng)
    eng.close()
    assert a[0] == b[0] and np.array_equal(a[2], b[2]), "deterministic assembly + fixed-order partials: bitwise repeatable"
    (L0, _, gw0, _), (L1, _, gw1, _) = both_modes(case, "energy", deterministic=False, mode=tf32_mode)
    errs = leaf_errors(case, gw1, gw0)
    assert max(errs.values()) < TOL_TF32, errs


def test_tf32_field_backward_op(cuda_device, tf32_mode):
    """op-level VJP (pcx_field_backward) in TF32 mode against the fp64 mode on a random cotangent."""
    case = make_case("hex8_neohookean", n=5, nt=3)
    eng = make_engine(case)
    rng = np.random.default_rng(3)
    g = rng.standard_normal((case.mesh["coords"].shape[0], case.ndof))
    ref = eng.field_backward(case.weights, 2, g).cpu().numpy()
    eng.set_option("tf32_adjoint", tf32_mode)
    got = eng.field_backward(case.weights, 2, g).cpu().numpy()
    eng.close()
    errs = leaf_errors(case, got, ref)
    assert max(errs.values()) < TOL_TF32, errs


def test_tf32_full_field_data_loss(cuda_device, tf32_mode):
    """FullFieldDataLoss batch (1 024 rows, data_loss_functions.py:13-24) through pcx_field_data_loss_and_grad."""
    case = make_case("quad4_neohookean", n=8, nt=2)
    eng = make_engine(case)
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 1, size=(1024, 3))
    outs = 0.05 * rng.standard_normal((1024, 2))
    l0, g0 = eng.field_data_loss_and_grad(case.weights, pts, outs, weight=3.0)
    eng.set_option("tf32_adjoint", tf32_mode)
    l1, g1 = eng.field_data_loss_and_grad(case.weights, pts, outs, weight=3.0)
    l0, l1, g0, g1 = float(l0.cpu()[0]), float(l1.cpu()[0]), g0.cpu().numpy(), g1.cpu().numpy()
    eng.close()
    assert l0 == l1
    errs = leaf_errors(case, g1, g0)
    assert max(errs.values()) < TOL_TF32, errs


def test_tcgen05_adjoint_large_batch(cuda_device):
    """Many 128-row tiles per CTA (TMEM accumulators flushed several times, both h ring buffers recycled, a ragged last
    tile): Hex8 40^3 (68 921 nodes x 2 steps = 1 077 tiles), tcgen05 kernel against the fp64 mode."""
    case = make_case("hex8_neohookean", n=40, nt=2)
    (L0, _, gw0, _), (L2, _, gw2, _) = both_modes(case, "energy", mode=2)
    assert L2 == L0
    errs = leaf_errors(case, gw2, gw0)
    assert max(errs.values()) < TOL_TF32, errs
    a = both_modes(case, "energy", mode=2)[1]
    assert np.array_equal(a[2], gw2), "one owner thread per partial entry, reductions in program order: bitwise repeatable"
