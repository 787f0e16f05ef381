"""CPU: the C-ABI library loads and exports every symbol include/pcx_abi.h declares (no compute
calls), and the host-side mirror of the reference interface behaves like the reference's."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol():
    from pancax_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pcx_abi.h")).read()
    declared = sorted(set(re.findall(r"\b(pcx_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pcx_abi.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.pcx_abi_version() == 5
    assert lib.pcx_last_error() is not None


def test_null_handle_is_an_error_not_a_crash():
    from pancax_b200 import _lib
    lib = _lib.load()
    assert lib.pcx_num_weights(None) == -1
    assert lib.pcx_create(None, None) != 0
    assert b"null" in lib.pcx_last_error()


def test_engine_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pancax_b200 import Engine
    with pytest.raises(RuntimeError, match="no CPU path"):
        Engine(np.zeros((4, 2)), np.zeros((1, 4), dtype=np.int32), "quad4", 2, 2, np.array([0.0, 1.0]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pancax_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                assert "/root/reference" not in src


# ------------------------------------------------------------------ fem
def test_read_reference_fixture_arrays_and_roundtrip(tmp_path):
    import pancax_b200 as px
    g = np.load(os.path.join(G, "function_space.npz"))
    for elem in ("hex8", "quad4", "tri3"):
        nodesets = {k[len(elem) + 4:]: g[k] for k in g.files if k.startswith(f"{elem}_ns_")}
        m = px.Mesh(g[f"{elem}_coords"], g[f"{elem}_conns"], elem, nodesets)
        p = str(tmp_path / f"{elem}.g")
        px.write_exodus_mesh(p, m)
        m2 = px.read_exodus_mesh(p)
        assert m2.elem_type == elem
        assert np.array_equal(m2.coords, m.coords) and np.array_equal(m2.conns, m.conns)
        assert sorted(m2.nodeSets) == sorted(m.nodeSets)
        for k in m.nodeSets:
            assert np.array_equal(m2.nodeSets[k], m.nodeSets[k])
        assert m2.conns.dtype == np.int32 and m2.conns.min() == 0   # connect1 - 1, read_exodus_mesh.py:90-93


def test_structured_mesh():
    import pancax_b200 as px
    m = px.create_structured_mesh("hex8", (3, 4, 5), lengths=(1.0, 2.0, 3.0), origin=(0.0, 0.0, 1.0))
    assert m.num_elements == 60 and m.num_nodes == 4 * 5 * 6
    assert np.allclose(m.coords.min(0), [0, 0, 1]) and np.allclose(m.coords.max(0), [1, 2, 4])
    assert np.allclose(m.coords[m.nodeSets["nset_5"], 1], 2.0) and np.allclose(m.coords[m.nodeSets["nset_1"], 2], 4.0)
    mp = px.create_structured_mesh("quad4", 6, permute_seed=3)
    assert np.allclose(mp.coords[mp.nodeSets["nset_1"], 1], 1.0)
    # element areas are unchanged by the permutation
    X = mp.coords[mp.conns]
    area = 0.5 * np.abs((X[:, 2, 0] - X[:, 0, 0]) * (X[:, 3, 1] - X[:, 1, 1]) - (X[:, 3, 0] - X[:, 1, 0]) * (X[:, 2, 1] - X[:, 0, 1]))
    assert np.allclose(area, 1.0 / 36.0)
    with pytest.raises(ValueError):
        px.Mesh(np.zeros((3, 2)), np.zeros((1, 4), dtype=np.int32), "hex8")


def test_domain_time_checks():
    import pancax_b200 as px
    from pancax_b200.domains import (SimulationTimesNotStrictlyIncreasingException,
                                     SimulationTimesNotUniqueException)
    m = px.create_structured_mesh("quad4", 2)
    with pytest.raises(SimulationTimesNotUniqueException):
        px.VariationalDomain(m, np.array([0.0, 0.5, 0.5]))
    with pytest.raises(SimulationTimesNotStrictlyIncreasingException):
        px.VariationalDomain(m, np.array([0.0, 0.7, 0.5]))
    d = px.VariationalDomain(m, np.linspace(0, 1, 3))
    assert d.q_order == 2 and d.conns.shape == (4, 4)


def _problem(elem="quad4", n=3, nt=3):
    import pancax_b200 as px
    m = px.create_structured_mesh(elem, n)
    d = px.VariationalDomain(m, np.linspace(0, 1, nt))
    nd = m.num_dimensions
    if nd == 2:
        phys = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=px.BoundedProperty(0.01, 5.0, 0)),
                                 px.PlaneStrain())
    else:
        phys = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_1", "nset_3") for c in range(nd)]
    return px.ForwardProblem(d, phys, dirichlet_bcs=bcs)


def test_problem_and_parameters_shapes():
    import pancax_b200 as px
    prob = _problem("quad4")
    assert prob.n_dims == 2 and prob.n_dofs == 2
    assert prob.dof_manager.get_unknown_size() == 2 * 16 - 2 * 8
    params = px.Parameters(prob, key=0, device="cpu")
    net = params.fields.networks
    # test/networks/test_mlp.py, test_fields.py: shapes; default n_layers=3, n_neurons=50 (fields.py:28-29)
    assert (net.n_inputs, net.n_outputs, net.n_neurons, net.n_layers) == (3, 2, 50, 3)
    assert net.num_weights == 3 * 50 + 50 + 2 * (50 * 50 + 50) + 50 * 2 == 5400
    assert px.mlp_num_weights(4, 3, 50, 4) == 8050 and px.mlp_num_weights(3, 1, 50, 3) == 5350
    layers = net.layers()
    assert layers[0][0].shape == (50, 3) and layers[-1][0].shape == (2, 50) and layers[-1][1] is None
    assert float(layers[0][0].abs().max()) <= 1 / np.sqrt(3)
    assert params.prop_raw.numel() == 1
    assert params.fields.normalize_time is True
    assert set(params.properties()) == {"bulk_modulus", "shear_modulus"}
    p3 = _problem("hex8", 2)
    assert px.Parameters(p3, key=0, n_layers=4, device="cpu").fields.networks.num_weights == 8050
    # a static problem (all times equal is impossible: unique) -> single time disables normalisation
    m = px.create_structured_mesh("quad4", 2)
    d1 = px.VariationalDomain(m, np.array([0.0]))
    prob1 = px.ForwardProblem(d1, px.Poisson(lambda x: x[..., 0]), dirichlet_bcs=[])
    assert px.Parameters(prob1, key=0, device="cpu").fields.normalize_time is False


def test_ensemble_parameters_structure():
    """parameters.py:89-100: a 2-D key makes an ensemble, one independently initialised member per key row."""
    import pancax_b200 as px
    prob = _problem("quad4")
    keys = np.array([[0, 1], [0, 2], [7, 7]], dtype=np.uint32)
    params = px.Parameters(prob, key=keys, device="cpu")
    assert params.is_ensemble and params.n_ensemble == 3 and len(params.members) == 3
    W = params.stacked_weights
    assert W.shape == (3, 5400)
    assert not np.array_equal(W[0].numpy(), W[1].numpy()) and not np.array_equal(W[1].numpy(), W[2].numpy())
    single = px.Parameters(prob, key=keys[1], device="cpu")
    assert np.array_equal(single.fields.networks.weights.numpy(), W[1].numpy())      # member k == the single run with key k
    assert params.stacked_prop_raw.shape == (3, 1)
    assert len(params.properties()) == 3
    assert not px.Parameters(prob, key=0, device="cpu").is_ensemble


def test_incompatible_problem_pieces_raise():
    import pancax_b200 as px
    from pancax_b200.problems import DomainPhysicsCompatabilityError
    m = px.create_structured_mesh("quad4", 2)
    d = px.VariationalDomain(m, np.linspace(0, 1, 2))
    with pytest.raises(DomainPhysicsCompatabilityError):
        px.ForwardProblem(d, object())
    from pancax_b200.physics_kernels import PlaneStress
    ps = PlaneStress()                       # solid_mechanics.py:32-71: on the path since PCX_FORM_PLANE_STRESS
    assert ps.n_dimensions == 2 and ps.name == "plane_stress"
    assert px.SolidMechanics(px.NeoHookean(bulk_modulus=10.0, shear_modulus=1.0), ps).n_dofs == 2
    assert px.BlatzKo(shear_modulus=0.4).properties() == ("shear_modulus",) or list(px.BlatzKo(shear_modulus=0.4).properties()) == ["shear_modulus"]
    with pytest.raises(NotImplementedError):
        px.MLP(3, 2, 50, 3, activation="relu", device="cpu")


def test_tabulate_bc_and_refusal_of_non_affine_wrappers():
    import pancax_b200 as px
    from pancax_b200.loss_functions import _wrap_bc
    m = px.create_structured_mesh("quad4", 3)
    times = np.linspace(0, 1, 3)
    f = px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2)
    a, b = px.tabulate_bc(f, m.coords, times, 2)
    z = np.random.default_rng(0).standard_normal((m.num_nodes, 2))
    assert np.allclose(a[2] + b[2] * z, f(m.coords, 1.0, z))
    assert np.all(a[0] == 0) and np.all(b[0] == 0)          # ramp: u(x, 0) == 0 exactly
    with pytest.raises(NotImplementedError):
        px.tabulate_bc(lambda x, t, z: z ** 2, m.coords, times, 2)
    with pytest.raises(NotImplementedError):
        px.tabulate_bc(lambda x, t, z: z[:, ::-1], m.coords, times, 2)   # not diagonal

    # a per-point wrapper written for pancax (jax-style .at[].set) is tabulated through the adapter
    def per_point(xs, t, nn):
        y = xs[1]
        u_out = nn
        u_out = u_out.at[0].set(y * (y - 1.0) * t * nn[0])
        u_out = u_out.at[1].set(y * t * 1.0 + y * (y - 1.0) * t * nn[1])
        return u_out
    a2, b2 = px.tabulate_bc(_wrap_bc(per_point, 2), m.coords, times, 2)
    assert np.allclose(a2, a) and np.allclose(b2, b)


def test_partition_mesh_closure_and_coverage():
    import pancax_b200 as px
    from pancax_b200.parallel import partition_mesh
    m = px.create_structured_mesh("hex8", 4, permute_seed=1)
    seen = np.zeros(m.num_elements, dtype=int)
    total_nodes = 0
    for r in range(3):
        sub, l2g = partition_mesh(m, r, 3)
        lo, hi = (m.num_elements * r) // 3, (m.num_elements * (r + 1)) // 3
        seen[lo:hi] += 1
        assert np.array_equal(l2g[sub.conns], m.conns[lo:hi])          # bit-exact connectivity
        assert np.array_equal(sub.coords, m.coords[l2g])
        for k, v in sub.nodeSets.items():
            assert set(l2g[v]) <= set(m.nodeSets[k])
        total_nodes += sub.num_nodes
    assert np.all(seen == 1) and total_nodes >= m.num_nodes


def test_adam_restatement_first_step_is_lr_sign():
    from oracle.optim import AdamState, adam_step
    g = np.array([0.5, -2.0, 1e-3])
    p = adam_step(np.zeros(3), g, AdamState(3), 1e-3)
    assert np.allclose(p, -1e-3 * np.sign(g), rtol=1e-6)     # |mhat|/sqrt(vhat) = 1 on the first update


def test_exodus_results_writer_roundtrip(tmp_path):
    """post_processor.py:156-354 layout: nodal vals_nod_var{k}, element vals_elem_var{k}eb1, names, times."""
    import pancax_b200 as px
    from pancax_b200.post_processor import output_names, read_exodus_results, write_exodus_results
    mesh = px.create_structured_mesh("quad4", 3)
    times = np.linspace(0.0, 1.0, 4)
    rng = np.random.default_rng(0)
    nod = {"displ_x": rng.standard_normal((4, 16)), "displ_y": rng.standard_normal((4, 16))}
    el = {f"{c}_{q + 1}": rng.standard_normal((4, 9)) for c in output_names("F", "full_tensor")[:2] for q in range(4)}
    path = str(tmp_path / "out.e")
    write_exodus_results(path, mesh, times, nod, el)
    t2, nod2, el2 = read_exodus_results(path)
    assert np.array_equal(t2, times)
    assert list(nod2) == list(nod) and list(el2) == list(el)       # F_xx_1..F_xx_4, F_xy_1.. (components outer)
    for k in nod:
        assert np.array_equal(nod2[k], nod[k])
    for k in el:
        assert np.array_equal(el2[k], el[k])
    m2 = px.read_exodus_mesh(path)                                  # still a valid mesh file
    assert np.array_equal(m2.conns, mesh.conns) and np.allclose(m2.coords, mesh.coords)
    assert output_names("I1_bar", "scalar") == ("I1_bar",)
    with pytest.raises(ValueError):
        output_names("x", "vector")


def test_roofline_formulas_match_the_survey_numbers():
    """algorithmic work per unit as formulas of the shape (used by bench.py)."""
    from pancax_b200 import roofline as r
    assert r.mlp_pw(4, 3, 50, 4) == 7850 and r.mlp_pw(3, 1, 50, 3) == 5200
    assert r.mlp_forward_flops_per_row(4, 3, 50, 4) == 15700
    assert r.mlp_backward_flops_per_row(4, 3, 50, 4) == 31000
    assert r.activation_bytes_per_row(50, 4) == 1792
    assert r.mlp_padded_dmma_per_tile(50, 4) == dict(NT=7, KS=13, forward_per_8_rows=293, adjoint_per_64_rows=4816)
    assert 700 <= r.element_flops_per_qp("hex8", 3) <= 900 and 1200 <= r.element_flops_per_qp("hex8", 3, True) <= 1400
    assert abs(r.element_bytes_per_qp("hex8", 3, 8) - 16.0) < 1e-12
    total = r.step_flops_per_qp_eval("hex8", 4, 3, 50, 4, 2146689 / 16777216)
    assert 6500 <= total <= 7100          # "~6.8 kflop per qp-eval" for EnergyLoss on Hex8 128^3


def _tanh_rep_numpy(x):
    """The arithmetic of pcx_mlp.cuh:tanh_rep replayed in numpy (fma emulated in 80-bit extended precision, the
    reciprocal seed truncated to ~20 bits like MUFU.RCP64H), with the 2^(j/64) table parsed from the header."""
    src = open(os.path.join(ROOT, "pancax_b200", "csrc", "pcx_mlp.cuh")).read()
    tab = src[src.index("g_exp2_tab[EXP2_TAB_N] = {"):]
    tab = np.array([float.fromhex(t) for t in re.findall(r"0x1\.[0-9a-f]+p\+0", tab[:tab.index("};")])])
    assert tab.size == 128 and abs(tab[64] - np.sqrt(2.0)) < 1e-15
    T64 = tab[::2]
    L = np.longdouble

    def fma(a, b, c):
        return np.asarray(L(a) * L(b) + L(c), dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    a = np.minimum(np.abs(x), 22.0 + 2.0 ** -16)          # integer clamp on the high word: at most 22 + 2^-16
    MAGIC, C = 6755399441055744.0, float.fromhex("-0x1.71547652b82fep+7")
    t = fma(a, C, MAGIC)
    k = (t - MAGIC).astype(np.int64)
    s = fma(a, C, -k.astype(np.float64))
    coef = [float.fromhex(h) for h in ("0x1.5d87fe78a6731p-40", "0x1.3b2ab6fba4e77p-31", "0x1.c6b08d704a0c0p-23",
                                       "0x1.ebfbdff82c58fp-15", "0x1.62e42fefa39efp-7")]
    p = np.full_like(x, coef[0])
    for c in coef[1:] + [1.0]:
        p = fma(p, s, c)
    m = T64[k & 63] * p
    n = k >> 6
    E = np.ldexp(m, n)
    den = 1.0 + E
    r0 = np.ldexp(np.floor(np.ldexp(1.0 / den, 20)), -20)      # ~2^-20 seed
    e = fma(-den, r0, 1.0)
    r1 = fma(r0, fma(e, e, e), r0)
    q = fma(-2.0 * E, r1, 1.0)
    return np.copysign(q, x)


def test_forward_kernel_tanh_arithmetic_is_accurate_to_3e16():
    xs = np.concatenate([np.linspace(-25.0, 25.0, 200001), np.array([0.0, -0.0, 1e-300, -1e-300, 1e-9, 40.0, -1e6, 21.9999, 22.0001]),
                         np.random.default_rng(0).standard_normal(100000) * 3.0])
    ref = np.tanh(xs.astype(np.longdouble)).astype(np.float64)
    err = np.abs(_tanh_rep_numpy(xs) - ref)
    assert err.max() <= 3.0e-16, err.max()


def test_data_containers_from_csv(tmp_path):
    """pancax/data.py:11-35,86-105,118-243: CSV headers are stripped, n_time_steps counts the distinct times, the loader
    yields shuffled FULL batches, GlobalData resamples the record onto n_time_steps and refuses repeated / decreasing times."""
    import pancax_b200 as px
    from pancax_b200.data import (FullFieldData, FullFieldDataLoader, GlobalData, GlobalDataTimesNotStrictlyIncreasingException,
                                  GlobalDataTimesNotUniqueException)
    rng = np.random.default_rng(0)
    ff = tmp_path / "dic.csv"
    rows = np.column_stack([rng.uniform(0, 1, (30, 2)), np.repeat([0.0, 0.5, 1.0], 10), rng.standard_normal((30, 2))])
    ff.write_text(" x , y ,t, u_x ,u_y\n" + "\n".join(",".join(f"{v:.17g}" for v in r) for r in rows))
    data = FullFieldData(str(ff), ["x", "y", "t"], ["u_x", "u_y"])
    assert len(data) == 30 and data.n_time_steps == 3 and data.inputs.flags.c_contiguous
    # pandas' default (fast) float parser, like the reference: within an ulp or two of the written value
    assert np.allclose(data.inputs, rows[:, :3], rtol=1e-13, atol=1e-15) and np.allclose(data.outputs, rows[:, 3:], rtol=1e-13, atol=1e-15)
    batches = list(FullFieldDataLoader(data).dataloader(8))
    assert len(batches) == 3 and all(b[0].shape == (8, 3) and b[1].shape == (8, 2) for b in batches)
    seen = np.concatenate([b[0] for b in batches])
    assert len({tuple(r) for r in seen}) == 24                      # no row twice within an epoch
    mesh = px.create_structured_mesh("quad4", 3)
    gd_file = tmp_path / "global.csv"
    gd_file.write_text("time, disp , force\n0.0,0.0,0.0\n1.0,0.1,2.0\n2.0,0.2,3.0\n")
    gd = GlobalData(str(gd_file), "time", "disp", "force", mesh, 1, "y", 5, interpolate=True)
    assert gd.n_time_steps == 5 and gd.reaction_dof == 1 and gd.n_nodes == len(mesh.nodeSets[list(mesh.nodeSets)[0]])
    assert np.allclose(gd.times, np.linspace(0, 2, 5)) and np.allclose(gd.outputs, [0.0, 1.0, 2.0, 2.5, 3.0])
    assert np.allclose(gd.displacements, np.linspace(0, 0.2, 5))
    with pytest.raises(ValueError):
        GlobalData(str(gd_file), "time", "disp", "force", mesh, 1, 1, 5)
    gd_file.write_text("time,disp,force\n0.0,0.0,0.0\n1.0,0.1,2.0\n1.0,0.2,3.0\n")
    with pytest.raises(GlobalDataTimesNotUniqueException):
        GlobalData(str(gd_file), "time", "disp", "force", mesh, 1, "x", 3)
    gd_file.write_text("time,disp,force\n0.0,0.0,0.0\n2.0,0.1,2.0\n1.0,0.2,3.0\n")
    with pytest.raises(GlobalDataTimesNotStrictlyIncreasingException):
        GlobalData(str(gd_file), "time", "disp", "force", mesh, 1, "x", 3)


def test_traced_loss_arithmetic_is_the_chain_rule():
    """loss_functions.TracedLoss (what UserDefinedLossFunction differentiates): + - * / between traced values and numbers."""
    import torch
    from pancax_b200.loss_functions import Grads, TracedLoss
    a = TracedLoss(torch.tensor(2.0, dtype=torch.float64), Grads(torch.tensor([1.0, 0.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64)))
    b = TracedLoss(torch.tensor(-3.0, dtype=torch.float64), Grads(torch.tensor([0.0, 2.0], dtype=torch.float64), torch.tensor([1.0], dtype=torch.float64)))
    f = 2.0 * a - b / 4.0 + 1.5 - a * b + (1.0 - a) / b
    va, vb = 2.0, -3.0
    assert abs(float(f) - (2 * va - vb / 4 + 1.5 - va * vb + (1 - va) / vb)) < 1e-15
    dfa = 2.0 - vb - 1.0 / vb
    dfb = -0.25 - va - (1 - va) / vb ** 2
    assert torch.allclose(f.grads.fields, dfa * a.grads.fields + dfb * b.grads.fields, rtol=1e-15, atol=1e-15)
    assert torch.allclose(f.grads.props, dfa * a.grads.props + dfb * b.grads.props, rtol=1e-15, atol=1e-15)
    assert isinstance(sum([a, b]), TracedLoss) and float(-a) == -2.0


def test_simple_fefv_state_bookkeeping():
    """test/constitutive_models/test_hyperviscoelastic.py:18-118 of the reference: initial_state is one identity per
    Prony branch, extract_tensor slices the flat state vector branch by branch."""
    import pancax_b200 as px
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=0.855),
                          px.PronySeries(moduli=[1.0, 2.0, 3.0], relaxation_times=[1.0, 10.0, 100.0]),
                          px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
    assert model.num_prony_terms() == 3 and model.num_state_variables == 27
    assert np.array_equal(model.initial_state(), np.tile([1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0], 3))
    state = np.linspace(1.0, 27.0, 27)
    for k in range(3):
        assert np.array_equal(model.extract_tensor(state, 9 * k), np.arange(9 * k + 1.0, 9 * k + 10.0).reshape(3, 3))
        assert np.array_equal(state.reshape((3, 3, 3))[k], model.extract_tensor(state, 9 * k))
    assert model.extract_scalar(state, 4) == 5.0 and model.bulk_modulus == 1000.0


def test_history_writers_and_file_finders(tmp_path):
    """pancax/history_writer.py:44-151 and utils.py:14-55: the loss history of a training loop as CSV (scalar aux entries,
    property_k columns from ``props``, one file per ensemble member) and the example scripts' file look-up."""
    import torch
    import pancax_b200 as px
    hw = px.HistoryWriter(str(tmp_path / "hist" / "history.csv"), log_every=2, write_every=4)
    for epoch in range(6):
        loss = (torch.tensor(1.0 / (epoch + 1), dtype=torch.float64),
                {"energy": torch.tensor(2.0 * epoch, dtype=torch.float64), "reactions": torch.zeros(3),
                 "props": {"bulk_modulus": 10.0, "shear_modulus": torch.tensor(0.5 + epoch)}})
        hw.write_history(loss, epoch)
    rows = (tmp_path / "hist" / "history.csv").read_text().strip().splitlines()
    assert rows[0] == "energy,property_0,property_1,epoch,loss" and len(rows) == 1 + 3          # epochs 0, 2, 4 written at 4
    assert rows[-1].split(",") == ["8.0", "10.0", "4.5", "4", "0.2"]
    ew = px.EnsembleHistoryWriter(str(tmp_path / "ens"), 2, log_every=1, write_every=1)
    ew.write_history((torch.tensor([1.0, 2.0]), {"energy": torch.tensor([3.0, 4.0])}), 0)
    assert (tmp_path / "ens_1.csv").read_text().strip().splitlines() == ["energy,epoch,loss", "4.0,0,2.0"]
    # file finders look next to the CALLING file and in its data / mesh sub-directory
    here = os.path.dirname(os.path.abspath(__file__))
    assert str(px.find_data_file("cases.py")) == os.path.join(here, "cases.py")          # next to the caller
    assert str(px.find_mesh_file("cases.py")) == os.path.join(here, "cases.py")
    with pytest.raises(px.utils.MeshFileNotFoundException):
        px.find_mesh_file("no_such_mesh.g")
    with pytest.raises(px.utils.DataFileNotFoundException):
        px.find_data_file("no_such_data.csv")
