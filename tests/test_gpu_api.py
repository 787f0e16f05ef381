"""GPU: (a) the CUDA path against vectors produced by the reference's own source (tests/golden),
(b) the pancax-named host API (Problem / Parameters / losses / Adam / Trainer) end to end against the
oracle on the same inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10

SW = [10.0, 0.93074321, -0.07673672, 0.0, -0.5, 0.10448312, 1.71691036, 0.01]
PROPS = {"neohookean": ("neohookean", [0.833, 0.3846]), "neohookean_stiff": ("neohookean", [1000.0, 1.0]),
         "gent": ("gent", [0.833, 0.3846, 3.0]), "swanson": ("swanson", SW), "hencky": ("hencky", [0.833, 0.3846])}


@pytest.mark.parametrize("elem,mname", [("hex8", "neohookean_stiff"), ("hex8", "gent"), ("hex8", "swanson"),
                                        ("hex8", "hencky"), ("quad4", "neohookean_stiff"), ("quad4", "gent"),
                                        ("quad4", "swanson"), ("quad4", "hencky"), ("tri3", "neohookean")])
def test_cuda_potential_energy_matches_reference_source(cuda_device, elem, mname):
    """Pi computed by the CUDA sweep on the reference's fixture meshes == Pi computed by the
    reference's own potential_energy (physics_kernels/base.py:349-354) through oracle/refshim."""
    from pancax_b200 import Engine
    g, fsg = np.load(os.path.join(G, "potential_energy.npz")), np.load(os.path.join(G, "function_space.npz"))
    ndof = 3 if elem == "hex8" else 2
    eng = Engine(fsg[f"{elem}_coords"], fsg[f"{elem}_conns"], elem, 2, ndof, np.array([0.0, 1.0]))
    model, props = PROPS[mname]
    eng.set_physics("solid", model, "three_dimensional" if elem == "hex8" else "plane_strain", props)
    eng.set_field(ndof + 1, ndof, 8, 1)
    Pi, _, _ = eng.energy_force(g[f"{elem}_us"], want_force=False)
    ref = float(g[f"{elem}_{mname}_Pi"])
    assert abs(float(Pi.cpu()[0]) - ref) <= 1e-11 * abs(ref)
    eng.close()


def test_cuda_geometry_matches_reference_source(cuda_device):
    from pancax_b200 import Engine
    fsg = np.load(os.path.join(G, "function_space.npz"))
    for elem in ("hex8", "quad4", "tri3"):
        eng = Engine(fsg[f"{elem}_coords"], fsg[f"{elem}_conns"], elem, 2, 1, np.array([0.0, 1.0]))
        gN, JxW, _ = eng.element_geometry()
        sel = fsg[f"{elem}_sel"]
        assert np.allclose(gN.cpu().numpy()[sel], fsg[f"{elem}_gradN"], rtol=1e-12, atol=1e-13)
        assert np.allclose(JxW.cpu().numpy()[sel], fsg[f"{elem}_JxW"], rtol=1e-13)
        assert abs(float(JxW.sum()) - float(fsg[f"{elem}_volume"])) < 1e-12
        eng.close()


def _build(elem="quad4", n=6, nt=3, bounded=False, inverse=False, t0=0.25):
    import pancax_b200 as px
    m = px.create_structured_mesh(elem, n, permute_seed=5)
    times = np.linspace(t0, 1.0, nt)
    d = px.VariationalDomain(m, times, q_order=2)
    nd = m.num_dimensions
    G_ = px.BoundedProperty(0.01, 5.0, 0.9) if bounded else 1.0
    K_ = 10.0 if bounded else 1000.0
    phys = px.SolidMechanics(px.NeoHookean(bulk_modulus=K_, shear_modulus=G_),
                             px.PlaneStrain() if nd == 2 else px.ThreeDimensional())
    top, bot = ("nset_1", "nset_3") if nd == 2 else ("nset_5", "nset_3")
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in (top, bot) for c in range(nd)]
    if inverse:
        gd = px.GlobalData.from_arrays(times, 0.3 * times, m.nodeSets[top], 1)
        prob = px.InverseProblem(d, phys, None, gd, dirichlet_bcs=bcs)
    else:
        prob = px.ForwardProblem(d, phys, dirichlet_bcs=bcs)
    params = px.Parameters(prob, key=3, n_layers=3, n_neurons=50,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", nd))
    return prob, params


def _oracle_problem(prob, params):
    from oracle import bvps as obvps
    from oracle import losses as ol
    net = params.fields.networks
    model = prob.physics.constitutive_model
    props = []
    import pancax_b200 as px
    for name in model.properties():
        p = getattr(model, name)
        if isinstance(p, px.BoundedProperty):
            props.append(ol.PropSpec(name, float(p.prop_val[0]), True, p.prop_min, p.prop_max))
        else:
            props.append(ol.PropSpec(name, float(p)))
    gd = getattr(prob, "global_data", None)
    nd = prob.n_dims
    return ol.OracleProblem(
        coords=prob.coords, conns=prob.conns, elem=prob.mesh.elem_type, q_order=2, times=prob.times, kind="solid",
        ndof=prob.n_dofs, mlp=ol.MLPSpec(net.n_inputs, net.n_outputs, net.n_neurons, net.n_layers),
        model=model.name, formulation=prob.physics.formulation.name, props=props,
        bc_func=obvps.UniaxialTensionLinearRamp(0.5, 1.0, "y", nd),
        unknown_idx=prob.dof_manager.unknownIndices,
        reaction_nodes=None if gd is None else gd.reaction_nodes, reaction_dof=0 if gd is None else gd.reaction_dof,
        global_outputs=None if gd is None else gd.outputs)


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def test_energy_loss_through_the_public_api(cuda_device):
    import pancax_b200 as px
    from oracle import losses as ol
    prob, params = _build("hex8", n=4, nt=2, t0=0.0)
    loss_fn = px.EnergyLoss(weight=2.5)
    val, aux = loss_fn(params, prob)                       # value only
    (val2, aux2), grads = loss_fn.value_and_grad(params, prob)
    w = params.fields.networks.weights.cpu().numpy()
    Lo, auxo, gwo, _ = ol.loss_and_grad(_oracle_problem(prob, params), dict(kind="energy", energy_weight=2.5), w)
    assert abs(float(val) - Lo) <= TOL * abs(Lo) and float(val) == float(val2)
    assert abs(float(aux["energy"]) - float(auxo["energy"])) <= TOL * abs(float(auxo["energy"]))
    assert rel(grads.fields.cpu().numpy(), gwo) < TOL


def test_inverse_problem_combined_loss_and_adam(cuda_device):
    """C4-style: EnergyResidualAndReactionLoss + FullFieldDataLoss combined, trainable shear modulus,
    three Adam steps; parameters track the oracle + restated optax Adam."""
    import pancax_b200 as px
    from oracle import losses as ol
    from oracle.optim import AdamState, adam_step
    prob, params = _build("quad4", n=6, nt=3, bounded=True, inverse=True)
    rng = np.random.default_rng(0)
    nb = 256
    inputs = np.column_stack([rng.uniform(0, 1, (nb, 2)), rng.choice(prob.times, nb)])
    outputs = 0.05 * rng.standard_normal((nb, 2))
    loss_fn = px.CombineLossFunctions(
        px.EnergyResidualAndReactionLoss(energy_weight=1.0, residual_weight=25.0, reaction_weight=2.5),
        px.FullFieldDataLoss(weight=10.0), with_props=True)
    opt = px.Adam(loss_fn, learning_rate=1e-3, has_aux=True, transition_steps=2)
    st = opt.init(params)
    op = _oracle_problem(prob, params)
    w = params.fields.networks.weights.cpu().numpy().copy()
    pr = params.prop_raw.cpu().numpy().copy()
    sw, sp = AdamState(w.size), AdamState(pr.size)
    for it in range(3):
        params, st, (loss, aux) = opt.step(params, st, prob, inputs, outputs)
        L1, a1, gw1, gp1 = ol.loss_and_grad(op, dict(kind="energy_residual_reaction", energy_weight=1.0,
                                                     residual_weight=25.0, reaction_weight=2.5), w, pr)
        L2, gw2 = ol.full_field_data_loss_and_grad(op, w, inputs, outputs, weight=10.0)
        assert abs(float(loss) - (L1 + L2)) <= TOL * abs(L1 + L2), it
        assert rel(aux["reactions"].cpu().numpy(), a1["reactions"]) < TOL
        assert "props" in aux and abs(aux["props"]["bulk_modulus"] - 10.0) < 1e-15
        w = adam_step(w, gw1 + gw2, sw, 1e-3, transition_steps=2)
        pr = adam_step(pr, gp1, sp, 1e-3, transition_steps=2)
        assert rel(params.fields.networks.weights.cpu().numpy(), w) < 1e-9, it
        assert rel(params.prop_raw.cpu().numpy(), pr) < 1e-9, it
    assert st.epoch == 3


def test_user_defined_loss_function_of_builtin_losses(cuda_device):
    """loss_functions/utils.py:30-34 as the reference's inverse examples use it (examples/inverse_problems/mechanics/
    path-dependent/script.py:121-127, vanilla/example_v2.py): a Python function that adds the built-in losses.  The
    function is executed, the built-in losses inside it are the fused CUDA calls, the arithmetic between their values is
    differentiated by the chain rule (TracedLoss): same value / aux / gradient as CombineLossFunctions, and weighted /
    product combinations against the hand-applied chain rule; a loss that is not built from them is refused."""
    import torch
    import pancax_b200 as px
    prob, params = _build("quad4", n=6, nt=3, bounded=True, inverse=True)
    rng = np.random.default_rng(1)
    nb = 128
    inputs = np.column_stack([rng.uniform(0, 1, (nb, 2)), rng.choice(prob.times, nb)])
    outputs = 0.05 * rng.standard_normal((nb, 2))
    phys = px.EnergyResidualAndReactionLoss(energy_weight=1.0, residual_weight=25.0, reaction_weight=2.5)
    data = px.FullFieldDataLoss(weight=10.0)

    def loss_function(params, problem, inputs, outputs):
        loss_1, aux_1 = phys(params, problem)
        loss_2, aux_2 = data(params, problem, inputs, outputs)
        aux_1.update(aux_2)
        return loss_1 + loss_2, aux_1

    user = px.UserDefinedLossFunction(loss_function)
    comb = px.CombineLossFunctions(phys, data)
    (Lu, au), gu = user.value_and_grad(params, prob, inputs, outputs)
    (Lc, ac), gc = comb.value_and_grad(params, prob, inputs, outputs)
    assert float(Lu) == float(Lc) and set(au) == set(ac)
    assert torch.equal(gu.fields, gc.fields) and torch.equal(gu.props, gc.props)
    Lv, av = user(params, prob, inputs, outputs)                     # plain call: values only
    assert abs(float(Lv) - float(Lc)) <= 1e-14 * abs(float(Lc)) and not isinstance(Lv, px.loss_functions.TracedLoss)

    def weighted(params, problem, inputs, outputs):
        l1, a1 = phys(params, problem)
        l2, a2 = data(params, problem, inputs, outputs)
        return 0.5 * l1 - l2 / 4.0 + 3.0 + l1 * l2, a1

    (Lw, _), gw = px.UserDefinedLossFunction(weighted).value_and_grad(params, prob, inputs, outputs)
    (L1, _), g1 = phys.value_and_grad(params, prob)
    (L2, _), g2 = data.value_and_grad(params, prob, inputs, outputs)
    L1f, L2f = float(L1), float(L2)
    assert abs(float(Lw) - (0.5 * L1f - L2f / 4.0 + 3.0 + L1f * L2f)) <= 1e-13 * abs(float(Lw))
    exp_w = (0.5 + L2f) * g1.fields + (-0.25 + L1f) * g2.fields
    exp_p = (0.5 + L2f) * g1.props + (-0.25 + L1f) * g2.props
    assert rel(gw.fields.cpu().numpy(), exp_w.cpu().numpy()) < 1e-13 and rel(gw.props.cpu().numpy(), exp_p.cpu().numpy()) < 1e-13
    # the optimiser takes it like any other loss (optimizers.py:77-107)
    opt = px.Adam(user, learning_rate=1e-3, has_aux=True)
    st = opt.init(params)
    params, st, (loss, aux) = opt.step(params, st, prob, inputs, outputs)
    assert torch.isfinite(loss) and "reactions" in aux and st.epoch == 1
    with pytest.raises(NotImplementedError):
        px.UserDefinedLossFunction(lambda p, pr: (torch.zeros(()), {})).value_and_grad(params, prob)


def test_full_field_data_loss_batch_rows_shard_over_ranks(cuda_device):
    """FullFieldDataLoss on a sharded problem takes rows rank, rank + R, ... of the batch with the weight
    scaled by n_r / N, so the ranks' partial (loss, gradient) ADD UP to the whole-batch values in the packed allreduce
    (ranks faked here by swapping the shard descriptor; 5 ranks on 13 rows: ragged and short shards)."""
    import torch
    import pancax_b200 as px
    from pancax_b200.parallel import ShardInfo, ShardPlan
    prob, params = _build("quad4", n=5, nt=3, bounded=False, inverse=True)
    rng = np.random.default_rng(2)
    for nb, R in ((13, 5), (3, 4)):                      # the second: fewer rows than ranks (an empty shard)
        inputs = np.column_stack([rng.uniform(0, 1, (nb, 2)), rng.choice(prob.times, nb)])
        outputs = 0.05 * rng.standard_normal((nb, 2))
        lf = px.FullFieldDataLoss(weight=7.0)
        prob.shard = None
        (L, _), g = lf.value_and_grad(params, prob, inputs, outputs)
        Ls, gs = 0.0, 0.0
        for r in range(R):
            prob.shard = ShardInfo(ShardPlan(None, None, 0, None, R, r), None)
            (l, _), gr = lf.value_and_grad(params, prob, inputs, outputs)
            Ls, gs = Ls + float(l), gs + gr.fields
        prob.shard = None
        assert abs(Ls - float(L)) <= 1e-13 * abs(float(L))
        assert float((gs - g.fields).norm() / g.fields.norm()) < 1e-12


def test_trainer_reduces_energy_and_checkpoints(cuda_device, tmp_path):
    import pancax_b200 as px
    prob, params = _build("quad4", n=8, nt=2, t0=0.0)
    opt = px.Adam(px.EnergyLoss(), learning_rate=1e-3, has_aux=True)
    tr = px.Trainer(prob, opt, log_every=5, serialise_every=10, workdir=str(tmp_path))
    l0 = float(px.EnergyLoss()(params, prob)[0])
    params = tr.train(params, 30)
    l1 = float(px.EnergyLoss()(params, prob)[0])
    assert l1 < l0
    assert os.path.exists(tmp_path / "history" / "history.csv")
    ck = sorted(os.listdir(tmp_path / "checkpoint"))
    assert len(ck) == 3
    w = params.fields.networks.weights.clone()
    params.fields.networks.weights.zero_()
    params.deserialise(str(tmp_path / "checkpoint" / ck[-1]))
    assert float(params.fields.networks.weights.abs().sum()) > 0 and not torch.equal(w, params.fields.networks.weights)


def test_ensemble_step_equals_independent_members(cuda_device, tmp_path):
    """optimizers.py:109-139 (ensemble_step = vmap of value_and_grad + one element-wise optax update): every member of
    the ensemble must follow exactly the trajectory of a single run started from the same key; losses come back
    stacked along axis 0."""
    import pancax_b200 as px
    keys = np.array([[0, 3], [0, 4], [5, 6]], dtype=np.uint32)
    prob, _ = _build("quad4", n=6, nt=2, t0=0.0)
    bc = px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 2)
    ens = px.Parameters(prob, key=keys, dirichlet_bc_func=bc)
    opt = px.Adam(px.EnergyLoss(), learning_rate=1e-3, has_aux=True)
    st = opt.init(ens)
    assert st.is_ensemble
    for _ in range(4):
        ens, st, (loss, aux) = opt.step(ens, st, prob)
    assert loss.shape == (3,) and aux["energy"].shape == (3,) and st.epoch == 4
    for k in range(3):
        single = px.Parameters(prob, key=keys[k], dirichlet_bc_func=bc)
        opt1 = px.Adam(px.EnergyLoss(), learning_rate=1e-3, has_aux=True)
        s1 = opt1.init(single)
        for _ in range(4):
            single, s1, (l1, _) = opt1.step(single, s1, prob)
        assert torch.equal(single.fields.networks.weights, ens.members[k].fields.networks.weights), k
        assert float(l1) == float(loss[k])
    ens.serialise(str(tmp_path / "ens"), 4)
    w = ens.stacked_weights.clone()
    for m in ens.members:
        m.fields.networks.weights.zero_()
    ens.deserialise(str(tmp_path / "ens_0000004.npz"))
    assert torch.equal(ens.stacked_weights, w)


def test_lbfgs_descends_monotonically_and_beats_adam(cuda_device):
    """optimizers.py:189-417 interface; the update rule restates optimistix.LBFGS's published algorithm (parity
    unpinned: third-party, not vendored).  Property checks: every accepted step satisfies the Armijo condition (the
    loss never increases), and on the small Poisson-like energy it gets lower than Adam in the same number of steps."""
    import pancax_b200 as px
    prob, params = _build("quad4", n=8, nt=2, t0=0.0)
    w0 = params.fields.networks.weights.clone()
    opt = px.LBFGS(px.EnergyLoss(), has_aux=True)
    st = opt.init(params)
    losses = []
    for _ in range(25):
        params, st, (loss, aux) = opt.step(params, st, prob)
        losses.append(float(loss))
        if st.opt_state["done"]:
            break
    assert all(b <= a + 1e-14 * abs(a) for a, b in zip(losses, losses[1:])), losses
    assert "energy" in aux and len(st.opt_state["s"]) <= 10
    params.fields.networks.weights.copy_(w0)
    adam = px.Adam(px.EnergyLoss(), learning_rate=1e-3, has_aux=True)
    sa = adam.init(params)
    for _ in range(len(losses)):
        params, sa, (la, _) = adam.step(params, sa, prob)
    assert losses[-1] < float(la)
    with pytest.raises(NotImplementedError):
        opt.init(px.Parameters(prob, key=np.array([[0, 1], [0, 2]], dtype=np.uint32)))


def test_poisson_example_through_the_api(cuda_device):
    """C1: the reference's Poisson example (examples/forward_problems/poisson/variational_example.py)."""
    import math
    import pancax_b200 as px
    from oracle import bvps as obvps
    from oracle import losses as ol
    fsg = np.load(os.path.join(G, "function_space.npz"))
    nodesets = {k[len("quad4_ns_"):]: fsg[k] for k in fsg.files if k.startswith("quad4_ns_")}
    mesh = px.Mesh(fsg["quad4_coords"], fsg["quad4_conns"], "quad4", nodesets)
    times = np.linspace(0.0, 1.0, 2)
    domain = px.VariationalDomain(mesh, times)
    physics = px.Poisson(lambda x: 2 * math.pi ** 2 * np.sin(2.0 * math.pi * x[..., 0]) * np.sin(2.0 * math.pi * x[..., 1]))

    def bc_func(x, t, z):           # per-point, exactly as in the example
        x, y = x[0], x[1]
        return x * (1. - x) * y * (1. - y) * z
    bcs = [px.DirichletBC(nset_name=f"nset_{i}", component=0) for i in (1, 2, 3, 4)]
    problem = px.ForwardProblem(domain, physics, [], bcs, [])
    params = px.Parameters(problem, key=100, dirichlet_bc_func=bc_func)
    assert params.fields.networks.num_weights == 5350
    (loss, aux), grads = px.EnergyLoss().value_and_grad(params, problem)
    op = ol.OracleProblem(coords=mesh.coords, conns=mesh.conns, elem="quad4", q_order=2, times=times, kind="poisson",
                          ndof=1, mlp=ol.MLPSpec(3, 1, 50, 3), bc_func=obvps.poisson_example_bc,
                          source=obvps.poisson_example_source)
    Lo, _, gwo, _ = ol.loss_and_grad(op, dict(kind="energy", energy_weight=1.0),
                                     params.fields.networks.weights.cpu().numpy())
    assert abs(float(loss) - Lo) <= TOL * abs(Lo)
    assert rel(grads.fields.cpu().numpy(), gwo) < TOL


def test_adam_kernel_matches_restated_optax(cuda_device):
    from oracle.optim import AdamState, adam_step
    from pancax_b200 import _lib
    from pancax_b200.engine import _ptr, _stream
    rng = np.random.default_rng(0)
    n = 10007
    p = rng.standard_normal(n)
    st = AdamState(n)
    dp = torch.as_tensor(p).cuda()
    mu, nu = torch.zeros_like(dp), torch.zeros_like(dp)
    lib = _lib.load()
    for k in range(1, 6):
        g = rng.standard_normal(n)
        lr = 1e-2 * 0.99 ** ((k - 1) / 500)
        p = adam_step(p, g, st, 1e-2)
        _lib.check(lib.pcx_adam_step(_ptr(dp), _ptr(torch.as_tensor(g).cuda()), _ptr(mu), _ptr(nu), n, k, lr,
                                     0.9, 0.999, 1e-8, _stream()))
    assert rel(dp.cpu().numpy(), p) < 1e-14


def test_sharded_residual_loss_on_all_visible_gpus(cuda_device):
    """SURVEY 8e on real devices: one element shard per GPU, NCCL, through the host mirror
    (tests/mgpu_residual_check.py).  Needs >= 2 GPUs; with one GPU the same phased path is covered by
    test_gpu_parity.py::test_sharded_residual_losses_sum_to_whole_mesh."""
    import json
    import subprocess
    import sys
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(ng, 8)}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(root, "tests", "mgpu_residual_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


def test_post_processor_writes_reference_layout(cuda_device, tmp_path):
    """PostProcessor.init / write_outputs (post_processor.py:141-354) through the CUDA library;
    the values in the file equal the oracle's fields, the variable names / order are the reference's."""
    import pancax_b200 as px
    from oracle import losses as ol
    from pancax_b200.post_processor import read_exodus_results
    from tests.cases import make_case, oracle_problem
    case = make_case("quad4_neohookean", n=4, nt=3, width=20, depth=2, permute=False)
    mesh = px.Mesh(case.mesh["coords"], case.mesh["conns"], "quad4", case.mesh["nodesets"])
    domain = px.VariationalDomain(mesh, case.times, q_order=2)
    physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_1", "nset_3") for c in range(2)]
    problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=2, n_neurons=20,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 2))
    params.fields.networks.set_weights(case.weights)
    pp = px.PostProcessor(None)
    with pytest.raises(ValueError):
        pp.init(params, problem, str(tmp_path / "x.e"), ["not_a_variable"], [])
    out = str(tmp_path / "out.e")
    pp.init(params, problem, out, ["field_values", "internal_force"], ["deformation_gradient", "I1_bar", "pk1_stress"])
    pp.write_outputs(params, problem)
    pp.close()
    times, nod, el = read_exodus_results(out)
    assert np.array_equal(times, case.times)
    assert list(nod) == ["displ_x", "displ_y", "internal_force_x", "internal_force_y"]
    assert list(el)[:5] == ["F_xx_1", "F_xx_2", "F_xx_3", "F_xx_4", "F_xy_1"] and len(el) == (9 + 1 + 9) * 4
    prob = oracle_problem(case)
    for n, t in enumerate(case.times):
        us = ol.field_values(prob, case.weights, t)
        assert np.abs(nod["displ_y"][n] - us[:, 1]).max() <= 1e-13 * max(1.0, np.abs(us).max())
        _, f = ol.energy_and_force(prob, us)
        assert np.linalg.norm(nod["internal_force_x"][n] - f[:, 0]) <= 1e-10 * np.linalg.norm(f) + 1e-13
        F, I1b, P = ol.element_fields(prob, us)
        assert np.abs(el["F_yy_3"][n] - F[:, 2, 1, 1]).max() <= 1e-13
        assert np.abs(el["I1_bar_2"][n] - I1b[:, 1]).max() <= 1e-12 * np.abs(I1b).max()
        assert np.linalg.norm(el["P_zz_4"][n] - P[:, 3, 2, 2]) <= 1e-10 * np.linalg.norm(P) + 1e-13


@pytest.mark.parametrize("name,kind,first_step", [
    ("quad4_neohookean", "energy_residual_reaction", 0),
    ("quad4_neohookean_pstress", "energy_residual_reaction", 0),     # PlaneStress: root workspace sized at set-up
    ("hex8_visco_neohookean", "energy", 1),                          # SimpleFeFv: state history sized at set-up
    ("hex8_visco_neohookean", "energy_residual_reaction", 1),        # ... with the per-step v_t buffers of the residual losses
    ("quad4_visco_neohookean_pstress", "energy", 1),                 # ... and the per-step roots / scratch of PlaneStress with state
    ("hex8_neohookean", "energy_residual_reaction", 0),
])
def test_loss_and_grad_is_cuda_graph_capturable(cuda_device, name, kind, first_step):
    """The C ABI never allocates, synchronises or reads back inside a compute call (pcx_abi.h), so a whole
    loss+gradient step can be captured in a CUDA graph by the caller.  The captured handle has NEVER run before the
    capture (a second handle of the same case warms the kernels up), so any workspace allocated lazily inside a
    compute call -- or a host buffer staged inside it -- would break the capture."""
    import torch
    from tests.cases import make_case, make_engine
    case = make_case(name, n=4 if name.startswith("hex8") else 6, nt=3, width=20, depth=2)
    if first_step == 0:
        case.times = np.linspace(0.25, 1.0, 3)
    warm, eng = make_engine(case), make_engine(case)
    w = torch.as_tensor(case.weights, device=eng.device)
    pr = torch.as_tensor(case.prop_raw, device=eng.device) if eng.n_train else None
    gout = torch.as_tensor(case.global_outputs, device=eng.device)       # device pointer in pcx_loss_desc (ABI v4)
    out = tuple(torch.empty(k, dtype=torch.float64, device=eng.device) for k in (1, 4 + 3, eng.n_weights, 1))
    kw = dict(global_outputs=gout, energy_weight=1.3, residual_weight=0.7, reaction_weight=2.1, first_step=first_step)
    ref = warm.loss_and_grad(kind, w, pr, **kw)
    ref = [o.clone() for o in ref]
    torch.cuda.synchronize()
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            eng.loss_and_grad(kind, w, pr, out=out, **kw)
    torch.cuda.current_stream().wait_stream(s)
    for o in out:
        o.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1]) and torch.equal(out[2], ref[2])
    # new weights and new global data through the same graph: the captured pointers are the caller's buffers
    w2, gout2 = w * 1.01, gout * 0.5 + 0.01
    w.copy_(w2)
    gout.copy_(gout2)
    g.replay()
    torch.cuda.synchronize()
    L2, a2, g2, _ = warm.loss_and_grad(kind, w2, pr, **dict(kw, global_outputs=gout2))
    assert torch.equal(out[0], L2) and torch.equal(out[1], a2) and torch.equal(out[2], g2)
    eng.close()
    warm.close()


def test_strong_form_workspace_is_reserved_at_setup(cuda_device):
    """pcx_reserve: the strong-form workspace is a set-up allocation; a compute call on more points than reserved is
    refused (PCX_ERR_STATE) instead of allocating."""
    import torch
    from pancax_b200._lib import PcxError, check
    from tests.cases import make_case, make_engine
    import ctypes as C
    case = make_case("poisson_quad4", n=4, nt=2, width=20, depth=2)
    eng = make_engine(case)
    w = torch.as_tensor(case.weights, device=eng.device)
    pts = torch.rand(300, 3, dtype=torch.float64, device=eng.device)
    u = torch.empty(300, 1, dtype=torch.float64, device=eng.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (eng.h, C.c_void_p(w.data_ptr()), C.c_void_p(pts.data_ptr()), None, 300, C.c_void_p(u.data_ptr()), None, None, None, st)
    with pytest.raises(PcxError, match="pcx_reserve"):
        check(eng.lib.pcx_points_jet(*args))
    check(eng.lib.pcx_reserve(eng.h, 1, 300))
    check(eng.lib.pcx_points_jet(*args))
    u_ref = eng.points_forward(w, pts)
    torch.cuda.synchronize()
    assert torch.allclose(u, u_ref, rtol=1e-12, atol=1e-14)
    eng.close()


def test_graph_replayed_training_matches_eager_training(cuda_device, tmp_path):
    """Trainer.train as a pure device loop.  One captured step (loss + gradient + Adam with the
    schedule / counter on the device, pcx_adam_step_sched) replayed N times == N eager opt.step calls."""
    import torch
    import pancax_b200 as px
    finals = []
    for use_graph in (False, True):
        mesh = px.create_structured_mesh("quad4", 6)
        times = np.linspace(0.0, 1.0, 3)
        domain = px.VariationalDomain(mesh, times, q_order=2)
        model = px.NeoHookean(bulk_modulus=10.0, shear_modulus=px.BoundedProperty(0.01, 5.0, 1.0))
        physics = px.SolidMechanics(model, px.PlaneStrain())
        bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_1", "nset_3") for c in range(2)]
        gd = px.GlobalData.from_arrays(times, np.linspace(0.0, 0.1, 3), mesh.nodeSets["nset_1"], 1)
        problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
        params = px.Parameters(problem, key=0, n_layers=2, n_neurons=20,
                               dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.3, 1.0, "y", 2))
        opt = px.Adam(px.EnergyResidualAndReactionLoss(1.0, False, 2.0, 3.0), learning_rate=1e-3, has_aux=True,
                      transition_steps=7, decay_rate=0.9)
        tr = px.Trainer(problem, opt, workdir=str(tmp_path / ("g" if use_graph else "e")), log_every=5,
                        serialise_every=10 ** 9)
        params = tr.train(params, 25, cuda_graph=use_graph)
        torch.cuda.synchronize()
        finals.append((params.fields.networks.weights.cpu().numpy().copy(), params.prop_raw.cpu().numpy().copy()))
    (we, pe), (wg, pg) = finals
    assert np.linalg.norm(wg - we) <= 1e-12 * np.linalg.norm(we)
    assert np.abs(pg - pe).max() <= 1e-12 * max(1.0, np.abs(pe).max())
