"""Device-time measurements of BASELINE.json's other configurations (bench.py measures C3), through the public
host-mirror API, in both adjoint modes (fp64 DMMA / TF32 tensor cores).  One JSON line per case:

  C2  2D plane-strain NeoHookean uniaxial tension, Quad4 512^2 (1 048 576 qp), nt = 11, MLP 3->50x3->2,
      EnergyResidualAndReactionLoss (reaction on the top node set) and EnergyAndResidualLoss
  C4  inverse problem, scaled synthetic variant: Quad4 512^2, shear modulus trainable (BoundedProperty),
      EnergyResidualAndReactionLoss(250e9, 250e9) + FullFieldDataLoss(10e9) on a batch of 1 024 synthetic DIC rows
      (analytic uniaxial field + N(0, 1e-4) noise, seed 0) + synthetic global force-displacement data
  C5  3D Swanson / Hencky Hex8 n^3 sweep (1 M .. 128 M qp), EnergyLoss, MLP 4->50x4->3

  R3D 3D residual / reaction loss on Hex8 100^3 (NeoHookean / Swanson / Hencky): the stiffness-action sweep in 3D

    python tools/bench_configs.py [--cases c2,c4,c5,visco,visco2d,r3d,pstress] [--c5-n 50,100,159,252] [--steps 5]

Every (time step, node) row is evaluated (null-step culling off), deterministic assembly, CUDA events on the
launching stream, inputs resident.  Not the bench.py contract: no e2e / CPU arm here."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

SWANSON = dict(bulk_modulus=10.0, A1=0.93074321, P1=-0.07673672, B1=0.0, Q1=0.5, C1=0.10448312, R1=1.71691036,
               cutoff_strain=0.01)   # test/constitutive_models/test_swanson.py:8-14


def timed(step, steps, warmup=3):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def measure(label, problem, params, loss_fn, extra_args, steps, evals):
    eng = get_engine(problem, params)
    eng.set_option("cull_null_steps", 0)
    out = {"case": label, "evals_per_step": int(evals), "elements": int(eng.ne), "nodes": int(eng.nn), "nt": int(eng.nt)}
    grads = {}
    for mode, name in ((0, "fp64"), (1, "tf32_adjoint")):
        eng.set_option("tf32_adjoint", mode)

        def step():
            return loss_fn.value_and_grad(params, problem, *extra_args)
        (loss, aux), g = step()
        grads[name] = (float(loss), g.fields.clone(), g.props.clone())
        eng.profile_enable(True)
        eng.profile_collect()
        ms = timed(step, steps)
        prof = eng.profile_collect()
        eng.profile_enable(False)
        n_steps_prof = steps + 3
        out[name] = {"ms_per_step": ms, "evals_per_s": evals / (ms * 1e-3),
                     "kernel_ms_per_step": {k: round(v[0] / n_steps_prof, 4) for k, v in prof.items() if v[1]}}
    l64, g64, p64 = grads["fp64"]
    l32, g32, p32 = grads["tf32_adjoint"]
    out["loss"] = l64
    out["tf32_vs_fp64"] = {"loss_identical": l64 == l32,
                           "grad_rel_err": float((torch.linalg.norm(g32 - g64) / torch.linalg.norm(g64)).cpu()),
                           "prop_grad_rel_err": float((torch.linalg.norm(p32 - p64) / torch.linalg.norm(p64).clamp_min(1e-300)).cpu())
                           if p64.numel() and float(torch.linalg.norm(p64)) > 0 else 0.0,
                           "speedup": out["fp64"]["ms_per_step"] / out["tf32_adjoint"]["ms_per_step"]}
    print(json.dumps(out), flush=True)
    eng.set_option("tf32_adjoint", 0)
    problem._engines.clear()
    eng.close()


def c2(steps, device):
    n = 512
    mesh = px.create_structured_mesh("quad4", (n, n), lengths=(1.0, 1.0))
    times = np.linspace(0.0, 1.0, 11)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
    gd = px.GlobalData.from_arrays(times, np.zeros(11), mesh.nodeSets["nset_1"], 1)
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
    evals = n * n * 4 * 11
    measure("C2 Quad4 512^2 NeoHookean plane strain, nt=11, EnergyResidualAndReactionLoss, MLP 3->50x3->2",
            problem, params, px.EnergyResidualAndReactionLoss(), (), steps, evals)
    measure("C2 same mesh, EnergyAndResidualLoss", problem, params, px.EnergyAndResidualLoss(), (), steps, evals)


def c4(steps, device):
    n = 512
    mesh = px.create_structured_mesh("quad4", (n, n), lengths=(1.0, 1.0))
    times = np.linspace(0.0, 1.0, 11)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    model = px.NeoHookean(bulk_modulus=0.833, shear_modulus=px.BoundedProperty(0.01, 5.0, key=0))
    physics = px.SolidMechanics(model, px.PlaneStrain())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
    gd = px.GlobalData.from_arrays(times, 0.3 * times, mesh.nodeSets["nset_1"], 1)
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
    rng = np.random.default_rng(0)
    bs = 1024
    xy, t = rng.uniform(0, 1, size=(bs, 2)), rng.choice(times, size=bs)
    inputs = np.column_stack([xy, t])
    outputs = np.column_stack([-0.3 * (xy[:, 0] - 0.5) * t, xy[:, 1] * t]) + 1e-2 * rng.standard_normal((bs, 2))
    loss_fn = px.CombineLossFunctions(px.EnergyResidualAndReactionLoss(energy_weight=1.0, residual_weight=250e9,
                                                                      reaction_weight=250e9),
                                      px.FullFieldDataLoss(weight=10e9), with_props=False)
    evals = n * n * 4 * 11
    measure("C4 (synthetic, scaled) Quad4 512^2, NeoHookean K=0.833 + trainable G in [0.01,5], nt=11, "
            "EnergyResidualAndReaction(250e9,250e9) + FullFieldData(10e9) on 1024 DIC rows, MLP 3->50x3->2",
            problem, params, loss_fn, (inputs, outputs), steps, evals)


def c5(steps, device, ns):
    for model_name in ("swanson", "hencky"):
        for n in ns:
            mesh = px.create_structured_mesh("hex8", (n, n, n), lengths=(1.0, 1.0, 1.0))
            times = np.linspace(0.0, 1.0, 2)
            domain = px.VariationalDomain(mesh, times, q_order=2)
            model = px.Swanson(**SWANSON) if model_name == "swanson" else px.Hencky(bulk_modulus=0.833, shear_modulus=0.3846)
            physics = px.SolidMechanics(model, px.ThreeDimensional())
            bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
            problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
            params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                                   dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
            evals = n ** 3 * 8 * 2
            measure(f"C5 Hex8 {n}^3 ({n ** 3 * 8} qp) {model_name}, nt=2, EnergyLoss, MLP 4->50x4->3",
                    problem, params, px.EnergyLoss(), (), steps, evals)
            del problem, params, domain, mesh
            torch.cuda.empty_cache()


def r3d(steps, device, n=100, nt=2):
    """3D residual / reaction loss (the inverse-problem loss on a Hex8 block): adds the stiffness-action sweep K v."""
    mesh = px.create_structured_mesh("hex8", (n, n, n), lengths=(1.0, 1.0, 1.0))
    times = np.linspace(0.0, 1.0, nt)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    for model_name, model in (("neohookean", px.NeoHookean(bulk_modulus=0.833, shear_modulus=0.3846)),
                              ("swanson", px.Swanson(**SWANSON)),
                              ("hencky", px.Hencky(bulk_modulus=0.833, shear_modulus=0.3846))):
        physics = px.SolidMechanics(model, px.ThreeDimensional())
        bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
        gd = px.GlobalData.from_arrays(times, 0.1 * times, mesh.nodeSets["nset_5"], 1)
        problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
        params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                               dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
        measure(f"R3D Hex8 {n}^3 ({n ** 3 * 8} qp) {model_name}, nt={nt}, EnergyResidualAndReactionLoss, MLP 4->50x4->3",
                problem, params, px.EnergyResidualAndReactionLoss(), (), steps, n ** 3 * 8 * nt)
        problem._engines.clear()


def pstress(steps, device, n=512, nt=11):
    """C2's mesh under PlaneStress (per-point root find of P_33 = 0, solid_mechanics.py:32-71)."""
    mesh = px.create_structured_mesh("quad4", (n, n), lengths=(1.0, 1.0))
    times = np.linspace(0.0, 1.0, nt)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    for model_name, model in (("neohookean", px.NeoHookean(bulk_modulus=10.0, shear_modulus=1.0)),
                              ("hencky", px.Hencky(bulk_modulus=10.0, shear_modulus=1.0))):
        physics = px.SolidMechanics(model, px.PlaneStress())
        bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
        gd = px.GlobalData.from_arrays(times, np.zeros(nt), mesh.nodeSets["nset_1"], 1)
        problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
        params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                               dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
        measure(f"PSTRESS Quad4 {n}^2 {model_name} plane stress, nt={nt}, EnergyResidualAndReactionLoss, MLP 3->50x3->2",
                problem, params, px.EnergyResidualAndReactionLoss(), (), steps, n * n * 4 * nt)
        problem._engines.clear()


def visco(steps, device, n=64, nt=11):
    """examples/forward_problems/mechanics/example_hyper_visco_3d.py at scale: SimpleFeFv(NeoHookean, one Prony branch),
    path-dependent EnergyLoss: forward sweep over the steps with the states kept + reverse sweep + one MLP adjoint."""
    mesh = px.create_structured_mesh("hex8", (n, n, n), lengths=(1.0, 1.0, 1.0))
    times = np.linspace(0.0, 1.0, nt)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=10.0, shear_modulus=0.855),
                          px.PronySeries(moduli=[1.0], relaxation_times=[10.0]), px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
    physics = px.SolidMechanics(model, px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
    evals = n ** 3 * 8 * (nt - 1)     # the scan starts at step 1
    measure(f"visco: Hex8 {n}^3 ({n ** 3 * 8} qp) SimpleFeFv(NeoHookean, 1 Prony branch), nt={nt}, path-dependent EnergyLoss, "
            "MLP 4->50x4->3", problem, params, px.EnergyLoss(is_path_dependent=True), (), steps, evals)
    # examples/inverse_problems/mechanics/path-dependent/script.py at scale: the residual + reaction loss on the same
    # model (forward sweep with the step's internal force, reverse sweep with the second-order terms)
    gd = px.GlobalData.from_arrays(times, 0.1 * times, mesh.nodeSets["nset_5"], 1)
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
    measure(f"visco: Hex8 {n}^3 ({n ** 3 * 8} qp) SimpleFeFv(NeoHookean, 1 Prony branch), nt={nt}, path-dependent "
            "EnergyResidualAndReactionLoss, MLP 4->50x4->3", problem, params,
            px.EnergyResidualAndReactionLoss(is_path_dependent=True), (), steps, evals)


def visco_pstress(steps, device, n=512, nt=11):
    """C2's mesh with SimpleFeFv under PlaneStress (per-point root find on the total stress at fixed old state) and plane
    strain, path-dependent EnergyLoss."""
    mesh = px.create_structured_mesh("quad4", (n, n), lengths=(1.0, 1.0))
    times = np.linspace(0.0, 1.0, nt)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    model = px.SimpleFeFv(px.NeoHookean(bulk_modulus=10.0, shear_modulus=0.855),
                          px.PronySeries(moduli=[1.0], relaxation_times=[10.0]), px.WLF(C1=17.44, C2=51.6, theta_ref=60.0))
    for form_name, form in (("plane strain", px.PlaneStrain()), ("plane stress", px.PlaneStress())):
        physics = px.SolidMechanics(model, form)
        bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
        problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
        params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                               dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.5, 1.0, "y", 2))
        measure(f"visco 2D: Quad4 {n}^2 ({n * n * 4} qp) SimpleFeFv(NeoHookean, 1 Prony branch) {form_name}, nt={nt}, "
                "path-dependent EnergyLoss, MLP 3->50x3->2", problem, params, px.EnergyLoss(is_path_dependent=True), (), steps,
                n * n * 4 * (nt - 1))
        problem._engines.clear()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c2,c4,c5")
    ap.add_argument("--c5-n", default="50,100,159,252")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a B200"
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cases = args.cases.split(",")
    if "c2" in cases:
        c2(args.steps, device)
    if "c4" in cases:
        c4(args.steps, device)
    if "c5" in cases:
        c5(args.steps, device, [int(x) for x in args.c5_n.split(",")])
    if "visco" in cases:
        visco(args.steps, device)
    if "visco2d" in cases:
        visco_pstress(args.steps, device)
    if "r3d" in cases:
        r3d(args.steps, device)
    if "pstress" in cases:
        pstress(args.steps, device)


if __name__ == "__main__":
    main()
