"""A/B of the forward-kernel tuning options on config C2 (Quad4 n^2, plane-strain NeoHookean, nt = 11, MLP 3->50x3->2,
EnergyResidualAndReactionLoss): per-kernel-class device times from pcx_profile_*.
    python tools/kernel_ab_c2.py [n] [steps] > gpurun_out/kernel_ab_c2.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
device = torch.device("cuda", 0)
mesh = px.create_structured_mesh("quad4", (n, n))
times = np.linspace(0.0, 1.0, 11)
domain = px.VariationalDomain(mesh, times, q_order=2)
physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
gd = px.GlobalData.from_arrays(times, 0.3 * times, mesh.nodeSets["nset_1"], 1)
problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                       dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
eng = get_engine(problem, params)
w = params.fields.networks.weights
gout = torch.as_tensor(gd.outputs, device=device)
VARIANTS = [("default", {}), ("static groups", dict(tune_fwd_dynamic=0)), ("single exp2 table", dict(tune_fwd_dynamic=1, tune_fwd_tanh=0)),
            ("edge tile on DMMA", dict(tune_fwd_tanh=1, tune_fwd_edge=0)), ("MT=2 MINB=2", dict(tune_fwd_edge=1, tune_fwd_variant=22)),
            ("default again", dict(tune_fwd_variant=12))]
for name, opts in VARIANTS:
    for k, v in opts.items():
        eng.set_option(k, v)
    for _ in range(3):
        out = eng.loss_and_grad("energy_residual_reaction", w, None, global_outputs=gout)
    torch.cuda.synchronize()
    eng.profile_enable(True)
    eng.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = eng.loss_and_grad("energy_residual_reaction", w, None, global_outputs=gout)
    e1.record()
    torch.cuda.synchronize()
    prof = eng.profile_collect()
    eng.profile_enable(False)
    print(json.dumps({"variant": name, "options": opts, "n": n, "ms_per_step": e0.elapsed_time(e1) / steps, "loss": float(out[0][0]),
                      "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}}), flush=True)
