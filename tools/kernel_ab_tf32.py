"""(TF32 modes) A/B device-time measurements of kernel variants on the bench workload (C3: Hex8 n^3, nt = 2, every row evaluated).
One engine, the variants switched with pcx_set_option(PCX_OPT_TUNE_*); per-kernel-class times from CUDA events on the
launching stream (pcx_profile_*).  Also checks that every variant returns the same loss / gradient as the baseline.

    python tools/kernel_ab.py [n] [steps] > gpurun_out/kernel_ab.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
device = torch.device("cuda", 0)
mesh = px.create_structured_mesh("hex8", (n, n, n))
times = np.linspace(0.0, 1.0, 2)
domain = px.VariationalDomain(mesh, times, q_order=2)
physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.ThreeDimensional())
bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                       dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 3))
eng = get_engine(problem, params)
eng.set_option("cull_null_steps", 0)
w = params.fields.networks.weights
VARIANTS = [("fp64", dict(tf32_adjoint=0)), ("tf32_mma_sync", dict(tf32_adjoint=1)), ("tf32_tcgen05", dict(tf32_adjoint=2)),
            ("tf32_mma_sync_again", dict(tf32_adjoint=1)), ("tf32_tcgen05_again", dict(tf32_adjoint=2))]
ref = None


for name, opts in VARIANTS:
    for k, v in opts.items():
        eng.set_option(k, v)
    for _ in range(3):
        out = eng.loss_and_grad("energy", w)
    torch.cuda.synchronize()
    eng.profile_enable(True)
    eng.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = eng.loss_and_grad("energy", w)
    e1.record()
    torch.cuda.synchronize()
    prof = eng.profile_collect()
    eng.profile_enable(False)
    L, g = float(out[0][0]), out[2].clone()
    if ref is None:
        ref = (L, g)
    rec = {"variant": name, "options": opts, "n": n, "ms_per_step": e0.elapsed_time(e1) / steps,
           "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]},
           "loss_rel_vs_baseline": abs(L - ref[0]) / abs(ref[0]),
           "grad_rel_vs_baseline": float((g - ref[1]).norm() / ref[1].norm())}
    print(json.dumps(rec), flush=True)
