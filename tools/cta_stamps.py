"""Per-CTA (start, end, SM id) stamps of ONE launch of the forward / adjoint MLP kernel on the bench workload: how evenly
do the persistent CTAs finish?  (tuning; PCX_OPT_TUNE_STAMPS)    python tools/cta_stamps.py [n] > gpurun_out/cta_stamps.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
device = torch.device("cuda", 0)
mesh = px.create_structured_mesh("hex8", (n, n, n))
times = np.linspace(0.0, 1.0, 2)
domain = px.VariationalDomain(mesh, times, q_order=2)
physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.ThreeDimensional())
problem = px.ForwardProblem(domain, physics, dirichlet_bcs=[])
params = px.Parameters(problem, key=0, n_layers=4, n_neurons=50, device=device,
                       dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 3))
eng = get_engine(problem, params)
eng.set_option("cull_null_steps", 0)
w = params.fields.networks.weights
out = {}
for rep in range(2):
    for kind in ("forward", "adjoint"):
        stamps = torch.zeros(3 * 2 * 148 * 2, dtype=torch.int64, device=device)
        us = eng.field_forward(w, 1)
        f = torch.ones_like(us)
        torch.cuda.synchronize()
        eng.set_option("tune_stamps", stamps.data_ptr())
        if kind == "forward":
            eng.field_forward(w, 1, out=us)
        else:
            eng.field_backward(w, 1, f)       # forward (296 CTAs) then adjoint (148 CTAs): the adjoint's stamps overwrite
        torch.cuda.synchronize()
        eng.set_option("tune_stamps", 0)
        st = stamps.cpu().numpy().reshape(-1, 3)
        st = st[:148] if kind == "adjoint" else st[st[:, 1] > 0]
        dur = (st[:, 1] - st[:, 0]) / 1e3
        t0 = st[:, 0].min()
        by_sm = {}
        for (a, b, sm), d in zip(st, dur):
            by_sm.setdefault(int(sm), []).append(round(float(d), 1))
        out[f"{kind}_{rep}"] = {"ctas": int(st.shape[0]), "kernel_us": float(st[:, 1].max() - t0) / 1e3,
                                "cta_us_min": float(dur.min()), "cta_us_mean": float(dur.mean()), "cta_us_max": float(dur.max()),
                                "cta_us_sorted_by_sm": [by_sm[k] for k in sorted(by_sm)]}
print(json.dumps(out))
