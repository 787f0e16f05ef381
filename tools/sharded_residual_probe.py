"""Where does the time of the sharded residual loss go?  C2 (Quad4 n^2, nt = 11) cut into 2 element shards driven in
lock-step on ONE GPU: CUDA-event time of each phase of the phased ABI per shard, next to the whole-mesh call.
    python tools/sharded_residual_probe.py [n]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200 import parallel                          # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nt = 11
device = torch.device("cuda", 0)
times = np.linspace(0.0, 1.0, nt)
full = px.create_structured_mesh("quad4", (n, n))
gout = 0.05 * times


def build(mesh):
    domain = px.VariationalDomain(mesh, times, q_order=2)
    physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
    gd = px.GlobalData.from_arrays(times, gout, mesh.nodeSets["nset_1"], 1)
    return px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)


def params_for(problem):
    return px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device, dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fp = build(full)
fparams = params_for(fp)
w = fparams.fields.networks.weights
eng = get_engine(fp, fparams)
eng.set_option("cull_null_steps", 0)
kw = dict(global_outputs=gout, energy_weight=1.0, residual_weight=1.0, reaction_weight=1.0)
out = {"n": n, "whole_mesh_ms": timed(lambda: eng.loss_and_grad("energy_residual_reaction", w, **kw))}
eng.profile_enable(True); eng.profile_collect()
eng.loss_and_grad("energy_residual_reaction", w, **kw)
out["whole_mesh_kernels_ms"] = {k: round(v[0], 4) for k, v in eng.profile_collect().items() if v[1]}
eng.profile_enable(False)
subs = [parallel.partition_mesh(full, r, 2) for r in range(2)]
plans = parallel.make_shard_plans([l2g for _, l2g in subs])
lo, hi = full.coords.min(axis=0), full.coords.max(axis=0)
engs = []
for (sub, _), plan in zip(subs, plans):
    sp = build(sub)
    sp.shard = parallel.ShardInfo(plan, None, lo, hi)
    sp._engines = {}
    e = get_engine(sp, params_for(sp))
    e.set_option("cull_null_steps", 0)
    engs.append(e)
e0 = engs[0]
out["shard0_begin_ms"] = timed(lambda: e0.loss_begin("energy_residual_reaction", w, None, **kw))
e0.loss_begin("energy_residual_reaction", w, None, **kw)
out["shard0_mid_ms"] = timed(lambda: (setattr(e0, "_x", None), e0.lib.pcx_loss_mid)[1] and None) if False else None


def mid():
    e0.loss_begin("energy_residual_reaction", w, None, **kw)
    e0.loss_mid()


def full_phases():
    e0.loss_begin("energy_residual_reaction", w, None, **kw)
    e0.loss_mid()
    e0.loss_finish(2)


out["shard0_begin_mid_ms"] = timed(mid)
out["shard0_begin_mid_finish_ms"] = timed(full_phases)
e0.profile_enable(True); e0.profile_collect()
full_phases()
out["shard0_kernels_ms"] = {k: round(v[0], 4) for k, v in e0.profile_collect().items() if v[1]}
e0.profile_enable(False)
out["both_shards_in_process_ms"] = timed(lambda: parallel.run_shards_in_process(engs, plans, "energy_residual_reaction", w, None, **kw))
print(json.dumps(out))
