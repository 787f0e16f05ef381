"""torchrun worker: cumulative time of the phases of the sharded residual loss (C2 cut into one shard per GPU).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_residual_probe_mgpu.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pancax_b200 as px                                  # noqa: E402
from pancax_b200 import parallel                          # noqa: E402
from pancax_b200.loss_functions import get_engine          # noqa: E402

rank, world, local_rank = parallel.init_distributed()
device = torch.device("cuda", local_rank)
n, nt = 512, 11
times = np.linspace(0.0, 1.0, nt)
full = px.create_structured_mesh("quad4", (n, n))
gout = 0.05 * times
sub, l2g = parallel.partition_mesh(full, rank, world)
domain = px.VariationalDomain(sub, times, q_order=2)
physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
gd = px.GlobalData.from_arrays(times, gout, sub.nodeSets["nset_1"], 1)
sp = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
parallel.shard_problem(sp, l2g, device)
params = px.Parameters(sp, key=0, n_layers=3, n_neurons=50, device=device, dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
eng = get_engine(sp, params)
eng.set_option("cull_null_steps", 0)
w = params.fields.networks.weights
ex = sp.shard.exchange
kw = dict(global_outputs=gout, energy_weight=1.0, residual_weight=1.0, reaction_weight=1.0)
loss_fn = px.EnergyResidualAndReactionLoss()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return {"device_ms": round(e0.elapsed_time(e1) / reps, 3), "host_issue_ms": round(t_issue, 3)}


def p1():
    return eng.loss_begin("energy_residual_reaction", w, None, **kw)


def p2():
    ex.sum_shared(p1())


def p3():
    p2(); ex.sum_scalars(eng.loss_mid())


def p4():
    p3(); eng.loss_finish(world)


def p5():
    return loss_fn.value_and_grad(params, sp)


out = {"world": world, "rank": rank, "begin": timed(p1), "begin+exchange": timed(p2), "begin+exchange+mid+scalars": timed(p3),
       "all_phases": timed(p4), "value_and_grad": timed(p5),
       "allreduce_64KB": timed(lambda: dist.all_reduce(torch.zeros(8192, dtype=torch.float64, device=device)))}
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
