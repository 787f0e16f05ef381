"""torchrun worker: EnergyResidualAndReactionLoss (and EnergyAndResidualLoss) through the host mirror on N
element shards (one per GPU, NCCL) against the same loss on the whole mesh on rank 0's GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/mgpu_residual_check.py [n]
Prints one JSON line on rank 0 and exits non-zero on mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(px, mesh, times, device, gd_nodes, lo=None, hi=None):
    domain = px.VariationalDomain(mesh, times, q_order=2)
    model = px.NeoHookean(bulk_modulus=10.0, shear_modulus=px.BoundedProperty(0.01, 5.0, 1.0))
    physics = px.SolidMechanics(model, px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    gd = px.GlobalData.from_arrays(times, np.linspace(0.0, 0.3, len(times)), gd_nodes, 1)
    problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(0.4, 1.0, "y", 3))
    if lo is not None:
        params.fields.x_mins, params.fields.x_maxs = lo, hi
    return problem, params


def main():
    import pancax_b200 as px
    from pancax_b200 import parallel
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rank, world, local_rank = parallel.init_distributed()
    device = torch.device("cuda", local_rank)
    times = np.linspace(0.2, 1.0, 3)
    full = px.create_structured_mesh("hex8", (n, n, n))
    sub, l2g = parallel.partition_mesh(full, rank, world)
    lo, hi = parallel.global_bounds(sub.coords)
    problem, params = build(px, sub, times, device, sub.nodeSets["nset_5"], lo, hi)
    parallel.shard_problem(problem, l2g, device)
    out = {}
    ok = True
    for name, lf in (("energy_residual_reaction", px.EnergyResidualAndReactionLoss(1.3, False, 0.7, 2.1)),
                     ("energy_residual", px.EnergyAndResidualLoss(1.3, False, 0.7))):
        (loss, aux), grads = lf.value_and_grad(params, problem)
        res = parallel.PackedResult(2, grads.fields.numel(), grads.props.numel(), device)
        res.loss.copy_(loss.reshape(1)); res.aux[0] = aux["energy"]; res.aux[1] = aux["residual"]
        res.gw.copy_(grads.fields); res.gp[:grads.props.numel()].copy_(grads.props)
        res.allreduce()
        torch.cuda.synchronize()
        if rank == 0:
            fp, fparams = build(px, full, times, device, full.nodeSets["nset_5"])
            (L1, a1), g1 = lf.value_and_grad(fparams, fp)
            rl = abs(float(res.loss[0]) - float(L1)) / abs(float(L1))
            rg = float((res.gw - g1.fields).norm() / g1.fields.norm())
            rp = float((res.gp[:1] - g1.props).norm() / g1.props.norm())
            rr = abs(float(res.aux[1]) - float(a1["residual"])) / abs(float(a1["residual"]))
            out[name] = dict(loss=float(L1), rel_loss=rl, rel_grad=rg, rel_gprop=rp, rel_residual=rr)
            ok = ok and max(rl, rg, rp, rr) < 1e-10
    if rank == 0:
        print(json.dumps({"world": world, "n": n, "shared_nodes": problem.shard.plan.n_shared, "ok": ok, **out}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
