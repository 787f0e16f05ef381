"""CPU, world_size 2, gloo: the multi-rank path of the host logic -- element partition with closure
nodes, global normalisation bounds and the single packed allreduce -- with the oracle standing in for
the per-rank CUDA evaluation.  sum over ranks == single-rank result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from oracle import losses as ol
    from pancax_b200 import parallel
    from pancax_b200.fem import Mesh
    from tests.cases import make_case, oracle_problem
    r, w, _ = parallel.init_distributed()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    case = make_case("hex8_neohookean", n=4, nt=2, width=12, depth=2)
    full = Mesh(case.mesh["coords"], case.mesh["conns"], "hex8", case.mesh["nodesets"])
    sub, l2g = parallel.partition_mesh(full, rank, world)
    lo, hi = parallel.global_bounds(sub.coords)
    assert np.allclose(lo, full.coords.min(0)) and np.allclose(hi, full.coords.max(0))
    prob = oracle_problem(case)
    prob.coords, prob.conns = sub.coords, sub.conns
    # per-rank evaluation with GLOBAL normalisation constants
    wts = torch.as_tensor(case.weights).clone().requires_grad_(True)
    fld, phys, pe = ol._build(prob, wts, torch.zeros(0))
    fld.x_mins, fld.x_maxs = torch.as_tensor(lo), torch.as_tensor(hi)
    from oracle.physics import potential_energy
    tot = 0.0
    for t in case.times:
        us = fld(torch.as_tensor(sub.coords), float(t))
        tot = tot + potential_energy(phys, pe, torch.as_tensor(sub.coords),
                                     torch.as_tensor(sub.conns, dtype=torch.long), us)
    (g,) = torch.autograd.grad(tot, wts)
    res = parallel.PackedResult(4 + 2, g.numel(), 0, "cpu")
    res.loss[0] = tot.detach()
    res.aux[0] = tot.detach()
    res.gw.copy_(g)
    res.allreduce()
    if rank == 0:
        q.put((float(res.loss[0]), res.gw.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_sums_to_single_rank():
    from tests.cases import make_case, run_oracle_loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss2, g2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    case = make_case("hex8_neohookean", n=4, nt=2, width=12, depth=2)
    L1, _, g1, _ = run_oracle_loss(case, "energy", energy_weight=1.0)
    assert abs(loss2 - L1) <= 1e-12 * abs(L1)
    assert np.linalg.norm(g2 - g1) <= 1e-12 * np.linalg.norm(g1)
