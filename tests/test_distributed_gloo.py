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


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from pancax_b200 import parallel
    from pancax_b200.fem import Mesh
    from tests.cases import make_case
    parallel.init_distributed()
    case = make_case("hex8_neohookean", n=4, nt=2, width=12, depth=2)
    full = Mesh(case.mesh["coords"], case.mesh["conns"], "hex8", case.mesh["nodesets"])
    sub, l2g = parallel.partition_mesh(full, rank, world)
    plan = parallel.make_shard_plan(l2g)                     # all_gather_object of the node maps
    ex = parallel.ShardExchange(plan, "cpu")
    nt, ndof = 2, 3
    # every rank contributes f_r[n] = (rank+1) * g(global node id) on its closure nodes
    g = torch.as_tensor(l2g, dtype=torch.float64)
    f = ((rank + 1) * (g[None, :, None] + 1.0) * torch.ones(nt, 1, ndof, dtype=torch.float64)).contiguous()
    ex.sum_shared(f)
    rr = torch.full((nt, 2), float(rank + 1), dtype=torch.float64)
    ex.sum_scalars(rr)
    unk, n_uo, rn, n_ro = plan.order_owned_first(np.arange(sub.coords.shape[0] * ndof), np.arange(5), ndof)
    q.put((rank, l2g, f.numpy(), rr.numpy(), plan.owned.copy(), plan.n_shared, n_uo))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_exchange():
    """Host logic of the sharded residual losses: shared-node plan from gathered node maps,
    packed allreduce of the shared rows, ownership (each global node owned exactly once)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, l2g0, f0, rr0, own0, ns0, nuo0), (_, l2g1, f1, rr1, own1, ns1, nuo1) = res
    shared = np.intersect1d(l2g0, l2g1)
    assert ns0 == ns1 == shared.size > 0
    for l2g, f, rank in ((l2g0, f0, 0), (l2g1, f1, 1)):
        is_sh = np.isin(l2g, shared)
        want = np.where(is_sh, 3.0, rank + 1.0) * (l2g + 1.0)          # (1 + 2) * g on shared nodes
        assert np.array_equal(f[0, :, 0], want) and np.array_equal(f[1, :, 2], want)
    assert np.array_equal(rr0, np.full((2, 2), 3.0)) and np.array_equal(rr1, rr0)
    owners = np.zeros(max(l2g0.max(), l2g1.max()) + 1, dtype=int)
    np.add.at(owners, l2g0[own0], 1)
    np.add.at(owners, l2g1[own1], 1)
    assert np.array_equal(owners[np.union1d(l2g0, l2g1)], np.ones(np.union1d(l2g0, l2g1).size, dtype=int))
    assert nuo0 == own0.sum() * 3 and nuo1 == own1.sum() * 3


def _exchange3_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from pancax_b200 import parallel
    from tests.cases import make_case
    parallel.init_distributed()
    case = make_case("quad4_neohookean", n=7, nt=3, width=12, depth=2)
    conns = case.mesh["conns"][np.random.default_rng(5).permutation(case.mesh["conns"].shape[0])]   # shards interleave in space
    ne = conns.shape[0]
    lo, hi = (ne * rank) // world, (ne * (rank + 1)) // world
    l2g = np.unique(conns[lo:hi])
    plan = parallel.make_shard_plan(l2g)
    ex = parallel.ShardExchange(plan, "cpu")
    nt, ndof = 3, 2
    rng = np.random.default_rng(100 + rank)
    f = torch.as_tensor(rng.standard_normal((nt, l2g.size, ndof)))       # this shard's partial rows
    part = f.clone()
    ex.sum_shared(f)
    q.put((rank, l2g, part.numpy(), f.numpy(), sorted(plan.peers)))
    dist.barrier()
    dist.destroy_process_group()


def test_three_rank_neighbour_exchange_sums_in_ascending_shard_order():
    """Neighbour exchange (isend / irecv of the rows shared with each peer): nodes shared by two AND by three shards end
    with the sum of all sharers' partial rows, accumulated in ascending shard order -- bit-identical on every sharer."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange3_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nglob = max(r[1].max() for r in res) + 1
    total = np.zeros((3, nglob, 2))
    count = np.zeros(nglob, dtype=int)
    for rank, l2g, part, f, peers in res:          # ascending shard order, from zero: the canonical sum
        total[:, l2g] += part
        count[l2g] += 1
    assert (count == 3).any() and (count == 2).any()
    for rank, l2g, part, f, peers in res:
        sh = count[l2g] > 1
        assert np.array_equal(f[:, sh], total[:, l2g[sh]])          # bitwise
        assert np.array_equal(f[:, ~sh], part[:, ~sh])
        assert len(peers) == 2
