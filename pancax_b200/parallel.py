"""Multi-GPU plumbing: one process per GPU (torchrun), quadrature points / elements sharded across
ranks, ONE collective per step: allreduce(sum) of [loss | aux | weight grads | property grads]
(SURVEY.md 8e).

EnergyLoss needs no halo exchange: u at a node is a pure function of (x, t, theta), so every rank
evaluates the field on the closure nodes of its own elements, accumulates the partial internal force
f_r of its own elements only and back-propagates it; the VJP is linear in the cotangent, hence
sum_r grads_r is exact.  Residual / reaction losses need the assembled f on the nodes shared between
shards before the norm, and the global norm / reactions before v: two more small allreduces
(shared-node rows of f, then nt*2 scalars) between the phases of the C ABI's phased loss
(pcx_loss_begin / _mid / _finish); see ShardPlan / ShardExchange."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from .fem import Mesh


def init_distributed():
    """(rank, world, local_rank); initialises torch.distributed from torchrun's environment."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def partition_mesh(mesh: Mesh, rank: int, world: int):
    """Contiguous element range of ``mesh`` for ``rank`` plus its closure nodes, renumbered locally.
    Returns (sub_mesh, local_to_global_node)."""
    ne = mesh.num_elements
    lo = (ne * rank) // world
    hi = (ne * (rank + 1)) // world
    conns = mesh.conns[lo:hi]
    nodes, inv = np.unique(conns, return_inverse=True)
    local_conns = inv.reshape(conns.shape).astype(np.int32)
    g2l = -np.ones(mesh.num_nodes, dtype=np.int64)
    g2l[nodes] = np.arange(nodes.shape[0])
    nodesets = {}
    for k, v in mesh.nodeSets.items():
        loc = g2l[v]
        nodesets[k] = loc[loc >= 0].astype(np.int32)
    return Mesh(mesh.coords[nodes], local_conns, mesh.elem_type, nodesets), nodes


def global_bounds(coords: np.ndarray):
    """Field normalisation constants over ALL ranks (fields.py:70-71 uses the whole mesh)."""
    lo, hi = coords.min(axis=0), coords.max(axis=0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
        tlo, thi = torch.as_tensor(lo, device=dev), torch.as_tensor(hi, device=dev)
        dist.all_reduce(tlo, op=dist.ReduceOp.MIN)
        dist.all_reduce(thi, op=dist.ReduceOp.MAX)
        lo, hi = tlo.cpu().numpy(), thi.cpu().numpy()
    return lo, hi


class PackedResult:
    """One contiguous device buffer [loss(1) | aux(na) | g_weights(nw) | g_props(np)] so a step needs a
    single allreduce (64 KB: latency-bound on NVLink, issued on the compute stream right after the
    last reduction kernel)."""

    def __init__(self, na, nw, npr, device):
        self.na, self.nw, self.np = na, nw, max(npr, 1)
        self.buf = torch.zeros(1 + na + nw + self.np, dtype=torch.float64, device=device)
        o = 0
        self.loss = self.buf[o:o + 1]; o += 1
        self.aux = self.buf[o:o + na]; o += na
        self.gw = self.buf[o:o + nw]; o += nw
        self.gp = self.buf[o:o + self.np]

    def as_out(self):
        return (self.loss, self.aux, self.gw, self.gp)

    def allreduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            # aux[PCX_AUX_FLAGS] is a bit mask (only bit0 is defined), not a quantity: the sum over ranks counts the
            # ranks that raised it -- fold it back to the OR
            if self.na > 3:
                self.aux[3:4].clamp_(max=1.0)
        return self


# ------------------------------------------------------------------------------------------------
# residual / reaction losses on shards: who shares which node, who owns it
# ------------------------------------------------------------------------------------------------
class ShardPlan:
    """Static exchange plan of ONE shard, derived from every shard's local->global node map.

    shared_local : local indices of this shard's nodes that also live on another shard
    shared_slot  : their slot in the (sorted) global list of shared nodes
    n_shared     : length of that global list
    owned        : bool mask over local nodes; a shared node is owned by the lowest shard holding it
    peers        : {other shard -> local indices of the nodes shared with it, in ascending global-id order (the same
                   order on both sides)}: the neighbour exchange sends / receives exactly these rows
    """

    def __init__(self, shared_local, shared_slot, n_shared, owned, n_shards, index, peers=None):
        self.shared_local, self.shared_slot = shared_local, shared_slot
        self.n_shared, self.owned, self.n_shards, self.index = n_shared, owned, n_shards, index
        self.peers = {} if peers is None else peers

    def order_owned_first(self, unknown_idx, reaction_nodes, ndof):
        """Reorder dof / node lists so the entries owned by this shard come first (what
        pcx_set_ownership expects).  Returns (unknown_idx, n_unknown_owned, reaction_nodes, n_reaction_owned)."""
        out = []
        for arr, div in ((unknown_idx, ndof), (reaction_nodes, 1)):
            if arr is None:
                out += [None, 0]
                continue
            arr = np.asarray(arr)
            own = self.owned[arr // div]
            out += [np.concatenate([arr[own], arr[~own]]), int(own.sum())]
        return tuple(out)


def make_shard_plans(l2g_all):
    """Plans for all shards from their local->global node maps (list of int arrays)."""
    n = len(l2g_all)
    ids, counts = np.unique(np.concatenate([np.asarray(a, dtype=np.int64) for a in l2g_all]), return_counts=True)
    shared_ids = ids[counts > 1]
    locs, slots = [], []
    for l2g in l2g_all:
        l2g = np.asarray(l2g, dtype=np.int64)
        if shared_ids.size:
            pos = np.clip(np.searchsorted(shared_ids, l2g), 0, shared_ids.size - 1)
            hit = shared_ids[pos] == l2g
        else:
            pos, hit = np.zeros(l2g.shape, dtype=np.int64), np.zeros(l2g.shape, dtype=bool)
        loc = np.nonzero(hit)[0]
        locs.append(loc)
        slots.append(pos[loc])
    owner = np.full(shared_ids.size, -1, dtype=np.int64)
    for r in reversed(range(n)):
        owner[slots[r]] = r
    # pairwise neighbour lists: slots present on both shards, ascending (= ascending global id), as local indices
    peers = [dict() for _ in range(n)]
    for r in range(n):
        for q in range(r + 1, n):
            common, ir, iq = np.intersect1d(slots[r], slots[q], assume_unique=True, return_indices=True)
            if common.size:
                peers[r][q] = locs[r][ir].astype(np.int64)
                peers[q][r] = locs[q][iq].astype(np.int64)
    plans = []
    for r, l2g in enumerate(l2g_all):
        owned = np.ones(len(l2g), dtype=bool)
        owned[locs[r]] = owner[slots[r]] == r
        plans.append(ShardPlan(locs[r], slots[r], int(shared_ids.size), owned, n, r, peers[r]))
    return plans


def make_shard_plan(local_to_global):
    """This rank's plan; the node maps of all ranks are gathered once at set-up."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return make_shard_plans([local_to_global])[0]
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, np.asarray(local_to_global, dtype=np.int64))
    return make_shard_plans(gathered)[dist.get_rank()]


class ShardExchange:
    """The two exchanges of the phased loss over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

    sum_shared: NEIGHBOUR exchange of the shared-node rows of f -- each shard packs the rows it shares with a peer
    (library kernel pcx_gather_rows), sends them to that peer only and receives the peer's (ncclSend / ncclRecv through
    batch_isend_irecv: one fused NCCL group), then rebuilds the rows as the sum of all sharers' contributions in ASCENDING
    SHARD ORDER (pcx_scatter_rows), so every sharer holds bit-identical values.  A z-slab partition talks to two
    neighbours about one plane each; round 1 allreduced the global list of all shared nodes on every rank.
    sum_scalars: ONE allreduce of nt*2 scalars.  Everything runs on the current stream without host synchronisation."""

    def __init__(self, plan: ShardPlan, device):
        self.plan, self.n_shards = plan, plan.n_shards
        self.device = torch.device(device)
        self.loc = torch.as_tensor(plan.shared_local, dtype=torch.long, device=device)
        self.peer_idx = {q: torch.as_tensor(ix, dtype=torch.long, device=device) for q, ix in sorted(plan.peers.items())}
        self._bufs = None

    # ---- pack / unpack: library kernels on the GPU, torch indexing on the host (gloo tests)
    def _gather(self, f, idx, out):
        if f.is_cuda:
            from . import _lib
            import ctypes as C
            nt, nn, ndof = f.shape
            _lib.check(_lib.load().pcx_gather_rows(C.c_void_p(f.data_ptr()), nt, nn, ndof, C.c_void_p(idx.data_ptr()), idx.numel(),
                                                   C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            out.copy_(f[:, idx])
        return out

    def _scatter(self, f, idx, src, accumulate):
        if f.is_cuda:
            from . import _lib
            import ctypes as C
            nt, nn, ndof = f.shape
            _lib.check(_lib.load().pcx_scatter_rows(C.c_void_p(f.data_ptr()), nt, nn, ndof, C.c_void_p(idx.data_ptr()), idx.numel(),
                                                    None if src is None else C.c_void_p(src.data_ptr()), 1 if accumulate else 0,
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        elif src is None:
            f[:, idx] = 0.0
        elif accumulate:
            f[:, idx] += src
        else:
            f[:, idx] = src

    def sum_shared(self, f):
        if self.n_shards == 1 or self.plan.n_shared == 0:
            return f
        nt, _, ndof = f.shape
        if self._bufs is None or self._bufs[0].shape[0] != nt:
            mine = torch.empty(nt, self.loc.numel(), ndof, dtype=f.dtype, device=f.device)
            send = {q: torch.empty(nt, ix.numel(), ndof, dtype=f.dtype, device=f.device) for q, ix in self.peer_idx.items()}
            recv = {q: torch.empty_like(b) for q, b in send.items()}
            self._bufs = (mine, send, recv)
        mine, send, recv = self._bufs
        self._gather(f, self.loc, mine)                       # every send uses the ORIGINAL partial values
        ops = []
        for q, ix in self.peer_idx.items():
            self._gather(f, ix, send[q])
            ops.append(dist.P2POp(dist.isend, send[q], q))
            ops.append(dist.P2POp(dist.irecv, recv[q], q))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self._scatter(f, self.loc, None, False)               # 0, then contributions in ascending shard order
        me, done_me = self.plan.index, False
        for q, ix in self.peer_idx.items():
            if q > me and not done_me:
                self._scatter(f, self.loc, mine, True)
                done_me = True
            self._scatter(f, ix, recv[q], True)
        if not done_me:
            self._scatter(f, self.loc, mine, True)
        return f

    def sum_scalars(self, rr):
        if self.n_shards > 1:
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        return rr


class ShardInfo:
    def __init__(self, plan, exchange, x_mins=None, x_maxs=None):
        self.plan, self.exchange = plan, exchange
        self.x_mins, self.x_maxs = x_mins, x_maxs      # Field normalisation bounds of the WHOLE mesh


def shard_problem(problem, local_to_global, device=None, bounds=None):
    """Declare ``problem`` (built on this rank's element shard, nodes numbered locally) to be one shard
    of a larger mesh: ``local_to_global`` maps its nodes to mesh-wide node ids.  Residual / reaction
    losses evaluated on it then go through the phased ABI with the two exchanges; EnergyLoss needs
    nothing.  The Field normalisation bounds of the WHOLE mesh (fields.py:70-71) are computed here
    (``bounds = (x_mins, x_maxs)`` overrides them: shards driven from one process) and used by every Field
    built on, or evaluated with, the sharded problem.  Collective: every rank must call it."""
    plan = make_shard_plan(local_to_global)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
    if bounds is None:
        bounds = global_bounds(np.asarray(problem.coords))     # MIN / MAX allreduce over the ranks
    problem.shard = ShardInfo(plan, ShardExchange(plan, device), np.asarray(bounds[0], dtype=np.float64),
                              np.asarray(bounds[1], dtype=np.float64))
    problem._engines = {}
    return problem


def run_shards_in_process(engines, plans, kind, weights, prop_raw=None, **kw):
    """All shards driven in lock-step from ONE process (tests; one GPU): the same three phases and two
    exchanges as the distributed path, the sums done with plain tensor adds.  Returns the summed
    (loss, aux, g_weights, g_props)."""
    n = len(engines)
    fs = [e.loss_begin(kind, weights, prop_raw, **kw) for e in engines]
    dev = fs[0].device
    tot = torch.zeros(fs[0].shape[0], plans[0].n_shared, fs[0].shape[2], dtype=fs[0].dtype, device=dev)
    idx = [(torch.as_tensor(p.shared_local, device=dev), torch.as_tensor(p.shared_slot, device=dev)) for p in plans]
    for f, (loc, slot) in zip(fs, idx):
        tot[:, slot] += f[:, loc]
    for f, (loc, slot) in zip(fs, idx):
        f[:, loc] = tot[:, slot]
    rrs = [e.loss_mid() for e in engines]
    rr = torch.stack(rrs).sum(0)
    for r in rrs:
        r.copy_(rr)
    outs = [e.loss_finish(n) for e in engines]
    return tuple(sum(o[i] for o in outs) for i in range(4))
