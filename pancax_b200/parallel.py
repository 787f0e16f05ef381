"""Multi-GPU plumbing: one process per GPU (torchrun), quadrature points / elements sharded across
ranks, ONE collective per step: allreduce(sum) of [loss | aux | weight grads | property grads]
(SURVEY.md 8e).

EnergyLoss needs no halo exchange: u at a node is a pure function of (x, t, theta), so every rank
evaluates the field on the closure nodes of its own elements, accumulates the partial internal force
f_r of its own elements only and back-propagates it; the VJP is linear in the cotangent, hence
sum_r grads_r is exact.  Residual / reaction losses need the assembled f on interface nodes before
the norm and are refused on more than one rank for now."""
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
        return self
