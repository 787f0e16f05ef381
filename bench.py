#!/usr/bin/env python
"""bench.py -- quadrature-point loss+gradient evaluations per second (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one rank per GPU)
  python bench.py --impl reference --gpus N ...             the CPU restatement of pancax's path

A "step" is one evaluation of loss and d(loss)/d(theta) over every quadrature point and time step of
the workload (what pancax's eqx.filter_value_and_grad(loss.filtered_loss) computes inside
opt.step, pancax/optimizers.py:77-88).  One eval = one quadrature point at one time step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's own banner / debug lines ("NCCL version ...", NCCL_DEBUG=INFO output)
# go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402
import torch  # noqa: E402

MLP_IN, MLP_OUT, WIDTH, DEPTH = 4, 3, 50, 4
K_BULK, G_SHEAR = 1000.0, 1.0


def algorithmic_flops_per_row_bwd():
    # pancax_b200/roofline.py (SURVEY 8d formulas): 4 P_w - 2 in0 width
    from pancax_b200 import roofline
    return roofline.mlp_backward_flops_per_row(MLP_IN, MLP_OUT, WIDTH, DEPTH)


def algorithmic_flops_per_row_fwd():
    from pancax_b200 import roofline
    return roofline.mlp_forward_flops_per_row(MLP_IN, MLP_OUT, WIDTH, DEPTH)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


SWANSON = dict(bulk_modulus=10.0, A1=0.93074321, P1=-0.07673672, B1=0.0, Q1=0.5, C1=0.10448312, R1=1.71691036,
               cutoff_strain=0.01)   # test/constitutive_models/test_swanson.py:8-14 (config C5)


def strong_grid(world):
    """Near-cubic process grid (px, py, pz) for cutting ONE cube into `world` blocks: fewer duplicated closure nodes
    than z slabs (128^3 over 8 ranks: 65^3 / 64^3 = +4.8 % nodes per rank instead of 129*129*17 / (128*128*16) = +7.9 %)."""
    g = [1, 1, 1]
    k, w = 2, world
    while w > 1:          # world is a power of two on one box (1, 2, 4, 8); peel factors largest axis last
        g[k] *= 2
        w //= 2
        k = (k - 1) % 3
    return tuple(g)


def build_problem(n, nt, rank, world, device, strong=False, model="neohookean"):
    """--scaling strong: ONE n^3 unit cube cut into `world` near-cubic blocks (BASELINE config 3 read literally:
    "16M quadrature points ... sharded across 8 GPUs").  Default (weak):
    unit-cube slab [0,1]^2 x [rank, rank+1] of n^3 Hex8 elements per rank (weak scaling: the bar
    grows along z with the GPU count), 3D NeoHookean, uniaxial-tension Dirichlet wrapper."""
    import pancax_b200 as px
    from pancax_b200 import parallel
    if strong:
        gx, gy, gz = strong_grid(world)
        assert n % gx == 0 and n % gy == 0 and n % gz == 0, "--scaling strong needs n divisible by the process grid"
        ix, iy, iz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
        mesh = px.create_structured_mesh("hex8", (n // gx, n // gy, n // gz), lengths=(1.0 / gx, 1.0 / gy, 1.0 / gz),
                                         origin=(ix / gx, iy / gy, iz / gz))
    else:
        mesh = px.create_structured_mesh("hex8", (n, n, n), lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, float(rank)))
    times = np.linspace(0.0, 1.0, nt)
    domain = px.VariationalDomain(mesh, times, q_order=2)
    cm = {"neohookean": lambda: px.NeoHookean(bulk_modulus=K_BULK, shear_modulus=G_SHEAR),
          "gent": lambda: px.Gent(bulk_modulus=0.833, shear_modulus=0.3846, Jm_parameter=3.0),
          "swanson": lambda: px.Swanson(**SWANSON),
          "hencky": lambda: px.Hencky(bulk_modulus=0.833, shear_modulus=0.3846)}[model]()
    physics = px.SolidMechanics(cm, px.ThreeDimensional())
    bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    problem = px.ForwardProblem(domain, physics, dirichlet_bcs=bcs)
    params = px.Parameters(problem, key=0, n_layers=DEPTH, n_neurons=WIDTH, device=device,
                           dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0 if model == "neohookean" else 0.5, 1.0, "y", 3))
    lo, hi = parallel.global_bounds(mesh.coords)
    params.fields.x_mins, params.fields.x_maxs = lo, hi
    return problem, params


def _max_over_ranks(ms, device, world):
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu()[0])


def _timed(step, steps, warmup, device, world):
    """warmup untimed steps, then `steps` steps between barrier + synchronize on both sides, CUDA events on the
    launching stream, max over ranks.  Returns total ms."""
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    return _max_over_ranks(e0.elapsed_time(e1), device, world)


def _try_graph(step, device):
    """Capture one step (the library's launches + the NCCL allreduce) in a CUDA graph; returns a replay callable or
    None when the capture is refused (then the eager step is timed).  The C ABI never allocates / synchronises inside a
    compute call, so the whole step is capturable; it matters where the step is short (strong scaling at 8 GPUs:
    ~1.5 ms, 9 launches + one 64 KB allreduce)."""
    if os.environ.get("PCX_BENCH_NO_GRAPH"):
        return None
    try:
        g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=device)
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            step()                                  # same stream as the capture: NCCL's per-stream set-up happens here
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
                step()
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        return g.replay
    except Exception as e:      # noqa: BLE001
        print(f"[bench] CUDA-graph capture of the step refused ({type(e).__name__}: {e}); timing the eager step", file=sys.stderr)
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        return None


def strong_scaling_leg(args, rank, world, device, weak_value):
    """BASELINE config 3 read literally: ONE n^3 block (16.8 M quadrature points at n = 128) cut over the N GPUs.
    Reported at every N next to the (weak) headline so the driver's 1/2/4/8 runs carry the strong curve too."""
    import pancax_b200 as px
    from pancax_b200 import parallel
    from pancax_b200.loss_functions import get_engine
    n, nt = args.n, args.nt
    if world == 1:
        return None      # same workload as the headline line: filled in by the caller
    problem, params = build_problem(n, nt, rank, world, device, True, args.model)
    eng = get_engine(problem, params)
    eng.set_option("cull_null_steps", 0)
    w = params.fields.networks.weights
    res = parallel.PackedResult(4 + nt, eng.n_weights, eng.n_train, device)

    def step():
        eng.loss_and_grad("energy", w, out=res.as_out())
        res.allreduce()
    ms_eager = _timed(step, args.steps, 3, device, world)
    replay = _try_graph(step, device)
    ok = torch.tensor([1.0 if replay is not None else 0.0], device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)        # replay only if EVERY rank captured
    ms_graph = _timed(replay, args.steps, 3, device, world) if float(ok.cpu()[0]) > 0 else None
    ms = min(ms_eager, ms_graph) if ms_graph else ms_eager
    total = n ** 3 * 8 * nt
    value = total * args.steps / (ms * 1e-3)
    out = {"value": value, "unit": "evals/s", "ms_per_step": ms / args.steps, "ms_per_step_eager": ms_eager / args.steps,
           "ms_per_step_cuda_graph": (ms_graph / args.steps) if ms_graph else None, "n_gpus": world,
           "workload": f"ONE Hex8 {n}^3 block ({n ** 3 * 8} qp x nt={nt}) cut into {'x'.join(map(str, strong_grid(world)))} "
                       f"blocks, {eng.ne * eng.nq} qp and {eng.nn} closure nodes per GPU",
           "per_gpu_rate_vs_weak_line": value / weak_value if weak_value else None,
           "note": "strong scaling of the literal C3; per_gpu_rate_vs_weak_line = this value / the headline (weak) value of "
                   "the same run = per-GPU throughput on a 1/N-size shard relative to a full 128^3 shard (includes the "
                   "duplicated closure nodes and the exposed allreduce)"}
    eng.close()
    problem._engines.clear()
    del problem, params, res
    torch.cuda.empty_cache()
    return out


def residual_sharded_leg(args, rank, world, device):
    """Config C2 on N element shards: 2D plane-strain NeoHookean Quad4 n^2 (512^2 = 1 048 576 qp), nt = 11,
    EnergyResidualAndReactionLoss through the phased ABI (pcx_loss_begin / _mid / _finish) with the two exchanges
    between the phases, against the SAME loss on the whole mesh (rank 0).  At N = 1 the whole-mesh step is the timed
    one and the sharded path is checked with two shards driven in lock-step on the one GPU."""
    import torch.distributed as dist
    import pancax_b200 as px
    from pancax_b200 import parallel
    from pancax_b200.loss_functions import get_engine
    n, nt = args.c2_n, 11
    times = np.linspace(0.0, 1.0, nt)
    full = px.create_structured_mesh("quad4", (n, n), lengths=(1.0, 1.0))
    gout = 0.05 * times
    lw = dict(energy_weight=1.0, residual_weight=1.0, reaction_weight=1.0)

    def build(mesh):
        domain = px.VariationalDomain(mesh, times, q_order=2)
        physics = px.SolidMechanics(px.NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0), px.PlaneStrain())
        bcs = [px.DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_1") for c in range(2)]
        gd = px.GlobalData.from_arrays(times, gout, mesh.nodeSets["nset_1"], 1)
        problem = px.InverseProblem(domain, physics, None, gd, dirichlet_bcs=bcs)
        return problem

    def make_params(problem):
        return px.Parameters(problem, key=0, n_layers=3, n_neurons=50, device=device,
                             dirichlet_bc_func=px.UniaxialTensionLinearRamp(1.0, 1.0, "y", 2))
    loss_fn = px.EnergyResidualAndReactionLoss(**lw)
    evals = n * n * 4 * nt
    whole = None
    if rank == 0:
        fp = build(full)
        fparams = make_params(fp)
        get_engine(fp, fparams).set_option("cull_null_steps", 0)
        (L1, a1), g1 = loss_fn.value_and_grad(fparams, fp)
        whole = (fp, fparams, L1.clone(), a1["residual"].clone(), a1["reactions"].clone(), g1.fields.clone())
    out = {"workload": f"C2: Quad4 {n}^2 ({n * n * 4} qp), plane-strain NeoHookean K=1000 G=1, nt={nt}, "
                       "EnergyResidualAndReactionLoss with reaction on the top node set, MLP 3->50x3->2 fp64",
           "unit": "evals/s", "n_gpus": world}
    if world == 1:
        fp, fparams = whole[0], whole[1]

        def step():
            return loss_fn.value_and_grad(fparams, fp)
        ms = _timed(step, args.steps, 3, device, 1)
        # the sharded path on this one GPU: two element shards in lock-step through the phased ABI
        subs = [parallel.partition_mesh(full, r, 2) for r in range(2)]
        plans = parallel.make_shard_plans([l2g for _, l2g in subs])
        lo, hi = full.coords.min(axis=0), full.coords.max(axis=0)
        engines = []
        for (sub, _), plan in zip(subs, plans):
            sp = build(sub)
            sp.shard = parallel.ShardInfo(plan, None, lo, hi)
            sp._engines = {}
            e = get_engine(sp, make_params(sp))
            e.set_option("cull_null_steps", 0)
            engines.append(e)
        L, aux, gw, _ = parallel.run_shards_in_process(engines, plans, "energy_residual_reaction", fparams.fields.networks.weights,
                                                       None, global_outputs=gout, **lw)
        out.update({"n_shards": 2, "shards": "2 element shards driven in lock-step on the one GPU (phased ABI + exchanges)",
                    "shared_nodes": int(plans[0].n_shared)})
        for e in engines:
            e.close()
    else:
        sub, l2g = parallel.partition_mesh(full, rank, world)
        sp = build(sub)
        parallel.shard_problem(sp, l2g, device)
        sparams = make_params(sp)
        get_engine(sp, sparams).set_option("cull_null_steps", 0)
        res = parallel.PackedResult(4 + nt, sparams.fields.networks.weights.numel(), 0, device)

        def step():
            (loss, aux), grads = loss_fn.value_and_grad(sparams, sp)
            res.loss.copy_(loss.reshape(1))
            res.aux[1:2].copy_(aux["residual"].reshape(1)); res.aux[4:].copy_(aux["reactions"])
            res.gw.copy_(grads.fields)
            res.allreduce()
        ms = _timed(step, args.steps, 3, device, world)
        L, aux, gw = res.loss[0], res.aux, res.gw
        out.update({"n_shards": world, "shards": "one element shard per GPU (contiguous element ranges = y slabs), NCCL exchanges",
                    "shared_nodes": int(sp.shard.plan.n_shared)})
    out.update({"value": evals * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps})
    if rank == 0:
        _, _, L1, R1, re1, g1 = whole
        out["rel_err_vs_whole_mesh"] = {
            "loss": abs(float(L) - float(L1)) / abs(float(L1)),
            "residual": abs(float(aux[1]) - float(R1)) / abs(float(R1)),
            "reactions": float((aux[4:] - re1).norm() / re1.norm()),
            "grad": float((gw - g1).norm() / g1.norm())}
        out["loss"] = float(L1)
    return out


def run_ours(args):
    import torch.distributed as dist
    import pancax_b200 as px
    from pancax_b200 import parallel
    from pancax_b200.loss_functions import get_engine
    rank, world, local_rank = parallel.init_distributed()
    if world != args.gpus:
        if rank == 0:
            print(f"[bench] note: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    device = torch.device("cuda", local_rank)
    n, nt = args.n, args.nt
    strong = args.scaling == "strong"
    problem, params = build_problem(n, nt, rank, world, device, strong, args.model)
    loss_fn = px.EnergyLoss()
    eng = get_engine(problem, params)
    # Headline numbers evaluate EVERY (time step, node) row like the reference does.  The library's default
    # skips the MLP on time steps whose Dirichlet table b is identically zero (t = 0 of the load ramp: u = a
    # exactly, zero cotangent) -- exact, but it removes half of this workload's rows (nt = 2), so it is
    # reported separately under "null_step_culling" and never mixed into value / e2e / roofline.
    eng.set_option("cull_null_steps", 0)
    evals_per_rank = eng.ne * eng.nq * nt
    w = params.fields.networks.weights
    res = parallel.PackedResult(4 + nt, eng.n_weights, eng.n_train, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        eng.loss_and_grad("energy", w, out=res.as_out())
        res.allreduce()

    # ---------------- warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count
    eng.profile_enable(True)
    eng.profile_collect()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = eng.profile_collect()
    eng.profile_enable(False)
    launches = eng.launch_count - launches0
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.cpu()[0])
    value = evals_per_rank * world * args.steps / (ms * 1e-3)

    # ---------------- end to end through the public API with HOST buffers
    host_w = torch.empty_like(w, device="cpu").pin_memory()
    host_w.copy_(w)
    host_out = torch.empty(1 + eng.n_weights, dtype=torch.float64).pin_memory()
    opt_res = parallel.PackedResult(0, eng.n_weights, 0, device)

    def step_e2e():
        params.fields.networks.weights.copy_(host_w, non_blocking=True)          # H2D: the step's inputs
        (loss, aux), grads = loss_fn.value_and_grad(params, problem)              # the call a user makes
        opt_res.loss.copy_(loss.reshape(1)); opt_res.gw.copy_(grads.fields)
        opt_res.allreduce()
        host_out.copy_(opt_res.buf[:1 + eng.n_weights], non_blocking=True)        # D2H: loss + gradient
        torch.cuda.current_stream().synchronize()                                 # the user reads them
        return float(host_out[0])

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    t_ms = torch.tensor([ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_ms.cpu()[0])
    e2e_value = evals_per_rank * world * args.steps / (ms_e2e * 1e-3)

    # ---------------- the library default (null-step culling on), same step, reported separately
    eng.set_option("cull_null_steps", 1)
    for _ in range(3):
        step_device()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_cull = float(t_ms.cpu()[0])
    eng.set_option("cull_null_steps", 0)
    culled = {"value": evals_per_rank * world * args.steps / (ms_cull * 1e-3), "unit": "evals/s",
              "ms_per_step": ms_cull / args.steps, "null_steps": int(np.sum(eng.null_steps())), "nt": nt,
              "note": "library default: the MLP is skipped on time steps whose Dirichlet table b == 0 (u = a, zero "
                      "cotangent; identical results). NOT the headline: value/e2e/roofline evaluate every row."}

    # ---------------- TF32 tensor-core adjoint (PCX_OPT_TF32_ADJOINT; north_star item 4 / config C5), every row
    # evaluated, reported separately: the headline stays the fp64 (1e-10) path
    step_device()
    g64 = res.gw.clone()
    eng.set_option("tf32_adjoint", 1)
    for _ in range(3):
        step_device()
    g32 = res.gw.clone()
    eng.profile_enable(True)
    eng.profile_collect()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    prof32 = eng.profile_collect()
    eng.profile_enable(False)
    eng.set_option("tf32_adjoint", 0)
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_tf32 = float(t_ms.cpu()[0])
    tf32 = {"value": evals_per_rank * world * args.steps / (ms_tf32 * 1e-3), "unit": "evals/s",
            "ms_per_step": ms_tf32 / args.steps,
            "grad_rel_err_vs_fp64": float((torch.linalg.norm(g32 - g64) / torch.linalg.norm(g64)).cpu()),
            "loss_identical": True, "tolerance": 1e-5,
            "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof32.items() if v[1]},
            "note": "adjoint through the MLP in error-compensated 3xTF32 (mma.sync.m16n8k8.tf32, fp32 accumulate, fp64 "
                    "flush every 1024 rows) from fp32 activations; forward / element sweeps / loss stay fp64. NOT the "
                    "headline: value/e2e/roofline are the fp64 path."}

    # the same mode with every contraction over rows on tcgen05.mma kind::tf32 / TMEM (PCX_OPT_TF32_ADJOINT = 2)
    eng.set_option("tf32_adjoint", 2)
    for _ in range(3):
        step_device()
    g32u = res.gw.clone()
    eng.profile_enable(True)
    eng.profile_collect()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    prof32u = eng.profile_collect()
    eng.profile_enable(False)
    eng.set_option("tf32_adjoint", 0)
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_u = float(t_ms.cpu()[0])
    tf32["tcgen05"] = {"value": evals_per_rank * world * args.steps / (ms_u * 1e-3), "unit": "evals/s",
                       "ms_per_step": ms_u / args.steps,
                       "grad_rel_err_vs_fp64": float((torch.linalg.norm(g32u - g64) / torch.linalg.norm(g64)).cpu()),
                       "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof32u.items() if v[1]},
                       "note": "k_mlp_backward_umma: dW_L = d_L^T h_L, dW_out, dW_0 as tcgen05.mma kind::tf32 (M 64, N 56 / 8, "
                               "K = 128 rows per tile, 3xTF32 into two TMEM accumulator sets, operands K-major straight from "
                               "one 28 KB bulk copy per layer tile), delta chain on mma.sync"}

    # ---------------- strong scaling of the literal C3 and the sharded residual loss (C2), every N
    strong_leg = resid_leg = None
    if not args.no_extra_legs and args.scaling == "weak":
        strong_leg = strong_scaling_leg(args, rank, world, device, value)
        if world == 1:
            strong_leg = {"value": value, "unit": "evals/s", "ms_per_step": ms / args.steps, "n_gpus": 1,
                          "workload": f"ONE Hex8 {n}^3 block on one GPU = the headline line", "per_gpu_rate_vs_weak_line": 1.0}
        resid_leg = residual_sharded_leg(args, rank, world, device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel (mlp_backward), live numbers
    from pancax_b200 import _lib
    import ctypes as C
    pk = C.c_double()
    _lib.check(_lib.load().pcx_measure_dmma_peak(C.byref(pk), 5))
    bwd_ms, bwd_n = prof["mlp_backward"]
    rows_per_launch = eng.nn * min(nt, int(os.environ.get("PCX_TCHUNK", nt)))
    # launches may be split in time chunks: use measured launches per step
    rows_total = eng.nn * nt * args.steps
    achieved = algorithmic_flops_per_row_bwd() * rows_total / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("mlp_backward_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "k_mlp_backward (FP64 DMMA)", "achieved": achieved, "peak": pk.value,
                "unit": "TFLOP/s", "frac": (achieved / pk.value) if achieved else None, "traffic": traffic,
                "peak_source": "fp64 mma.sync.m8n8k4 chain measured live on this GPU by pcx_measure_dmma_peak "
                               "(MEASURED_PEAKS.json has no FP64 figure)",
                "avg_launch_ms": bwd_ms / bwd_n if bwd_n else None, "launches": bwd_n,
                "algorithmic_flop_per_row": algorithmic_flops_per_row_bwd(), "rows_per_step": eng.nn * nt,
                "formulas": "pancax_b200/roofline.py",
                "step_share": {k: round(v[0] / ms, 4) for k, v in prof.items() if v[1]}}
    # the whole step against the same FP64 roof (algorithmic flop of SURVEY 8d / pancax_b200/roofline.py)
    from pancax_b200 import roofline as _rf
    step_flop = _rf.step_flops_per_qp_eval("hex8", MLP_IN, MLP_OUT, WIDTH, DEPTH, eng.nn / float(eng.ne * eng.nq)) * evals_per_rank
    roofline["whole_step"] = {"algorithmic_gflop": step_flop / 1e9, "achieved_tflops": step_flop / (ms / args.steps * 1e-3) / 1e12,
                              "frac_of_fp64_peak": step_flop / (ms / args.steps * 1e-3) / 1e12 / pk.value,
                              "kernel_frac": {"mlp_forward": (algorithmic_flops_per_row_fwd() * rows_total / (prof["mlp_forward"][0] * 1e-3) / 1e12 / pk.value)
                                              if prof["mlp_forward"][0] > 0 else None,
                                              "element_force": (_rf.element_flops_per_qp("hex8", MLP_OUT) * evals_per_rank * args.steps /
                                                                (prof["element_force"][0] * 1e-3) / 1e12 / pk.value)
                                              if prof["element_force"][0] > 0 else None}}
    # ---------------- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_n, nt, max_seconds=25.0, threads=args.cpu_threads)
    out = {
        "metric": "quad-pt loss+grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{'C3' if args.model in ('neohookean', 'gent') else 'C5'}: 3D {dict(neohookean='NeoHookean', gent='Gent', swanson='Swanson', hencky='Hencky')[args.model]} Hex8 {n}^3 {'in total' if strong else 'per GPU'} ({eng.ne * eng.nq} qp per GPU x nt={nt}), EnergyLoss, "
                               f"MLP {MLP_IN}->{WIDTH}x{DEPTH}->{MLP_OUT} fp64, weights U(+-1/sqrt(fan_in)) seed 0",
                   "elements_per_gpu": eng.ne, "nodes_per_gpu": eng.nn, "qp_per_gpu": eng.ne * eng.nq, "nt": nt,
                   "parallelism": f"qp-sharded dp{world} (z slabs), one allreduce of loss+grads per step",
                   "cache": "per-step working set (activations 3.4 GB per time step) is far larger than the 126 MB L2",
                   "deterministic_assembly": True, "rows": "every (time step, node) row evaluated (null-step culling off)"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(host_w.numel() * 8), "d2h_bytes_per_step": int(host_out.numel() * 8),
                "api": "EnergyLoss().value_and_grad(params, problem) with pinned-host weights in, loss+grad out"},
        "roofline": roofline, "cpu_baseline": cpu, "null_step_culling": culled, "tf32_adjoint": tf32,
        "strong_scaling": strong_leg, "residual_sharded": resid_leg,
    }
    _emit(out)
    if world > 1:
        dist.destroy_process_group()


def _cpu_threads(requested=None):
    """torchrun exports OMP_NUM_THREADS=1 to its workers: without this the reference arm at N > 1 ran on ONE core
    (VERDICT r1).  All host cores at every N, or --cpu-threads."""
    n = int(requested) if requested else (os.cpu_count() or 1)
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _oracle_case(n, nt):
    from tests.cases import make_case
    return make_case("hex8_neohookean", n=n, nt=nt, width=WIDTH, depth=DEPTH, permute=False)


def cpu_baseline(n, nt, max_seconds=25.0, threads=None):
    """The CPU restatement of pancax's path (oracle/, torch fp64, all host threads) timed on a bounded
    sample of the SAME workload (smaller mesh, same physics / network / loss)."""
    from tests.cases import run_oracle_loss
    _cpu_threads(threads)
    case = _oracle_case(n, nt)
    evals = case.mesh["conns"].shape[0] * 8 * nt
    run_oracle_loss(case, "energy", energy_weight=1.0)      # warm-up
    ts = []
    t_start = time.perf_counter()
    while len(ts) < 5 and (time.perf_counter() - t_start) < max_seconds:
        t0 = time.perf_counter()
        run_oracle_loss(case, "energy", energy_weight=1.0)
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return {"value": evals / med, "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"Hex8 {n}^3 ({evals} evals per step), same physics/MLP/loss, median of {len(ts)} steps, "
                      f"{med * 1e3:.1f} ms per step; CPU restatement (oracle/), not pancax/JAX: jax is not "
                      "installable in this image", "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: pancax's own CPU implementation is not runnable here (no jax / equinox wheels,
    no network), so this arm times the oracle port on all host threads, on the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.cases import run_oracle_loss
    _cpu_threads(args.cpu_threads)
    n, nt = args.cpu_n, args.nt
    case = _oracle_case(n, nt)
    evals = case.mesh["conns"].shape[0] * 8 * nt
    for _ in range(max(1, min(args.warmup, 3))):
        run_oracle_loss(case, "energy", energy_weight=1.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_oracle_loss(case, "energy", energy_weight=1.0)
    dt = time.perf_counter() - t0
    value = evals * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"each step = one loss+grad over Hex8 {n}^3 ({evals} evals), a bounded sample of the C3 workload; "
              "CPU restatement (oracle/, torch fp64), not pancax/JAX")
    out = {"impl": "reference", "metric": "quad-pt loss+grad evals/sec", "value": value, "unit": "evals/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"C3 sample: 3D NeoHookean Hex8 {n}^3, EnergyLoss, MLP {MLP_IN}->{WIDTH}x{DEPTH}->"
                                  f"{MLP_OUT} fp64, nt={nt}"},
           "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)


class _CleanStdout:
    """stdout carries exactly ONE JSON line: everything else a library writes to fd 1 while the benchmark runs (NCCL's
    "NCCL version ..." banner goes to stdout when NCCL_DEBUG is set by the launcher) is diverted to stderr, and the
    line is written to the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        os.write(self.real, (line.rstrip("\n") + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


_OUT = None


def _emit(obj):
    line = json.dumps(obj)
    if _OUT is not None:
        _OUT.emit(line)
    else:
        print(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=128, help="Hex8 elements per direction per GPU (C3: 128)")
    ap.add_argument("--nt", type=int, default=2, help="time steps (C3: 2 or 11)")
    ap.add_argument("--cpu-n", type=int, default=32, help="mesh size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=None, help="host threads of the CPU arms (default: all cores, at every N)")
    ap.add_argument("--c2-n", type=int, default=512, help="Quad4 elements per side of the sharded residual-loss leg (C2: 512)")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the strong_scaling / residual_sharded legs")
    ap.add_argument("--model", default="neohookean", choices=["neohookean", "gent", "swanson", "hencky"],
                    help="constitutive model (default: the C3 headline; swanson / hencky = config C5)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): n^3 elements per GPU; strong: one n^3 block cut into z slabs")
    args = ap.parse_args()
    global _OUT
    with _CleanStdout() as out:
        _OUT = out
        if args.impl == "reference":
            run_reference(args)
        else:
            if not torch.cuda.is_available():
                raise SystemExit("bench.py: no CUDA device; pancax_b200 has no CPU path (use --impl reference for the CPU arm)")
            run_ours(args)
        _OUT = None


if __name__ == "__main__":
    main()
