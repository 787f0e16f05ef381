"""Launch-bound regime: the reference's own CPU-runnable config C1 (2D Poisson deep-energy PINN on the
23x23 Quad4 mesh, MLP 3->50x3->1, nt = 2: 2 116 quadrature points).  One loss+gradient step is ~12
kernel launches of a few microseconds each, so the step is bound by launch latency; the C ABI never
allocates or synchronises inside a compute call, which makes the whole step capturable in a CUDA graph.
    python tools/bench_small.py          (on a B200)  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from tests.cases import make_case, make_engine
    case = make_case("poisson_quad4", n=23, nt=2, width=50, depth=3, permute=False)
    eng = make_engine(case)
    w = torch.as_tensor(case.weights, device=eng.device)
    out = (torch.empty(1, dtype=torch.float64, device=eng.device), torch.empty(4 + 2, dtype=torch.float64, device=eng.device),
           torch.empty(eng.n_weights, dtype=torch.float64, device=eng.device), torch.empty(1, dtype=torch.float64, device=eng.device))
    evals = eng.ne * eng.nq * eng.nt

    def step():
        eng.loss_and_grad("energy", w, out=out)

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    eager_us = (time.perf_counter() - t0) / n * 1e6
    ref = [o.clone() for o in out]
    l0 = eng.launch_count
    step()
    launches = eng.launch_count - l0

    # capture the same step on a side stream and replay it
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        with torch.cuda.graph(g, stream=s):
            step()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay()
    torch.cuda.synchronize()
    graph_us = (time.perf_counter() - t0) / n * 1e6
    same = all(torch.equal(a, b) for a, b in zip(out, ref))
    print(json.dumps({"workload": "C1: 2D Poisson Quad4 23x23 (2116 qp x nt=2), EnergyLoss, MLP 3->50x3->1 fp64",
                      "kernel_launches_per_step": int(launches), "eager_us_per_step": eager_us,
                      "cuda_graph_us_per_step": graph_us, "evals_per_s_eager": evals / (eager_us * 1e-6),
                      "evals_per_s_graph": evals / (graph_us * 1e-6), "graph_replay_bitwise_equal_to_eager": bool(same)}))
    eng.close()


if __name__ == "__main__":
    main()
