"""Measures the post-processing element-field kernel (k_element_pp) on the C3 mesh:
HBM-bound, 19 doubles written per quadrature point (+ connectivity / gathered rows read).
    python tools/bench_pp.py > profiles/rNN_pp_kernel.json      (on a B200)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import pancax_b200 as px
    from pancax_b200 import Engine
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    mesh = px.create_structured_mesh("hex8", n)
    eng = Engine(mesh.coords, mesh.conns, "hex8", 2, 3, np.linspace(0.0, 1.0, 2))
    eng.set_physics("solid", "neohookean", "three_dimensional", [1000.0, 1.0])
    eng.set_field(4, 3, 50, 4)
    rng = np.random.default_rng(0)
    us = torch.as_tensor(0.01 * rng.standard_normal((eng.nn, 3)), device=eng.device)
    for _ in range(3):
        eng.element_fields(us)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        F, I1, P = eng.element_fields(us)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nqp = eng.ne * eng.nq
    algo_bytes = nqp * 19 * 8 + eng.ne * 8 * 4 + eng.nn * 6 * 8        # outputs + connectivity + node rows once
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6452.8}
    print(json.dumps({"kernel": "k_element_pp<8,3,3,neohookean> (F, I1_bar, P at every quadrature point)",
                      "workload": f"Hex8 {n}^3, {nqp} quadrature points", "ms_per_launch": ms,
                      "qp_per_s": nqp / (ms * 1e-3), "algorithmic_bytes_per_launch": algo_bytes,
                      "achieved_gbs": algo_bytes / (ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                      "frac": algo_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "bound": "hbm"}))
    eng.close()


if __name__ == "__main__":
    main()
