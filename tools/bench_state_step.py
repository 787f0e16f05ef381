"""Measures pcx_state_step (one time step of a SimpleFeFv model at a given old state: Pi, state_new, internal force,
PK1 stress at every quadrature point -- k_visco_qp MODE 2 + scatter + assembly) on a Hex8 n^3 mesh.
    python tools/bench_state_step.py [n] [n_prony] > profiles/rNN_state_step.json      (on a B200)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import pancax_b200 as px
    from pancax_b200 import Engine
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    mesh = px.create_structured_mesh("hex8", n)
    eng = Engine(mesh.coords, mesh.conns, "hex8", 2, 3, np.linspace(0.0, 1.0, 3))
    eng.set_physics("solid", "neohookean", "three_dimensional", [10.0, 0.855],
                    prony=([1.0 / (b + 1) for b in range(nb)], [10.0 / (b + 1) for b in range(nb)]))
    eng.set_field(4, 3, 50, 4)
    rng = np.random.default_rng(0)
    us = torch.as_tensor(0.01 * rng.standard_normal((eng.nn, 3)), device=eng.device)
    z = eng.initial_state()
    _, z, _, _ = eng.state_step(us * 0.5, z, 0.5)          # a non-trivial old state
    out = {}
    for label, kw in (("Pi_state_f_P", {}), ("Pi_state_f", {"want_P": False})):
        for _ in range(3):
            eng.state_step(us, z, 0.5, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            eng.state_step(us, z, 0.5, **kw)
        e1.record()
        torch.cuda.synchronize()
        out[label] = e0.elapsed_time(e1) / reps
    nqp = eng.ne * eng.nq
    print(json.dumps({"op": "pcx_state_step (k_visco_qp MODE 2 + k_visco_scatter + k_assemble)",
                      "workload": f"Hex8 {n}^3, {nqp} quadrature points, NeoHookean + {nb} Prony branch(es), fp64",
                      "ms_per_call": out, "qp_per_s": nqp / (out["Pi_state_f_P"] * 1e-3),
                      "bytes_per_qp": {"state_old_read": 72 * nb, "state_new_write": 72 * nb, "flux_write+read": 144, "P_write": 72}}))
    eng.close()


if __name__ == "__main__":
    main()
