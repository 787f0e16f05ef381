"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
fp64 loss + gradient (energy and residual / reaction losses), both TF32 adjoint kernels (mma.sync and tcgen05), the
op-level calls; sizes of a few thousand quadrature points so that the instrumented run finishes in a minute."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.cases import make_case, make_engine, run_engine_loss      # noqa: E402

torch.cuda.set_device(0)
for name, n, nt, kind in (("hex8_neohookean", 7, 3, "energy"), ("quad4_neohookean", 19, 3, "energy_residual_reaction"),
                          ("tri3_neohookean_bounded", 8, 2, "energy_residual")):
    case = make_case(name, n=n, nt=nt)
    eng = make_engine(case)
    ref = None
    for mode in (0, 1, 2):
        eng.set_option("tf32_adjoint", mode)
        L, aux, gw, gp = run_engine_loss(case, kind, eng=eng)
        if ref is None:
            ref = gw
        print(name, kind, "tf32_adjoint", mode, "loss", L, "grad rel", float(np.linalg.norm(gw - ref) / np.linalg.norm(ref)), flush=True)
    eng.close()
# models with state: path-dependent energy / residual + reaction losses (k_visco_qp MODE 0-3), PlaneStress with state (MODE 4 / 5)
for name, n, nt, kind in (("hex8_visco_neohookean", 3, 3, "energy"), ("hex8_visco_neohookean", 3, 3, "energy_residual_reaction"),
                          ("quad4_visco_gent_bounded", 5, 3, "energy_residual"), ("quad4_visco_neohookean_pstress", 5, 3, "energy")):
    case = make_case(name, n=n, nt=nt, width=16, depth=2)
    L, aux, gw, gp = run_engine_loss(case, kind, first_step=1)
    print(name, kind, "loss", L, "|grad|", float(np.linalg.norm(gw)), flush=True)
torch.cuda.synchronize()
print("done", flush=True)
