"""Profiling target for the state-model residual losses: two loss + gradient calls of the path-dependent
EnergyResidualAndReactionLoss with SimpleFeFv(NeoHookean, 2 Prony branches) on Hex8 n^3, nt = 11.  Per call k_visco_qp
runs nt-1 MODE 2 launches (forward sweep + force flux) and then nt-1 MODE 3 launches (reverse sweep, second order):
    ncu --set full -k k_visco_qp --launch-skip 25 -c 1 ... python tools/visco_profile_target.py   # a MODE 2 launch
    ncu --set full -k k_visco_qp --launch-skip 35 -c 1 ... python tools/visco_profile_target.py   # a MODE 3 launch
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                   # noqa: E402
from tests.cases import make_case, make_engine, LOSS_WEIGHTS   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
case = make_case("hex8_visco_neohookean", n=n, nt=11, width=50, depth=4, permute=False)
eng = make_engine(case)
for _ in range(2):
    out = eng.loss_and_grad("energy_residual_reaction", case.weights, None, global_outputs=case.global_outputs, first_step=1,
                            **LOSS_WEIGHTS)
torch.cuda.synchronize()
print(float(out[0].cpu()[0]))
eng.close()
