"""Quick device-time probe of the fused loss+grad at bench size (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.cases import make_case, make_engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
det = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t0 = time.time()
case = make_case("hex8_neohookean", n=n, nt=nt, permute=False)
print("case built", time.time() - t0, flush=True)
eng = make_engine(case, deterministic=bool(det))
print("engine built", time.time() - t0, flush=True)
w = torch.as_tensor(case.weights).cuda()
out = None
for i in range(3):
    out = eng.loss_and_grad("energy", w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for i in range(K):
    out = eng.loss_and_grad("energy", w)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
evals = eng.ne * eng.nq * nt
print(f"n={n} nt={nt} det={det}: {ms:.3f} ms/step, {evals/ms/1e6:.3f} G evals/s, loss={float(out[0][0]):.12e}")
# per-op timing
def timeit(f, K=5):
    f(); torch.cuda.synchronize()
    e0.record()
    for i in range(K): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
us = eng.field_forward(w, 1)
print("field_forward (pack+fwd, 1 t): %.3f ms" % timeit(lambda: eng.field_forward(w, 1, out=us)))
Pi, f, _ = eng.energy_force(us)
print("energy_force (sweep+assemble+reduce, 1 t): %.3f ms" % timeit(lambda: eng.energy_force(us)))
gw = eng.field_backward(w, 1, f)
print("field_backward (pack+fwd+bwd+reduce, 1 t): %.3f ms" % timeit(lambda: eng.field_backward(w, 1, f, out=gw)))
print("geometry: %.3f ms" % timeit(lambda: eng.element_geometry()))
