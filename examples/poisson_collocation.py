"""pancax's examples/forward_problems/poisson/collocation_example.py on the B200 path: a Poisson problem solved in STRONG
form on the nodes of a Quad4 mesh -- residual -lap(u) - f at batches of collocation points, the Laplacian of the network
from the fused forward-mode jet kernels (no nested jacfwd / hessian), Adam.  The reference wraps
``mean((strong_form_residual - outputs)^2)`` in a UserDefinedLossFunction; that expression is StrongFormResidualLoss's
batch form here (its jet / adjoint are the library's kernels).

    python examples/poisson_collocation.py [n] [epochs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pancax_b200 import (Adam, CollocationDataLoader, CollocationDomain, DirichletBC, ForwardProblem, Parameters, Poisson,   # noqa: E402
                         StrongFormResidualLoss, create_structured_mesh)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 200

mesh = create_structured_mesh("quad4", n)
times = np.linspace(0.0, 0.0, 1)
domain = CollocationDomain(mesh, times)
physics = Poisson(lambda x: 2 * np.pi ** 2 * np.sin(2.0 * np.pi * x[..., 0]) * np.sin(2.0 * np.pi * x[..., 1]))


def bc_func(x, t, z):
    x, y = x[0], x[1]
    return x * (1.0 - x) * y * (1.0 - y) * z


essential_bcs = [DirichletBC(component=0, nset_name=f"nset_{k}") for k in (1, 2, 3, 4)]
problem = ForwardProblem(domain, physics, [], essential_bcs, [])
params = Parameters(problem, 10, dirichlet_bc_func=bc_func)

loss_function = StrongFormResidualLoss()          # (params, problem, inputs, outputs) -> mean((residual - outputs)^2)
opt = Adam(loss_function, learning_rate=1e-3, has_aux=True)
opt_st = opt.init(params)
dataloader = CollocationDataLoader(problem.domain, num_fields=1)
first = None
for epoch in range(epochs):
    for inputs, outputs in dataloader.dataloader(128):
        params, opt_st, (loss, aux) = opt.step(params, opt_st, problem, inputs, outputs)
    first = float(loss) if first is None else first
    if epoch % 50 == 0 or epoch == epochs - 1:
        print(epoch, float(loss), flush=True)
print("residual loss", first, "->", float(loss))
