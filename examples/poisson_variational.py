"""pancax's examples/forward_problems/poisson/variational_example.py on the B200 path: the script reads like the
reference's (same classes, same argument meaning); numpy replaces jax.numpy in the user callables, which are
tabulated once on the host.  Needs a B200 and the built library (python -c "import __graft_entry__ as g; g.build()").

    python examples/poisson_variational.py [epochs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pancax_b200 import (Adam, DirichletBC, EnergyLoss, ForwardProblem, LBFGS, Parameters, Poisson,   # noqa: E402
                         VariationalDomain, create_structured_mesh)

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 500

# domain: the reference reads mesh_quad4.g (23 x 23 Quad4 of the unit square); read_exodus_mesh() takes that file
# unchanged, a structured mesh of the same shape is generated here so that the example has no file dependency
mesh = create_structured_mesh("quad4", 23)
times = np.linspace(0.0, 1.0, 2)
domain = VariationalDomain(mesh, times)

physics = Poisson(lambda x: 2 * np.pi ** 2 * np.sin(2. * np.pi * x[..., 0]) * np.sin(2. * np.pi * x[..., 1]))


def bc_func(x, t, z):
    x, y = x[0], x[1]
    return x * (1. - x) * y * (1. - y) * z


dirichlet_bcs = [DirichletBC(nset_name=f"nset_{i}", component=0) for i in (1, 2, 3, 4)]
problem = ForwardProblem(domain, physics, [], dirichlet_bcs, [])

loss_function = EnergyLoss()
params = Parameters(problem, 100, dirichlet_bc_func=bc_func)

opt = Adam(loss_function, learning_rate=1e-3, has_aux=True)
opt_st = opt.init(params)
for epoch in range(epochs):
    params, opt_st, loss = opt.step(params, opt_st, problem)
    if epoch % 100 == 0:
        print(epoch, float(loss[0]))

# switch to LBFGS (commented out in the reference's script)
opt = LBFGS(loss_function, has_aux=True)
opt_st = opt.init(params)
for epoch in range(20):
    params, opt_st, loss = opt.step(params, opt_st, problem)
    if opt_st.opt_state["done"]:
        break
print("final energy", float(loss[0]), " (continuum minimum for u = sin(2 pi x) sin(2 pi y): -pi^2 / 2 = %.4f)" % (-np.pi ** 2 / 2))
