"""pancax's examples/forward_problems/mechanics/example_hyper_visco_3d.py on the B200 path: SimpleFeFv (Neo-Hookean
equilibrium branch + one Prony branch) under the path-dependent energy loss -- forward sweep over the time steps with
the viscous states kept, reverse sweep, one MLP adjoint.

    python examples/viscoelastic_3d.py [n] [epochs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pancax_b200 import (Adam, DirichletBC, EnergyLoss, ForwardProblem, NeoHookean, Parameters, PostProcessor,   # noqa: E402
                         PronySeries, SimpleFeFv, SolidMechanics, ThreeDimensional, UniaxialTensionLinearRamp, VariationalDomain, WLF,
                         create_structured_mesh)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 100

mesh = create_structured_mesh("hex8", n)
times = np.linspace(0.0, 1.0, 11)
domain = VariationalDomain(mesh, times)
model = SimpleFeFv(NeoHookean(bulk_modulus=10.0, shear_modulus=0.855),
                   PronySeries(moduli=[1.0], relaxation_times=[10.0]), WLF(C1=17.44, C2=51.6, theta_ref=60.0))
physics = SolidMechanics(model, ThreeDimensional())
dirichlet_bcs = [DirichletBC(nset_name=s, component=c) for s in ("nset_3", "nset_5") for c in range(3)]
problem = ForwardProblem(domain, physics, [], dirichlet_bcs, [])

loss_function = EnergyLoss(is_path_dependent=True)          # PathDependentEnergyLoss in the reference's script
params = Parameters(problem, 0, dirichlet_bc_func=UniaxialTensionLinearRamp(0.5, 1.0, "y", 3))
opt = Adam(loss_function, learning_rate=1.0e-3, has_aux=True, clip_gradients=False)
opt_st = opt.init(params)
for epoch in range(epochs):
    params, opt_st, loss = opt.step(params, opt_st, problem)
    if epoch % 10 == 0:
        print(epoch, float(loss[0]), flush=True)

# results like the reference's script: nodal displacements, deformation gradient and the state variables (Fv); plus the
# internal force and the PK1 stress, evaluated step by step from each step's old state (pcx_state_step)
out = os.path.join(os.environ.get("PANCAX_WORKDIR", "/tmp/pancax_b200_example"), "output_visco.e")
os.makedirs(os.path.dirname(out), exist_ok=True)
pp = PostProcessor(mesh, "exodus")
pp.init(params, problem, out, node_variables=["field_values", "internal_force"],
        element_variables=["deformation_gradient", "state_variables", "pk1_stress"])
pp.write_outputs(params, problem)
pp.close()
print("wrote", out)
