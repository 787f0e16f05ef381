"""A 3D hyperelastic forward problem (uniaxial tension of a Neo-Hookean block, deep-energy loss) in the style of
pancax's examples/forward_problems/mechanics/: EnergyLoss, Adam inside Trainer, results written as Exodus.

    python examples/hyperelastic_3d.py [n] [epochs]          n^3 Hex8 elements (default 32), 8 n^3 quadrature points
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pancax_b200 import (Adam, DirichletBC, EnergyLoss, ForwardProblem, NeoHookean, Parameters, PostProcessor,   # noqa: E402
                         SolidMechanics, ThreeDimensional, Trainer, UniaxialTensionLinearRamp, VariationalDomain,
                         create_structured_mesh)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 200

mesh = create_structured_mesh("hex8", n)
times = np.linspace(0.0, 1.0, 11)
domain = VariationalDomain(mesh, times)
physics = SolidMechanics(NeoHookean(bulk_modulus=10.0, shear_modulus=0.855), ThreeDimensional())
dirichlet_bcs = [DirichletBC(nset_name=s, component=c) for s in ("nset_3", "nset_5") for c in range(3)]
problem = ForwardProblem(domain, physics, [], dirichlet_bcs, [])

params = Parameters(problem, 0, n_layers=4, n_neurons=50,
                    dirichlet_bc_func=UniaxialTensionLinearRamp(final_displacement=0.5, length=1.0, direction="y", n_dimensions=3))
opt = Adam(EnergyLoss(), learning_rate=1e-3, has_aux=True)
trainer = Trainer(problem, opt, log_every=50, serialise_every=10 ** 9, workdir=os.environ.get("PANCAX_WORKDIR", "/tmp/pancax_b200_example"))
params = trainer.train(params, epochs)
print("energy after", epochs, "epochs:", float(EnergyLoss()(params, problem)[0]))

pp = PostProcessor(mesh, "exodus")
out = os.path.join(os.environ.get("PANCAX_WORKDIR", "/tmp/pancax_b200_example"), "output.e")
pp.init(params, problem, out, node_variables=["field_values"], element_variables=["deformation_gradient", "pk1_stress"])
pp.write_outputs(params, problem)
pp.close()
print("wrote", out)
