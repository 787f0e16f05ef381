"""pancax's examples/inverse_problems/mechanics/path-dependent/script.py on the B200 path: identify the equilibrium shear
modulus of a finite-strain viscoelastic model (SimpleFeFv: Neo-Hookean + one Prony branch, plane strain) from a
force-displacement record and full-field displacement data, with the loss of that script --

    EnergyResidualAndReactionLoss(residual_weight, reaction_weight, is_path_dependent=True) + FullFieldDataLoss

wrapped in a UserDefinedLossFunction, a load-hold Dirichlet wrapper (ramp until t = 1, then held), Adam.  The reference
reads its DIC / load-frame CSVs from examples/.../path-dependent/data; here the same kind of files are synthesised
first (the reference tree is not on the GPU box), then read back through FullFieldData / GlobalData.

    python examples/inverse_path_dependent.py [n] [epochs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pancax_b200 import (Adam, BoundedProperty, DirichletBC, EnergyResidualAndReactionLoss, FullFieldData,   # noqa: E402
                         FullFieldDataLoader, FullFieldDataLoss, GlobalData, InverseProblem, NeoHookean, Parameters, PlaneStrain,
                         PronySeries, SimpleFeFv, SolidMechanics, UserDefinedLossFunction, VariationalDomain, WLF,
                         create_structured_mesh)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 50
work = os.environ.get("PANCAX_WORKDIR", "/tmp/pancax_b200_example")
os.makedirs(work, exist_ok=True)

# ---- the two data files of the reference's example, synthesised: load to 1.0 until t = 1, then hold until t = 3
times_1 = np.linspace(0.0, 1.0, 6)
times_2 = np.linspace(1.0, 3.0, 6)
times = np.hstack((times_1, times_2[1:]))
disps = np.minimum(times, 1.0) * 0.1
forces = 0.35 * disps * (1.0 + 0.4 * np.exp(-np.maximum(times - 1.0, 0.0)))     # relaxing reaction
global_file = os.path.join(work, "global_data.csv")
with open(global_file, "w") as f:
    f.write("times, disps, forces\n" + "\n".join(f"{t:.16g}, {d:.16g}, {r:.16g}" for t, d, r in zip(times, disps, forces)))
rng = np.random.default_rng(0)
xy = rng.uniform(0.0, 1.0, (400, 2))
rows = [(x, y, t, -0.3 * (x - 0.5) * d, y * d) for t, d in zip(times, disps) for x, y in xy]
field_file = os.path.join(work, "full_field_data.csv")
with open(field_file, "w") as f:
    f.write("x, y, t, u_x, u_y\n" + "\n".join(", ".join(f"{v:.16g}" for v in r) for r in rows))

mesh = create_structured_mesh("quad4", n)
field_data = FullFieldData(field_file, ["x", "y", "t"], ["u_x", "u_y"])
global_data = GlobalData(global_file, "times", "disps", "forces", mesh, 1, "y", n_time_steps=len(times), interpolate=False)
domain = VariationalDomain(mesh, global_data.times, q_order=2)

model = SimpleFeFv(NeoHookean(bulk_modulus=10.0, shear_modulus=BoundedProperty(0.01, 5.0, key=0)),
                   PronySeries(moduli=[1.0], relaxation_times=[10.0]), WLF(C1=17.44, C2=51.6, theta_ref=60.0))
physics = SolidMechanics(model, PlaneStrain())


def dirichlet_bc_func(xs, t, nn):
    """Batched form of the reference script's wrapper: u_x = y (y - L) t z_x / L^2, u_y ramps with t until t = 1 and is
    held afterwards (its jax.lax.cond on t > 1), plus the same bubble times the network output."""
    length, final_displacement = 1.0, 0.1
    y = xs[:, 1]
    t = np.broadcast_to(np.asarray(t, dtype=np.float64), y.shape)
    bubble = y * (y - length) * t / length ** 2
    out = np.empty_like(nn)
    out[:, 0] = bubble * nn[:, 0]
    out[:, 1] = y * np.minimum(t, 1.0) * final_displacement / length + bubble * nn[:, 1]
    return out


dirichlet_bcs = [DirichletBC(component=c, nset_name=s) for s in ("nset_1", "nset_3") for c in range(2)]
problem = InverseProblem(domain, physics, field_data, global_data, [], dirichlet_bcs, [])
params = Parameters(problem, 0, dirichlet_bc_func=dirichlet_bc_func)

physics_and_global_loss = EnergyResidualAndReactionLoss(residual_weight=250.0, reaction_weight=250.0, is_path_dependent=True)
full_field_data_loss = FullFieldDataLoss(weight=10.0)


def loss_function(params, problem, inputs, outputs):
    loss_1, aux_1 = physics_and_global_loss(params, problem)
    loss_2, aux_2 = full_field_data_loss(params, problem, inputs, outputs)
    aux_1.update(aux_2)
    return loss_1 + loss_2, aux_1


loss_function = UserDefinedLossFunction(loss_function)
opt = Adam(loss_function, learning_rate=1.0e-3, has_aux=True, transition_steps=50000)
opt_st = opt.init(params)
loader = FullFieldDataLoader(field_data)
for epoch in range(epochs):
    for inputs, outputs in loader.dataloader(1024):
        params, opt_st, (loss, aux) = opt.step(params, opt_st, problem, inputs, outputs)
    if epoch % 10 == 0 or epoch == epochs - 1:
        print(epoch, float(loss), float(aux["energy"]), float(aux["residual"]), float(aux["global_data_loss"]),
              float(aux["field_data_loss"]), params.properties(), flush=True)
print("done", epochs)
