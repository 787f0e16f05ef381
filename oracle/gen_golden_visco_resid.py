"""Generates tests/golden/visco_resid.npz: the reference's OWN path-dependent residual losses for a model with state
variables -- EnergyAndResidualLoss.path_dependent_call (weak_form_loss_functions.py:151-192) and
EnergyResidualAndReactionLoss.path_dependent_call (:242-276) with SimpleFeFv (hyperviscoelasticity/simple_fefv.py)
-- executed end to end on the torch stand-in of oracle/refshim_torch.py: the lax.scan over the steps threads the
viscous state, physics.potential_energy_residual_and_reaction_force (base.py:373-385) differentiates the step's
energy w.r.t. the nodal values at FIXED old state, and the outer gradient goes a second time through the state
update (mtk_log_sqrt's registered rule differentiated through the closed-form eigenvector algorithm, expm).

    python -m oracle.gen_golden_visco_resid          # build container only (needs /root/reference)

This is the call the reference's examples/inverse_problems/mechanics/path-dependent uses.  Both the oracle
(tests/test_loss_grad_golden.py, CPU) and the CUDA path (GPU) are compared against the file.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .gen_golden_ad import OUT, setup

# (case name, n, nt, width, depth, q_order)
CASES = [
    ("hex8_visco_neohookean", 2, 4, 12, 2, 2),
    ("quad4_visco_gent", 3, 3, 12, 2, 2),
    ("hex8_visco_hencky", 2, 3, 8, 2, 2),
    ("tri3_visco_neohookean_bounded", 3, 3, 10, 2, 2),
]


def main():
    S = setup()
    build, flat_grad, make_case, LW = S["build"], S["flat_grad"], S["make_case"], S["LOSS_WEIGHTS"]
    out = {}
    for name, n, nt, width, depth, q in CASES:
        case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
        tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
        out[f"{tag}__meta"] = np.array([n, nt, width, depth, q])
        for kind in ("energy_residual", "energy_residual_reaction"):
            params, problem, leaves, raws = build(case)
            if kind == "energy_residual":
                lf = S["EnergyAndResidualLoss"](energy_weight=LW["energy_weight"], is_path_dependent=True,
                                                residual_weight=LW["residual_weight"])
                # __call__ asserts for the path-dependent case (weak_form_loss_functions.py:145-147); the method is:
                loss, aux = lf.path_dependent_call(params, problem)
            else:
                lf = S["EnergyResidualAndReactionLoss"](energy_weight=LW["energy_weight"], is_path_dependent=True,
                                                        residual_weight=LW["residual_weight"],
                                                        reaction_weight=LW["reaction_weight"])
                loss, aux = lf(params, problem)
            gw, gp = flat_grad(loss, leaves, raws)
            out[f"{tag}__{kind}__loss"], out[f"{tag}__{kind}__gw"], out[f"{tag}__{kind}__gp"] = np.array(float(loss)), gw, gp
            for k, v in aux.items():
                out[f"{tag}__{kind}__aux_{k}"] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v)
            print(f"{tag:48s} {kind:26s} loss {float(loss): .12e} |gw| {np.linalg.norm(gw):.6e}"
                  + (f" gp {gp}" if gp.size else ""), flush=True)
    path = os.path.join(OUT, "visco_resid.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
