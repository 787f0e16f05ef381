"""Generates tests/golden/state_step.npz: the reference's OWN SimpleFeFv (hyperviscoelasticity/simple_fefv.py) evaluated
point by point at a GIVEN old state on the torch stand-in of oracle/refshim_torch.py -- energy(grad_u, theta, Z_old, dt)
-> (psi, Z_new) (:20-40) and pk1_stress = jax.jacfwd(energy, argnums=0) (constitutive_models/mechanics/base.py:112-118),
what the post-processor's element_pp(pk1_stress) and physics.potential_energy(..., state_old, dt) evaluate per
quadrature point (physics_kernels/base.py:24-74,349-354).

    python -m oracle.gen_golden_state_step          # build container only (needs /root/reference)

Both the oracle (tests/test_oracle_golden.py, CPU) and pcx_state_step on one-element meshes with an affine displacement
(GPU) are compared against the file.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .gen_golden_ad import OUT, setup

MODELS = ["neohookean", "gent", "swanson", "hencky"]
NPTS = 6


def points(seed):
    """grad_u (NPTS, 3, 3) of moderate size, old viscous deformation gradients (NPTS, nb, 3, 3) near the identity with
    unit determinant scale, dt per point."""
    rng = np.random.default_rng(seed)
    H = 0.15 * rng.standard_normal((NPTS, 3, 3))
    Z = np.eye(3) + 0.08 * rng.standard_normal((NPTS, 2, 3, 3))
    dt = np.array([0.05, 0.3, 1.0, 2.5, 0.3, 1.0])
    return H, Z, dt


def main():
    S = setup()
    build, make_case = S["build"], S["make_case"]
    out = {}
    for m, model in enumerate(MODELS):
        case = make_case(f"hex8_visco_{model}", n=1, nt=3, width=8, depth=2, q_order=2)
        params, problem, leaves, raws = build(case)
        cm = problem.physics.constitutive_model
        H, Z, dt = points(100 + m)
        psi, Zn, P = [], [], []
        for k in range(NPTS):
            gu = torch.as_tensor(H[k])
            zo = torch.as_tensor(Z[k]).reshape(-1)
            e, zn = cm.energy(gu, 60.0, zo, float(dt[k]))
            p = cm.pk1_stress(gu, 60.0, zo, float(dt[k]))
            p = p[0] if isinstance(p, (tuple, list)) else p          # jacfwd of (psi, Z): the energy's part
            psi.append(float(e)); Zn.append(np.asarray(zn.detach()).reshape(2, 3, 3)); P.append(np.asarray(p.detach()).reshape(3, 3))
        out[f"{model}__grad_u"], out[f"{model}__Z_old"], out[f"{model}__dt"] = H, Z, dt
        out[f"{model}__psi"], out[f"{model}__Z_new"], out[f"{model}__P"] = np.array(psi), np.array(Zn), np.array(P)
        out[f"{model}__props"] = np.array([float(p) for p in case.props])
        out[f"{model}__prony"] = np.array(case.prony)
        print(model, "psi", psi[:3], "|P|", float(np.linalg.norm(np.array(P))), flush=True)
    path = os.path.join(OUT, "state_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
