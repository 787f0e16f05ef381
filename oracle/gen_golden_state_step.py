"""Generates tests/golden/state_step.npz: the reference's OWN SimpleFeFv (hyperviscoelasticity/simple_fefv.py) evaluated
point by point at a GIVEN old state on the torch stand-in of oracle/refshim_torch.py -- energy(grad_u, theta, Z_old, dt)
-> (psi, Z_new) (:20-40) and pk1_stress = jax.jacfwd(energy, argnums=0) (constitutive_models/mechanics/base.py:112-118),
what the post-processor's element_pp(pk1_stress) and physics.potential_energy(..., state_old, dt) evaluate per
quadrature point (physics_kernels/base.py:24-74,349-354) -- and the same through the reference's PlaneStress formulation
(solid_mechanics.py:32-71: scalar_root_find.find_root on sigma_33 of the whole model at fixed old state).

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


NPS = 3


def ps_points(seed):
    """in-plane grad_u (NPS, 2, 2); old viscous deformation gradients with the block structure the plane-stress path
    keeps ([A 0; 0 c], A 2x2), as they evolve from the identity."""
    rng = np.random.default_rng(seed)
    H = 0.15 * rng.standard_normal((NPS, 2, 2))
    Z = np.zeros((NPS, 2, 3, 3))
    for k in range(NPS):
        for b in range(2):
            Z[k, b, :2, :2] = np.eye(2) + 0.08 * rng.standard_normal((2, 2))
            Z[k, b, 2, 2] = 1.0 + 0.05 * rng.standard_normal()
    return H, Z, np.array([0.05, 0.3, 1.0])


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
        # PlaneStress with state (solid_mechanics.py:49-71): the reference's own find_root on the Cauchy stress of the whole
        # model at fixed old state, point by point (~20 s per point on the stand-in), then energy / pk1_stress at that root
        case2 = make_case(f"quad4_visco_{model}_pstress", n=1, nt=3, width=8, depth=2, q_order=2)
        _, problem2, _, _ = build(case2)
        cm2, form = problem2.physics.constitutive_model, problem2.physics.formulation
        assert type(form).__name__ == "PlaneStress"
        H2, Z2, dt2 = ps_points(200 + m)
        x33, psi, Zn, P = [], [], [], []
        for k in range(NPS):
            zo = torch.as_tensor(Z2[k]).reshape(-1)
            H3 = form.modify_field_gradient(cm2, torch.as_tensor(H2[k]), 60.0, zo, float(dt2[k])).detach()
            e, zn = cm2.energy(H3, 60.0, zo, float(dt2[k]))
            pk = cm2.pk1_stress(H3, 60.0, zo, float(dt2[k]))
            pk = pk[0] if isinstance(pk, (tuple, list)) else pk
            x33.append(float(H3[2, 2])); psi.append(float(e.detach()))
            Zn.append(np.asarray(zn.detach()).reshape(2, 3, 3)); P.append(np.asarray(pk.detach()).reshape(3, 3))
        out[f"{model}__ps_grad_u"], out[f"{model}__ps_Z_old"], out[f"{model}__ps_dt"] = H2, Z2, dt2
        out[f"{model}__ps_grad_u_33"], out[f"{model}__ps_psi"] = np.array(x33), np.array(psi)
        out[f"{model}__ps_Z_new"], out[f"{model}__ps_P"] = np.array(Zn), np.array(P)
        out[f"{model}__ps_props"] = np.array([float(p) for p in case2.props])
        print(model, "plane stress: grad_u_33", x33, "P_33", [float(p[2, 2]) for p in P], flush=True)
    path = os.path.join(OUT, "state_step.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
