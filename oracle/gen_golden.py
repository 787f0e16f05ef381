"""Generates tests/golden/*.npz by EXECUTING the reference's own Python source (pancax 0.0.16 under
/root/reference) for the forward pieces of the hot path, through the numpy stand-in for jax/equinox
in oracle/refshim (jax itself cannot be installed in this image).

    python -m oracle.gen_golden            # run in the build container (needs /root/reference)

The vectors pin the oracle (tests/test_oracle_golden.py, CPU) and, transitively, the CUDA path
(tests/test_gpu_parity.py compares CUDA against the oracle and directly against some of these).
Derivatives: the stand-in has no AD, so P = dW/dF goldens are complex-step derivatives of the
reference's energy code (NeoHookean, Swanson: exact to rounding) or central differences (Gent,
Hencky: ~1e-9), and the log-sqrt rule is the reference's registered custom-JVP function called
directly.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, ROOT)
    from oracle import refshim
    refshim.import_reference(REF)
    from pancax.fem.elements import Hex8Element, Quad4Element, SimplexTriElement
    from pancax.fem.quadrature_rules import QuadratureRule
    from pancax.fem.function_space import NonAllocatedFunctionSpace
    from pancax.fem.dof_manager import DofManager
    from pancax.bcs import DirichletBC
    from pancax.constitutive_models import Gent, Hencky, NeoHookean, Swanson, BoundedProperty
    from pancax.math import tensor_math
    from pancax.physics_kernels import Poisson, SolidMechanics, PlaneStrain, ThreeDimensional
    from pancax.bvps.uniaxial_tension import UniaxialTensionLinearRamp
    from pancax.bvps.simple_shear import SimpleShearLinearRamp
    from pancax.bvps.biaxial_tension import BiaxialLinearRamp
    from pancax_b200.fem import read_exodus_mesh

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(2024)

    # ------------------------------------------------------------------ 1. parent tables / rules
    tabs = {}
    elems = {"hex8": Hex8Element(), "quad4": Quad4Element(), "tri3": SimplexTriElement(1)}
    for name, q_orders in (("hex8", (1, 2)), ("quad4", (1, 2, 3)), ("tri3", (1, 2))):
        el = elems[name]
        tabs[f"{name}_nodes"] = np.asarray(el.coordinates)
        for q in q_orders:
            qr = QuadratureRule(el, q)
            sf = el.compute_shapes(el.coordinates, qr.xigauss)
            tabs[f"{name}_q{q}_xi"] = np.asarray(qr.xigauss)
            tabs[f"{name}_q{q}_w"] = np.asarray(qr.wgauss)
            tabs[f"{name}_q{q}_N"] = np.asarray(sf.values)
            tabs[f"{name}_q{q}_dN"] = np.asarray(sf.gradients)
    errs = {}
    for name, q in (("hex8", 3), ("quad4", 4)):
        try:
            QuadratureRule(elems[name], q)
            errs[name] = ""
        except ValueError as e:
            errs[name] = str(e)
    tabs["hex8_q3_error"] = np.array(errs["hex8"])
    tabs["quad4_q4_error"] = np.array(errs["quad4"])
    np.savez(os.path.join(OUT, "tables.npz"), **tabs)

    # ------------------------------------------------------------------ 2. fixture meshes + function space
    meshes = {
        "hex8": os.path.join(REF, "test/mesh.g"),
        "quad4": os.path.join(REF, "test/fem/mesh_quad4.g"),
        "tri3": os.path.join(REF, "examples/inverse_problems/mechanics/vanilla/mesh/mesh.g"),
    }
    fs = {}
    loaded = {}
    for name, path in meshes.items():
        m = read_exodus_mesh(path)
        loaded[name] = m
        fs[f"{name}_coords"], fs[f"{name}_conns"] = m.coords, m.conns
        for k, v in m.nodeSets.items():
            fs[f"{name}_ns_{k}"] = v
        mesh_ns = types.SimpleNamespace(parentElement=elems[name])
        space = NonAllocatedFunctionSpace(mesh_ns, QuadratureRule(elems[name], 2))
        sel = np.arange(0, m.conns.shape[0], max(1, m.conns.shape[0] // 40))[:40]
        gN = np.stack([np.asarray(space.shape_function_gradients(m.coords[m.conns[e]])) for e in sel])
        JxW = np.stack([np.asarray(space.JxWs(m.coords[m.conns[e]])) for e in sel])
        fs[f"{name}_sel"], fs[f"{name}_gradN"], fs[f"{name}_JxW"] = sel, gN, JxW
        # total volume over ALL elements
        fs[f"{name}_volume"] = np.array(sum(float(np.sum(space.JxWs(m.coords[m.conns[e]])))
                                            for e in range(m.conns.shape[0])))
    # DofManager (fem/dof_manager.py) on the hex fixture
    hm = loaded["hex8"]
    mesh_ns = types.SimpleNamespace(num_nodes=hm.num_nodes, nodeSets=hm.nodeSets, conns=hm.conns)
    bcs = [DirichletBC(component=c, nset_name=s) for s in ("nset_3", "nset_5") for c in range(3)]
    dm = DofManager(mesh_ns, 3, bcs)
    fs["hex8_unknownIndices"] = np.asarray(dm.unknownIndices)
    np.savez_compressed(os.path.join(OUT, "function_space.npz"), **fs)

    # ------------------------------------------------------------------ 3. constitutive energies
    con = {}
    SW = dict(bulk_modulus=10.0, A1=0.93074321, P1=-0.07673672, B1=0.0, Q1=-0.5, C1=0.10448312, R1=1.71691036,
              cutoff_strain=0.01)
    SW2 = dict(SW, B1=0.21, Q1=0.5)
    models = {
        "neohookean": NeoHookean(bulk_modulus=0.833, shear_modulus=0.3846),
        "neohookean_stiff": NeoHookean(bulk_modulus=1000.0, shear_modulus=1.0),
        "gent": Gent(bulk_modulus=0.833, shear_modulus=0.3846, Jm_parameter=3.0),
        "swanson": Swanson(**SW),
        "swanson_b": Swanson(**SW2),
        "hencky": Hencky(bulk_modulus=0.833, shear_modulus=0.3846),
    }
    H = 0.25 * rng.standard_normal((48, 3, 3))
    H[:8] *= 0.01                       # near-identity cases
    H[8:16, 2, :] = 0.0                 # plane-strain embeddings
    H[8:16, :, 2] = 0.0
    H[16] = 0.0                         # F = I exactly
    H[17] = np.diag([0.3, 0.3, -0.1])   # repeated eigenvalue
    H[18] = np.diag([1.4, 0.05, 0.02])  # beyond the Gent guard rail (I1bar > Jm + 3 - 0.001)
    con["grad_u"] = H
    state = np.zeros(0)

    def W_of(model, g):
        return model.energy(g, 60.0, state, 1.0)[0]

    for name, model in models.items():
        con[f"{name}_W"] = np.array([float(W_of(model, h)) for h in H])
        P = np.zeros_like(H)
        if name.startswith("neohookean") or name.startswith("swanson"):
            h = 1e-30                   # complex step through the reference's own energy code
            for n in range(H.shape[0]):
                for i in range(3):
                    for j in range(3):
                        g = H[n].astype(complex)
                        g[i, j] += 1j * h
                        P[n, i, j] = np.imag(W_of(model, g)) / h
            con[f"{name}_P_method"] = np.array("complex-step")
        else:
            h = 1e-6                    # central differences (comparisons / sorting are not analytic)
            for n in range(H.shape[0]):
                for i in range(3):
                    for j in range(3):
                        gp, gm = H[n].copy(), H[n].copy()
                        gp[i, j] += h
                        gm[i, j] -= h
                        P[n, i, j] = (float(W_of(model, gp)) - float(W_of(model, gm))) / (2 * h)
            con[f"{name}_P_method"] = np.array("central-difference h=1e-6")
        con[f"{name}_P"] = P
        con[f"{name}_J"] = np.array([float(model.jacobian(h)) for h in H])
        con[f"{name}_I1bar"] = np.array([float(model.I1_bar(h)) for h in H])
    con["swanson_I2bar"] = np.array([float(models["swanson"].I2_bar(h)) for h in H])
    # BoundedProperty.__call__ (properties.py:33-38)
    bp = BoundedProperty.__new__(BoundedProperty)
    bp.prop_min, bp.prop_max = 0.01, 5.0
    raws = np.linspace(-3, 3, 13)
    vals = []
    for r in raws:
        bp.prop_val = np.array([r])
        vals.append(float(bp()))
    con["bounded_raw"], con["bounded_val"] = raws, np.array(vals)
    np.savez(os.path.join(OUT, "constitutive.npz"), **con)

    # ------------------------------------------------------------------ 4. sym-3x3 eigen / log-sqrt + its JVP rule
    tm = {}
    A = rng.standard_normal((24, 3, 3))
    C = np.einsum("nki,nkj->nij", A, A) + 0.2 * np.eye(3)
    C[0] = np.eye(3)                                      # triple eigenvalue
    C[1] = np.diag([2.0, 2.0, 0.5])                       # double eigenvalue
    Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    C[2] = Q @ np.diag([1.3, 1.3, 0.7]) @ Q.T             # rotated double eigenvalue
    C[3] = Q @ np.diag([1.0, 1.02, 1.04]) @ Q.T           # inside the 5 % Taylor switch
    C[4] = 1e-6 * C[4]
    C[5] = np.eye(3) + 1e-9 * (A[5] + A[5].T)             # F ~ I to rounding
    Hs = rng.standard_normal((24, 3, 3))
    Hs = 0.5 * (Hs + Hs.transpose(0, 2, 1))
    tm["C"], tm["H"] = C, Hs
    ev, evec, ls, jv = [], [], [], []
    for n in range(C.shape[0]):
        lam, V = tensor_math.eigen_sym33_unit(C[n])
        ev.append(np.asarray(lam)); evec.append(np.asarray(V))
        ls.append(np.asarray(tensor_math.mtk_log_sqrt(C[n])))
        prim, tang = tensor_math.mtk_log_sqrt.jvp_rule((C[n],), (Hs[n],))
        jv.append(np.asarray(tang))
    tm["evals"], tm["evecs"], tm["log_sqrt"], tm["log_sqrt_jvp"] = map(np.stack, (ev, evec, ls, jv))
    xs = np.linspace(0, 1, 21)
    tm["cos3_x"], tm["cos3"] = xs, np.array([float(tensor_math.cos_of_acos_divided_by_3(x)) for x in xs])
    l1 = np.array([1.0, 1.0, 1.0, 2.0, 0.3, 1.0])
    l2 = np.array([1.0, 1.04, 1.06, 0.5, 0.31, 1.0 + 1e-12])
    tm["rld_l1"], tm["rld_l2"] = l1, l2
    tm["rld"] = np.array([float(tensor_math.relative_log_difference(a, b)) for a, b in zip(l1, l2)])
    np.savez(os.path.join(OUT, "tensor_math.npz"), **tm)

    # ------------------------------------------------------------------ 5. potential energy on the fixtures
    pe = {}

    def run_pi(name, physics, us, nd):
        m = loaded[name]
        mesh_ns = types.SimpleNamespace(parentElement=elems[name])
        space = NonAllocatedFunctionSpace(mesh_ns, QuadratureRule(elems[name], 2))
        dom = types.SimpleNamespace(coords=m.coords, conns=m.conns, fspace=space)
        nq = len(space.quadrature_rule)
        state_old = np.zeros((m.conns.shape[0], nq, 0))
        pi, _ = physics.potential_energy(physics, dom, 1.0, us, state_old, 1.0)
        return float(pi)

    def smooth_u(x, ndof):
        cols = [0.11 * np.sin(1.3 * x[:, 0] + 0.4 * x[:, 1]) + 0.07 * x[:, 1] ** 2,
                0.09 * x[:, 0] * x[:, 1] - 0.05 * np.cos(2.0 * x[:, 0]),
                0.06 * np.sin(x[:, 0] - x[:, -1]) + 0.03 * x[:, -1]]
        return np.column_stack(cols[:ndof])

    for mname, model in (("neohookean_stiff", models["neohookean_stiff"]), ("gent", models["gent"]),
                         ("swanson", models["swanson"]), ("hencky", models["hencky"])):
        us = smooth_u(loaded["hex8"].coords, 3)
        pe[f"hex8_{mname}_Pi"] = np.array(run_pi("hex8", SolidMechanics(model, ThreeDimensional()), us, 3))
        us2 = smooth_u(loaded["quad4"].coords, 2)
        pe[f"quad4_{mname}_Pi"] = np.array(run_pi("quad4", SolidMechanics(model, PlaneStrain()), us2, 2))
    us3 = smooth_u(loaded["tri3"].coords, 2) * 0.05
    pe["tri3_neohookean_Pi"] = np.array(run_pi("tri3", SolidMechanics(models["neohookean"], PlaneStrain()), us3, 2))
    pe["hex8_us"], pe["quad4_us"], pe["tri3_us"] = smooth_u(loaded["hex8"].coords, 3), \
        smooth_u(loaded["quad4"].coords, 2), us3
    import jax.numpy as jnp

    def src(x):
        return 2 * jnp.pi ** 2 * jnp.sin(2.0 * jnp.pi * x[0]) * jnp.sin(2.0 * jnp.pi * x[1])
    upo = smooth_u(loaded["quad4"].coords, 1)
    pe["quad4_poisson_us"] = upo
    pe["quad4_poisson_Pi"] = np.array(run_pi("quad4", Poisson(src), upo, 2))
    np.savez_compressed(os.path.join(OUT, "potential_energy.npz"), **pe)

    # ------------------------------------------------------------------ 6. Dirichlet wrappers
    bv = {}
    pts2, pts3 = rng.uniform(0, 1, (16, 2)), rng.uniform(0, 1, (16, 3))
    z2, z3 = rng.standard_normal((16, 2)), rng.standard_normal((16, 3))
    bv["pts2"], bv["pts3"], bv["z2"], bv["z3"], bv["t"] = pts2, pts3, z2, z3, np.array(0.7)

    def ev(f, pts, z):
        return np.stack([np.asarray(f(refshim.ShimArray.__new__(refshim.ShimArray, p.shape, buffer=p.copy()), 0.7,
                                     np.array(zz).view(refshim.ShimArray))) for p, zz in zip(pts, z)])
    for d in ("x", "y"):
        bv[f"uniaxial2_{d}"] = ev(UniaxialTensionLinearRamp(0.8, 1.5, d, 2), pts2, z2)
    for d in ("x", "y", "z"):
        bv[f"uniaxial3_{d}"] = ev(UniaxialTensionLinearRamp(0.8, 1.5, d, 3), pts3, z3)
    bv["simple_shear"] = ev(SimpleShearLinearRamp(0.8, 1.5, "xy"), pts2, z2)
    bv["biaxial"] = ev(BiaxialLinearRamp(0.8, 0.3, 1.5, 1.2), pts2, z2)
    np.savez(os.path.join(OUT, "bvps.npz"), **bv)
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
