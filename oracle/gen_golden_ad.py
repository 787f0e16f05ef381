"""Generates tests/golden/loss_grad.npz: END-TO-END loss / aux / gradient vectors obtained by
EXECUTING the reference's own loss functions, physics kernels, Field, function space, constitutive
models and Dirichlet wrappers (pancax 0.0.16 under /root/reference) on the torch-backed stand-in
for jax/equinox in oracle/refshim_torch.py (every jax.grad / value_and_grad on the path becomes
torch.autograd.grad(create_graph=True); jax itself cannot be installed in this image).

    python -m oracle.gen_golden_ad          # build container only (needs /root/reference)

Inputs are the seeded cases of tests/cases.py (weights injected, not re-drawn), so the oracle
(tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_api.py, GPU) are both compared
against these files.  The only restated piece on this path is equinox.nn.MLP/Linear (un-vendored
third-party; see refshim_torch.py).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

# (case name, n, nt, width, depth, q_order, losses)
CASES = [
    ("hex8_neohookean", 3, 2, 50, 4, 2, ("energy", "energy_residual", "energy_residual_reaction")),
    ("hex8_gent", 2, 2, 16, 2, 2, ("energy", "energy_residual_reaction")),
    ("hex8_swanson", 2, 2, 16, 2, 2, ("energy", "energy_residual_reaction")),
    ("hex8_hencky", 2, 2, 16, 2, 2, ("energy", "energy_residual", "energy_residual_reaction")),
    ("quad4_neohookean", 5, 3, 50, 3, 2, ("energy", "energy_residual", "energy_residual_reaction")),
    ("quad4_neohookean_bounded", 4, 2, 20, 2, 2, ("energy", "energy_residual_reaction")),
    ("quad4_gent_bounded", 4, 2, 20, 2, 2, ("energy", "energy_residual_reaction")),
    ("quad4_swanson", 4, 2, 20, 2, 3, ("energy",)),
    ("tri3_neohookean_bounded", 5, 3, 20, 3, 2, ("energy", "energy_residual_reaction")),
    ("tri3_hencky", 4, 2, 20, 2, 1, ("energy",)),
    ("tri3_hencky_bounded", 4, 3, 20, 2, 2, ("energy_residual_reaction",)),
    ("quad4_hencky", 3, 3, 20, 2, 2, ("energy_residual",)),
    ("poisson_quad4", 6, 2, 50, 3, 2, ("energy",)),
    # PlaneStress (solid_mechanics.py:32-71): the reference's own root find + custom_root, per quadrature point
    ("quad4_neohookean_pstress", 3, 2, 16, 2, 2, ("energy", "energy_residual", "energy_residual_reaction")),
    ("tri3_gent_pstress_bounded", 3, 3, 16, 2, 2, ("energy_residual_reaction",)),
    ("quad4_hencky_pstress", 2, 2, 12, 2, 2, ("energy", "energy_residual")),
    # Blatz-Ko (hyperelasticity/blatz_ko.py)
    ("hex8_blatz_ko", 2, 2, 16, 2, 2, ("energy", "energy_residual_reaction")),
    ("quad4_blatz_ko_bounded", 3, 3, 16, 2, 2, ("energy_residual_reaction",)),
]


# State-variable model: SimpleFeFv around the named equilibrium model, EnergyLoss(is_path_dependent=True)
# (weak_form_loss_functions.py:48-72): (case name, n, nt, width, depth, q_order)
VISCO_CASES = [
    ("hex8_visco_neohookean", 2, 4, 12, 2, 2),
    ("quad4_visco_gent", 3, 3, 12, 2, 2),
    ("hex8_visco_hencky", 2, 3, 8, 2, 2),
]

# first time of the load ramp.  At t = 0 the ramped Dirichlet wrapper gives u = 0, hence f = 0 EXACTLY for
# Swanson, and the reference's d||f||/df is NaN there; that case starts at t0 > 0 instead.
# Hencky: at t = 0 (F = I, repeated stretches) the reference's second derivative through eigen_sym33_unit is NaN as
# well -- its residual-loss cases start at t0 > 0 too.
T0 = {"hex8_swanson": 0.25, "hex8_blatz_ko": 0.25, "quad4_blatz_ko_bounded": 0.2, "hex8_hencky": 0.25, "tri3_hencky_bounded": 0.2, "quad4_hencky": 0.2,
      "quad4_hencky_pstress": 0.2}


def setup():
    """Import the reference on the torch stand-in; returns build(case) -> (params, problem, leaves, raws), flat_grad and
    the reference's loss classes (shared with oracle/gen_golden_visco_resid.py)."""
    sys.path.insert(0, ROOT)
    from oracle import refshim_torch as shim
    shim.import_reference(REF)
    from pancax.fem.elements import Hex8Element, Quad4Element, SimplexTriElement
    from pancax.fem.quadrature_rules import QuadratureRule
    from pancax.fem.function_space import NonAllocatedFunctionSpace
    from pancax.constitutive_models import BlatzKo, Gent, Hencky, NeoHookean, Swanson, BoundedProperty
    from pancax.physics_kernels import Poisson, SolidMechanics, PlaneStrain, PlaneStress, ThreeDimensional
    from pancax.bvps.uniaxial_tension import UniaxialTensionLinearRamp
    from pancax.loss_functions.weak_form_loss_functions import (EnergyLoss, EnergyAndResidualLoss,
                                                                EnergyResidualAndReactionLoss)
    from pancax.loss_functions.data_loss_functions import FullFieldDataLoss
    from pancax.networks.parameters import Parameters
    import jax.numpy as jnp
    from tests.cases import LOSS_WEIGHTS, make_case, SWANSON_PROPS

    elems = {"hex8": Hex8Element(), "quad4": Quad4Element(), "tri3": SimplexTriElement(1)}
    model_cls = {"neohookean": NeoHookean, "gent": Gent, "swanson": Swanson, "hencky": Hencky, "blatz_ko": BlatzKo}
    prop_names = {"neohookean": ("bulk_modulus", "shear_modulus"),
                  "gent": ("bulk_modulus", "shear_modulus", "Jm_parameter"),
                  "swanson": tuple(SWANSON_PROPS.keys()), "hencky": ("bulk_modulus", "shear_modulus"),
                  "blatz_ko": ("shear_modulus",)}

    def build(case):
        mesh = case.mesh
        el = elems[mesh["elem"]]
        space = NonAllocatedFunctionSpace(types.SimpleNamespace(parentElement=el), QuadratureRule(el, case.q_order))
        coords = torch.as_tensor(mesh["coords"], dtype=torch.float64)
        conns = torch.as_tensor(mesh["conns"], dtype=torch.long)
        times = torch.as_tensor(case.times, dtype=torch.float64)
        raws = []
        if case.kind == "solid":
            kw = {}
            for nme, p in zip(prop_names[case.model], case.props):
                if isinstance(p, tuple):
                    bp = BoundedProperty.__new__(BoundedProperty)      # skip the jax.random draw: inject
                    object.__setattr__(bp, "prop_min", p[0])
                    object.__setattr__(bp, "prop_max", p[1])
                    raw = torch.tensor([p[2]], dtype=torch.float64, requires_grad=True)
                    object.__setattr__(bp, "prop_val", raw)
                    raws.append(raw)
                    kw[nme] = bp
                else:
                    kw[nme] = float(p)
            model = model_cls[case.model](**kw)
            if case.prony is not None:      # simple_fefv.py:11-14 (the converters of PronySeries' fields applied by hand)
                from pancax.constitutive_models import PronySeries, SimpleFeFv, WLF
                model = SimpleFeFv(model, PronySeries(moduli=torch.tensor(case.prony[0], dtype=torch.float64),
                                                      relaxation_times=torch.tensor(case.prony[1], dtype=torch.float64)),
                                   WLF(C1=17.44, C2=51.6, theta_ref=60.0))
            form = {"three_dimensional": ThreeDimensional, "plane_strain": PlaneStrain,
                    "plane_stress": PlaneStress}[case.formulation]()
            physics = SolidMechanics(model, form)
            bc = UniaxialTensionLinearRamp(case.bc[1], case.bc[2], case.bc[3], coords.shape[1])
        else:
            physics = Poisson(lambda x: 2 * jnp.pi ** 2 * jnp.sin(2.0 * jnp.pi * x[0]) * jnp.sin(2.0 * jnp.pi * x[1]))

            def bc(x, t, z):                 # examples/forward_problems/poisson/variational_example.py:30-32
                return x[0] * (1.0 - x[0]) * x[1] * (1.0 - x[1]) * z
        dofm = types.SimpleNamespace(unknownIndices=None if case.unknown_idx is None
                                     else torch.as_tensor(case.unknown_idx, dtype=torch.long))
        domain = types.SimpleNamespace(coords=coords, conns=conns, fspace=space, dof_manager=dofm, times=times)
        gd = None
        if case.reaction_nodes is not None:
            gd = types.SimpleNamespace(outputs=torch.as_tensor(case.global_outputs, dtype=torch.float64),
                                       reaction_nodes=torch.as_tensor(case.reaction_nodes, dtype=torch.long),
                                       reaction_dof=int(case.reaction_dof))
        problem = types.SimpleNamespace(coords=coords, times=times, domain=domain, physics=physics, global_data=gd,
                                        n_dims=coords.shape[1], n_dofs=case.ndof, conns=conns, fspace=space)
        n_in, n_out, width, depth = case.mlp
        import pancax.networks.fields as ref_fields
        # Field.__init__ (fields.py:21-77) with its defaults n_layers=3/n_neurons=50 overridden per case
        fld = ref_fields.Field(problem, np.array([0, 0], dtype=np.uint32), n_layers=depth, n_neurons=width,
                               dirichlet_bc_func=bc)
        params = Parameters.__new__(Parameters)
        from pancax.networks.parameters import _Parameters
        inner = _Parameters.__new__(_Parameters)
        for k, v in (("fields", fld), ("physics", physics), ("state", None)):
            object.__setattr__(inner, k, v)
        for k, v in (("is_ensemble", False), ("n_ensemble", 1), ("parameters", inner)):
            object.__setattr__(params, k, v)
        # inject the case's weights (equinox leaf order)
        leaves = shim.mlp_leaves(fld.networks.mlp)
        off = 0
        with torch.no_grad():
            for leaf in leaves:
                n = leaf.numel()
                leaf.copy_(torch.as_tensor(case.weights[off:off + n]).reshape(leaf.shape))
                off += n
        assert off == case.weights.size, (off, case.weights.size)
        return params, problem, leaves, raws

    def flat_grad(loss, leaves, raws):
        gs = torch.autograd.grad(loss, leaves + raws, allow_unused=True)
        gw = torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(gs[:len(leaves)], leaves)])
        gp = torch.cat([g.reshape(-1) for g in gs[len(leaves):]]) if raws else torch.zeros(0, dtype=torch.float64)
        return gw.numpy(), gp.numpy()

    return dict(locals())


def main():
    S = setup()
    build, flat_grad, make_case, LOSS_WEIGHTS = S["build"], S["flat_grad"], S["make_case"], S["LOSS_WEIGHTS"]
    EnergyLoss, EnergyAndResidualLoss = S["EnergyLoss"], S["EnergyAndResidualLoss"]
    EnergyResidualAndReactionLoss, FullFieldDataLoss = S["EnergyResidualAndReactionLoss"], S["FullFieldDataLoss"]
    out = {}

    for name, n, nt, width, depth, q, kinds in CASES:
        case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q, t0=T0.get(name, 0.0))
        tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
        out[f"{tag}__meta"] = np.array([n, nt, width, depth, q])
        for kind in kinds:
            params, problem, leaves, raws = build(case)
            if kind == "energy":
                lf = EnergyLoss(weight=LOSS_WEIGHTS["energy_weight"])
                loss, aux = lf(params, problem)
            elif kind == "energy_residual":
                lf = EnergyAndResidualLoss(energy_weight=LOSS_WEIGHTS["energy_weight"],
                                           residual_weight=LOSS_WEIGHTS["residual_weight"])
                # __call__ drops the value (weak_form_loss_functions.py:145-149); the path it runs is:
                loss, aux = lf.path_independent_call(params, problem)
            else:
                lf = EnergyResidualAndReactionLoss(energy_weight=LOSS_WEIGHTS["energy_weight"],
                                                   residual_weight=LOSS_WEIGHTS["residual_weight"],
                                                   reaction_weight=LOSS_WEIGHTS["reaction_weight"])
                loss, aux = lf(params, problem)
            gw, gp = flat_grad(loss, leaves, raws)
            out[f"{tag}__{kind}__loss"] = np.array(float(loss))
            out[f"{tag}__{kind}__gw"] = gw
            out[f"{tag}__{kind}__gp"] = gp
            for k, v in aux.items():
                out[f"{tag}__{kind}__aux_{k}"] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v)
            print(f"{tag:48s} {kind:26s} loss {float(loss): .12e} |gw| {np.linalg.norm(gw):.6e}"
                  + (f" gp {gp}" if gp.size else ""), flush=True)
        # internal force / residual / reaction of the first time step > 0 (base.py:356-385)
        if case.kind == "solid":
            params, problem, leaves, raws = build(case)
            fld, physics, _ = params
            t = problem.times[-1]
            us = physics.vmap_field_values(fld, problem.coords, t).detach()
            ne, nq = problem.domain.conns.shape[0], len(problem.domain.fspace.quadrature_rule)
            state = torch.zeros((ne, nq, 0), dtype=torch.float64)
            (pi, _), f = physics.potential_energy_and_internal_force(params, problem.domain, t, us, state, 1.0)
            out[f"{tag}__us_last"] = us.numpy()
            out[f"{tag}__Pi_last"] = np.array(float(pi))
            out[f"{tag}__f_last"] = f.detach().numpy()
            # element variables of the post-processor (physics_kernels/base.py:24-158 element_pp wrappers
            # registered in solid_mechanics.py:115-170): F, I1_bar, pk1_stress at the quadrature points
            pp = physics.update_var_name_to_method().var_name_to_method
            for key, name in (("deformation_gradient", "F"), ("I1_bar", "I1bar"), ("pk1_stress", "P")):
                val = pp[key]["method"](params, problem.domain, t, us, 60.0, state, 1.0)
                out[f"{tag}__pp_{name}"] = val.detach().numpy()

    # FullFieldDataLoss (data_loss_functions.py:13-24) on a seeded batch
    case = make_case("quad4_neohookean", n=5, nt=3, width=50, depth=3)
    params, problem, leaves, raws = build(case)
    rng = np.random.default_rng(5)
    inputs = np.column_stack([rng.uniform(0, 1, (64, 2)), rng.uniform(0, 1, 64)])
    outputs = 0.05 * rng.standard_normal((64, 2))
    lf = FullFieldDataLoss(weight=3.0)
    loss, aux = lf(params, problem, torch.as_tensor(inputs), torch.as_tensor(outputs))
    gw, _ = flat_grad(loss, leaves, [])
    out["ffd_quad4_n5_nt3_w50_d3__inputs"], out["ffd_quad4_n5_nt3_w50_d3__outputs"] = inputs, outputs
    out["ffd_quad4_n5_nt3_w50_d3__loss"], out["ffd_quad4_n5_nt3_w50_d3__gw"] = np.array(float(loss)), gw
    print("FullFieldDataLoss", float(loss), np.linalg.norm(gw))

    # ---------------------------------------------------------------- path-dependent calls (stateless models)
    # weak_form_loss_functions.py:48-72 (EnergyLoss.path_dependent_call) and :242-276
    # (EnergyResidualAndReactionLoss.path_dependent_call): the scan starts at time step 1.
    case = make_case("quad4_neohookean", n=5, nt=4, width=20, depth=2, t0=0.1)
    for kind, lf in (("energy", EnergyLoss(is_path_dependent=True, weight=LOSS_WEIGHTS["energy_weight"])),
                     ("energy_residual_reaction", EnergyResidualAndReactionLoss(
                         energy_weight=LOSS_WEIGHTS["energy_weight"], is_path_dependent=True,
                         residual_weight=LOSS_WEIGHTS["residual_weight"], reaction_weight=LOSS_WEIGHTS["reaction_weight"]))):
        params, problem, leaves, raws = build(case)
        loss, aux = lf(params, problem)
        gw, _ = flat_grad(loss, leaves, raws)
        out[f"pathdep_quad4_neohookean_n5_nt4_w20_d2__{kind}__loss"] = np.array(float(loss))
        out[f"pathdep_quad4_neohookean_n5_nt4_w20_d2__{kind}__gw"] = gw
        for k, v in aux.items():
            out[f"pathdep_quad4_neohookean_n5_nt4_w20_d2__{kind}__aux_{k}"] = np.asarray(v.detach().numpy() if isinstance(v, torch.Tensor) else v)
        print("path-dependent", kind, float(loss), np.linalg.norm(gw))

    # ---------------------------------------------------------------- state variables: SimpleFeFv, path-dependent EnergyLoss
    for name, n, nt, width, depth, q in VISCO_CASES:
        case = make_case(name, n=n, nt=nt, width=width, depth=depth, q_order=q)
        tag = f"{name}_n{n}_nt{nt}_w{width}_d{depth}_q{q}"
        params, problem, leaves, raws = build(case)
        loss, aux = EnergyLoss(is_path_dependent=True, weight=LOSS_WEIGHTS["energy_weight"])(params, problem)
        gw, _ = flat_grad(loss, leaves, raws)
        out[f"{tag}__energy__loss"], out[f"{tag}__energy__gw"] = np.array(float(loss)), gw
        out[f"{tag}__energy__aux_energy"] = np.asarray(aux["energy"].detach().numpy())
        print("visco path-dependent", tag, float(loss), np.linalg.norm(gw), flush=True)

    # ---------------------------------------------------------------- strong form
    # Poisson: StrongFormResidualLoss.__call__ (strong_form_loss_functions.py:15-27) -> physics.strong_form_residual
    # (poisson.py:26-29) -> field_laplacians = trace(jax.hessian(field_values)) (base.py:213-226), with the
    # collocation example's Dirichlet wrapper x(1-x)y(1-y)z (examples/.../poisson/collocation_example.py:11-13).
    from pancax.loss_functions.strong_form_loss_functions import StrongFormResidualLoss
    from pancax.physics_kernels.heat_equation import HeatEquation
    sfo = {}
    case = make_case("poisson_quad4", n=4, nt=2, width=20, depth=3)
    params, problem, leaves, raws = build(case)
    dom = types.SimpleNamespace(coords=problem.coords, times=problem.times)
    loss, aux = StrongFormResidualLoss(weight=2.5)(params, dom)
    gw, _ = flat_grad(loss, leaves, [])
    fld, physics, _ = params
    res = torch.stack([torch.stack([physics.strong_form_residual(fld, x, t) for x in problem.coords])
                       for t in problem.times])                       # (nt, nn, 1)
    sfo["poisson_quad4_n4_nt2_w20_d3__loss"], sfo["poisson_quad4_n4_nt2_w20_d3__gw"] = np.array(float(loss)), gw
    sfo["poisson_quad4_n4_nt2_w20_d3__residuals"] = res.detach().numpy()
    # the jet itself at the nodes, last time step: base.py:205-229
    tl = problem.times[-1]
    sfo["poisson_quad4_n4_nt2_w20_d3__u"] = torch.stack([physics.field_values(fld, x, tl) for x in problem.coords]).detach().numpy()
    sfo["poisson_quad4_n4_nt2_w20_d3__grad_u"] = torch.stack([physics.field_gradients(fld, x, tl) for x in problem.coords]).detach().numpy()
    sfo["poisson_quad4_n4_nt2_w20_d3__lap_u"] = torch.stack([physics.field_laplacians(fld, x, tl) for x in problem.coords]).detach().numpy()
    sfo["poisson_quad4_n4_nt2_w20_d3__u_t"] = torch.stack([physics.field_time_derivatives(fld, x, tl) for x in problem.coords]).detach().numpy()
    print("strong-form Poisson", float(loss), np.linalg.norm(gw))
    # Heat: HeatEquation.strong_form_residual (heat_equation.py:11-17) evaluated per point with params =
    # (field, (rho, c, k)); StrongFormResidualLoss passes params.fields only, so the reference's own loss cannot
    # reach this residual -- the mean of squares below is assembled here, the residuals are the reference's.
    params, problem, leaves, raws = build(case)
    fld, _, _ = params
    heat = HeatEquation()
    rck = (1.7, 0.9, 0.35)
    res = torch.stack([torch.stack([heat.strong_form_residual((fld, rck), x, t) for x in problem.coords])
                       for t in problem.times])
    loss = 1.5 * torch.stack([torch.square(res[n]).mean() for n in range(res.shape[0])]).mean()
    gw, _ = flat_grad(loss, leaves, [])
    sfo["heat_quad4_n4_nt2_w20_d3__loss"], sfo["heat_quad4_n4_nt2_w20_d3__gw"] = np.array(float(loss)), gw
    sfo["heat_quad4_n4_nt2_w20_d3__residuals"] = res.detach().numpy()
    print("strong-form heat", float(loss), np.linalg.norm(gw))
    np.savez_compressed(os.path.join(OUT, "strong_form.npz"), **sfo)

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "loss_grad.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
