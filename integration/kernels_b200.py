"""pancax/kernels_b200.py -- the file a pancax maintainer adds to make the UNCHANGED pancax Python API call the B200
library (libpancax_b200.so) through XLA custom calls (jax.ffi).  Companion of integration/pcx_xla_ffi.cc.

STATUS: source only.  jax / jaxlib are not installable in this repository's build image (no wheels, no network), so this
module has never been imported with a real jax; it is syntax-checked (tests/test_integration_sources.py) and every
custom call it issues is type-checked on the C++ side against the binding it names.  The SAME C ABI is exercised end to
end by this repository's ctypes binding (pancax_b200/_lib.py) in the GPU test suite.

What it does
  * set-up (host, ctypes): pcx_create / pcx_set_physics / pcx_set_field / pcx_reserve from a pancax problem + Field --
    mesh arrays and the tabulated Dirichlet wrapper are device arrays owned here (XLA never frees them under the handle);
  * ops (inside jit): jax.ffi.ffi_call on the handlers of pcx_xla_ffi.cc, wrapped in jax.custom_vjp where pancax
    differentiates through them;
  * hooks S1-S3: opt-in monkey patches of BasePhysics.vmap_field_values,
    BaseEnergyFormPhysics.potential_energy_and_internal_force and AbstractOptimizer._single_step.
Nothing here falls back to the jnp implementation: a configuration the library refuses raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import jax
import jax.numpy as jnp
import numpy as np

# descriptors of include/pcx_abi.h -- identical to pancax_b200/_lib.py (kept in one place when vendored into pancax)
from pancax_b200._lib import (ELEM, FORM, LOSS, MODEL, PCX_AUX_FIXED, PHYS_POISSON, PHYS_SOLID, FieldDesc, MeshDesc,
                              PhysicsDesc, check, load as load_abi)

FFI_TARGETS = ("PcxElementGeometry", "PcxFieldForward", "PcxFieldBackward", "PcxPointsForward", "PcxPointsBackward",
               "PcxEnergyForce", "PcxStiffnessAction", "PcxResidualReaction", "PcxElementFields", "PcxLossAndGrad",
               "PcxLossBegin", "PcxLossMid", "PcxLossFinish", "PcxFieldDataLossAndGrad", "PcxPointsJet",
               "PcxStrongFormLossAndGrad", "PcxGetState", "PcxStateStep", "PcxAdamStep", "PcxAdamStepSched")
_registered = False


def register(ffi_so: str = "libpcx_xla_ffi.so"):
    """jax.ffi.register_ffi_target for every handler symbol of pcx_xla_ffi.cc (once per process)."""
    global _registered
    if _registered:
        return
    so = C.CDLL(ffi_so)
    for name in FFI_TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(so, name)), platform="CUDA")
    _registered = True


def _dev_ptr(x) -> int:
    return int(x.unsafe_buffer_pointer())


def flat_weights(networks):
    """Leaves of params.fields.networks in equinox order [W0 (out,in), b0, W1, b1, ..., Wout]: pcx_field_desc's layout."""
    leaves = [l for l in jax.tree_util.tree_leaves(networks) if hasattr(l, "dtype") and jnp.issubdtype(l.dtype, jnp.inexact)]
    return jnp.concatenate([l.ravel() for l in leaves])


def unflatten_like(networks, flat):
    """Inverse of flat_weights: a pytree shaped like `networks` whose inexact-array leaves are slices of `flat`."""
    leaves, treedef = jax.tree_util.tree_flatten(networks)
    out, off = [], 0
    for l in leaves:
        if hasattr(l, "dtype") and jnp.issubdtype(l.dtype, jnp.inexact):
            out.append(flat[off:off + l.size].reshape(l.shape))
            off += l.size
        else:
            out.append(l)
    return jax.tree_util.tree_unflatten(treedef, out)


@dataclass
class Handle:
    """One pcx_handle (one GPU) + the device arrays it borrows."""
    ptr: int
    nn: int
    ne: int
    nq: int
    nnpe: int
    nd: int
    ndof: int
    nt: int
    n_weights: int
    n_train: int
    keep: tuple          # device arrays the library reads through raw pointers: must outlive the handle

    def close(self):
        if self.ptr:
            load_abi().pcx_destroy(C.c_void_p(self.ptr))
            self.ptr = 0


def make_handle(problem, field, deterministic: bool = True) -> Handle:
    """pcx_create / pcx_set_physics / pcx_set_field from a pancax ForwardProblem / InverseProblem and a Field.
    Static inputs: domain.coords / conns / times (pancax/domains.py:132-156), dof_manager.unknownIndices
    (fem/dof_manager.py:31-42), global_data (data.py:159-243), the constitutive model's properties, Field's
    normalisation constants (networks/fields.py:66-76) and its Dirichlet wrapper tabulated at (time, node) -- with jax
    itself, so the table is what the reference's own wrapper returns."""
    from pancax.constitutive_models.properties import BoundedProperty
    lib = load_abi()
    dom = problem.domain
    elem = {8: "hex8", 4: "quad4", 3: "tri3"}[int(dom.conns.shape[1])]
    coords = jnp.asarray(dom.coords, dtype=jnp.float64)
    conns = jnp.asarray(dom.conns, dtype=jnp.int32)
    times = np.ascontiguousarray(np.asarray(dom.times), dtype=np.float64)
    nn, nd = coords.shape
    ndof, nt = int(problem.n_dofs), times.shape[0]
    keep = [coords, conns]
    md = MeshDesc()
    md.elem_type, md.q_order, md.ndof = ELEM[elem], int(getattr(dom, "q_order", 2)), ndof
    md.ne, md.nn, md.nt, md.deterministic = int(conns.shape[0]), int(nn), int(nt), 1 if deterministic else 0
    md.coords, md.conns = _dev_ptr(coords), _dev_ptr(conns)
    if getattr(dom, "dof_manager", None) is not None:
        unk = jnp.asarray(dom.dof_manager.unknownIndices, dtype=jnp.int64)
        keep.append(unk)
        md.unknown_idx, md.n_unknown = _dev_ptr(unk), int(unk.shape[0])
    gd = getattr(problem, "global_data", None)
    if gd is not None:
        rn = jnp.asarray(gd.reaction_nodes, dtype=jnp.int32)
        keep.append(rn)
        md.reaction_nodes, md.n_reaction, md.reaction_dof = _dev_ptr(rn), int(rn.shape[0]), int(gd.reaction_dof)
    md.times = times.ctypes.data_as(C.POINTER(C.c_double))
    h = C.c_void_p()
    check(lib.pcx_create(C.byref(h), C.byref(md)))

    pd = PhysicsDesc()
    phys = problem.physics
    n_train = 0
    if type(phys).__name__ == "Poisson":
        pd.physics = PHYS_POISSON
        xq = jnp.zeros((md.ne, lib.pcx_nq(h), nd), dtype=jnp.float64)
        check(lib.pcx_element_geometry(h, None, None, C.c_void_p(_dev_ptr(xq)), None))
        src = jnp.asarray(jax.vmap(jax.vmap(phys.f))(xq), dtype=jnp.float64)      # poisson.py:17 tabulated once
        keep.append(src)
        pd.source_q = _dev_ptr(src)
    else:
        model = phys.constitutive_model
        prony = None
        if type(model).__name__ == "SimpleFeFv":
            prony, model = model.prony_series, model.eq_model
        pd.physics, pd.model = PHYS_SOLID, MODEL[{"NeoHookean": "neohookean", "Gent": "gent", "Swanson": "swanson",
                                                  "Hencky": "hencky", "BlatzKo": "blatz_ko"}[type(model).__name__]]
        pd.formulation = FORM[{"PlaneStrain": "plane_strain", "ThreeDimensional": "three_dimensional",
                               "PlaneStress": "plane_stress"}[type(phys.formulation).__name__]]
        names = [f for f in model.__dataclass_fields__]
        pd.nprops = len(names)
        for i, name in enumerate(names):
            p = getattr(model, name)
            if isinstance(p, BoundedProperty):
                pd.bounded[i], pd.prop_min[i], pd.prop_max[i] = 1, float(p.prop_min), float(p.prop_max)
                n_train += 1
            else:
                pd.value[i] = float(p)
        if prony is not None:
            pd.n_prony = int(np.asarray(prony.moduli).shape[0])
            for b in range(pd.n_prony):
                pd.prony_moduli[b], pd.prony_times[b] = float(prony.moduli[b]), float(prony.relaxation_times[b])
    check(lib.pcx_set_physics(h, C.byref(pd)))

    net = field.networks
    fd = FieldDesc()
    sizes = [l.shape for l in jax.tree_util.tree_leaves(net) if hasattr(l, "ndim") and l.ndim == 2]
    fd.n_in, fd.n_out, fd.width, fd.depth = nd + 1, ndof, int(sizes[0][0]), len(sizes) - 1
    fd.use_final_bias = 0                                               # networks/mlp.py:61 default
    fd.normalize_time = 1 if bool(field.normalize_time) else 0
    for j in range(nd):
        fd.x_min[j], fd.x_max[j] = float(field.x_mins[j]), float(field.x_maxs[j])
    fd.t_min, fd.t_max = float(field.t_min), float(field.t_max)
    bc = field.dirichlet_bc_func
    zeros, ones = jnp.zeros(ndof), jnp.ones(ndof)
    a = jnp.stack([jax.vmap(lambda x: bc(x, t, zeros))(coords) for t in dom.times])          # (nt, nn, ndof)
    b = jnp.stack([jax.vmap(lambda x: bc(x, t, ones))(coords) for t in dom.times]) - a
    probe = jnp.asarray(np.random.default_rng(0).standard_normal(ndof))
    chk = jnp.stack([jax.vmap(lambda x: bc(x, t, probe))(coords) for t in dom.times])
    if float(jnp.max(jnp.abs(chk - (a + b * probe)))) > 1e-9 * max(1.0, float(jnp.max(jnp.abs(chk)))):
        raise NotImplementedError("dirichlet_bc_func is not diagonal-affine in the network output z (SURVEY F7)")
    a, b = jnp.asarray(a, dtype=jnp.float64), jnp.asarray(b, dtype=jnp.float64)
    keep += [a, b]
    fd.bc_a, fd.bc_b = _dev_ptr(a), _dev_ptr(b)
    check(lib.pcx_set_field(h, C.byref(fd)))
    return Handle(int(h.value), int(nn), int(md.ne), int(lib.pcx_nq(h)), int(conns.shape[1]), int(nd), ndof, int(nt),
                  int(lib.pcx_num_weights(h)), n_train, tuple(keep))


def _f64(*shape):
    return jax.ShapeDtypeStruct(shape, jnp.float64)


_EMPTY = None


def _empty():
    global _EMPTY
    if _EMPTY is None:
        _EMPTY = jnp.zeros((0,), dtype=jnp.float64)
    return _EMPTY


# ------------------------------------------------------------------------------------------------ ops
def field_values(h: Handle, weights, t_idx: int):
    """S1: BasePhysics.vmap_field_values(field, domain.coords, times[t_idx]) with its VJP."""
    @jax.custom_vjp
    def f(w):
        return jax.ffi.ffi_call("PcxFieldForward", _f64(h.nn, h.ndof))(w, handle=np.int64(h.ptr), t_idx=np.int32(t_idx))

    def fwd(w):
        return f(w), w

    def bwd(w, g):
        return (jax.ffi.ffi_call("PcxFieldBackward", _f64(h.n_weights))(w, g, handle=np.int64(h.ptr), t_idx=np.int32(t_idx)),)
    f.defvjp(fwd, bwd)
    return f(weights)


def potential_energy_and_internal_force(h: Handle, prop_raw, us):
    """S2: (Pi, f) of base.py:349-361.  custom_vjp: the cotangent on Pi returns ct f (and ct dPi/dprops), the cotangent
    v on f returns K v (the tangent is symmetric for a hyperelastic W) and d<f, v>/dprops."""
    outs = (_f64(), _f64(h.nn, h.ndof), _f64(h.n_train))

    @jax.custom_vjp
    def pf(p, u):
        Pi, f, _ = jax.ffi.ffi_call("PcxEnergyForce", outs)(p, u, handle=np.int64(h.ptr))
        return Pi, f

    def fwd(p, u):
        Pi, f, gp = jax.ffi.ffi_call("PcxEnergyForce", outs)(p, u, handle=np.int64(h.ptr))
        return (Pi, f), (p, u, f, gp)

    def bwd(res, ct):
        p, u, f, gp = res
        ct_Pi, ct_f = ct
        Kv, gp2 = jax.ffi.ffi_call("PcxStiffnessAction", (_f64(h.nn, h.ndof), _f64(h.n_train)))(p, u, ct_f, handle=np.int64(h.ptr))
        return ct_Pi * gp + gp2, ct_Pi * f + Kv
    pf.defvjp(fwd, bwd)
    return pf(prop_raw, us)


def element_fields(h: Handle, prop_raw, us):
    """(F (ne, nq, 3, 3), I1_bar (ne, nq), P (ne, nq, 3, 3)) at the quadrature points: the element_pp wrappers behind the
    post-processor's deformation_gradient / I1_bar / pk1_stress (physics_kernels/base.py:24-158)."""
    outs = (_f64(h.ne, h.nq, 3, 3), _f64(h.ne, h.nq), _f64(h.ne, h.nq, 3, 3))
    return jax.ffi.ffi_call("PcxElementFields", outs)(prop_raw, us, handle=np.int64(h.ptr))


def state_step(h: Handle, prop_raw, us, state_old, dt, n_state: int):
    """One step of a state-variable model at a given old state (PcxStateStep, no VJP: the post-processor's and the
    op-level forward calls): (Pi, state_new (ne, nq, n_state), f (nn, ndof), P (ne, nq, 3, 3)) of
    physics.potential_energy / potential_energy_and_internal_force / element_pp(pk1_stress), base.py:24-74,349-361."""
    outs = (_f64(h.ne, h.nq, n_state), _f64(), _f64(h.nn, h.ndof), _f64(h.ne, h.nq, 3, 3))
    z_new, Pi, f, P = jax.ffi.ffi_call("PcxStateStep", outs)(prop_raw, us, jnp.reshape(state_old, (h.ne, h.nq, n_state)),
                                                            handle=np.int64(h.ptr), dt=np.float64(dt))
    return Pi, z_new, f, P


def residual_and_reaction(h: Handle, f):
    out = jax.ffi.ffi_call("PcxResidualReaction", _f64(2))(f, handle=np.int64(h.ptr))
    return out[0], out[1]


def loss_and_grad(h: Handle, kind: str, weights, prop_raw=None, global_outputs=None, energy_weight=1.0,
                  residual_weight=1.0, reaction_weight=1.0, first_step=0, want_grad=True):
    """S3: eqx.filter_value_and_grad(loss.filtered_loss, has_aux=True) for the built-in weak-form losses, fused.
    Returns (loss, aux[4 + nt], g_weights, g_props)."""
    p = _empty() if prop_raw is None else prop_raw
    g = _empty() if global_outputs is None else jnp.asarray(global_outputs, dtype=jnp.float64)
    outs = (_f64(), _f64(PCX_AUX_FIXED + h.nt), _f64(h.n_weights), _f64(h.n_train))
    return jax.ffi.ffi_call("PcxLossAndGrad", outs)(
        weights, p, g, handle=np.int64(h.ptr), kind=np.int32(LOSS[kind]), first_step=np.int32(first_step),
        want_grad=np.int32(1 if want_grad else 0), energy_weight=np.float64(energy_weight),
        residual_weight=np.float64(residual_weight), reaction_weight=np.float64(reaction_weight))


def sharded_loss_and_grad(h: Handle, kind: str, weights, prop_raw, global_outputs, shared_rows, n_shards, axis_name,
                          energy_weight=1.0, residual_weight=1.0, reaction_weight=1.0):
    """Residual / reaction losses on one element shard per device, inside a shard_map over `axis_name`: the phased ABI
    with jax.lax.psum at the two exchange points.  shared_rows: (local node ids, slots in the global shared-node list,
    length of that list) of this shard (pancax_b200.parallel.ShardPlan)."""
    loc, slot, n_shared = shared_rows
    p = _empty() if prop_raw is None else prop_raw
    g = _empty() if global_outputs is None else jnp.asarray(global_outputs, dtype=jnp.float64)
    kw = dict(handle=np.int64(h.ptr), kind=np.int32(LOSS[kind]), energy_weight=np.float64(energy_weight),
              residual_weight=np.float64(residual_weight), reaction_weight=np.float64(reaction_weight))
    f = jax.ffi.ffi_call("PcxLossBegin", _f64(h.nt, h.nn, h.ndof))(weights, p, g, **kw)
    buf = jnp.zeros((h.nt, n_shared, h.ndof)).at[:, slot].set(f[:, loc])
    f = f.at[:, loc].set(jax.lax.psum(buf, axis_name)[:, slot])
    rr = jax.lax.psum(jax.ffi.ffi_call("PcxLossMid", _f64(h.nt, 2))(f, handle=np.int64(h.ptr)), axis_name)
    outs = (_f64(), _f64(PCX_AUX_FIXED + h.nt), _f64(h.n_weights), _f64(h.n_train))
    loss, aux, gw, gp = jax.ffi.ffi_call("PcxLossFinish", outs)(g, f, rr, replicated_scale=np.float64(1.0 / n_shards), **kw)
    return tuple(jax.lax.psum(v, axis_name) for v in (loss, aux, gw, gp))


def field_data_loss_and_grad(h: Handle, weights, inputs, outputs, a=None, b=None, weight=1.0):
    """FullFieldDataLoss (data_loss_functions.py:13-24) on a batch: (loss, g_weights)."""
    return jax.ffi.ffi_call("PcxFieldDataLossAndGrad", (_f64(), _f64(h.n_weights)))(
        weights, inputs, _empty() if a is None else a, _empty() if b is None else b, outputs,
        handle=np.int64(h.ptr), weight=np.float64(weight))


def adam_step(params, grads, mu, nu, count, lr, b1=0.9, b2=0.999, eps=1e-8):
    """optax.chain(scale_by_adam, scale_by_schedule, scale(-1)) + apply_updates on one flat leaf, in place."""
    shp = _f64(*params.shape)
    return jax.ffi.ffi_call("PcxAdamStep", (shp, shp, shp), input_output_aliases={0: 0, 2: 1, 3: 2})(
        params, grads, mu, nu, count=np.int64(count), lr=np.float64(lr), b1=np.float64(b1), b2=np.float64(b2),
        eps=np.float64(eps))


# ------------------------------------------------------------------------------------------------ hooks S1-S4
def install(problem, params, deterministic: bool = True, ffi_so: str = "libpcx_xla_ffi.so") -> Handle:
    """Opt-in: after this call the pancax objects of `problem` evaluate on the B200 library.
    S1 BasePhysics.vmap_field_values   (physics_kernels/base.py:248-250)  -> field_values
    S2 BaseEnergyFormPhysics.potential_energy_and_internal_force (base.py:356-361) -> potential_energy_and_internal_force
    S3 AbstractOptimizer._single_step  (optimizers.py:77-107) -> loss_and_grad when the loss is a built-in weak-form loss
    S4 physics.var_name_to_method[...]["method"] of the element variables (solid_mechanics.py:115-170) -> element_fields /
       state_step (the post-processor's per-step calls, post_processor.py:305-316)
    The reference classes are patched per instance (object.__setattr__ on the eqx.Module), nothing global."""
    import equinox as eqx
    from pancax.loss_functions.weak_form_loss_functions import (EnergyAndResidualLoss, EnergyLoss,
                                                                EnergyResidualAndReactionLoss)
    register(ffi_so)
    field = params.fields
    h = make_handle(problem, field, deterministic)
    physics = problem.physics
    times = np.asarray(problem.domain.times)
    ref_vmap_field_values = type(physics).vmap_field_values
    ref_pe_force = getattr(type(physics), "potential_energy_and_internal_force", None)

    def t_index(t):
        return int(np.argmin(np.abs(times - float(t))))

    def vmap_field_values(fld, xs, t, *args):                                        # S1
        if xs is problem.domain.coords:
            return field_values(h, flat_weights(fld.networks), t_index(t))
        return ref_vmap_field_values(physics, fld, xs, t, *args)                     # other point sets: reference path

    def pe_and_force(prm, domain, t, us, state_old, dt, *args):                      # S2
        raw = jnp.concatenate([jnp.ravel(p.prop_val) for p in jax.tree_util.tree_leaves(
            prm.physics.constitutive_model, is_leaf=lambda x: hasattr(x, "prop_val")) if hasattr(p, "prop_val")]) \
            if h.n_train else _empty()
        n_state = int(getattr(physics.constitutive_model, "num_state_variables", 0))
        if n_state > 0:      # SimpleFeFv: forward evaluation at the given old state (the losses go through S3)
            Pi, state_new, f, _ = state_step(h, raw, us, state_old, dt, n_state)
            return (Pi, state_new), f
        Pi, f = potential_energy_and_internal_force(h, raw, us)
        return (Pi, state_old), f
    object.__setattr__(physics, "vmap_field_values", vmap_field_values)
    if ref_pe_force is not None:
        object.__setattr__(physics, "potential_energy_and_internal_force", pe_and_force)

    # S4: the element variables of the post-processor (post_processor.py:305-316 calls
    # physics.var_name_to_method[var]["method"](params, problem, time, us, theta, state_old, dt))
    def _raw(prm):
        return jnp.concatenate([jnp.ravel(p.prop_val) for p in jax.tree_util.tree_leaves(
            prm.physics.constitutive_model, is_leaf=lambda x: hasattr(x, "prop_val")) if hasattr(p, "prop_val")]) \
            if h.n_train else _empty()

    def _pp(which):
        def method(prm, prob, t, us, theta, state_old, dt, *args):
            n_state = int(getattr(physics.constitutive_model, "num_state_variables", 0))
            if which == "state_variables":
                return state_step(h, _raw(prm), us, state_old, dt, n_state)[1]
            if which == "pk1_stress" and n_state > 0:
                return state_step(h, _raw(prm), us, state_old, dt, n_state)[3]
            F, I1b, P = element_fields(h, _raw(prm), us)
            return {"deformation_gradient": F, "I1_bar": I1b, "pk1_stress": P}[which]
        return method
    table = getattr(physics, "var_name_to_method", None)
    if isinstance(table, dict) and hasattr(physics, "constitutive_model"):
        for var in ("deformation_gradient", "I1_bar", "pk1_stress", "state_variables"):
            if var in table:
                table[var]["method"] = _pp(var)

    kinds = {EnergyLoss: "energy", EnergyAndResidualLoss: "energy_residual",
             EnergyResidualAndReactionLoss: "energy_residual_reaction"}

    def fused_value_and_grad(loss_fn, prm, prob):                                    # S3
        kind = kinds.get(type(loss_fn))
        if kind is None:
            raise NotImplementedError(f"{type(loss_fn).__name__} is not a built-in weak-form loss: no B200 fast path")
        w = flat_weights(prm.fields.networks)
        gd = getattr(prob, "global_data", None)
        loss, aux, gw, gp = loss_and_grad(
            h, kind, w, None, None if gd is None else gd.outputs,
            energy_weight=getattr(loss_fn, "energy_weight", getattr(loss_fn, "weight", 1.0)),
            residual_weight=getattr(loss_fn, "residual_weight", 1.0), reaction_weight=getattr(loss_fn, "reaction_weight", 1.0),
            first_step=1 if getattr(loss_fn, "is_path_dependent", False) else 0)
        grads = eqx.tree_at(lambda q: q.fields.networks, prm, unflatten_like(prm.fields.networks, gw))
        keys = {"energy": ("energy",), "energy_residual": ("energy", "residual"),
                "energy_residual_reaction": ("energy", "residual", "global_data_loss")}[kind]
        auxd = {k: aux[i] for i, k in enumerate(keys)}
        if kind == "energy_residual_reaction":
            auxd["reactions"] = aux[PCX_AUX_FIXED:]
        return (loss, auxd), grads
    h.fused_value_and_grad = fused_value_and_grad       # used by the patched AbstractOptimizer._single_step
    return h
