"""Engine: one pcx_handle (one GPU) + the device buffers it borrows.

Thin host-side owner of the static inputs of the hot path (mesh arrays, tabulated Dirichlet
wrapper, tabulated source term) and of the output buffers; every compute method is one call
into the C ABI on the current torch CUDA stream.  torch is used for device memory and streams
only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (ELEM, FORM, LOSS, MODEL, PCX_AUX_FIXED, PCX_MAX_PROPS, PHYS_POISSON,
                   PHYS_SOLID, FieldDesc, LossDesc, MeshDesc, PhysicsDesc, check)

NNPE = {"hex8": 8, "quad4": 4, "tri3": 3}
NDIM = {"hex8": 3, "quad4": 2, "tri3": 2}

MODEL_PROPS = {
    "neohookean": ("bulk_modulus", "shear_modulus"),
    "gent": ("bulk_modulus", "shear_modulus", "Jm_parameter"),
    "swanson": ("bulk_modulus", "A1", "P1", "B1", "Q1", "C1", "R1", "cutoff_strain"),
    "hencky": ("bulk_modulus", "shear_modulus"),
    "blatz_ko": ("shear_modulus",),
}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mlp_num_weights(n_in, n_out, width, depth, use_final_bias=False):
    sizes = [n_in] + [width] * depth + [n_out]
    n = 0
    for i in range(len(sizes) - 1):
        n += sizes[i] * sizes[i + 1]
        if i < len(sizes) - 2 or use_final_bias:
            n += sizes[i + 1]
    return n


def tabulate_bc(bc_func, coords, times, ndof, check_affine=True):
    """Tabulate a Dirichlet wrapper g(x, t, z) that is diagonal-affine in z:
    a = g(x,t,0), b_i = g(x,t,e_i)_i - a_i for every (time, node).  ``bc_func`` takes
    (x (N,nd), t float, z (N,ndof)) numpy arrays.  Non-affine / non-diagonal wrappers are
    refused (there is no fallback path that could evaluate Python inside a kernel)."""
    coords = np.asarray(coords, dtype=np.float64)
    nn = coords.shape[0]
    nt = len(times)
    a = np.empty((nt, nn, ndof))
    b = np.empty((nt, nn, ndof))
    rng = np.random.default_rng(0)
    for it, t in enumerate(times):
        z0 = np.zeros((nn, ndof))
        a[it] = np.asarray(bc_func(coords, float(t), z0))
        ones = np.ones((nn, ndof))
        g1 = np.asarray(bc_func(coords, float(t), ones))
        b[it] = g1 - a[it]
        if check_affine:
            z = rng.standard_normal((nn, ndof))
            gz = np.asarray(bc_func(coords, float(t), z))
            err = np.abs(gz - (a[it] + b[it] * z)).max()
            scale = max(1.0, np.abs(gz).max())
            if err > 1e-9 * scale:
                raise NotImplementedError(
                    "dirichlet_bc_func is not diagonal-affine in the network output z; "
                    "pancax_b200 tabulates u = a(x,t) + b(x,t)*z and cannot run arbitrary Python "
                    f"inside a kernel (max deviation {err:.3e})")
    return a, b


class Engine:
    def __init__(self, coords, conns, elem, q_order, ndof, times, unknown_idx=None,
                 reaction_nodes=None, reaction_dof=0, deterministic=True, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pancax_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        conns = np.ascontiguousarray(conns, dtype=np.int32)
        self.elem, self.q_order, self.ndof = elem, int(q_order), int(ndof)
        self.nn, self.nd = coords.shape
        self.ne, self.nnpe = conns.shape
        if self.nd != NDIM[elem] or self.nnpe != NNPE[elem]:
            raise ValueError(f"mesh arrays do not match element type {elem}")
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.nt = len(self.times)
        dev = self.device
        with torch.cuda.device(dev):
            self.coords = torch.from_numpy(coords).to(dev)
            self.conns = torch.from_numpy(conns).to(dev)
            # an EMPTY index list is not a missing one (every dof constrained: R = 0; no reaction nodes: reaction = 0,
            # as jnp.linalg.norm / jnp.sum of an empty slice give): the library gets a valid pointer and a zero count
            def _dev(a, dt):
                a = np.ascontiguousarray(a, dtype=dt)
                return torch.from_numpy(a if a.size else np.zeros(1, dtype=dt)).to(dev)
            self.n_unknown = 0 if unknown_idx is None else int(np.size(unknown_idx))
            self.n_reaction = 0 if reaction_nodes is None else int(np.size(reaction_nodes))
            self.unknown_idx = None if unknown_idx is None else _dev(unknown_idx, np.int64)
            self.reaction_nodes = None if reaction_nodes is None else _dev(reaction_nodes, np.int32)
            md = MeshDesc()
            md.elem_type, md.q_order, md.ndof = ELEM[elem], self.q_order, self.ndof
            md.reaction_dof = int(reaction_dof)
            md.ne, md.nn = self.ne, self.nn
            md.n_unknown, md.n_reaction = self.n_unknown, self.n_reaction
            md.coords, md.conns = self.coords.data_ptr(), self.conns.data_ptr()
            md.unknown_idx = None if self.unknown_idx is None else self.unknown_idx.data_ptr()
            md.reaction_nodes = None if self.reaction_nodes is None else self.reaction_nodes.data_ptr()
            md.nt, md.deterministic = self.nt, 1 if deterministic else 0
            md.times = self.times.ctypes.data_as(C.POINTER(C.c_double))
            h = C.c_void_p()
            check(self.lib.pcx_create(C.byref(h), C.byref(md)))
        self.h = h
        self.nq = self.lib.pcx_nq(self.h)
        self.n_weights = None
        self.n_train = 0
        self._keep = []

    # ------------------------------------------------------------------ setup
    def set_physics(self, kind, model=None, formulation=None, props=None, source_q=None, prony=None):
        """kind: "poisson" | "solid".  props: list of floats or (prop_min, prop_max) tuples
        (trainable BoundedProperty) in the model's property order.  prony = (moduli, relaxation_times): the model
        is the equilibrium branch of a SimpleFeFv viscoelastic model with these Prony branches (state variables:
        path-dependent EnergyLoss only)."""
        pd = PhysicsDesc()
        if kind == "poisson":
            pd.physics, pd.model, pd.formulation, pd.nprops = PHYS_POISSON, 0, 0, 0
            if source_q is not None:
                sq = torch.as_tensor(np.ascontiguousarray(source_q, dtype=np.float64)).to(self.device)
                if sq.numel() != self.ne * self.nq:
                    raise ValueError("source_q must have ne*nq entries")
                self._keep.append(sq)
                pd.source_q = sq.data_ptr()
        elif kind == "solid":
            pd.physics, pd.model, pd.formulation = PHYS_SOLID, MODEL[model], FORM[formulation]
            pd.nprops = len(props)
            if len(props) != len(MODEL_PROPS[model]):
                raise ValueError(f"{model} takes properties {MODEL_PROPS[model]}")
            for i, p in enumerate(props):
                if isinstance(p, (tuple, list)):
                    pd.bounded[i] = 1
                    pd.prop_min[i], pd.prop_max[i] = float(p[0]), float(p[1])
                else:
                    pd.value[i] = float(p)
            if prony is not None:
                moduli, taus = [float(v) for v in prony[0]], [float(v) for v in prony[1]]
                if len(moduli) != len(taus):
                    raise ValueError("PronySeries: moduli and relaxation_times differ in length")
                pd.n_prony = len(moduli)
                for b, (g, t) in enumerate(zip(moduli, taus)):
                    pd.prony_moduli[b], pd.prony_times[b] = g, t
        else:
            raise ValueError(kind)
        with torch.cuda.device(self.device):
            check(self.lib.pcx_set_physics(self.h, C.byref(pd)))
        self.n_train = self.lib.pcx_num_trainable_props(self.h)
        self.kind = kind

    def set_field(self, n_in, n_out, width, depth, use_final_bias=False, bc_a=None, bc_b=None,
                  x_min=None, x_max=None, t_min=None, t_max=None, normalize_time=None):
        fd = FieldDesc()
        fd.n_in, fd.n_out, fd.width, fd.depth = n_in, n_out, width, depth
        fd.use_final_bias = 1 if use_final_bias else 0
        cmin = self.coords.min(dim=0).values.cpu().numpy() if x_min is None else np.asarray(x_min)
        cmax = self.coords.max(dim=0).values.cpu().numpy() if x_max is None else np.asarray(x_max)
        for j in range(self.nd):
            fd.x_min[j], fd.x_max[j] = float(cmin[j]), float(cmax[j])
        fd.t_min = float(self.times.min()) if t_min is None else float(t_min)
        fd.t_max = float(self.times.max()) if t_max is None else float(t_max)
        if normalize_time is None:   # fields.py:73-76
            normalize_time = not np.isclose(fd.t_min, fd.t_max, rtol=1e-5, atol=1e-8)
        fd.normalize_time = 1 if normalize_time else 0
        self._bc_b_host_zero = None if bc_b is None else \
            ~np.any(np.asarray(bc_b).reshape(self.nt, -1) != 0.0, axis=1)
        for name, tab in (("bc_a", bc_a), ("bc_b", bc_b)):
            if tab is not None:
                t = torch.as_tensor(np.ascontiguousarray(tab, dtype=np.float64)).to(self.device)
                if t.numel() != self.nt * self.nn * self.ndof:
                    raise ValueError(f"{name} must have shape (nt, nn, ndof)")
                self._keep.append(t)
                setattr(fd, name, t.data_ptr())
        with torch.cuda.device(self.device):
            check(self.lib.pcx_set_field(self.h, C.byref(fd)))
        self.n_weights = self.lib.pcx_num_weights(self.h)
        self.field_desc = fd

    # ------------------------------------------------------------------ helpers
    def _dev(self, x, dtype=torch.float64):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    def _empty(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    @property
    def launch_count(self):
        return self.lib.pcx_launch_count(self.h)

    # ------------------------------------------------------------------ ops
    def element_geometry(self, want_grad_N=True, want_JxW=True, want_xq=True):
        gN = self._empty(self.ne, self.nq, self.nnpe, self.nd) if want_grad_N else None
        JxW = self._empty(self.ne, self.nq) if want_JxW else None
        xq = self._empty(self.ne, self.nq, self.nd) if want_xq else None
        check(self.lib.pcx_element_geometry(self.h, _ptr(gN), _ptr(JxW), _ptr(xq), _stream()))
        return gN, JxW, xq

    def field_forward(self, weights, t_idx, out=None):
        w = self._dev(weights)
        us = self._empty(self.nn, self.ndof) if out is None else out
        check(self.lib.pcx_field_forward(self.h, _ptr(w), int(t_idx), _ptr(us), _stream()))
        return us

    def field_backward(self, weights, t_idx, g_us, out=None):
        w, g = self._dev(weights), self._dev(g_us)
        gw = self._empty(self.n_weights) if out is None else out
        check(self.lib.pcx_field_backward(self.h, _ptr(w), int(t_idx), _ptr(g), _ptr(gw), _stream()))
        return gw

    def points_forward(self, weights, pts, a=None, b=None):
        w, p, a, b = self._dev(weights), self._dev(pts), self._dev(a), self._dev(b)
        n = p.shape[0]
        u = self._empty(n, self.ndof)
        check(self.lib.pcx_points_forward(self.h, _ptr(w), _ptr(p), _ptr(a), _ptr(b), n, _ptr(u), _stream()))
        return u

    def points_backward(self, weights, pts, g_u, b=None):
        w, p, g, b = self._dev(weights), self._dev(pts), self._dev(g_u), self._dev(b)
        n = p.shape[0]
        gw = self._empty(self.n_weights)
        check(self.lib.pcx_points_backward(self.h, _ptr(w), _ptr(p), _ptr(b), _ptr(g), n, _ptr(gw), _stream()))
        return gw

    def energy_force(self, us, prop_raw=None, want_force=True, want_gprops=False):
        u, pr = self._dev(us), self._dev(prop_raw)
        Pi = self._empty(1)
        f = self._empty(self.nn, self.ndof) if want_force else None
        gp = self._empty(max(self.n_train, 1)) if want_gprops else None
        check(self.lib.pcx_energy_force(self.h, _ptr(pr), _ptr(u), _ptr(Pi), _ptr(f), _ptr(gp), _stream()))
        return Pi, f, (gp[: self.n_train] if gp is not None else None)

    def stiffness_action(self, us, v, prop_raw=None, want_gprops=False):
        u, vv, pr = self._dev(us), self._dev(v), self._dev(prop_raw)
        Kv = self._empty(self.nn, self.ndof)
        gp = self._empty(max(self.n_train, 1)) if want_gprops else None
        check(self.lib.pcx_stiffness_action(self.h, _ptr(pr), _ptr(u), _ptr(vv), _ptr(Kv), _ptr(gp), _stream()))
        return Kv, (gp[: self.n_train] if gp is not None else None)

    def element_fields(self, us, prop_raw=None, want_F=True, want_I1bar=True, want_P=True):
        """Per-quadrature-point deformation gradient (ne, nq, 3, 3), I1_bar (ne, nq) and PK1 stress
        (ne, nq, 3, 3): the element variables pancax's post-processor writes."""
        u, pr = self._dev(us), self._dev(prop_raw)
        F = self._empty(self.ne, self.nq, 3, 3) if want_F else None
        I1 = self._empty(self.ne, self.nq) if want_I1bar else None
        P = self._empty(self.ne, self.nq, 3, 3) if want_P else None
        check(self.lib.pcx_element_fields(self.h, _ptr(pr), _ptr(u), _ptr(F), _ptr(I1), _ptr(P), _stream()))
        return F, I1, P

    # ---- strong-form / collocation path
    def reserve_points(self, n):
        """Set-up: workspace of the strong-form path for batches of up to ``n`` points (pcx_reserve; the compute
        calls never allocate).  Grows only; a no-op once enough is reserved."""
        if n > getattr(self, "_points_reserved", 0):
            with torch.cuda.device(self.device):
                check(self.lib.pcx_reserve(self.h, 1, int(n)))
            self._points_reserved = int(n)

    def points_jet(self, weights, pts, tab=None, want=("u", "grad_u", "lap_u", "u_t")):
        """(u (n,ndof), grad_u (n,ndof,nd), lap_u (n,ndof), u_t (n,ndof)) of the Field at points
        pts (n, nd+1) = (x..., t); ``tab`` (n, ndof, 2*(nd+3)) = Dirichlet wrapper tables or None."""
        w, p, tb = self._dev(weights), self._dev(pts), self._dev(tab)
        n = p.shape[0]
        self.reserve_points(n)
        u = self._empty(n, self.ndof) if "u" in want else None
        gu = self._empty(n, self.ndof, self.nd) if "grad_u" in want else None
        lu = self._empty(n, self.ndof) if "lap_u" in want else None
        ut = self._empty(n, self.ndof) if "u_t" in want else None
        check(self.lib.pcx_points_jet(self.h, _ptr(w), _ptr(p), _ptr(tb), n, _ptr(u), _ptr(gu), _ptr(lu), _ptr(ut),
                                      _stream()))
        return u, gu, lu, ut

    def strong_form_loss_and_grad(self, weights, pts, c_t, c_lap, weight=1.0, tab=None, src=None, target=None,
                                  want_grad=True, want_resid=False):
        """weight * mean((c_t u_t - c_lap lap u - src - target)^2) over the batch and its weight gradient."""
        w, p, tb, sr, tg = (self._dev(weights), self._dev(pts), self._dev(tab), self._dev(src), self._dev(target))
        n = p.shape[0]
        self.reserve_points(n)
        loss = self._empty(1)
        res = self._empty(n, self.ndof) if want_resid else None
        gw = self._empty(self.n_weights) if want_grad else None
        check(self.lib.pcx_strong_form_loss_and_grad(self.h, _ptr(w), _ptr(p), _ptr(tb), _ptr(sr), _ptr(tg), n,
                                                     float(c_t), float(c_lap), float(weight), _ptr(loss), _ptr(res),
                                                     _ptr(gw), _stream()))
        return loss, res, gw

    def residual_reaction(self, f):
        f = self._dev(f)
        out = self._empty(2)
        check(self.lib.pcx_residual_reaction(self.h, _ptr(f), _ptr(out), _stream()))
        return out

    def loss_and_grad(self, kind, weights, prop_raw=None, energy_weight=1.0, residual_weight=1.0,
                      reaction_weight=1.0, global_outputs=None, out=None, want_grad=True, first_step=0):
        """Returns device tensors (loss[1], aux[4+nt], g_weights, g_props).  ``first_step=1``: the
        path-dependent call of a model without state variables (the time loop starts at step 1)."""
        w, pr = self._dev(weights), self._dev(prop_raw)
        ld = LossDesc()
        ld.kind = LOSS[kind]
        ld.first_step = int(first_step)
        ld.energy_weight, ld.residual_weight, ld.reaction_weight = \
            float(energy_weight), float(residual_weight), float(reaction_weight)
        if global_outputs is not None:
            ld.global_outputs = self._dev_outputs(global_outputs).data_ptr()
        if out is None:
            loss = self._empty(1)
            aux = self._empty(PCX_AUX_FIXED + self.nt)
            gw = self._empty(self.n_weights) if want_grad else None
            gp = self._empty(max(self.n_train, 1))
        else:
            loss, aux, gw, gp = out
        check(self.lib.pcx_loss_and_grad(self.h, C.byref(ld), _ptr(w), _ptr(pr), _ptr(loss), _ptr(aux),
                                         _ptr(gw), _ptr(gp) if self.n_train > 0 else None, _stream()))
        return loss, aux, gw, gp[: self.n_train]

    def null_steps(self):
        """Which time steps have an identically-zero Dirichlet table b (u = a there)."""
        if self._bc_b_host_zero is None:
            return np.zeros(self.nt, dtype=bool)
        return self._bc_b_host_zero

    def _dev_outputs(self, global_outputs):
        """Device copy of problem.global_data.outputs (ABI v4: pcx_loss_desc.global_outputs is a device pointer the
        kernels read directly).  A device tensor is used as is; host data is uploaded once into a buffer that lives
        as long as the engine and re-uploaded only when its values change, so a captured CUDA graph keeps reading
        the same address."""
        if isinstance(global_outputs, torch.Tensor) and global_outputs.is_cuda:
            go = global_outputs.to(dtype=torch.float64).contiguous()
            if go.numel() != self.nt:
                raise ValueError("global_outputs must have nt entries")
            self._go_user = go          # keep it alive while work that reads it may be in flight
            return go
        go = np.ascontiguousarray(global_outputs.cpu().numpy() if isinstance(global_outputs, torch.Tensor)
                                  else global_outputs, dtype=np.float64).reshape(-1)
        if go.shape[0] != self.nt:
            raise ValueError("global_outputs must have nt entries")
        cur = getattr(self, "_go_host", None)
        if cur is None or not np.array_equal(cur, go):
            self._go_host = go.copy()
            if getattr(self, "_go_dev", None) is None:
                self._go_dev = torch.from_numpy(self._go_host).to(self.device)
            else:
                self._go_dev.copy_(torch.from_numpy(self._go_host))
        return self._go_dev

    def get_state(self, t_idx):
        """State variables (ne, nq, n_prony*9) of time step ``t_idx`` after a path-dependent call (SimpleFeFv)."""
        ns = int(self.lib.pcx_num_state_variables(self.h))
        if ns <= 0:
            raise ValueError("the constitutive model has no state variables")
        out = self._empty(self.ne, self.nq, ns)
        check(self.lib.pcx_get_state(self.h, int(t_idx), _ptr(out), _stream()))
        return out

    def initial_state(self):
        """``constitutive_model.initial_state()`` at every quadrature point (simple_fefv.py:72-75): Fv = I per Prony
        branch, (ne, nq, n_prony*9)."""
        ns = int(self.lib.pcx_num_state_variables(self.h))
        if ns <= 0:
            raise ValueError("the constitutive model has no state variables")
        z = self._empty(self.ne, self.nq, ns // 9, 9)
        z.zero_()
        z[..., 0::4] = 1.0
        return z.reshape(self.ne, self.nq, ns)

    def state_step(self, us, state_old, dt, prop_raw=None, want_force=True, want_P=True):
        """One time step of a state-variable model at a given old state (pcx_state_step): (Pi, state_new, f, P) =
        ``physics.potential_energy(params, domain, t, us, state_old, dt)`` (base.py:349-354), the internal force at
        fixed old state (base.py:356-361) and ``element_pp(pk1_stress)`` (ne, nq, 3, 3)."""
        u, pr, zo = self._dev(us), self._dev(prop_raw), self._dev(state_old)
        ns = int(self.lib.pcx_num_state_variables(self.h))
        if ns <= 0:
            raise ValueError("the constitutive model has no state variables (use energy_force / element_fields)")
        if zo.numel() != self.ne * self.nq * ns:
            raise ValueError(f"state_old has {zo.numel()} entries, expected (ne, nq, {ns})")
        zn = self._empty(self.ne, self.nq, ns)
        Pi = self._empty(1)
        f = self._empty(self.nn, self.ndof) if want_force else None
        P = self._empty(self.ne, self.nq, 3, 3) if want_P else None
        check(self.lib.pcx_state_step(self.h, _ptr(pr), _ptr(u), _ptr(zo), float(dt), _ptr(zn), _ptr(Pi), _ptr(f), _ptr(P),
                                      _stream()))
        return Pi, zn, f, P

    def set_option(self, name, value):
        """``cull_null_steps``: skip the MLP on time steps whose Dirichlet table b is identically zero (exact).
        ``tf32_adjoint``: adjoint through the MLP on the TF32 tensor cores (3xTF32, fp32 activations; 1 = mma.sync
        kernel, 2 = contractions on tcgen05.mma kind::tf32 / TMEM); loss and aux
        unchanged, weight gradients within 1e-5 relative per leaf (PCX_OPT_TF32_ADJOINT)."""
        codes = {"cull_null_steps": 1, "tf32_adjoint": 2, "tune_fwd_tanh": 100, "tune_fwd_edge": 101, "tune_bwd_edge": 102,
                 "tune_fwd_variant": 103, "tune_stamps": 104, "tune_fwd_dynamic": 105, "tune_l2pf": 107}
        check(self.lib.pcx_set_option(self.h, codes[name], int(value)))

    # ---- sharded residual / reaction losses: pcx_loss_and_grad cut at the two exchange points
    def set_ownership(self, n_unknown_owned, n_reaction_owned):
        check(self.lib.pcx_set_ownership(self.h, int(n_unknown_owned), int(n_reaction_owned)))

    def _loss_desc(self, kind, energy_weight, residual_weight, reaction_weight, global_outputs):
        ld = LossDesc()
        ld.kind = LOSS[kind]
        ld.energy_weight, ld.residual_weight, ld.reaction_weight = \
            float(energy_weight), float(residual_weight), float(reaction_weight)
        go = None
        if global_outputs is not None:
            go = self._dev_outputs(global_outputs)
            ld.global_outputs = go.data_ptr()
        return ld, go

    def loss_begin(self, kind, weights, prop_raw=None, energy_weight=1.0, residual_weight=1.0,
                   reaction_weight=1.0, global_outputs=None):
        """Phase 1: field forward + force sweep.  Returns this shard's partial f [nt, nn, ndof]."""
        w, pr = self._dev(weights), self._dev(prop_raw)
        self._ld, self._go = self._loss_desc(kind, energy_weight, residual_weight, reaction_weight, global_outputs)
        if not hasattr(self, "_f_full"):
            self._f_full = self._empty(self.nt, self.nn, self.ndof)
            self._rr = self._empty(self.nt, 2)
        check(self.lib.pcx_loss_begin(self.h, C.byref(self._ld), _ptr(w), _ptr(pr), _ptr(self._f_full), _stream()))
        return self._f_full

    def loss_mid(self):
        """Phase 2 (after the shared rows of f were summed over shards, in place): partial
        (sum f^2 over owned unknown dofs, reaction over owned nodes) [nt, 2]."""
        check(self.lib.pcx_loss_mid(self.h, _ptr(self._f_full), _ptr(self._rr), _stream()))
        return self._rr

    def loss_finish(self, n_shards, out=None):
        """Phase 3 (after rr was summed over shards, in place): PARTIAL (loss, aux, g_weights, g_props)."""
        if out is None:
            loss = self._empty(1)
            aux = self._empty(PCX_AUX_FIXED + self.nt)
            gw = self._empty(self.n_weights)
            gp = self._empty(max(self.n_train, 1))
        else:
            loss, aux, gw, gp = out
        check(self.lib.pcx_loss_finish(self.h, C.byref(self._ld), _ptr(self._f_full), _ptr(self._rr),
                                       1.0 / n_shards, _ptr(loss), _ptr(aux), _ptr(gw),
                                       _ptr(gp) if self.n_train > 0 else None, _stream()))
        return loss, aux, gw, gp[: self.n_train]

    def sharded_loss_and_grad(self, kind, weights, prop_raw, exchange, out=None, **kw):
        """One shard's part of a residual / reaction loss.  ``exchange.sum_shared(f)`` adds, in place,
        the other shards' rows of the shared nodes; ``exchange.sum_scalars(rr)`` sums rr over shards in
        place; ``exchange.n_shards`` is the shard count.  The sum over shards of the returned
        (loss, aux, g_weights, g_props) is the whole-mesh result of ``loss_and_grad``."""
        exchange.sum_shared(self.loss_begin(kind, weights, prop_raw, **kw))
        exchange.sum_scalars(self.loss_mid())
        return self.loss_finish(exchange.n_shards, out=out)

    def field_data_loss_and_grad(self, weights, pts, outputs, a=None, b=None, weight=1.0):
        w, p, y, a, b = (self._dev(weights), self._dev(pts), self._dev(outputs), self._dev(a),
                         self._dev(b))
        n = p.shape[0]
        loss = self._empty(1)
        gw = self._empty(self.n_weights)
        check(self.lib.pcx_field_data_loss_and_grad(self.h, _ptr(w), _ptr(p), _ptr(a), _ptr(b), _ptr(y),
                                                    n, float(weight), _ptr(loss), _ptr(gw), _stream()))
        return loss, gw

    def profile_enable(self, on=True):
        check(self.lib.pcx_profile_enable(self.h, 1 if on else 0))

    def profile_collect(self):
        """{class: (total ms, launches)} since the last collect (synchronises)."""
        ms = (C.c_double * 8)()
        cnt = (C.c_int64 * 8)()
        check(self.lib.pcx_profile_collect(self.h, ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PROF_CLASSES)}

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
