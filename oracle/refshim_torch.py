"""Test infrastructure: a TORCH-backed stand-in for the ``jax`` / ``equinox`` / ``jaxtyping``
namespaces with automatic differentiation, so that the reference's own loss functions
(pancax/loss_functions/*.py), physics kernels (pancax/physics_kernels/base.py:248-385), Field
(pancax/networks/fields.py) and constitutive models are EXECUTED here -- including every
``jax.grad`` / ``value_and_grad`` on the path, mapped onto ``torch.autograd.grad(create_graph=True)``
-- to produce end-to-end loss / aux / gradient golden vectors (oracle/gen_golden_ad.py).

jax itself cannot be installed in this image.  What is NOT the reference's own code here:
  * ``equinox.nn.MLP`` / ``Linear`` (equinox ~=0.13, un-vendored): restated from its published
    algorithm -- ``Linear``: ``weight (out,in) @ x + bias``; ``MLP``: ``depth`` hidden Linears with
    ``activation`` then a final Linear (bias iff ``use_final_bias``) and ``final_activation``;
  * ``vmap`` is a Python loop + ``torch.stack``; ``lax.cond`` is an ``if`` on the concrete predicate
    (same value and same selected-branch derivative as jax);
  * ``custom_jvp`` functions are differentiated through the reference's REGISTERED jvp rule (VJP
    assembled from the rule applied to basis tangents, out of differentiable ops, so a second
    differentiation goes through the rule's own code like jax's does).

Only oracle/gen_golden_ad.py imports this module; nothing at test or run time does.
"""
from __future__ import annotations

import copy
import sys
import types

import numpy as _np
import torch

F64 = torch.float64


# --------------------------------------------------------------------------- tensor helpers
def _t(x):
    """Anything -> torch tensor (float64 for floats, int64 for ints), keeping autograd links."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)):
        if any(isinstance(v, (torch.Tensor, list, tuple)) for v in x):
            return torch.stack([_t(v).to(F64) if not isinstance(v, torch.Tensor) else v for v in x])
        a = _np.asarray(x)
    else:
        a = _np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a, dtype=F64)
    if a.dtype.kind in "iu":
        return torch.as_tensor(a, dtype=torch.int64)
    if a.dtype.kind == "b":
        return torch.as_tensor(a, dtype=torch.bool)
    return torch.as_tensor(a)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        a = self.a

        class _Op:
            def set(self, v):
                b = a.clone()
                b[idx] = _t(v).to(b.dtype) if not isinstance(v, (int, float)) else v
                return b

            def add(self, v):
                b = a.clone()
                b[idx] = b[idx] + v
                return b
        return _Op()


def _patch_tensor():
    if getattr(torch.Tensor, "_refshim_patched", False):
        return
    torch.Tensor.at = property(lambda self: _At(self))
    torch.Tensor.astype = lambda self, dt: self.to(dt)
    torch.Tensor._refshim_patched = True


def _axis_kw(k):
    if "axis" in k:
        k["dim"] = k.pop("axis")
    k.pop("dtype", None) if k.get("dtype", 1) is None else None
    return k


def _reduce(name):
    fn = getattr(torch, name)

    def f(x, axis=None, **k):
        x = _t(x)
        if axis is None:
            return fn(x)
        out = fn(x, dim=axis)
        return out.values if hasattr(out, "values") and not isinstance(out, torch.Tensor) else out
    return f


def _dot(a, b):
    a, b = _t(a), _t(b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    if a.dtype != b.dtype:
        a, b = a.to(F64), b.to(F64)
    return torch.matmul(a, b)


def _hstack(tup):
    return torch.cat([torch.atleast_1d(_t(v)).to(F64) for v in tup])


def _norm(x, ord=None, axis=None):
    x = _t(x)
    if ord is None:
        return torch.sqrt(torch.sum(x * x)) if axis is None else torch.sqrt(torch.sum(x * x, dim=axis))
    if ord == _np.inf and x.ndim == 2:
        return torch.max(torch.sum(torch.abs(x), dim=1))
    return torch.linalg.norm(x, ord=ord, dim=axis)


def _make_jnp():
    jnp = types.ModuleType("jax.numpy")
    unary = ["log", "exp", "sqrt", "sign", "square", "abs", "sin", "cos", "tanh", "arccos", "arcsin",
             "trace", "isnan", "isfinite", "ravel", "diag", "argsort", "sort", "cumsum", "floor", "log1p",
             "arctan", "sinh", "cosh", "fabs"]
    for nm in unary:
        if not hasattr(torch, nm):
            continue
        setattr(jnp, nm, (lambda f: (lambda x, *a, **k: f(_t(x), *a, **_axis_kw(k))))(getattr(torch, nm)))
    jnp.trace = lambda x, offset=0, axis1=0, axis2=1: torch.diagonal(_t(x), offset=offset, dim1=axis1, dim2=axis2).sum(-1)
    jnp.array = lambda x, dtype=None: _t(x)
    jnp.asarray = jnp.array
    jnp.zeros = lambda shape, dtype=None: torch.zeros(shape, dtype=F64)
    jnp.ones = lambda shape, dtype=None: torch.ones(shape, dtype=F64)
    jnp.zeros_like = lambda x: torch.zeros_like(_t(x))
    jnp.ones_like = lambda x: torch.ones_like(_t(x))
    jnp.eye = lambda n, dtype=None: torch.eye(n, dtype=F64)
    jnp.identity = jnp.eye
    jnp.arange = lambda *a, **k: torch.arange(*a)
    jnp.linspace = lambda a, b, n: torch.linspace(a, b, n, dtype=F64)
    for nm in ("sum", "mean", "max", "min", "prod", "amax", "amin"):
        setattr(jnp, nm, _reduce({"amax": "max", "amin": "min"}.get(nm, nm)))
    jnp.dot = _dot
    jnp.matmul = _dot
    jnp.outer = lambda a, b: torch.outer(_t(a), _t(b))
    jnp.power = lambda a, b: torch.pow(_t(a), b)
    jnp.clip = lambda x, a_min=None, a_max=None: torch.clamp(_t(x), a_min, a_max)
    jnp.tensordot = lambda a, b, axes=2: torch.tensordot(_t(a), _t(b), dims=axes)
    jnp.einsum = lambda s, *ops: torch.einsum(s, *[_t(o) for o in ops])
    jnp.hstack = _hstack
    jnp.vstack = lambda tup: torch.vstack([_t(v) for v in tup])
    jnp.stack = lambda tup, axis=0: torch.stack([_t(v) for v in tup], dim=axis)
    jnp.column_stack = lambda tup: torch.column_stack([_t(v) for v in tup])
    jnp.concatenate = lambda tup, axis=0: torch.cat([_t(v) for v in tup], dim=axis)
    def _where(c, a, b):
        a, b = _t(a), _t(b)
        if a.dtype == torch.bool and b.dtype == torch.bool:      # boolean selects stay boolean (rtsafe_'s `converged`)
            return torch.where(_t(c), a, b)
        return torch.where(_t(c), a.to(F64), b.to(F64))
    jnp.where = _where
    jnp.maximum = lambda a, b: torch.maximum(_t(a).to(F64), _t(b).to(F64))
    jnp.minimum = lambda a, b: torch.minimum(_t(a).to(F64), _t(b).to(F64))
    jnp.allclose = lambda a, b, **k: bool(torch.allclose(_t(a).to(F64), _t(b).to(F64)))
    jnp.isclose = lambda a, b, **k: torch.isclose(_t(a).to(F64), _t(b).to(F64))
    jnp.reshape = lambda x, s: _t(x).reshape(s)
    jnp.transpose = lambda x, *a: _t(x).permute(*a) if a else _t(x).T
    jnp.pi, jnp.inf, jnp.nan, jnp.e = _np.pi, _np.inf, _np.nan, _np.e
    jnp.float64, jnp.int32, jnp.int64, jnp.bool = F64, torch.int32, torch.int64, torch.bool
    jnp.ndarray = torch.Tensor
    lin = types.ModuleType("jax.numpy.linalg")
    lin.norm = _norm
    lin.det = lambda x: torch.linalg.det(_t(x))
    lin.inv = lambda x: torch.linalg.inv(_t(x))
    lin.solve = lambda a, b: torch.linalg.solve(_t(a), _t(b))
    lin.eigh = lambda x: tuple(torch.linalg.eigh(_t(x)))
    jnp.linalg = lin

    # anything else (table construction at init time: polynomial roots, Vandermonde solves, ...): run
    # numpy on detached values and hand tensors back.  Not differentiable -- and never on the AD path.
    def _np_fallback(target):
        def getattr_(item):
            obj = getattr(target, item)
            if not callable(obj) or isinstance(obj, type):
                return obj

            def to_np(v):
                if isinstance(v, torch.Tensor):
                    return v.detach().numpy()
                if isinstance(v, (list, tuple)):
                    return type(v)(to_np(u) for u in v)
                return v

            def to_t(v):
                if isinstance(v, _np.ndarray) and v.dtype.kind in "fiub":
                    return _t(v)
                if isinstance(v, tuple):
                    return tuple(to_t(u) for u in v)
                return v

            def f(*a, **k):
                return to_t(obj(*[to_np(v) for v in a], **{kk: to_np(v) for kk, v in k.items()}))
            return f
        return getattr_
    jnp.__getattr__ = _np_fallback(_np)
    lin.__getattr__ = _np_fallback(_np.linalg)
    return jnp, lin


# --------------------------------------------------------------------------- vmap / control flow
def _stack_tree(outs):
    first = outs[0]
    if first is None:
        return None
    if isinstance(first, tuple) and hasattr(first, "_fields"):
        return type(first)(*[_stack_tree([o[i] for o in outs]) for i in range(len(first))])
    if isinstance(first, (tuple, list)):
        return type(first)(_stack_tree([o[i] for o in outs]) for i in range(len(first)))
    if isinstance(first, dict):
        return {k: _stack_tree([o[k] for o in outs]) for k in first}
    return torch.stack([_t(o) if not isinstance(o, torch.Tensor) else o for o in outs])


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _t(a).shape[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else _t(a).select(ax, i) for a, ax in zip(args, axes)]
            outs.append(fun(*sl))
        return _stack_tree(outs)
    return mapped


def _fresh(x):
    """A new autograd node standing for `x` (so d/dx is taken through the function only)."""
    x = _t(x)
    if x.requires_grad:
        return x * 1.0
    return x.clone().requires_grad_(True)


def value_and_grad(fun, argnums=0, has_aux=False):
    def vg(*args):
        args = list(args)
        x = _fresh(args[argnums])
        args[argnums] = x
        out = fun(*args)
        val, aux = out if has_aux else (out, None)
        (g,) = torch.autograd.grad(val, x, create_graph=True, allow_unused=True)
        if g is None:
            g = torch.zeros_like(x)
        return ((val, aux), g) if has_aux else (val, g)
    return vg


def grad(fun, argnums=0, has_aux=False):
    vg = value_and_grad(fun, argnums, has_aux)

    def g(*args):
        out, gr = vg(*args)
        return (gr, out[1]) if has_aux else gr
    return g


def jacfwd(fun, argnums=0, has_aux=False):
    """Jacobian w.r.t. one argument for every output of ``fun`` (tuple outputs give a tuple of
    Jacobians, as jax does); built from reverse-mode rows -- same values as forward mode."""
    def jf(*args):
        args = list(args)
        x = _fresh(args[argnums])
        args[argnums] = x
        out = fun(*args)

        def jac(o):
            o = _t(o)
            if o.numel() == 0 or not o.requires_grad:
                return torch.zeros(tuple(o.shape) + tuple(x.shape), dtype=F64)
            rows = []
            for k in range(o.numel()):
                (g,) = torch.autograd.grad(o.reshape(-1)[k], x, create_graph=True, allow_unused=True)
                rows.append(torch.zeros_like(x) if g is None else g)
            return torch.stack(rows).reshape(tuple(o.shape) + tuple(x.shape))
        if isinstance(out, tuple):
            return tuple(jac(o) for o in out)
        return jac(out)
    return jf


class custom_jvp:
    """Differentiates through the REGISTERED rule: VJP(ct) = sum_k <rule(e_k), ct> e_k."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.jvp_rule, self.nondiff_argnums = fun, None, nondiff_argnums
        self.__name__ = getattr(fun, "__name__", "custom_jvp")

    def defjvp(self, rule):
        self.jvp_rule = rule
        return rule

    def __call__(self, A):
        fun, rule = self.fun, self.jvp_rule
        if not (isinstance(A, torch.Tensor) and A.requires_grad):
            return fun(A)

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, a):
                ctx.save_for_backward(a)
                with torch.no_grad():
                    return fun(a.detach())

            @staticmethod
            def backward(ctx, ct):
                # built from differentiable ops on the saved primal: under create_graph=True (the outer jax.grad of a
                # loss that already contains jax.grad, base.py:356-385) autograd differentiates the registered rule
                # itself a second time, with respect to the primal AND the cotangent -- what jax does with a custom_jvp
                (a,) = ctx.saved_tensors
                gs = []
                for k in range(a.numel()):
                    e = torch.zeros(a.numel(), dtype=a.dtype)
                    e[k] = 1.0
                    _, tang = rule((a,), (e.reshape(a.shape),))
                    gs.append(torch.sum(tang * ct))
                return torch.stack(gs).reshape(a.shape)
        return Fn.apply(A)


# --------------------------------------------------------------------------- equinox
class _Field:
    def __init__(self, default=None, static=False, **kw):
        self.default, self.static = default, static


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "__init__" in cls.__dict__:
            return
        for klass in cls.__mro__[1:]:
            if klass is Module:
                break
            if "__init__" in klass.__dict__ and not getattr(klass.__dict__["__init__"], "_shim_auto", False):
                return
        names = []
        for klass in reversed(cls.__mro__):
            for n in getattr(klass, "__annotations__", {}):
                if n not in names:
                    names.append(n)
        defaults = {}
        for n in names:
            if hasattr(cls, n):
                d = getattr(cls, n)
                if isinstance(d, _Field):
                    if d.default is not None:
                        defaults[n] = d.default
                elif not callable(d) or isinstance(d, (int, float)):
                    defaults[n] = d

        def __init__(self, *args, **kwargs):
            vals = dict(defaults)
            for n, a in zip(names, args):
                vals[n] = a
            vals.update(kwargs)
            missing = [n for n in names if n not in vals]
            if missing:
                raise TypeError(f"{cls.__name__} missing fields {missing}")
            for n in names:
                object.__setattr__(self, n, vals[n])
        __init__._shim_auto = True
        cls.__init__ = __init__
        cls.__dataclass_fields__ = {n: None for n in names}


class _Path:
    def __init__(self, names=()):
        object.__setattr__(self, "_names", tuple(names))

    def __getattr__(self, n):
        return _Path(self._names + (n,))


def tree_at(where, pytree, replace=None, replace_fn=None):
    names = where(_Path())._names

    def rec(obj, names):
        if not names:
            return replace if replace_fn is None else replace_fn(obj)
        new = copy.copy(obj)
        object.__setattr__(new, names[0], rec(getattr(obj, names[0]), names[1:]))
        return new
    return rec(pytree, names)


class Linear(Module):
    """equinox.nn.Linear restated: y = weight @ x + bias; weight (out,in), bias (out,);
    default init U(-1/sqrt(in), 1/sqrt(in)) for both (drawn here from numpy's default_rng)."""
    weight: torch.Tensor
    bias: object

    def __init__(self, in_features, out_features, use_bias=True, *, key):
        lim = 1.0 / _np.sqrt(in_features)
        self.weight = torch.tensor(key.uniform(-lim, lim, (out_features, in_features)), dtype=F64, requires_grad=True)
        self.bias = torch.tensor(key.uniform(-lim, lim, (out_features,)), dtype=F64, requires_grad=True) \
            if use_bias else None

    def __call__(self, x):
        y = self.weight @ x
        return y + self.bias if self.bias is not None else y


class EqxMLP(Module):
    """equinox.nn.MLP restated (in_size, out_size, width_size, depth, activation, final_activation,
    use_bias=True, use_final_bias=True, key)."""
    layers: tuple
    activation: object
    final_activation: object

    def __init__(self, in_size, out_size, width_size, depth, activation=None, final_activation=lambda x: x,
                 use_bias=True, use_final_bias=True, *, key):
        rng = key if isinstance(key, _np.random.Generator) else _np.random.default_rng(int(_np.asarray(key).ravel()[-1]))
        layers = []
        if depth == 0:
            layers.append(Linear(in_size, out_size, use_final_bias, key=rng))
        else:
            layers.append(Linear(in_size, width_size, use_bias, key=rng))
            for _ in range(depth - 1):
                layers.append(Linear(width_size, width_size, use_bias, key=rng))
            layers.append(Linear(width_size, out_size, use_final_bias, key=rng))
        self.layers = tuple(layers)
        self.activation = activation
        self.final_activation = final_activation

    def __call__(self, x):
        for layer in self.layers[:-1]:
            x = self.activation(layer(x))
        return self.final_activation(self.layers[-1](x))


def mlp_leaves(mlp):
    """Leaves in equinox order: [W0, b0, W1, b1, ..., Wout (, bout)]."""
    out = []
    for layer in mlp.layers:
        out.append(layer.weight)
        if layer.bias is not None:
            out.append(layer.bias)
    return out


# --------------------------------------------------------------------------- install
def install():
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_is_refshim_torch", False):
        return
    if "jax" in sys.modules:
        raise RuntimeError("another jax (or the numpy stand-in) is already imported in this process")
    _patch_tensor()
    jax = types.ModuleType("jax")
    jax._is_refshim_torch = True
    jnp, lin = _make_jnp()
    jax.numpy = jnp
    lax = types.ModuleType("jax.lax")
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if bool(pred) else ff(*ops)
    lax.switch = lambda idx, branches, *ops: branches[int(idx)](*ops)

    def while_loop(cond, body, val):
        while bool(cond(val)):
            val = body(val)
        return val
    lax.while_loop = while_loop

    def scan(f, init, xs):
        carry, ys = init, []
        for x in xs:
            carry, y = f(carry, x)
            ys.append(y)
        return carry, ys
    lax.scan = scan

    def fori_loop(lo, hi, body, val):
        for i in range(lo, hi):
            val = body(i, val)
        return val
    lax.fori_loop = fori_loop

    def custom_root(f, initial_guess, solve, tangent_solve, has_aux=False):
        """jax.lax.custom_root: the root is found by the caller's ``solve`` (the reference's own rtsafe_) and differentiated by the implicit function theorem.  Here the derivatives are attached with
        two differentiable Newton steps from the converged root: Newton's map is a contraction of order 2 at the
        root, so k steps reproduce the derivatives of the implicit function up to order 2^k - 1 = 3 -- enough for the
        gradient of a residual loss (third-order in the energy)."""
        res = solve(f, initial_guess)                    # (its own jax.value_and_grad needs a tape; the result is detached)
        x_star, aux = res if has_aux else (res, None)
        x = _t(x_star).detach().clone().requires_grad_(True)
        for _ in range(2):
            fx = f(x)
            if not (isinstance(fx, torch.Tensor) and fx.requires_grad):
                break
            (dfx,) = torch.autograd.grad(fx, x, create_graph=True)
            x = x - fx / dfx
        return (x, aux) if has_aux else x
    lax.custom_root = custom_root
    jax.lax = lax
    jax.vmap, jax.grad, jax.value_and_grad, jax.custom_jvp = vmap, grad, value_and_grad, custom_jvp
    jax.jit = lambda f, *a, **k: f

    def _na(name):
        def f(*a, **k):
            raise NotImplementedError(f"jax.{name} is not provided by the torch stand-in")
        return f
    jax.jacfwd = jacfwd
    jax.jacrev = jacfwd
    jax.hessian = lambda f, argnums=0, has_aux=False: jacfwd(jacfwd(f, argnums=argnums), argnums=argnums)
    for nm in ("jvp",):
        setattr(jax, nm, lambda f, *a, _n=nm, **k: _na(_n))
    jsp = types.ModuleType("jax.scipy")
    jsp.linalg = types.ModuleType("jax.scipy.linalg")
    def _expm(A):
        # simple_fefv.py:94 (state update).  torch.linalg.matrix_exp on ONE small matrix picks too low a Taylor degree
        # (measured: 1e-10 absolute error at ||A|| ~ 0.01 .. 0.05 against eigh / a 30-term series; the batched code
        # path is accurate to 2e-15), so a single matrix goes through the batched path
        A = _t(A)
        if A.ndim == 2:
            return torch.linalg.matrix_exp(torch.stack([A, A]))[0]
        return torch.linalg.matrix_exp(A)
    jsp.linalg.expm = _expm
    jax.scipy = jsp
    jnn = types.ModuleType("jax.nn")
    jnn.sigmoid = lambda x: torch.sigmoid(_t(x))
    jnn.tanh = lambda x: torch.tanh(_t(x))
    jnn.softplus = lambda x: torch.nn.functional.softplus(_t(x))
    jax.nn = jnn
    jrandom = types.ModuleType("jax.random")
    jrandom.PRNGKey = lambda s: _np.array([0, s], dtype=_np.uint32)
    jax.random = jrandom
    jtu = types.ModuleType("jax.tree_util")
    jax.tree_util = jtu
    jax.tree = types.SimpleNamespace()
    jax.Array = torch.Tensor
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": lin, "jax.lax": lax, "jax.scipy": jsp,
            "jax.scipy.linalg": jsp.linalg, "jax.nn": jnn, "jax.random": jrandom, "jax.tree_util": jtu}

    eqx = types.ModuleType("equinox")
    eqx.Module = Module
    eqx.field = lambda *, static=False, default=None, **kw: _Field(default=default, static=static)
    eqx.filter_jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
    eqx.filter_vmap = lambda f=None, **k: (vmap(f) if f is not None else (lambda g: vmap(g)))
    eqx.is_array = lambda x: isinstance(x, torch.Tensor)
    eqx.is_inexact_array = eqx.is_array
    eqx.if_array = lambda ax: ax
    eqx.tree_at = tree_at
    eqx.combine = lambda diff, static: diff       # the stand-in never partitions
    eqx.nn = types.SimpleNamespace(MLP=EqxMLP, Linear=Linear)
    mods["equinox"] = eqx

    jt = types.ModuleType("jaxtyping")

    class _T:
        def __class_getitem__(cls, item):
            return cls
    for nm in ("Float", "Int", "Bool", "PyTree"):
        setattr(jt, nm, type(nm, (_T,), {}))
    jt.Array = torch.Tensor      # properties.py:_check_other_type does isinstance(other, Array)
    mods["jaxtyping"] = jt

    for nm in ("netCDF4", "optax", "optimistix", "matplotlib", "matplotlib.pyplot", "vtk"):
        if nm not in sys.modules:
            m = types.ModuleType(nm)
            m.__getattr__ = lambda item, _n=nm: (_ for _ in ()).throw(
                AttributeError(f"{_n}.{item} is not available (stub)"))
            mods[nm] = m
    sys.modules.update(mods)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def import_reference(root="/root/reference"):
    """Make ``pancax.<sub>`` importable from ``root`` without running pancax/__init__.py."""
    import os
    install()
    if "pancax" in sys.modules and getattr(sys.modules["pancax"], "_is_refshim_pkg", False):
        return
    pkg = types.ModuleType("pancax")
    pkg.__path__ = [os.path.join(root, "pancax")]
    pkg._is_refshim_pkg = True
    sys.modules["pancax"] = pkg
