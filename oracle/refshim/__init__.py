"""Test infrastructure: a numpy stand-in for the ``jax`` / ``equinox`` / ``jaxtyping`` namespaces,
just large enough to EXECUTE the reference's own pure-Python source for the forward pieces of the
hot path (parent tables, quadrature rules, function space, constitutive energies, sym-3x3 eigen /
log-sqrt, potential_energy) inside this container, where jax cannot be installed.

No automatic differentiation: ``vmap`` is a Python loop, ``lax.cond`` an ``if``, ``custom_jvp`` keeps
the primal function and exposes the registered rule as ``.jvp_rule``.  Only oracle/gen_golden.py
uses it (to produce tests/golden/*.npz); nothing at test or run time imports it.
"""
from __future__ import annotations

import sys
import types

import numpy as _np


class ShimArray(_np.ndarray):
    """ndarray with jax's functional ``.at[idx].set/add``."""

    class _At:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            a = self.a

            class _Op:
                def set(self, v):
                    b = _np.array(a, copy=True).view(ShimArray)
                    b[idx] = v
                    return b

                def add(self, v):
                    b = _np.array(a, copy=True).view(ShimArray)
                    b[idx] += v
                    return b
            return _Op()

    @property
    def at(self):
        return ShimArray._At(self)


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, ShimArray):
        return x.view(ShimArray)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    return x


class _NumpyProxy(types.ModuleType):
    """jax.numpy -> numpy, results re-wrapped as ShimArray."""

    def __init__(self, name, target):
        super().__init__(name)
        self._target = target

    def __getattr__(self, item):
        obj = getattr(self._target, item)
        if callable(obj) and not isinstance(obj, type):
            def f(*a, **k):
                k.pop("dtype", None) if item in ("array", "asarray") and False else None
                return _wrap(obj(*a, **k))
            f.__name__ = item
            return f
        return obj


def _tree_stack(outs):
    first = outs[0]
    if isinstance(first, tuple) and hasattr(first, "_fields"):      # NamedTuple
        return type(first)(*[_tree_stack([o[i] for o in outs]) for i in range(len(first))])
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([o[i] for o in outs]) for i in range(len(first)))
    return _wrap(_np.stack([_np.asarray(o) for o in outs]))


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax] if not isinstance(a, (tuple, list)) else len(a)
                break
        outs = []
        for i in range(n):
            sl = []
            for a, ax in zip(args, axes):
                if ax is None:
                    sl.append(a)
                else:
                    sl.append(_wrap(_np.take(_np.asarray(a), i, axis=ax)))
            outs.append(fun(*sl))
        return _tree_stack(outs)
    return mapped


class custom_jvp:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.jvp_rule, self.nondiff_argnums = fun, None, nondiff_argnums
        self.__name__ = getattr(fun, "__name__", "custom_jvp")

    def defjvp(self, rule):
        self.jvp_rule = rule
        return rule

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


def _no_ad(name):
    def f(*a, **k):
        raise NotImplementedError(f"jax.{name} is not available in the numpy stand-in (no AD)")
    return f


def install():
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_is_refshim", False):
        return
    jax = types.ModuleType("jax")
    jax._is_refshim = True
    jnp = _NumpyProxy("jax.numpy", _np)
    jnp.linalg = _NumpyProxy("jax.numpy.linalg", _np.linalg)
    jnp.ndarray = _np.ndarray
    jnp.float64, jnp.int32, jnp.bool = _np.float64, _np.int32, bool
    jax.numpy = jnp
    lax = types.ModuleType("jax.lax")
    lax.cond = lambda pred, tf, ff, *ops: tf(*ops) if bool(pred) else ff(*ops)
    lax.switch = lambda idx, branches, *ops: branches[int(idx)](*ops)

    def while_loop(cond, body, val):
        while cond(val):
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
    jax.lax = lax
    jax.vmap = vmap
    jax.custom_jvp = custom_jvp
    jax.jit = lambda f, *a, **k: f
    for nm in ("grad", "value_and_grad", "jacfwd", "jacrev", "hessian", "jvp"):
        setattr(jax, nm, lambda f, *a, _n=nm, **k: _no_ad(_n))
    import scipy.linalg as _sl
    jsp = types.ModuleType("jax.scipy")
    jsp.linalg = _NumpyProxy("jax.scipy.linalg", _sl)
    jax.scipy = jsp
    jnn = types.ModuleType("jax.nn")
    jnn.sigmoid = lambda x: _wrap(1.0 / (1.0 + _np.exp(-_np.asarray(x))))
    jnn.tanh = _np.tanh
    jnn.softplus = lambda x: _np.log1p(_np.exp(x))
    jax.nn = jnn
    jrandom = types.ModuleType("jax.random")
    jrandom.PRNGKey = lambda s: _np.array([0, s], dtype=_np.uint32)
    jax.random = jrandom
    jtu = types.ModuleType("jax.tree_util")
    jax.tree_util = jtu
    jax.Array = _np.ndarray
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp.linalg, "jax.lax": lax, "jax.scipy": jsp,
            "jax.scipy.linalg": jsp.linalg, "jax.nn": jnn, "jax.random": jrandom, "jax.tree_util": jtu}

    # ---- equinox
    eqx = types.ModuleType("equinox")

    class _Field:
        def __init__(self, default=None, static=False, **kw):
            self.default, self.static = default, static

    def field(*, static=False, default=None, **kw):
        return _Field(default=default, static=static)

    class Module:
        def __init_subclass__(cls, **kw):
            super().__init_subclass__(**kw)
            if "__init__" in cls.__dict__:
                return
            for klass in cls.__mro__[1:]:
                if klass is Module:
                    break
                # equinox semantics: a user-written __init__ of a parent is inherited as is
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

    eqx.Module, eqx.field = Module, field
    eqx.filter_jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
    eqx.filter_vmap = lambda f=None, **k: (vmap(f) if f is not None else (lambda g: vmap(g)))
    eqx.is_array = lambda x: isinstance(x, _np.ndarray)
    eqx.is_inexact_array = eqx.is_array
    eqx.if_array = lambda ax: ax
    eqx.tree_at = _no_ad("equinox.tree_at")
    eqx.nn = types.SimpleNamespace(MLP=None, Linear=None)
    mods["equinox"] = eqx

    # ---- jaxtyping
    jt = types.ModuleType("jaxtyping")

    class _T:
        def __class_getitem__(cls, item):
            return cls
    for nm in ("Array", "Float", "Int", "Bool", "PyTree"):
        setattr(jt, nm, type(nm, (_T,), {}))
    mods["jaxtyping"] = jt

    # ---- things imported at module level but never used on this path
    for nm in ("netCDF4", "optax", "optimistix", "matplotlib", "matplotlib.pyplot", "vtk"):
        if nm not in sys.modules:
            m = types.ModuleType(nm)
            m.__getattr__ = lambda item, _n=nm: (_ for _ in ()).throw(
                AttributeError(f"{_n}.{item} is not available (stub)"))
            mods[nm] = m
    sys.modules.update(mods)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def import_reference(root="/root/reference"):
    """Make ``pancax.<sub>`` importable from ``root`` WITHOUT running pancax/__init__.py (which
    imports optax, matplotlib, ...).  Returns nothing; use normal imports afterwards."""
    import os
    install()
    if "pancax" in sys.modules and getattr(sys.modules["pancax"], "_is_refshim_pkg", False):
        return
    pkg = types.ModuleType("pancax")
    pkg.__path__ = [os.path.join(root, "pancax")]
    pkg._is_refshim_pkg = True
    sys.modules["pancax"] = pkg
