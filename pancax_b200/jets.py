"""Second-order forward-mode arithmetic on numpy arrays, used ONCE at set-up to tabulate a Dirichlet
wrapper g(x, t, z) = a(x, t) + b(x, t) z together with the derivatives the strong-form path needs
(d_j a, lap a, a_t, d_j b, lap b, b_t): the wrapper is an arbitrary Python callable (fields.py:101) that
cannot run inside a kernel, so it is evaluated on the host with ``Jet`` numbers standing in for x and t.

A Jet carries, for N points at once: value v (N,), first derivatives g (N, nd+1) w.r.t. (x_0..x_{nd-1}, t)
and the spatial Laplacian l (N,) = sum_j d^2/dx_j^2.  Supported: + - * / ** (real exponent), unary -, and
the numpy ufuncs sin cos exp log sqrt tanh."""
from __future__ import annotations

import numpy as np


class Jet:
    __array_priority__ = 1000

    def __init__(self, v, g, l, nd):
        self.v, self.g, self.l, self.nd = v, g, l, nd

    # ---- construction helpers
    @staticmethod
    def const(c, like):
        v = np.broadcast_to(np.asarray(c, dtype=np.float64), like.v.shape).copy()
        return Jet(v, np.zeros_like(like.g), np.zeros_like(like.l), like.nd)

    def _lift(self, o):
        return o if isinstance(o, Jet) else Jet.const(o, self)

    def _chain(self, f0, f1, f2):
        """f(self) with f0 = f(v), f1 = f'(v), f2 = f''(v)."""
        sp = self.g[:, :self.nd]
        return Jet(f0, f1[:, None] * self.g, f1 * self.l + f2 * np.sum(sp * sp, axis=1), self.nd)

    # ---- arithmetic
    def __add__(self, o):
        if isinstance(o, Vec):
            return NotImplemented          # Vec.__radd__ broadcasts over the components
        o = self._lift(o)
        return Jet(self.v + o.v, self.g + o.g, self.l + o.l, self.nd)
    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.g, -self.l, self.nd)

    def __sub__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        return self + (-self._lift(o))

    def __rsub__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        return self._lift(o) - self

    def __mul__(self, o):
        if isinstance(o, Vec):
            return o * self
        o = self._lift(o)
        nd = self.nd
        cross = np.sum(self.g[:, :nd] * o.g[:, :nd], axis=1)
        return Jet(self.v * o.v, self.g * o.v[:, None] + self.v[:, None] * o.g,
                   self.l * o.v + 2.0 * cross + self.v * o.l, nd)
    __rmul__ = __mul__

    def reciprocal(self):
        v = self.v
        return self._chain(1.0 / v, -1.0 / v ** 2, 2.0 / v ** 3)

    def __truediv__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        return self * self._lift(o).reciprocal()

    def __rtruediv__(self, o):
        return self._lift(o) * self.reciprocal()

    def __pow__(self, p):
        if isinstance(p, Jet):
            raise NotImplementedError("Jet ** Jet")
        p = float(p)
        v = self.v
        return self._chain(v ** p, p * v ** (p - 1.0), p * (p - 1.0) * v ** (p - 2.0))

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__":
            return NotImplemented
        name = ufunc.__name__
        if name in ("add", "subtract", "multiply", "true_divide", "divide", "power"):
            a, b = inputs
            op = {"add": lambda x, y: x + y, "subtract": lambda x, y: x - y, "multiply": lambda x, y: x * y,
                  "true_divide": lambda x, y: x / y, "divide": lambda x, y: x / y, "power": lambda x, y: x ** y}[name]
            a = a if isinstance(a, Jet) else self._lift(a)
            return op(a, b)
        (x,) = inputs
        v = x.v
        if name == "sin":
            return x._chain(np.sin(v), np.cos(v), -np.sin(v))
        if name == "cos":
            return x._chain(np.cos(v), -np.sin(v), -np.cos(v))
        if name == "exp":
            e = np.exp(v)
            return x._chain(e, e, e)
        if name == "log":
            return x._chain(np.log(v), 1.0 / v, -1.0 / v ** 2)
        if name == "sqrt":
            s = np.sqrt(v)
            return x._chain(s, 0.5 / s, -0.25 / (s * v))
        if name == "tanh":
            t = np.tanh(v)
            return x._chain(t, 1.0 - t * t, -2.0 * t * (1.0 - t * t))
        if name == "negative":
            return -x
        if name == "square":
            return x * x
        return NotImplemented


class Vec:
    """A short vector of components (floats or Jets) that can be indexed like the reference's per-point
    arrays (``x[0]``) and like batched arrays (``x[:, 0]``), with elementwise arithmetic and jax's
    ``.at[i].set(v)``."""

    def __init__(self, comps):
        self.c = list(comps)

    def __len__(self):
        return len(self.c)

    def __iter__(self):
        return iter(self.c)

    def __getitem__(self, i):
        if isinstance(i, tuple):            # x[:, j]
            i = i[-1]
        if isinstance(i, slice):
            return Vec(self.c[i])
        return self.c[i]

    @property
    def shape(self):
        return (len(self.c),)

    @property
    def at(self):
        outer = self

        class _At:
            def __getitem__(self, i):
                class _S:
                    def set(self_inner, v):
                        c = list(outer.c)
                        c[i] = v
                        return Vec(c)

                    def add(self_inner, v):
                        c = list(outer.c)
                        c[i] = c[i] + v
                        return Vec(c)
                return _S()
        return _At()

    def _bin(self, o, f):
        if isinstance(o, Vec):
            return Vec([f(a, b) for a, b in zip(self.c, o.c)])
        return Vec([f(a, o) for a in self.c])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __neg__(self): return Vec([-a for a in self.c])


def tabulate_bc_jet(bc_func, pts, ndof):
    """Tables (n, ndof, 2*(nd+3)) = (a, d_j a.., lap a, a_t, b, d_j b.., lap b, b_t) of a diagonal-affine
    Dirichlet wrapper u = a(x,t) + b(x,t) z at the points pts (n, nd+1) = (x..., t).  ``bc_func(x, t, z)``
    may be written per point (``x[0]``) or batched (``x[:, 0]``); ``None`` -> identity (returns None)."""
    if bc_func is None:
        return None
    pts = np.asarray(pts, dtype=np.float64)
    n, nd = pts.shape[0], pts.shape[1] - 1
    K = nd + 3

    def var(j):
        g = np.zeros((n, nd + 1))
        g[:, j] = 1.0
        return Jet(pts[:, j].copy(), g, np.zeros(n), nd)
    x = Vec([var(j) for j in range(nd)])
    t = var(nd)

    def evaluate(zval):
        out = bc_func(x, t, Vec([zval] * ndof))
        comps = list(out) if isinstance(out, (Vec, list, tuple)) else [out]
        if len(comps) != ndof:
            raise ValueError(f"dirichlet_bc_func returned {len(comps)} components, expected {ndof}")
        return [c if isinstance(c, Jet) else Jet.const(c, t) for c in comps]
    g0, g1 = evaluate(0.0), evaluate(1.0)
    # affinity AND diagonality: a different probe value per component, so a wrapper that mixes components
    # (u_i depending on z_j, j != i) fails the check instead of passing with every z_j equal
    probe = [-0.7 + 0.45 * i for i in range(ndof)]
    out2 = bc_func(x, t, Vec(probe))
    comps2 = list(out2) if isinstance(out2, (Vec, list, tuple)) else [out2]
    g2 = [c if isinstance(c, Jet) else Jet.const(c, t) for c in comps2]
    tab = np.zeros((n, ndof, 2 * K))
    for i in range(ndof):
        a, b = g0[i], g1[i] - g0[i]
        chk = a + b * probe[i] - g2[i]
        scale = max(1.0, float(np.abs(g2[i].v).max()))
        if max(np.abs(chk.v).max(), np.abs(chk.l).max()) > 1e-9 * scale:
            raise NotImplementedError("dirichlet_bc_func is not diagonal-affine in the network output z; pancax_b200 "
                                      "tabulates u_i = a_i(x,t) + b_i(x,t) z_i and its derivatives")
        for off, q in ((0, a), (K, b)):
            tab[:, i, off] = q.v
            tab[:, i, off + 1:off + 1 + nd] = q.g[:, :nd]
            tab[:, i, off + 1 + nd] = q.l
            tab[:, i, off + 2 + nd] = q.g[:, nd]
    return tab
