"""Oracle (test infrastructure): parent-element tables, quadrature rules, Exodus I/O.

numpy only.  Each function cites the reference source (pancax 0.0.16 under
/root/reference) whose behaviour it restates.  Shapes follow what the reference *code*
builds: ``values (nq, nnpe)``, ``gradients (nq, nnpe, nd)`` (the docstring at
pancax/fem/elements/base_element.py:129-137 says (np, nd, nn); the code does not).
"""
from __future__ import annotations

import numpy as np
from scipy.io import netcdf_file

HEX8, QUAD4, TRI3 = "hex8", "quad4", "tri3"
NNPE = {HEX8: 8, QUAD4: 4, TRI3: 3}
NDIM = {HEX8: 3, QUAD4: 2, TRI3: 2}

# pancax/fem/elements/hex8_element.py:9-20
_HEX8_NODES = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)
# pancax/fem/elements/quad4_element.py:9-14
_QUAD4_NODES = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=np.float64)
# pancax/fem/elements/simplex_tri_element.py:19-35 with degree=1 -> (1,0),(0,1),(0,0)
_TRI3_NODES = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])


def quadrature_rule(elem: str, q_order: int):
    """(xi (nq, nd), w (nq,)).  pancax/fem/quadrature_rules.py:109-132 (hex),
    :135-180 (quad; degree-3 keeps the reference's 25/80 last weight, :174),
    :206-245 (triangle, degree 1-2)."""
    if elem == HEX8:
        if q_order == 1:
            return np.zeros((1, 3)), 8.0 * np.ones(1)
        if q_order == 2:
            return np.sqrt(1.0 / 3.0) * _HEX8_NODES, np.ones(8)
        raise ValueError(f"Unsupported quadrature degree {q_order} on hex element.")
    if elem == QUAD4:
        if q_order == 1:
            return np.zeros((1, 2)), 4.0 * np.ones(1)
        if q_order == 2:
            return np.sqrt(1.0 / 3.0) * _QUAD4_NODES, np.ones(4)
        if q_order == 3:
            s = np.sqrt(3.0 / 5.0)
            xi = np.array([[-s, -s], [0, -s], [s, -s], [-s, 0], [0, 0], [s, 0],
                           [-s, s], [0, s], [s, s]], dtype=np.float64)
            w = np.array([25 / 81, 40 / 81, 25 / 81, 40 / 81, 64 / 81, 40 / 81,
                          25 / 81, 40 / 81, 25 / 80], dtype=np.float64)
            return xi, w
        raise ValueError("Invalid quadrature degree: %s" % q_order)
    if elem == TRI3:
        if q_order == 1:
            return (np.array([[3.33333333333333333e-01, 3.33333333333333333e-01]]),
                    np.array([5.00000000000000000e-01]))
        if q_order == 2:
            xi = np.array([[6.66666666666666667e-01, 1.66666666666666667e-01],
                           [1.66666666666666667e-01, 6.66666666666666667e-01],
                           [1.66666666666666667e-01, 1.66666666666666667e-01]])
            w = np.array([1.66666666666666666e-01, 1.66666666666666667e-01,
                          1.66666666666666667e-01])
            return xi, w
        raise ValueError(f"Unsupported quadrature degree {q_order} on tri3 in the oracle.")
    raise ValueError(f"Unsupported element type: {elem}")


def shape_tables(elem: str, xi: np.ndarray):
    """Parent shape values N (nq, nnpe) and parent gradients dN (nq, nnpe, nd) at ``xi``.

    Hex8: pancax/fem/elements/hex8_element.py:39-107; Quad4: quad4_element.py:23-50;
    Tri3: simplex_tri_element.py:49-56 -- the reference solves a Vandermonde system in
    an orthonormal basis; for degree 1 the result is the barycentric functions for the
    nodal order (1,0),(0,1),(0,0), i.e. N=(x, y, 1-x-y), restated here in closed form
    (the Vandermonde route itself is exercised by tests/golden via oracle/refshim).
    """
    xi = np.asarray(xi, dtype=np.float64)
    nq = xi.shape[0]
    if elem == HEX8:
        s = _HEX8_NODES
        N = np.empty((nq, 8))
        dN = np.empty((nq, 8, 3))
        for a in range(8):
            fx = 1.0 + s[a, 0] * xi[:, 0]
            fy = 1.0 + s[a, 1] * xi[:, 1]
            fz = 1.0 + s[a, 2] * xi[:, 2]
            N[:, a] = 0.125 * (fx * fy * fz)
            dN[:, a, 0] = 0.125 * (s[a, 0] * fy * fz)
            dN[:, a, 1] = 0.125 * (s[a, 1] * fx * fz)
            dN[:, a, 2] = 0.125 * (s[a, 2] * fx * fy)
        return N, dN
    if elem == QUAD4:
        s = _QUAD4_NODES
        N = np.empty((nq, 4))
        dN = np.empty((nq, 4, 2))
        for a in range(4):
            fx = 1.0 + s[a, 0] * xi[:, 0]
            fy = 1.0 + s[a, 1] * xi[:, 1]
            N[:, a] = 0.25 * (fx * fy)
            dN[:, a, 0] = 0.25 * (s[a, 0] * fy)
            dN[:, a, 1] = 0.25 * (s[a, 1] * fx)
        return N, dN
    if elem == TRI3:
        N = np.stack([xi[:, 0], xi[:, 1], 1.0 - xi[:, 0] - xi[:, 1]], axis=1)
        dN = np.tile(np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, -1.0]]), (nq, 1, 1))
        return N, dN
    raise ValueError(f"Unsupported element type: {elem}")


class ParentElement:
    """What NonAllocatedFunctionSpace holds (pancax/fem/function_space.py:16-21)."""

    def __init__(self, elem: str, q_order: int):
        self.elem = elem
        self.q_order = q_order
        self.nd = NDIM[elem]
        self.nnpe = NNPE[elem]
        self.xi, self.w = quadrature_rule(elem, q_order)
        self.N, self.dN = shape_tables(elem, self.xi)
        self.nq = self.w.shape[0]


# --------------------------------------------------------------------------- Exodus
_ELEM_TYPE_MAP = {"hex8": HEX8, "hex": HEX8, "quad4": QUAD4, "quad": QUAD4,
                  "tri3": TRI3, "tri": TRI3}


def _names(ds, key, n, prefix):
    out = []
    if key in ds.variables:
        raw = ds.variables[key][:]
        for i in range(n):
            s = b"".join(raw[i].tolist()).split(b"\x00")[0].decode().strip()
            out.append(s)
    else:
        out = [""] * n
    return [s if s else f"{prefix}_{i + 1}" for i, s in enumerate(out)]


def read_exodus_mesh(path: str):
    """Reads coords (nn, nd) f64, conns (ne, nnpe) int32 0-based, element type and node
    sets from a NetCDF-3 Exodus file.  pancax/fem/read_exodus_mesh.py:16-71 (driver),
    :75-100 (coords, `connect1 - 1`), :99-102/:130-132 (auto names block_i/nodeset_i)."""
    with netcdf_file(path, "r", mmap=False) as ds:
        nd = ds.dimensions["num_dim"]
        cols = [np.array(ds.variables["coord" + c][:], dtype=np.float64)
                for c in "xyz"[:nd]]
        coords = np.ascontiguousarray(np.column_stack(cols))
        conn = ds.variables["connect1"]
        conns = np.ascontiguousarray(np.array(conn[:], dtype=np.int32) - 1)
        et = conn.elem_type
        et = (et.decode() if isinstance(et, bytes) else str(et)).strip().lower()
        if et not in _ELEM_TYPE_MAP:
            raise ValueError(f"Unsupported element type: {et}")
        nodesets = {}
        nns = ds.dimensions.get("num_node_sets", 0) or 0
        names = _names(ds, "ns_names", nns, "nodeset")
        for i in range(nns):
            key = f"node_ns{i + 1}"
            if key in ds.variables:
                nodesets[names[i]] = np.array(ds.variables[key][:], dtype=np.int32) - 1
    return dict(coords=coords, conns=conns, elem=_ELEM_TYPE_MAP[et], nodesets=nodesets)


def structured_mesh(elem: str, n, lengths=None, permute_seed=None):
    """Synthetic structured block (unit cube/square by default) with Cubit-like node sets
    (Quad4 nset_1: y=1, nset_2: x=0, nset_3: y=0, nset_4: x=1;
    Hex8 nset_1: z=1, nset_2: z=0, nset_3: y=0, nset_4: x=0, nset_5: y=1, nset_6: x=1).
    ``permute_seed`` applies a random node renumbering (gather/scatter test)."""
    nd = NDIM[elem]
    n = (n,) * nd if np.isscalar(n) else tuple(n)
    lengths = (1.0,) * nd if lengths is None else tuple(lengths)
    axes = [np.linspace(0.0, lengths[d], n[d] + 1) for d in range(nd)]
    if nd == 2:
        nx, ny = n
        X, Y = np.meshgrid(axes[0], axes[1], indexing="xy")  # node id = j*(nx+1)+i
        coords = np.column_stack([X.ravel(), Y.ravel()])
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        n0 = (j * (nx + 1) + i).ravel()
        quads = np.column_stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1])
        if elem == QUAD4:
            conns = quads
        else:  # two counter-clockwise triangles per cell
            conns = np.vstack([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]])
        tol = 1e-12
        nodesets = {
            "nset_1": np.where(np.abs(coords[:, 1] - lengths[1]) < tol)[0],
            "nset_2": np.where(np.abs(coords[:, 0]) < tol)[0],
            "nset_3": np.where(np.abs(coords[:, 1]) < tol)[0],
            "nset_4": np.where(np.abs(coords[:, 0] - lengths[0]) < tol)[0],
        }
    else:
        nx, ny, nz = n
        Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        coords = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        sx, sy = 1, nx + 1
        sz = (nx + 1) * (ny + 1)
        n0 = (k * sz + j * sy + i * sx).ravel()
        conns = np.column_stack([n0, n0 + sx, n0 + sx + sy, n0 + sy,
                                 n0 + sz, n0 + sx + sz, n0 + sx + sy + sz, n0 + sy + sz])
        tol = 1e-12
        nodesets = {
            "nset_1": np.where(np.abs(coords[:, 2] - lengths[2]) < tol)[0],
            "nset_2": np.where(np.abs(coords[:, 2]) < tol)[0],
            "nset_3": np.where(np.abs(coords[:, 1]) < tol)[0],
            "nset_4": np.where(np.abs(coords[:, 0]) < tol)[0],
            "nset_5": np.where(np.abs(coords[:, 1] - lengths[1]) < tol)[0],
            "nset_6": np.where(np.abs(coords[:, 0] - lengths[0]) < tol)[0],
        }
    conns = conns.astype(np.int32)
    if permute_seed is not None:
        rng = np.random.default_rng(permute_seed)
        perm = rng.permutation(coords.shape[0])          # new id of old node i = perm[i]
        new_coords = np.empty_like(coords)
        new_coords[perm] = coords
        coords = new_coords
        conns = perm[conns].astype(np.int32)
        nodesets = {k: np.sort(perm[v]) for k, v in nodesets.items()}
    nodesets = {k: v.astype(np.int32) for k, v in nodesets.items()}
    return dict(coords=np.ascontiguousarray(coords), conns=np.ascontiguousarray(conns),
                elem=elem, nodesets=nodesets)


def unknown_indices(nn: int, ndof: int, nodesets: dict, dirichlet_bcs):
    """Row-major dof ids (node*ndof+comp) that are NOT fixed.  ``dirichlet_bcs`` is a list
    of (nset_name, component).  pancax/fem/dof_manager.py:31-42."""
    is_bc = np.zeros((nn, ndof), dtype=bool)
    for nset, comp in dirichlet_bcs:
        is_bc[nodesets[nset], comp] = True
    ids = np.arange(nn * ndof).reshape(nn, ndof)
    return ids[~is_bc].astype(np.int64)
