from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

NNPE = {"hex8": 8, "quad4": 4, "tri3": 3}
NDIM = {"hex8": 3, "quad4": 2, "tri3": 2}
EXODUS_NAME = {"hex8": "HEX8", "quad4": "QUAD4", "tri3": "TRI3"}


@dataclass
class Mesh:
    """pancax/fem/mesh.py:10-60: coords (nn, nd) f64, conns (ne, nnpe) int32 0-based, one element
    type per mesh, node sets by name."""
    coords: np.ndarray
    conns: np.ndarray
    elem_type: str
    nodeSets: Dict[str, np.ndarray] = field(default_factory=dict)
    blocks: Optional[Dict[str, np.ndarray]] = None
    sideSets: Optional[Dict[str, np.ndarray]] = None

    def __post_init__(self):
        self.coords = np.ascontiguousarray(self.coords, dtype=np.float64)
        self.conns = np.ascontiguousarray(self.conns, dtype=np.int32)
        if self.elem_type not in NNPE:
            raise ValueError(f"Unsupported element type: {self.elem_type}")
        if self.conns.shape[1] != NNPE[self.elem_type] or self.coords.shape[1] != NDIM[self.elem_type]:
            raise ValueError("coords/conns shapes do not match the element type")

    @property
    def num_dimensions(self) -> int:
        return self.coords.shape[1]

    @property
    def num_elements(self) -> int:
        return self.conns.shape[0]

    @property
    def num_nodes(self) -> int:
        return self.coords.shape[0]


def create_structured_mesh(elem_type: str, n, lengths=None, origin=None, permute_seed=None) -> Mesh:
    """Structured block of n (per direction) elements with the node sets of the Cubit fixtures
    (Quad4: nset_1 y=max, nset_2 x=min, nset_3 y=min, nset_4 x=max; Hex8: nset_1 z=max,
    nset_2 z=min, nset_3 y=min, nset_4 x=min, nset_5 y=max, nset_6 x=max).  Vectorised: 2 M
    Hex8 elements take well under a second."""
    nd = NDIM[elem_type]
    n = (int(n),) * nd if np.isscalar(n) else tuple(int(v) for v in n)
    lengths = (1.0,) * nd if lengths is None else tuple(lengths)
    origin = (0.0,) * nd if origin is None else tuple(origin)
    axes = [origin[d] + np.linspace(0.0, lengths[d], n[d] + 1) for d in range(nd)]
    if nd == 2:
        nx, ny = n
        X, Y = np.meshgrid(axes[0], axes[1], indexing="xy")
        coords = np.column_stack([X.ravel(), Y.ravel()])
        i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        n0 = (j * (nx + 1) + i).ravel()
        quads = np.column_stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1])
        conns = quads if elem_type == "quad4" else np.vstack([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]])
        ix = np.arange(coords.shape[0]) % (nx + 1)
        iy = np.arange(coords.shape[0]) // (nx + 1)
        nodesets = {"nset_1": np.where(iy == ny)[0], "nset_2": np.where(ix == 0)[0],
                    "nset_3": np.where(iy == 0)[0], "nset_4": np.where(ix == nx)[0]}
    else:
        nx, ny, nz = n
        Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        coords = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
        n0 = (k * sz + j * sy + i * sx).ravel()
        conns = np.column_stack([n0, n0 + sx, n0 + sx + sy, n0 + sy,
                                 n0 + sz, n0 + sx + sz, n0 + sx + sy + sz, n0 + sy + sz])
        ids = np.arange(coords.shape[0])
        ix, iy, iz = ids % (nx + 1), (ids // (nx + 1)) % (ny + 1), ids // sz
        nodesets = {"nset_1": np.where(iz == nz)[0], "nset_2": np.where(iz == 0)[0],
                    "nset_3": np.where(iy == 0)[0], "nset_4": np.where(ix == 0)[0],
                    "nset_5": np.where(iy == ny)[0], "nset_6": np.where(ix == nx)[0]}
    if permute_seed is not None:
        rng = np.random.default_rng(permute_seed)
        perm = rng.permutation(coords.shape[0])
        new_coords = np.empty_like(coords)
        new_coords[perm] = coords
        coords, conns = new_coords, perm[conns]
        nodesets = {k: np.sort(perm[v]) for k, v in nodesets.items()}
    return Mesh(coords, conns.astype(np.int32), elem_type,
                {k: v.astype(np.int32) for k, v in nodesets.items()})


def write_exodus_mesh(path: str, mesh: Mesh, title: str = "pancax_b200 synthetic mesh") -> None:
    """Writes the subset of the Exodus II schema that pancax/fem/read_exodus_mesh.py:16-201 and
    pancax/data.py:222-225 touch as NetCDF-3 64-bit offset, so the real pancax can
    read the same bytes."""
    from scipy.io import netcdf_file
    nn, nd = mesh.coords.shape
    ne, nnpe = mesh.conns.shape
    names = sorted(mesh.nodeSets.keys())
    with netcdf_file(path, "w", version=2) as ds:
        ds.title = title
        ds.api_version = np.float32(8.25)
        ds.version = np.float32(8.25)
        ds.floating_point_word_size = np.int32(8)
        ds.file_size = np.int32(1)
        ds.maximum_name_length = np.int32(32)
        ds.int64_status = np.int32(0)
        for k, v in (("time_step", None), ("len_name", 256), ("num_dim", nd), ("num_nodes", nn),
                     ("num_elem", ne), ("num_el_blk", 1), ("num_el_in_blk1", ne),
                     ("num_nod_per_el1", nnpe), ("four", 4), ("len_string", 33), ("num_qa_rec", 1)):
            ds.createDimension(k, v)
        if names:
            ds.createDimension("num_node_sets", len(names))
        ds.createVariable("time_whole", "d", ("time_step",))
        for j, c in enumerate("xyz"[:nd]):
            v = ds.createVariable("coord" + c, "d", ("num_nodes",))
            v[:] = mesh.coords[:, j]
        v = ds.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
        v[:] = mesh.conns + 1
        v.elem_type = EXODUS_NAME[mesh.elem_type]

        def put_names(var, dim, strs):
            vv = ds.createVariable(var, "c", (dim, "len_name"))
            arr = np.zeros((len(strs), 256), dtype="S1")
            for i, s in enumerate(strs):
                arr[i, :len(s)] = list(s)
            vv[:] = arr
        put_names("eb_names", "num_el_blk", ["block_1"])
        put_names("coor_names", "num_dim", list("xyz"[:nd]))
        for nm, val in (("eb_status", [1]), ("eb_prop1", [1])):
            vv = ds.createVariable(nm, "i", ("num_el_blk",))
            vv[:] = np.array(val, dtype=np.int32)
        ds.variables["eb_prop1"].name = "ID"
        for nm, dim in (("elem_map", "num_elem"), ("elem_num_map", "num_elem"), ("node_num_map", "num_nodes")):
            vv = ds.createVariable(nm, "i", (dim,))
            vv[:] = np.arange(1, (ne if dim == "num_elem" else nn) + 1, dtype=np.int32)
        if names:
            put_names("ns_names", "num_node_sets", names)
            vv = ds.createVariable("ns_status", "i", ("num_node_sets",))
            vv[:] = np.ones(len(names), dtype=np.int32)
            vv = ds.createVariable("ns_prop1", "i", ("num_node_sets",))
            vv[:] = np.arange(1, len(names) + 1, dtype=np.int32)
            vv.name = "ID"
            for i, nm in enumerate(names):
                ds.createDimension(f"num_nod_ns{i + 1}", len(mesh.nodeSets[nm]))
                vv = ds.createVariable(f"node_ns{i + 1}", "i", (f"num_nod_ns{i + 1}",))
                vv[:] = np.asarray(mesh.nodeSets[nm], dtype=np.int32) + 1
                vd = ds.createVariable(f"dist_fact_ns{i + 1}", "d", (f"num_nod_ns{i + 1}",))
                vd[:] = np.ones(len(mesh.nodeSets[nm]))
