from __future__ import annotations

import numpy as np

from .mesh import Mesh

_ELEM = {"hex8": "hex8", "hex": "hex8", "quad4": "quad4", "quad": "quad4", "tri3": "tri3", "tri": "tri3"}


def _names(ds, key, n, prefix):
    out = [""] * n
    if key in ds.variables:
        raw = ds.variables[key][:]
        for i in range(n):
            out[i] = b"".join(raw[i].tolist()).split(b"\x00")[0].decode().strip()
    # unnamed sets get auto-generated names: read_exodus_mesh.py:99-102,130-132
    return [s if s else f"{prefix}_{i + 1}" for i, s in enumerate(out)]


def read_exodus_mesh(fileName: str) -> Mesh:
    """Same inputs/outputs as pancax/fem/read_exodus_mesh.py:16-71 for the element types on the
    B200 path (HEX8, QUAD4, TRI3), read with scipy's NetCDF-3 reader (the Exodus fixtures are
    64-bit-offset classic files).  Other element types raise ValueError like the reference does for
    unknown ones."""
    from scipy.io import netcdf_file
    with netcdf_file(fileName, "r", mmap=False) as ds:
        nd = ds.dimensions["num_dim"]
        coords = np.column_stack([np.array(ds.variables["coord" + c][:], dtype=np.float64) for c in "xyz"[:nd]])
        if ds.dimensions.get("num_el_blk", 1) != 1:
            raise ValueError("pancax_b200 handles one element block (the reference's potential_energy "
                             "is single-block too: physics_kernels/base.py:334,348)")
        conn = ds.variables["connect1"]
        conns = np.array(conn[:], dtype=np.int32) - 1
        et = conn.elem_type
        et = (et.decode() if isinstance(et, bytes) else str(et)).strip().lower()
        if et not in _ELEM:
            raise ValueError(f"Unsupported element type: {et}")
        nns = ds.dimensions.get("num_node_sets", 0) or 0
        names = _names(ds, "ns_names", nns, "nodeset")
        nodeSets = {}
        for i in range(nns):
            key = f"node_ns{i + 1}"
            if key in ds.variables:
                nodeSets[names[i]] = np.array(ds.variables[key][:], dtype=np.int32) - 1
        blocks = {_names(ds, "eb_names", 1, "block")[0]: np.arange(conns.shape[0])}
    return Mesh(coords, conns, _ELEM[et], nodeSets, blocks)
