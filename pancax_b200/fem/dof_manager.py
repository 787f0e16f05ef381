from __future__ import annotations

import numpy as np


class DofManager:
    """The part of pancax/fem/dof_manager.py:19-59 the hot path reads: isBc / isUnknown /
    unknownIndices (row-major dof id = node*dim + component, :39-41).  Vectorised; the Hessian
    bookkeeping of the reference (O(ne) Python loops) is not built."""

    def __init__(self, mesh, dim: int, DirichletBCs) -> None:
        self.fieldShape = mesh.num_nodes, dim
        self.isBc = np.full(self.fieldShape, False, dtype=bool)
        for ebc in DirichletBCs:
            self.isBc[mesh.nodeSets[ebc.nset_name], ebc.component] = True
        self.isUnknown = ~self.isBc
        self.ids = np.arange(self.isBc.size).reshape(self.fieldShape)
        self.unknownIndices = self.ids[self.isUnknown]
        self.bcIndices = self.ids[self.isBc]

    def get_bc_size(self) -> int:
        return int(np.sum(self.isBc))

    def get_unknown_size(self) -> int:
        return int(np.sum(self.isUnknown))

    def create_field(self, Uu, Ubc=0.0):
        U = np.zeros(self.isBc.shape)
        U[self.isBc] = Ubc
        U[self.isUnknown] = Uu
        return U
