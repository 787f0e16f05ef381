"""pancax/domains.py: VariationalDomain(mesh_file, times, p_order=1, q_order=2) (:132-156)."""
from __future__ import annotations

import numpy as np

from .fem import DofManager, Mesh, read_exodus_mesh


class SimulationTimesNotUniqueException(Exception):
    pass


class SimulationTimesNotStrictlyIncreasingException(Exception):
    pass


class VariationalDomain:
    def __init__(self, mesh_file, times, p_order: int = 1, q_order: int = 2):
        """``mesh_file``: path of an Exodus file, or an in-memory ``Mesh`` (synthetic meshes)."""
        if isinstance(mesh_file, Mesh):
            mesh, self.mesh_file = mesh_file, None
        else:
            mesh, self.mesh_file = read_exodus_mesh(mesh_file), mesh_file
        if p_order != 1:
            raise NotImplementedError("pancax_b200 handles first-order elements (Hex8/Quad4/Tri3)")
        times = np.asarray(times, dtype=np.float64)
        # domains.py:50-57
        if len(times) != len(set(times.tolist())):
            raise SimulationTimesNotUniqueException()
        for i in range(len(times) - 1):
            if times[i] >= times[i + 1]:
                raise SimulationTimesNotStrictlyIncreasingException()
        self.mesh = mesh
        self.coords = mesh.coords
        self.conns = mesh.conns
        self.times = times
        self.q_order = int(q_order)
        self.dof_manager = None

    def update_dof_manager(self, dirichlet_bcs, n_dofs):
        """domains.py:64-71"""
        self.dof_manager = DofManager(self.mesh, n_dofs, dirichlet_bcs)
        return self


class CollocationDomain:
    """pancax/domains.py:72-81: the mesh nodes are the collocation points of the strong-form losses."""

    def __init__(self, mesh_file, times, p_order: int = 1):
        if isinstance(mesh_file, Mesh):
            mesh, self.mesh_file = mesh_file, None
        else:
            mesh, self.mesh_file = read_exodus_mesh(mesh_file), mesh_file
        times = np.asarray(times, dtype=np.float64)
        if len(times) != len(set(times.tolist())):
            raise SimulationTimesNotUniqueException()
        for i in range(len(times) - 1):
            if times[i] >= times[i + 1]:
                raise SimulationTimesNotStrictlyIncreasingException()
        self.mesh, self.coords, self.conns, self.times = mesh, mesh.coords, mesh.conns, times
        self.q_order = 1
        self.dof_manager = None

    def update_dof_manager(self, dirichlet_bcs, n_dofs):
        self.dof_manager = DofManager(self.mesh, n_dofs, dirichlet_bcs)
        return self

    def collocation_points(self):
        """All (x..., t) rows, time-major: what CollocationDataLoader assembles (domains.py:96-110)."""
        nn = self.coords.shape[0]
        return np.column_stack([np.tile(self.coords, (len(self.times), 1)), np.repeat(self.times, nn)])


class CollocationDataLoader:
    """pancax/domains.py:88-130: shuffled full batches of the collocation rows (x..., t) of a CollocationDomain -- all
    mesh nodes at every time, time-major -- with zero targets of ``num_fields`` columns (the trailing partial batch is
    dropped)."""

    def __init__(self, domain: CollocationDomain, num_fields: int) -> None:
        self.inputs = domain.collocation_points()
        self.outputs = np.zeros((self.inputs.shape[0], num_fields))
        self.indices = np.arange(self.inputs.shape[0])

    def __len__(self):
        return self.inputs.shape[0]

    def dataloader(self, batch_size: int):
        order = np.random.permutation(self.indices)
        for b in range(len(self) // batch_size):
            rows = order[b * batch_size:(b + 1) * batch_size]
            yield self.inputs[rows], self.outputs[rows]

