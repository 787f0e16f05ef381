"""pancax/problems.py: ForwardProblem (:27-135), InverseProblem (:138-154)."""
from __future__ import annotations

from typing import Callable, List, Optional

from .domains import CollocationDomain, VariationalDomain
from .physics_kernels import BaseEnergyFormPhysics, BaseStrongFormPhysics


class DomainPhysicsCompatabilityError(Exception):
    pass


class ForwardProblem:
    def __init__(self, domain, physics, ics: Optional[List[Callable]] = None,
                 dirichlet_bcs: Optional[list] = None, neumann_bcs: Optional[list] = None) -> None:
        ok = (isinstance(domain, VariationalDomain) and isinstance(physics, BaseEnergyFormPhysics)) or \
            (isinstance(domain, CollocationDomain) and isinstance(physics, BaseStrongFormPhysics))   # problems.py:52-66
        if not ok:
            raise DomainPhysicsCompatabilityError(
                f"Incompatable domain and physics.\nGot domain of type = {type(domain)}\n"
                f"Got physics of type = {type(physics)}")
        dirichlet_bcs = [] if dirichlet_bcs is None else dirichlet_bcs
        if neumann_bcs:
            raise NotImplementedError("Neumann BCs are not on the B200 hot path")
        domain = domain.update_dof_manager(dirichlet_bcs, physics.n_dofs)   # problems.py:70
        self.domain = domain
        self.physics = physics
        self.ics = [] if ics is None else ics
        self.dirichlet_bcs = dirichlet_bcs
        self.neumann_bcs = []
        self.is_delta_pinn = False
        self.shard = None          # parallel.shard_problem(): this problem is one element shard of a larger mesh
        self._engines = {}

    conns = property(lambda self: self.domain.conns)
    coords = property(lambda self: self.domain.coords)
    dof_manager = property(lambda self: self.domain.dof_manager)
    mesh = property(lambda self: self.domain.mesh)
    mesh_file = property(lambda self: self.domain.mesh_file)
    n_dims = property(lambda self: self.domain.coords.shape[1])
    n_dofs = property(lambda self: self.physics.n_dofs)
    times = property(lambda self: self.domain.times)


class InverseProblem(ForwardProblem):
    def __init__(self, domain, physics, field_data, global_data, ics=None, dirichlet_bcs=None,
                 neumann_bcs=None) -> None:
        super().__init__(domain, physics, ics, dirichlet_bcs, neumann_bcs)
        self.field_data = field_data
        self.global_data = global_data
