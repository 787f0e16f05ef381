"""pancax/bcs.py:23-44: DirichletBC(component, function, block_name, nset_name, sset_name)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional


@dataclass
class DirichletBC:
    component: int
    function: Callable = lambda x, t: 0.0
    block_name: Optional[str] = None
    nset_name: Optional[str] = None
    sset_name: Optional[str] = None

    def coordinates(self, mesh):
        return mesh.coords[mesh.nodeSets[self.nset_name], :]
