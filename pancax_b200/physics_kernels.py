"""pancax/physics_kernels: the energy-form physics on the B200 path."""
from __future__ import annotations

from typing import Callable


class AbstractMechanicsFormulation:
    n_dimensions: int
    name: str


class PlaneStrain(AbstractMechanicsFormulation):
    """solid_mechanics.py:20-29"""
    n_dimensions = 2
    name = "plane_strain"


class ThreeDimensional(AbstractMechanicsFormulation):
    """solid_mechanics.py:74-83"""
    n_dimensions = 3
    name = "three_dimensional"


class PlaneStress(AbstractMechanicsFormulation):
    """solid_mechanics.py:32-71 needs a per-quadrature-point root find: not on the B200 path."""
    n_dimensions = 2
    name = "plane_stress"

    def __init__(self):
        raise NotImplementedError("PlaneStress is not on the B200 hot path (no fallback)")


class BaseEnergyFormPhysics:
    field_value_names: tuple = ()

    @property
    def n_dofs(self):
        return len(self.field_value_names)


class SolidMechanics(BaseEnergyFormPhysics):
    """solid_mechanics.py:86-109"""

    def __init__(self, constitutive_model, formulation) -> None:
        names = ("displ_x", "displ_y")
        if formulation.n_dimensions > 2:
            names = names + ("displ_z",)
        self.field_value_names = names
        self.constitutive_model = constitutive_model
        self.formulation = formulation


class Poisson(BaseEnergyFormPhysics):
    """poisson.py:8-19.  ``f`` is evaluated once at the quadrature points on the host:
    f(x) with x of shape (..., nd) (numpy), returning (...)."""
    field_value_names = ("u",)

    def __init__(self, f: Callable) -> None:
        self.f = f
