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
    """solid_mechanics.py:32-71: grad_u_33 is the per-quadrature-point root of sigma_33 = 0
    (math/scalar_root_find.py), found and implicitly differentiated inside the element kernels
    (PCX_FORM_PLANE_STRESS, csrc/pcx_models.cuh: plane_stress_solve / plane_stress_tangent)."""
    n_dimensions = 2
    name = "plane_stress"


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


class BaseStrongFormPhysics:
    """physics_kernels/base.py:430-440.  The strong-form residual of the physics on this path is linear in the
    jet of the field: r = c_t u_t - c_lap lap(u) - source(x)."""
    field_value_names: tuple = ()

    @property
    def n_dofs(self):
        return len(self.field_value_names)

    def strong_form_coefficients(self):
        raise NotImplementedError

    def strong_form_source(self, x):
        return None


class Poisson(BaseEnergyFormPhysics, BaseStrongFormPhysics):
    """poisson.py:8-29.  ``f`` is evaluated once at the quadrature / collocation points on the host:
    f(x) with x of shape (..., nd) (numpy), returning (...).  Strong form: -lap(u) - f(x)."""
    field_value_names = ("u",)

    def __init__(self, f: Callable) -> None:
        self.f = f

    def strong_form_coefficients(self):
        return 0.0, 1.0                      # (c_t, c_lap): poisson.py:26-29

    def strong_form_source(self, x):
        return self.f(x)


class HeatEquation(BaseStrongFormPhysics):
    """heat_equation.py:6-17: rho c du/dt - k lap(u).  The reference reads (rho, c, k) from ``params[1]``
    (a plain tuple next to the field); here they are constructor arguments."""
    field_value_names = ("temperature",)

    def __init__(self, density: float = 1.0, specific_heat: float = 1.0, conductivity: float = 1.0) -> None:
        self.density, self.specific_heat, self.conductivity = float(density), float(specific_heat), float(conductivity)

    def strong_form_coefficients(self):
        return self.density * self.specific_heat, self.conductivity
