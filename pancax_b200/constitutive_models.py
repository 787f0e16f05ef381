"""pancax/constitutive_models: property containers of the hyperelastic models on the B200 path.
The energies themselves are evaluated in csrc/pcx_models.cuh."""
from __future__ import annotations

import math
from dataclasses import dataclass, fields
from typing import Union

import numpy as np


class BoundedProperty:
    """pancax/constitutive_models/properties.py:10-38: trainable raw value ``prop_val`` with
    p = prop_min + (prop_max - prop_min) * sigmoid(prop_val).  ``key`` may be a numpy Generator /
    int seed (the initial value is drawn U(prop_min, prop_max) like properties.py:97-103) or a
    float: the initial *physical* value."""

    def __init__(self, prop_min: float, prop_max: float, key=None):
        self.prop_min, self.prop_max = float(prop_min), float(prop_max)
        if prop_min == prop_max:
            self.prop_val = np.zeros(1)
            return
        if isinstance(key, float):
            p_actual = key
        else:
            rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(key)
            p_actual = rng.uniform(prop_min, prop_max)
        y = (p_actual - prop_min) / (prop_max - prop_min)
        self.prop_val = np.array([math.log(y / (1.0 - y))])

    def __call__(self):
        s = 1.0 / (1.0 + math.exp(-float(self.prop_val[0])))
        return self.prop_min + (self.prop_max - self.prop_min) * s

    def __repr__(self):
        return str(self.__call__())


FixedProperty = float
Property = Union[BoundedProperty, FixedProperty]


class HyperelasticModel:
    name = None

    def properties(self):
        return [f.name for f in fields(self)]

    @property
    def num_state_variables(self):
        return 0


@dataclass
class NeoHookean(HyperelasticModel):
    """hyperelasticity/neohookean.py:6-34"""
    bulk_modulus: Property
    shear_modulus: Property
    name = "neohookean"


@dataclass
class Gent(HyperelasticModel):
    """hyperelasticity/gent.py:7-39"""
    bulk_modulus: Property
    shear_modulus: Property
    Jm_parameter: Property
    name = "gent"


@dataclass
class Swanson(HyperelasticModel):
    """hyperelasticity/swanson.py:7-65"""
    bulk_modulus: Property
    A1: Property
    P1: Property
    B1: Property
    Q1: Property
    C1: Property
    R1: Property
    cutoff_strain: float
    name = "swanson"


@dataclass
class Hencky(HyperelasticModel):
    """hyperelasticity/hencky.py:6-18"""
    bulk_modulus: Property
    shear_modulus: Property
    name = "hencky"


@dataclass
class BlatzKo(HyperelasticModel):
    """hyperelasticity/blatz_ko.py:6-24"""
    shear_modulus: Property
    name = "blatz_ko"


class PronySeries:
    """hyperviscoelasticity/base.py:20-31"""

    def __init__(self, moduli, relaxation_times) -> None:
        assert len(moduli) == len(relaxation_times)
        self.moduli = np.asarray(moduli, dtype=np.float64)
        self.relaxation_times = np.asarray(relaxation_times, dtype=np.float64)

    def __len__(self) -> int:
        return len(self.moduli)


class WLF:
    """hyperviscoelasticity/base.py:47-58.  Carried for interface compatibility: SimpleFeFv.energy uses a_T = 1
    (simple_fefv.py:26-27, the shift-factor call is commented out in the reference)."""

    def __init__(self, C1, C2, theta_ref):
        self.C1, self.C2, self.theta_ref = C1, C2, theta_ref

    def shift_factor(self, grad_u, theta, state, dt):
        return 10.0 ** (-self.C1 * (theta - self.theta_ref) / (self.C2 + (theta - self.theta_ref)))


class SimpleFeFv:
    """hyperviscoelasticity/simple_fefv.py:11-108: equilibrium hyperelastic model + Prony branches, one 3x3 viscous
    deformation gradient per branch and quadrature point as state (initial_state = identities)."""
    name = "simple_fefv"

    def __init__(self, eq_model, prony_series: PronySeries, shift_factor_model=None):
        self.eq_model, self.prony_series, self.shift_factor_model = eq_model, prony_series, shift_factor_model

    def num_prony_terms(self) -> int:
        return len(self.prony_series)

    @property
    def num_state_variables(self) -> int:
        return self.num_prony_terms() * 9

    def initial_state(self):
        return np.tile(np.eye(3).ravel(), self.num_prony_terms())

    # bookkeeping of the flat state vector (hyperviscoelasticity/base.py:11-16)
    def extract_scalar(self, state, start_index: int):
        return state[start_index]

    def extract_tensor(self, state, start_index: int):
        return np.asarray(state)[start_index:start_index + 9].reshape((3, 3))

    def properties(self):
        return self.eq_model.properties()

    def __getattr__(self, item):          # the equilibrium model's properties (bulk_modulus, ...)
        if item in ("eq_model", "prony_series", "shift_factor_model"):
            raise AttributeError(item)
        return getattr(self.eq_model, item)
