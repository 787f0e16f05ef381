"""pancax_b200 -- B200-native (sm_100a) implementation of pancax's training hot path:
deep-energy / FE-residual PINN loss + parameter gradient over every quadrature point of a
finite-element mesh, behind pancax's own Python interface names.

The compute path is ``lib/libpancax_b200.so`` (hand-written CUDA, C ABI in ``include/pcx_abi.h``).
There is no CPU fallback: importing the package works anywhere, using it needs the built
library and a B200.
"""
from . import _lib  # noqa: F401
from .engine import Engine, tabulate_bc, mlp_num_weights  # noqa: F401

__version__ = "0.1.0"
