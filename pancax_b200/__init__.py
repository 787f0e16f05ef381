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

# pancax's public names for the hot path (same constructors / argument meaning as the reference)
from .bcs import DirichletBC  # noqa: F401,E402
from .bvps import BiaxialLinearRamp, SimpleShearLinearRamp, UniaxialTensionLinearRamp  # noqa: F401,E402
from .constitutive_models import (BlatzKo, BoundedProperty, FixedProperty, Gent, Hencky, NeoHookean,  # noqa: F401,E402
                                  PronySeries, SimpleFeFv, Swanson, WLF)
from .data import FullFieldData, FullFieldDataLoader, GlobalData  # noqa: F401,E402
from .domains import CollocationDataLoader, CollocationDomain, VariationalDomain  # noqa: F401,E402
from .fem import DofManager, Mesh, create_structured_mesh, read_exodus_mesh, write_exodus_mesh  # noqa: F401,E402
from .history_writer import EnsembleHistoryWriter, HistoryWriter  # noqa: F401,E402
from .loss_functions import (CombineLossFunctions, EnergyAndResidualLoss, EnergyLoss,  # noqa: F401,E402
                             EnergyResidualAndReactionLoss, FullFieldDataLoss, StrongFormResidualLoss,
                             UserDefinedLossFunction)
from .networks import MLP, Field, Parameters  # noqa: F401,E402
from .optimizers import LBFGS, Adam  # noqa: F401,E402
from .physics_kernels import (BaseEnergyFormPhysics, BaseStrongFormPhysics, HeatEquation, PlaneStrain, PlaneStress,  # noqa: F401,E402
                              Poisson, SolidMechanics, ThreeDimensional)
from .post_processor import ExodusPostProcessor, PostProcessor  # noqa: F401,E402
from .problems import ForwardProblem, InverseProblem  # noqa: F401,E402
from .trainer import Trainer  # noqa: F401,E402
from .utils import find_data_file, find_mesh_file  # noqa: F401,E402
