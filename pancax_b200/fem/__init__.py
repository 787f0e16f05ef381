"""Host-side FE substrate of the hot path: mesh container, Exodus (NetCDF-3) reader/writer,
structured synthetic meshes and the dof bookkeeping the residual losses need.  Mirrors the
names of pancax/fem (Mesh, read_exodus_mesh, DofManager); the parent-element tables and
quadrature rules themselves live in the CUDA library (csrc/pcx_api.cu: build_tab)."""
from .mesh import Mesh, create_structured_mesh, write_exodus_mesh  # noqa: F401
from .read_exodus_mesh import read_exodus_mesh  # noqa: F401
from .dof_manager import DofManager  # noqa: F401
