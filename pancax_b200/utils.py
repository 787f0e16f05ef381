"""pancax/utils.py:14-55: locate a data / mesh file next to the calling script (or in its ``data`` / ``mesh``
sub-directory), as every reference example does."""
from __future__ import annotations

import inspect
from pathlib import Path


class DataFileNotFoundException(Exception):
    pass


class MeshFileNotFoundException(Exception):
    pass


def _find(name: str, subdir: str, exc, depth: int = 2) -> Path:
    here = Path(inspect.stack()[depth].filename).resolve().parent
    for cand in (here / name, here / subdir / name):
        if cand.is_file():
            print(f"Found {name} in {cand.parent}")
            return cand
    raise exc(f"Could not find {subdir} file {name} in either {here} or {here / subdir}")


def find_data_file(data_file_in: str) -> Path:
    return _find(data_file_in, "data", DataFileNotFoundException)


def find_mesh_file(mesh_file_in: str) -> Path:
    return _find(mesh_file_in, "mesh", MeshFileNotFoundException)
