"""pancax/data.py: FullFieldData (:11-35), FullFieldDataLoader (:86-105), GlobalData (:118-243)."""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np


class FullFieldData:
    def __init__(self, data_file: str, input_keys: List[str], output_keys: List[str]):
        import pandas
        df = pandas.read_csv(data_file)
        df.columns = df.columns.str.strip()
        self.inputs = np.ascontiguousarray(df[input_keys].values, dtype=np.float64)
        self.outputs = np.ascontiguousarray(df[output_keys].values, dtype=np.float64)
        self.n_time_steps = len(np.unique(self.inputs[:, -1]))

    def __len__(self):
        return self.inputs.shape[0]


class FullFieldDataLoader:
    def __init__(self, data: FullFieldData) -> None:
        self.data = data
        self.indices = np.arange(len(self.data))

    def __len__(self):
        return len(self.data)

    def dataloader(self, batch_size: int):
        perm = np.random.permutation(self.indices)
        start, end = 0, batch_size
        while end <= len(self):
            batch_perm = perm[start:end]
            yield self.data.inputs[batch_perm], self.data.outputs[batch_perm]
            start = end
            end = start + batch_size


class GlobalDataTimesNotUniqueException(Exception):
    pass


class GlobalDataTimesNotStrictlyIncreasingException(Exception):
    pass


class GlobalData:
    def __init__(self, data_file: str, times_key: str, disp_key: str, force_key: str, mesh_file,
                 nset_id: int, reaction_dof: Union[int, str], n_time_steps: int,
                 plotting: Optional[bool] = False, interpolate: Optional[bool] = False):
        import pandas
        from .fem import Mesh, read_exodus_mesh
        df = pandas.read_csv(data_file)
        df.columns = df.columns.str.strip()
        times_in = df[times_key].values
        disps_in = df[disp_key].values
        forces_in = df[force_key].values
        times = np.linspace(np.min(times_in), np.max(times_in), n_time_steps) if interpolate else times_in
        if len(times) != len(set(times)):
            raise GlobalDataTimesNotUniqueException()
        for i in range(len(times) - 1):
            if times[i] >= times[i + 1]:
                raise GlobalDataTimesNotStrictlyIncreasingException()
        mesh = mesh_file if isinstance(mesh_file, Mesh) else read_exodus_mesh(mesh_file)
        # data.py:222-225: node_ns{nset_id} - 1  (i.e. the nset_id-th node set of the file)
        names = list(mesh.nodeSets.keys())
        self.reaction_nodes = np.asarray(mesh.nodeSets[names[nset_id - 1]], dtype=np.int32)
        if reaction_dof == "x":
            reaction_dof = 0
        elif reaction_dof == "y":
            reaction_dof = 1
        elif reaction_dof == "z":
            reaction_dof = 2
        else:
            raise ValueError("reaction_dof needs to be either x or y.")
        self.times = np.asarray(times, dtype=np.float64)
        self.displacements = np.interp(times, times_in, disps_in)
        self.outputs = np.interp(times, times_in, forces_in)
        self.n_nodes = len(self.reaction_nodes)
        self.n_time_steps = len(times)
        self.reaction_dof = reaction_dof

    @classmethod
    def from_arrays(cls, times, outputs, reaction_nodes, reaction_dof, displacements=None):
        """Synthetic global data (no CSV): same fields as the CSV constructor."""
        self = cls.__new__(cls)
        self.times = np.asarray(times, dtype=np.float64)
        self.outputs = np.asarray(outputs, dtype=np.float64)
        self.displacements = np.zeros_like(self.times) if displacements is None else np.asarray(displacements)
        self.reaction_nodes = np.asarray(reaction_nodes, dtype=np.int32)
        self.reaction_dof = int(reaction_dof)
        self.n_nodes = len(self.reaction_nodes)
        self.n_time_steps = len(self.times)
        return self
