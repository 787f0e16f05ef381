"""Host-side data containers of the inverse problems: the interface of pancax/data.py (FullFieldData :11-35,
FullFieldDataLoader :86-105, GlobalData :118-243) on plain numpy arrays -- the library takes them as device buffers
(pcx_field_data_loss_and_grad, pcx_loss_desc.global_outputs)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np

_AXIS = {"x": 0, "y": 1, "z": 2}


def _csv_columns(path: str, keys: Sequence[str]) -> np.ndarray:
    """The named columns of a CSV file as a C-contiguous float64 (rows, len(keys)) array; header names are compared
    after stripping blanks (the DIC / load-frame exports pad them)."""
    import pandas
    table = pandas.read_csv(path)
    table = table.rename(columns={c: c.strip() for c in table.columns})
    return np.ascontiguousarray(table[list(keys)].to_numpy(dtype=np.float64))


class FullFieldData:
    """Rows of (x..., t) -> measured field values; ``n_time_steps`` = number of distinct times in the last input column."""

    def __init__(self, data_file: str, input_keys: List[str], output_keys: List[str]):
        self.inputs = _csv_columns(data_file, input_keys)
        self.outputs = _csv_columns(data_file, output_keys)
        self.n_time_steps = int(np.unique(self.inputs[:, -1]).size)

    def __len__(self):
        return self.inputs.shape[0]


class FullFieldDataLoader:
    """Shuffled full batches of a FullFieldData (the trailing partial batch is dropped, as in the reference)."""

    def __init__(self, data: FullFieldData) -> None:
        self.data = data
        self.indices = np.arange(len(data))

    def __len__(self):
        return len(self.data)

    def dataloader(self, batch_size: int):
        order = np.random.permutation(self.indices)
        for b in range(len(self) // batch_size):
            rows = order[b * batch_size:(b + 1) * batch_size]
            yield self.data.inputs[rows], self.data.outputs[rows]


class GlobalDataTimesNotUniqueException(Exception):
    pass


class GlobalDataTimesNotStrictlyIncreasingException(Exception):
    pass


class GlobalData:
    """Load-frame record (time, displacement, reaction force) resampled onto the problem's time steps, and the node set
    the reaction is summed over."""

    def __init__(self, data_file: str, times_key: str, disp_key: str, force_key: str, mesh_file,
                 nset_id: int, reaction_dof: Union[int, str], n_time_steps: int,
                 plotting: Optional[bool] = False, interpolate: Optional[bool] = False):
        from .fem import Mesh, read_exodus_mesh
        t_in, d_in, f_in = _csv_columns(data_file, (times_key, disp_key, force_key)).T
        times = np.linspace(t_in.min(), t_in.max(), n_time_steps) if interpolate else t_in
        steps = np.diff(times)
        if np.unique(times).size != times.size:
            raise GlobalDataTimesNotUniqueException()
        if np.any(steps <= 0.0):
            raise GlobalDataTimesNotStrictlyIncreasingException()
        mesh = mesh_file if isinstance(mesh_file, Mesh) else read_exodus_mesh(mesh_file)
        # the reference reads the Exodus variable node_ns{nset_id} (1-based ids): the nset_id-th node set of the file
        nset_name = list(mesh.nodeSets)[nset_id - 1]
        if reaction_dof not in _AXIS:
            raise ValueError("reaction_dof needs to be either x or y.")
        self._fill(times, np.interp(times, t_in, f_in), mesh.nodeSets[nset_name], _AXIS[reaction_dof],
                   np.interp(times, t_in, d_in))

    def _fill(self, times, outputs, reaction_nodes, reaction_dof, displacements):
        self.times = np.asarray(times, dtype=np.float64)
        self.outputs = np.asarray(outputs, dtype=np.float64)
        self.displacements = np.asarray(displacements, dtype=np.float64)
        self.reaction_nodes = np.asarray(reaction_nodes, dtype=np.int32)
        self.reaction_dof = int(reaction_dof)
        self.n_nodes = int(self.reaction_nodes.size)
        self.n_time_steps = int(self.times.size)

    @classmethod
    def from_arrays(cls, times, outputs, reaction_nodes, reaction_dof, displacements=None):
        """Synthetic global data (no CSV): same fields as the CSV constructor."""
        self = cls.__new__(cls)
        times = np.asarray(times, dtype=np.float64)
        self._fill(times, outputs, reaction_nodes, reaction_dof, np.zeros_like(times) if displacements is None else displacements)
        return self
