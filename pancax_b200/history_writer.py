"""Loss-history logging of the training loops: the interface of pancax/history_writer.py (HistoryWriter :44-105,
EnsembleHistoryWriter :108-151) -- ``write_history(loss, epoch)`` with ``loss = (value, aux)`` as the optimisers return
it, columns ``epoch``, ``loss``, every scalar aux entry, ``property_k`` / ``dproperty_k`` for the ``props`` / ``dprops``
entries -- on plain Python lists and the csv module (values may be torch tensors on the GPU: ``float()`` reads them)."""
from __future__ import annotations

import csv
from pathlib import Path
from typing import List


def _scalar(v):
    return float(v.item()) if hasattr(v, "item") else float(v)


def _prop_values(val):
    """``props`` as the losses report them: a dict name -> value (CombineLossFunctions(with_props=True)) or a sequence."""
    return list(val.values()) if isinstance(val, dict) else list(val)


class BaseHistoryWriter:
    def __init__(self, log_every: int, write_every: int) -> None:
        self.log_every, self.write_every = int(log_every), int(write_every)

    def write_history(self, loss, epoch: int) -> None:
        if epoch % self.log_every == 0:
            self.write_aux_values(loss)
            self.write_epoch(epoch)
            self.write_loss(loss)
        if epoch % self.write_every == 0:
            self.to_csv()


class HistoryWriter(BaseHistoryWriter):
    def __init__(self, history_file: str, log_every: int, write_every: int) -> None:
        super().__init__(log_every, write_every)
        self.history_file = Path(history_file)
        self.data_dict = {}

    def write_data(self, key: str, val: float) -> None:
        self.data_dict.setdefault(key, []).append(val)

    def write_aux_values(self, loss) -> None:
        for key, val in loss[1].items():
            if key in ("props", "dprops"):
                for k, p in enumerate(_prop_values(val)):
                    self.write_data(("property_" if key == "props" else "dproperty_") + str(k), _scalar(p))
            elif not hasattr(val, "numel") or val.numel() == 1:
                self.write_data(key, _scalar(val))          # vector-valued aux entries (reactions per step) are not columns

    def write_epoch(self, epoch) -> None:
        self.write_data("epoch", int(epoch))

    def write_loss(self, loss) -> None:
        self.write_data("loss", _scalar(loss[0]))

    def to_csv(self) -> None:
        keys = list(self.data_dict)
        n = max((len(v) for v in self.data_dict.values()), default=0)
        self.history_file.parent.mkdir(parents=True, exist_ok=True)
        with open(self.history_file, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keys)
            for i in range(n):
                w.writerow([self.data_dict[k][i] if i < len(self.data_dict[k]) else "" for k in keys])


class EnsembleHistoryWriter(BaseHistoryWriter):
    """One CSV per ensemble member (``{base_name}_{n}.csv``); ``loss`` = (stacked values, stacked aux) as
    ``Adam.ensemble_step`` returns them."""

    def __init__(self, base_name: str, n_pinns: int, log_every: int, write_every: int) -> None:
        super().__init__(log_every, write_every)
        self.history_writers: List[HistoryWriter] = [HistoryWriter(f"{base_name}_{n}.csv", log_every, write_every)
                                                     for n in range(n_pinns)]

    def write_aux_values(self, loss) -> None:
        for key, val in loss[1].items():
            for n, writer in enumerate(self.history_writers):
                if key == "props":
                    for k, p in enumerate(_prop_values(val[n])):
                        writer.write_data(f"property_{k}", _scalar(p))
                else:
                    v = val[n]
                    if not hasattr(v, "numel") or v.numel() == 1:
                        writer.write_data(key, _scalar(v))

    def write_epoch(self, epoch) -> None:
        for writer in self.history_writers:
            writer.write_epoch(epoch)

    def write_loss(self, loss) -> None:
        for n, writer in enumerate(self.history_writers):
            writer.write_data("loss", _scalar(loss[0][n]))

    def to_csv(self) -> None:
        for writer in self.history_writers:
            writer.to_csv()
