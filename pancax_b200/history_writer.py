"""Loss-history logging of the training loops, with the call interface of pancax/history_writer.py (HistoryWriter
:44-105, EnsembleHistoryWriter :108-151): ``write_history(loss, epoch)`` takes ``loss = (value, aux)`` as the optimisers
return it and keeps, every ``log_every`` epochs, the columns ``epoch``, ``loss``, every scalar aux entry and
``property_k`` / ``dproperty_k`` for the ``props`` / ``dprops`` entries; every ``write_every`` epochs the table goes to a
CSV file.  Values may be torch tensors on the GPU (one ``.item()`` read per logged value); vector-valued aux entries
(the reactions per time step) are not columns."""
from __future__ import annotations

import csv
from pathlib import Path

_PROP_COLUMNS = {"props": "property_", "dprops": "dproperty_"}


def _number(v) -> float:
    return float(v.item()) if hasattr(v, "item") else float(v)


def _is_scalar(v) -> bool:
    return not hasattr(v, "numel") or v.numel() == 1


def _columns(aux, member=None):
    """(column name, value) pairs of one aux dict; ``member`` selects an ensemble member's slice of stacked entries."""
    for key, val in aux.items():
        if member is not None:
            val = val[member]
        if key in _PROP_COLUMNS:
            items = val.values() if isinstance(val, dict) else val
            for k, p in enumerate(items):
                yield f"{_PROP_COLUMNS[key]}{k}", _number(p)
        elif _is_scalar(val):
            yield key, _number(val)


class HistoryWriter:
    def __init__(self, history_file: str, log_every: int, write_every: int) -> None:
        self.history_file = Path(history_file)
        self.log_every, self.write_every = int(log_every), int(write_every)
        self.data_dict = {}                  # column name -> values, in first-seen order

    # -- the reference's fine-grained calls
    def write_data(self, key: str, val) -> None:
        self.data_dict.setdefault(key, []).append(val)

    def write_aux_values(self, loss, member=None) -> None:
        for name, value in _columns(loss[1], member):
            self.write_data(name, value)

    def write_epoch(self, epoch) -> None:
        self.write_data("epoch", int(epoch))

    def write_loss(self, loss, member=None) -> None:
        self.write_data("loss", _number(loss[0] if member is None else loss[0][member]))

    # -- one call per epoch
    def write_history(self, loss, epoch: int) -> None:
        if epoch % self.log_every == 0:
            self.log(loss, epoch)
        if epoch % self.write_every == 0:
            self.to_csv()

    def log(self, loss, epoch: int, member=None) -> None:
        self.write_aux_values(loss, member)
        self.write_epoch(epoch)
        self.write_loss(loss, member)

    def to_csv(self) -> None:
        names = list(self.data_dict)
        depth = max((len(col) for col in self.data_dict.values()), default=0)
        self.history_file.parent.mkdir(parents=True, exist_ok=True)
        with open(self.history_file, "w", newline="") as f:
            out = csv.writer(f)
            out.writerow(names)
            out.writerows([self.data_dict[c][i] if i < len(self.data_dict[c]) else "" for c in names] for i in range(depth))


class EnsembleHistoryWriter:
    """One CSV per ensemble member (``{base_name}_{n}.csv``); ``loss`` = (stacked values, stacked aux) as
    ``Adam.ensemble_step`` returns them."""

    def __init__(self, base_name: str, n_pinns: int, log_every: int, write_every: int) -> None:
        self.log_every, self.write_every = int(log_every), int(write_every)
        self.history_writers = [HistoryWriter(f"{base_name}_{n}.csv", log_every, write_every) for n in range(n_pinns)]

    def write_history(self, loss, epoch: int) -> None:
        if epoch % self.log_every == 0:
            for n, member in enumerate(self.history_writers):
                member.log(loss, epoch, member=n)
        if epoch % self.write_every == 0:
            self.to_csv()

    def to_csv(self) -> None:
        for member in self.history_writers:
            member.to_csv()
