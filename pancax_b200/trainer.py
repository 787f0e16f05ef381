"""A working equivalent of pancax/trainer.py:9-87 (the reference's own is stale: it imports the
non-existent pancax.logging): epoch loop with loss history (CSV, like
pancax/history_writer.py:44-79) and checkpoints.  Exodus result output is out of scope (SURVEY
8f N2)."""
from __future__ import annotations

import os
from pathlib import Path


class Trainer:
    def __init__(self, problem, opt, checkpoint_base="checkpoint", history_file="history.csv",
                 log_file="pinn.log", output_file_base="output", output_node_variables=(),
                 output_element_variables=(), log_every=1000, output_every=10000,
                 serialise_every=10000, workdir=None) -> None:
        workdir = os.getcwd() if workdir is None else workdir
        for d in ("checkpoint", "history", "log", "results"):
            os.makedirs(Path(workdir) / d, exist_ok=True)
        self.problem, self.opt = problem, opt
        self.checkpoint_base = str(Path(workdir) / "checkpoint" / checkpoint_base)
        self.history_path = Path(workdir) / "history" / history_file
        self.log_path = Path(workdir) / "log" / log_file
        self.log_every, self.serialise_every = log_every, serialise_every
        self.epoch = 0
        self._header_written = False

    def init(self, params):
        return self.opt.init(params)

    def serialise(self, params):
        if self.epoch % self.serialise_every == 0:
            params.serialise(self.checkpoint_base, self.epoch)

    def _write_history(self, loss):
        if self.epoch % self.log_every != 0:
            return
        val, aux = loss if isinstance(loss, tuple) else (loss, {})
        row = {"epoch": self.epoch, "loss": float(val)}
        for k, v in aux.items():
            if hasattr(v, "numel") and v.numel() == 1:
                row[k] = float(v)
        with open(self.history_path, "a") as f:
            if not self._header_written:
                f.write(",".join(row.keys()) + "\n")
                self._header_written = True
            f.write(",".join(str(v) for v in row.values()) + "\n")
        with open(self.log_path, "a") as f:
            f.write(f"Epoch {self.epoch}: " + ", ".join(f"{k}={v}" for k, v in row.items()) + "\n")

    def step(self, params, opt_st, *args):
        params, opt_st, loss = self.opt.step(params, opt_st, self.problem, *args)
        self._write_history(loss)
        self.serialise(params)
        self.epoch += 1
        return params, opt_st

    def train(self, params, n_epochs, *args, cuda_graph: bool = False):
        """``cuda_graph=True``: the whole step (loss + gradient + Adam update, schedule and counter on the
        device) is captured once and replayed ``n_epochs`` times -- for launch-bound small problems; the
        loss history is still written every ``log_every`` epochs."""
        opt_st = self.init(params)
        if not cuda_graph:
            for _ in range(n_epochs):
                params, opt_st = self.step(params, opt_st, *args)
            return params
        graph, loss = self.opt.capture(params, opt_st, self.problem, *args)
        for _ in range(n_epochs):
            graph.replay()
            opt_st.epoch += 1
            opt_st.opt_state["count"] += 1
            self._write_history(loss)
            self.serialise(params)
            self.epoch += 1
        return params
