"""pancax/networks: MLP (mlp.py:49-85), Field (fields.py:10-101), Parameters (parameters.py:62-176).

Weights live in ONE flat float64 tensor in equinox leaf order [W0, b0, W1, b1, ..., Wout(, bout)]
(the layout the C ABI consumes; equinox.nn.Linear stores weight (out, in), bias (out,)), with
per-layer views for inspection and checkpoint compatibility."""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from .constitutive_models import BoundedProperty


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
        else torch.device("cpu")


def _rng(key):
    if isinstance(key, np.random.Generator):
        return key
    return np.random.default_rng(key)


class MLP:
    """Same constructor arguments as pancax.networks.MLP (mlp.py:52-63).  ``activation`` must be
    tanh (the only one the kernels implement); ``key`` is a numpy seed / Generator.  Initialisation
    follows equinox.nn.Linear: U(-1/sqrt(in), 1/sqrt(in)) for weights and biases."""

    def __init__(self, n_inputs: int, n_outputs: int, n_neurons: int, n_layers: int,
                 activation: Callable = "tanh", key=0, use_final_bias: Optional[bool] = False,
                 ensure_positivity: Optional[bool] = False, device=None):
        name = getattr(activation, "__name__", activation)
        if name != "tanh":
            raise NotImplementedError("pancax_b200 implements the tanh MLP only (no fallback)")
        if ensure_positivity:
            raise NotImplementedError("ensure_positivity (softplus head) is not on the B200 path")
        self.n_inputs, self.n_outputs = int(n_inputs), int(n_outputs)
        self.n_neurons, self.n_layers = int(n_neurons), int(n_layers)
        self.use_final_bias = bool(use_final_bias)
        rng = _rng(key)
        sizes = [self.n_inputs] + [self.n_neurons] * self.n_layers + [self.n_outputs]
        chunks, self._slices = [], []
        o = 0
        for i in range(len(sizes) - 1):
            fi, fo = sizes[i], sizes[i + 1]
            lim = 1.0 / math.sqrt(fi)
            W = rng.uniform(-lim, lim, size=(fo, fi))
            chunks.append(W.ravel())
            sl_w = (o, o + fi * fo, (fo, fi))
            o += fi * fo
            last = i == len(sizes) - 2
            if not last or self.use_final_bias:
                chunks.append(rng.uniform(-lim, lim, size=(fo,)))
                sl_b = (o, o + fo, (fo,))
                o += fo
            else:
                sl_b = None
            self._slices.append((sl_w, sl_b))
        self.weights = torch.from_numpy(np.concatenate(chunks)).to(device or _default_device())

    @property
    def num_weights(self):
        return self.weights.numel()

    def layers(self):
        """[(W, b|None)] views into the flat tensor."""
        out = []
        for sw, sb in self._slices:
            W = self.weights[sw[0]:sw[1]].view(*sw[2])
            b = None if sb is None else self.weights[sb[0]:sb[1]]
            out.append((W, b))
        return out

    def set_weights(self, flat):
        self.weights.copy_(torch.as_tensor(np.asarray(flat), dtype=torch.float64))


class Field:
    """fields.py:10-76: normalisation constants from the problem, one MLP (nd+1 -> ndof) and the
    Dirichlet wrapper."""

    def __init__(self, problem, key, ensure_positivity=False, seperate_networks=False, n_layers: int = 3,
                 n_neurons: int = 50, activation="tanh", dirichlet_bc_func: Optional[Callable] = None,
                 network_type=MLP, device=None):
        if seperate_networks:
            raise NotImplementedError("seperate_networks=True is not on the B200 path")
        if network_type is not MLP:
            raise NotImplementedError("only networks.MLP is on the B200 path")
        self.dirichlet_bc_func = dirichlet_bc_func      # None = identity (fields.py:33)
        self.networks = MLP(problem.n_dims + 1, problem.n_dofs, n_neurons, n_layers, activation, key,
                            ensure_positivity=ensure_positivity, device=device)
        self.seperate_networks = False
        self.t_min, self.t_max = float(np.min(problem.times)), float(np.max(problem.times))
        shard = getattr(problem, "shard", None)
        if shard is not None and getattr(shard, "x_mins", None) is not None:
            # one element shard of a larger mesh (parallel.shard_problem): the reference normalises with the bounds
            # of the whole mesh, never the shard's
            self.x_mins, self.x_maxs = shard.x_mins.copy(), shard.x_maxs.copy()
        else:
            self.x_mins = np.min(problem.coords, axis=0)
            self.x_maxs = np.max(problem.coords, axis=0)
        self.normalize_time = not np.allclose(self.t_min, self.t_max)   # fields.py:73-76


class Parameters:
    """parameters.py:62-110: (fields, physics, state).  Trainable BoundedProperty raw values are
    gathered in one small device tensor ``prop_raw`` in property order."""

    def __init__(self, problem, key=0, dirichlet_bc_func: Optional[Callable] = None, network_type=MLP,
                 seperate_networks: Optional[bool] = False, n_layers: int = 3, n_neurons: int = 50,
                 device=None):
        if isinstance(key, np.ndarray) and key.ndim == 2:
            # parameters.py:89-100: an ensemble = the same problem with one independently initialised parameter
            # set per key row (eqx.filter_vmap over keys).  Here: one member ``Parameters`` per row; every member
            # shares the problem's Engine (static mesh / physics / field shape), only the weight vectors differ.
            self.is_ensemble, self.n_ensemble = True, int(key.shape[0])
            self.members = [Parameters(problem, row, dirichlet_bc_func=dirichlet_bc_func, network_type=network_type,
                                       seperate_networks=seperate_networks, n_layers=n_layers, n_neurons=n_neurons,
                                       device=device) for row in key]
            self.physics, self.state = problem.physics, None
            return
        self.is_ensemble, self.n_ensemble = False, 1
        self.fields = Field(problem, key, dirichlet_bc_func=dirichlet_bc_func, network_type=network_type,
                            seperate_networks=seperate_networks, n_layers=n_layers, n_neurons=n_neurons,
                            device=device)
        self.physics = problem.physics
        self.state = None
        dev = self.fields.networks.weights.device
        raw = []
        model = getattr(self.physics, "constitutive_model", None)
        if model is not None:
            for name in model.properties():
                p = getattr(model, name)
                if isinstance(p, BoundedProperty):
                    raw.append(float(p.prop_val[0]))
        self.prop_raw = torch.tensor(raw, dtype=torch.float64, device=dev)

    def __iter__(self):
        return iter((self.fields, self.physics, self.state))

    # ---- ensembles: the stacked leaves of the reference's vmapped pytree, gathered on demand
    @property
    def stacked_weights(self):
        """(n_ensemble, n_weights) -- what ``params.fields.networks`` leaves look like under eqx.filter_vmap."""
        ms = self.members if self.is_ensemble else [self]
        return torch.stack([m.fields.networks.weights for m in ms])

    @property
    def stacked_prop_raw(self):
        ms = self.members if self.is_ensemble else [self]
        return torch.stack([m.prop_raw for m in ms])

    def properties(self):
        """Current physical property values (trainable ones read back from the device)."""
        if self.is_ensemble:
            return [m.properties() for m in self.members]
        out = {}
        model = getattr(self.physics, "constitutive_model", None)
        if model is None:
            return out
        raw = self.prop_raw.detach().cpu().numpy()
        k = 0
        for name in model.properties():
            p = getattr(model, name)
            if isinstance(p, BoundedProperty):
                s = 1.0 / (1.0 + math.exp(-raw[k]))
                out[name] = p.prop_min + (p.prop_max - p.prop_min) * s
                k += 1
            else:
                out[name] = float(p)
        return out

    # parameters.py:158-176: what is trainable by default = everything but the normalisation
    def freeze_physics_normalization_filter(self):
        return dict(fields=True, props=True)

    def freeze_physics_filter(self):
        return dict(fields=True, props=False)

    def freeze_fields_filter(self):
        return dict(fields=False, props=True)

    def _meta(self):
        """What a checkpoint must agree on with the Parameters it is restored onto: the net shape and the Field's
        normalisation constants (leaves of the reference's Parameters pytree, fields.py:66-76; restoring weights onto
        other bounds silently changes the field)."""
        m = self.members[0] if self.is_ensemble else self
        net, fld = m.fields.networks, m.fields
        return dict(net_shape=np.array([net.n_inputs, net.n_outputs, net.n_neurons, net.n_layers,
                                        int(net.use_final_bias)], dtype=np.int64),
                    x_mins=np.asarray(fld.x_mins, dtype=np.float64), x_maxs=np.asarray(fld.x_maxs, dtype=np.float64),
                    t_range=np.array([fld.t_min, fld.t_max], dtype=np.float64),
                    normalize_time=np.array([int(bool(fld.normalize_time))], dtype=np.int64))

    def serialise(self, base_name: str, epoch: int):
        """networks/base.py:80-96 equivalent: {base}_{epoch:07d}.npz with the flat leaves (weights in equinox leaf
        order, raw trainable properties) plus the net shape and normalisation constants.  NOT the reference's
        ``.eqx`` container (eqx.tree_serialise_leaves over the whole pytree); the two formats cannot read each
        other -- the flat ``weights`` vector is the leaves of ``params.fields.networks`` in equinox order, which is
        what an importer on the JAX side needs (integration/kernels_b200.py: flat_to_pytree)."""
        if self.is_ensemble:
            np.savez(f"{base_name}_{str(epoch).zfill(7)}.npz",
                     weights=self.stacked_weights.detach().cpu().numpy(),
                     prop_raw=self.stacked_prop_raw.detach().cpu().numpy(), **self._meta())
            return
        np.savez(f"{base_name}_{str(epoch).zfill(7)}.npz",
                 weights=self.fields.networks.weights.detach().cpu().numpy(),
                 prop_raw=self.prop_raw.detach().cpu().numpy(), **self._meta())

    def deserialise(self, f_name: str):
        d = np.load(f_name)
        if "net_shape" in d.files:       # checkpoints written before the metadata existed carry none: accepted as is
            meta = self._meta()
            if not np.array_equal(d["net_shape"], meta["net_shape"]):
                raise ValueError(f"checkpoint {f_name}: network shape {d['net_shape'].tolist()} != "
                                 f"{meta['net_shape'].tolist()} (n_in, n_out, n_neurons, n_layers, use_final_bias)")
            for k in ("x_mins", "x_maxs", "t_range", "normalize_time"):
                if not np.allclose(d[k], meta[k], rtol=1e-12, atol=0.0):
                    raise ValueError(f"checkpoint {f_name}: Field normalisation constant {k} = {d[k].tolist()} differs "
                                     f"from this problem's {np.asarray(meta[k]).tolist()}; the weights would describe a "
                                     "different field")
        if self.is_ensemble:
            for k, m in enumerate(self.members):
                m.fields.networks.set_weights(d["weights"][k])
                m.prop_raw.copy_(torch.as_tensor(d["prop_raw"][k]))
            return self
        self.fields.networks.set_weights(d["weights"])
        self.prop_raw.copy_(torch.as_tensor(d["prop_raw"]))
        return self
