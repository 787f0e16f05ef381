"""pancax/post_processor.py (ExodusPostProcessor :141-354, PostProcessor front end): nodal and
element (per-quadrature-point) variables of a trained model written to an Exodus II results file.

Same call sequence and variable names as the reference --

    pp = PostProcessor(mesh_file)
    pp.init(params, problem, "out.e", node_variables=["field_values", "internal_force"],
            element_variables=["deformation_gradient", "I1_bar", "pk1_stress"])
    pp.write_outputs(params, problem)
    pp.close()

-- but every number comes from the CUDA library (Engine.field_forward / energy_force / element_fields)
instead of jitted vmaps, and the file is written with scipy's NetCDF-3 writer (netCDF4 is not a
dependency).  Variable layout as post_processor.py:263-354: nodal ``vals_nod_var{k}`` (time_step,
num_nodes) named like ``physics.field_value_names`` / ``internal_force_x..``; element
``vals_elem_var{k}eb1`` (time_step, num_elem) named ``{component}_{q+1}``, components outer, quadrature
points inner (``F_xx_1 .. F_xx_nq, F_xy_1 ..``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .fem import Mesh, read_exodus_mesh
from .fem.mesh import EXODUS_NAME

_TENSOR = ("xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz")


def output_names(var_name: str, var_type: str):
    """physics_kernels/base.py:12-23"""
    if var_type == "full_tensor":
        return tuple(f"{var_name}_{c}" for c in _TENSOR)
    if var_type == "scalar":
        return (var_name,)
    raise ValueError(f"Unsupported variable type {var_type}")


def variable_names(problem) -> Dict[str, tuple]:
    """``physics.var_name_to_method[var]["names"]`` of the reference (base.py:186-196,
    solid_mechanics.py:115-170)."""
    names = {"field_values": tuple(problem.physics.field_value_names)}
    if hasattr(problem.physics, "constitutive_model"):
        nd = problem.physics.formulation.n_dimensions
        names["internal_force"] = ("internal_force_x", "internal_force_y") + (("internal_force_z",) if nd > 2 else ())
        names["deformation_gradient"] = output_names("F", "full_tensor")
        names["I1_bar"] = output_names("I1_bar", "scalar")
        names["pk1_stress"] = output_names("P", "full_tensor")
        ns = getattr(problem.physics.constitutive_model, "num_state_variables", 0)
        if ns:   # solid_mechanics.py:157-168
            names["state_variables"] = tuple(f"state_variable_{n + 1}" for n in range(ns))
    return names


def write_exodus_results(path: str, mesh: Mesh, times, node_vars: Dict[str, np.ndarray],
                         elem_vars: Dict[str, np.ndarray], title: str = "pancax_b200 results") -> None:
    """Mesh + time-dependent variables as Exodus II (NetCDF-3 64-bit offset).  ``node_vars[name]`` has
    shape (nt, num_nodes), ``elem_vars[name]`` (nt, num_elem); names are written in insertion order."""
    from scipy.io import netcdf_file
    nn, nd = mesh.coords.shape
    ne, nnpe = mesh.conns.shape
    nt = len(times)
    ns_names = sorted(mesh.nodeSets.keys())
    with netcdf_file(path, "w", version=2) as ds:
        ds.title = title
        ds.api_version = np.float32(8.25)
        ds.version = np.float32(8.25)
        ds.floating_point_word_size = np.int32(8)
        ds.file_size = np.int32(1)
        ds.maximum_name_length = np.int32(32)
        ds.int64_status = np.int32(0)
        for k, v in (("time_step", None), ("len_name", 256), ("num_dim", nd), ("num_nodes", nn), ("num_elem", ne),
                     ("num_el_blk", 1), ("num_el_in_blk1", ne), ("num_nod_per_el1", nnpe), ("four", 4),
                     ("len_string", 33)):
            ds.createDimension(k, v)
        if ns_names:
            ds.createDimension("num_node_sets", len(ns_names))
        if node_vars:
            ds.createDimension("num_nod_var", len(node_vars))
        if elem_vars:
            ds.createDimension("num_elem_var", len(elem_vars))

        def put_names(var, dim, strs):
            vv = ds.createVariable(var, "c", (dim, "len_name"))
            arr = np.zeros((len(strs), 256), dtype="S1")
            for i, s in enumerate(strs):
                arr[i, :len(s)] = list(s)
            vv[:] = arr

        for j, c in enumerate("xyz"[:nd]):
            v = ds.createVariable("coord" + c, "d", ("num_nodes",))
            v[:] = mesh.coords[:, j]
        v = ds.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
        v[:] = mesh.conns + 1
        v.elem_type = EXODUS_NAME[mesh.elem_type]
        put_names("eb_names", "num_el_blk", ["block_1"])
        put_names("coor_names", "num_dim", list("xyz"[:nd]))
        for nm, val in (("eb_status", [1]), ("eb_prop1", [1])):
            vv = ds.createVariable(nm, "i", ("num_el_blk",))
            vv[:] = np.array(val, dtype=np.int32)
        ds.variables["eb_prop1"].name = "ID"
        if ns_names:
            put_names("ns_names", "num_node_sets", ns_names)
            vv = ds.createVariable("ns_status", "i", ("num_node_sets",))
            vv[:] = np.ones(len(ns_names), dtype=np.int32)
            vv = ds.createVariable("ns_prop1", "i", ("num_node_sets",))
            vv[:] = np.arange(1, len(ns_names) + 1, dtype=np.int32)
            vv.name = "ID"
            for i, nm in enumerate(ns_names):
                ds.createDimension(f"num_nod_ns{i + 1}", len(mesh.nodeSets[nm]))
                vv = ds.createVariable(f"node_ns{i + 1}", "i", (f"num_nod_ns{i + 1}",))
                vv[:] = np.asarray(mesh.nodeSets[nm], dtype=np.int32) + 1
        if node_vars:
            put_names("name_nod_var", "num_nod_var", list(node_vars.keys()))
        if elem_vars:
            put_names("name_elem_var", "num_elem_var", list(elem_vars.keys()))
            tab = ds.createVariable("elem_var_tab", "i", ("num_el_blk", "num_elem_var"))
            tab[:] = np.ones((1, len(elem_vars)), dtype=np.int32)
        # record variables (unlimited time_step dimension) last
        tw = ds.createVariable("time_whole", "d", ("time_step",))
        nod = [ds.createVariable(f"vals_nod_var{k + 1}", "d", ("time_step", "num_nodes")) for k in range(len(node_vars))]
        elv = [ds.createVariable(f"vals_elem_var{k + 1}eb1", "d", ("time_step", "num_elem")) for k in range(len(elem_vars))]
        nvals, evals = list(node_vars.values()), list(elem_vars.values())
        for n in range(nt):
            tw[n] = float(times[n])
            for k, var in enumerate(nod):
                var[n, :] = np.asarray(nvals[k][n], dtype=np.float64)
            for k, var in enumerate(elv):
                var[n, :] = np.asarray(evals[k][n], dtype=np.float64)


def read_exodus_results(path: str):
    """(times, {node variable: (nt, nn)}, {element variable: (nt, ne)}) of a file written above (tests)."""
    from scipy.io import netcdf_file

    def names(ds, var):
        return ["".join(ch.decode() for ch in row if ch not in (b"", b"\x00")).strip() for row in ds.variables[var][:]]
    with netcdf_file(path, "r", mmap=False) as ds:
        times = np.array(ds.variables["time_whole"][:], dtype=np.float64)
        nod, el = {}, {}
        if "name_nod_var" in ds.variables:
            for k, nm in enumerate(names(ds, "name_nod_var")):
                nod[nm] = np.array(ds.variables[f"vals_nod_var{k + 1}"][:], dtype=np.float64)
        if "name_elem_var" in ds.variables:
            for k, nm in enumerate(names(ds, "name_elem_var")):
                el[nm] = np.array(ds.variables[f"vals_elem_var{k + 1}eb1"][:], dtype=np.float64)
    return times, nod, el


class BasePostProcessor:
    mesh_file = None
    node_variables: List[str] = None
    element_variables: List[str] = None
    output_file: str = None

    def check_variable_names(self, problem, variables) -> None:
        """post_processor.py:24-35"""
        known = variable_names(problem)
        for var in variables:
            if var not in known:
                msg = f"Unsupported variable requested for output {var}.\nSupported variables include:\n"
                msg += "".join(f"  {v}\n" for v in known)
                raise ValueError(msg)

    def get_node_variable_number(self, problem, variables) -> int:
        return sum(len(variable_names(problem)[v]) for v in variables)

    def get_element_variable_number(self, problem, variables) -> int:
        n = sum(len(variable_names(problem)[v]) for v in variables)
        return n * self._nq(problem) if n else 0

    @staticmethod
    def _nq(problem):
        return {("hex8", 1): 1, ("hex8", 2): 8, ("quad4", 1): 1, ("quad4", 2): 4, ("quad4", 3): 9,
                ("tri3", 1): 1, ("tri3", 2): 3}[(problem.domain.mesh.elem_type, problem.domain.q_order)]


class ExodusPostProcessor(BasePostProcessor):
    def __init__(self, mesh_file=None) -> None:
        self.mesh_file = mesh_file
        self.output_file = None

    def close(self) -> None:
        pass

    def init(self, params, problem, output_file: str, node_variables: List[str], element_variables: List[str]) -> None:
        if getattr(params, "is_ensemble", False):
            print("WARNING: post-processing is currently unsupported for ensembles")
            return
        self.output_file = output_file
        self.check_variable_names(problem, node_variables)
        self.check_variable_names(problem, element_variables)
        self.node_variables, self.element_variables = list(node_variables), list(element_variables)

    def collect(self, params, problem):
        """All requested variables for every time step as host arrays (the file-independent half of
        ``write_outputs``): ({nodal name: (nt, nn)}, {element name: (nt, ne)})."""
        from .loss_functions import get_engine
        eng = get_engine(problem, params)
        names = variable_names(problem)
        w = params.fields.networks.weights
        pr = params.prop_raw if eng.n_train else None
        nt, nq = len(problem.times), eng.nq
        node_vars = {n: np.empty((nt, eng.nn)) for v in self.node_variables for n in names[v]}
        elem_vars = {}
        for v in self.element_variables:
            for comp in names[v]:
                for q in range(nq):
                    elem_vars[f"{comp}_{q + 1}"] = np.empty((nt, eng.ne))
        want = set(self.element_variables)
        # models with state variables: the reference threads the state through the steps (post_processor.py:470-490:
        # dt = 0 at step 0, then times[n] - times[n-1]) and evaluates internal_force / pk1_stress / state_variables
        # from the step's OLD state (:262-266,273-279,312-316) -- one pcx_state_step per time step here
        ns = int(eng.lib.pcx_num_state_variables(eng.h))
        stepwise = ns > 0
        state_old = eng.initial_state() if stepwise else None
        for n in range(nt):
            us = eng.field_forward(w, n)
            f_state = P_state = None
            if stepwise:
                dt = 0.0 if n == 0 else float(problem.times[n] - problem.times[n - 1])
                _, state_new, f_state, P_state = eng.state_step(us, state_old, dt, pr,
                                                                want_force="internal_force" in self.node_variables,
                                                                want_P="pk1_stress" in want)
            if "field_values" in self.node_variables:
                u = us.cpu().numpy()
                for i, nm in enumerate(names["field_values"]):
                    node_vars[nm][n] = u[:, i]
            if "internal_force" in self.node_variables:
                f = f_state if stepwise else eng.energy_force(us, pr)[1]
                f = f.cpu().numpy()
                for i, nm in enumerate(names["internal_force"]):
                    node_vars[nm][n] = f[:, i]
            if "state_variables" in want:
                z = state_new.cpu().numpy()                                          # (ne, nq, ns)
                for c, comp in enumerate(names["state_variables"]):
                    for q in range(nq):
                        elem_vars[f"{comp}_{q + 1}"][n] = z[:, q, c]
            if want - {"state_variables"}:
                F, I1b, P = eng.element_fields(us, pr, want_F="deformation_gradient" in want,
                                               want_I1bar="I1_bar" in want, want_P="pk1_stress" in want and not stepwise)
                if stepwise:
                    P = P_state
                for v, arr in (("deformation_gradient", F), ("I1_bar", I1b), ("pk1_stress", P)):
                    if v not in want:
                        continue
                    a = arr.cpu().numpy().reshape(eng.ne, nq, -1)        # (ne, nq, ncomp)
                    for c, comp in enumerate(names[v]):
                        for q in range(nq):
                            elem_vars[f"{comp}_{q + 1}"][n] = a[:, q, c]
            if stepwise:
                state_old = state_new
        return node_vars, elem_vars

    def write_outputs(self, params, problem) -> None:
        node_vars, elem_vars = self.collect(params, problem)
        write_exodus_results(self.output_file, problem.domain.mesh, problem.times, node_vars, elem_vars)


class PostProcessor:
    """post_processor.py front end: chooses the writer from ``mesh_file_type`` ("exodus" only here; the
    VTK writer is outside the scope of this build)."""

    def __init__(self, mesh_file=None, mesh_file_type: str = "exodus") -> None:
        if mesh_file_type != "exodus":
            raise ValueError(f"Unsupported mesh_file_type {mesh_file_type} (exodus only)")
        self.pp = ExodusPostProcessor(mesh_file)

    def close(self):
        self.pp.close()

    def init(self, params, problem, output_file, node_variables, element_variables=()):
        self.pp.init(params, problem, output_file, list(node_variables), list(element_variables))

    def write_outputs(self, params, problem):
        self.pp.write_outputs(params, problem)
