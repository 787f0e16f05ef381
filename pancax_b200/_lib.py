"""ctypes binding of libpancax_b200.so (include/pcx_abi.h).

This is the binding a pancax maintainer would add next to the XLA-FFI registration (see
INTEGRATION.md).  There is NO fallback: if the shared library is missing or a call fails the
caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpancax_b200.so")

PCX_MAX_PROPS = 8
PCX_MAX_PRONY = 20
PCX_AUX_FIXED = 4
PROF_CLASSES = ["pack_weights", "mlp_forward", "element_force", "assemble", "element_tangent",
                "mlp_backward", "reduce_wgrad", "other"]

ELEM = {"hex8": 0, "quad4": 1, "tri3": 2}
PHYS_POISSON, PHYS_SOLID = 0, 1
FORM = {"plane_strain": 0, "three_dimensional": 1, "plane_stress": 2}
MODEL = {"none": 0, "neohookean": 1, "gent": 2, "swanson": 3, "hencky": 4, "blatz_ko": 5}
LOSS = {"energy": 0, "energy_residual": 1, "energy_residual_reaction": 2}


class MeshDesc(C.Structure):
    _fields_ = [
        ("elem_type", C.c_int32), ("q_order", C.c_int32), ("ndof", C.c_int32),
        ("reaction_dof", C.c_int32),
        ("ne", C.c_int64), ("nn", C.c_int64), ("n_unknown", C.c_int64), ("n_reaction", C.c_int64),
        ("coords", C.c_void_p), ("conns", C.c_void_p), ("unknown_idx", C.c_void_p),
        ("reaction_nodes", C.c_void_p),
        ("nt", C.c_int32), ("deterministic", C.c_int32),
        ("times", C.POINTER(C.c_double)),
    ]


class FieldDesc(C.Structure):
    _fields_ = [
        ("n_in", C.c_int32), ("n_out", C.c_int32), ("width", C.c_int32), ("depth", C.c_int32),
        ("use_final_bias", C.c_int32), ("normalize_time", C.c_int32),
        ("x_min", C.c_double * 3), ("x_max", C.c_double * 3),
        ("t_min", C.c_double), ("t_max", C.c_double),
        ("bc_a", C.c_void_p), ("bc_b", C.c_void_p),
    ]


class PhysicsDesc(C.Structure):
    _fields_ = [
        ("physics", C.c_int32), ("model", C.c_int32), ("formulation", C.c_int32),
        ("nprops", C.c_int32),
        ("value", C.c_double * PCX_MAX_PROPS), ("bounded", C.c_int32 * PCX_MAX_PROPS),
        ("prop_min", C.c_double * PCX_MAX_PROPS), ("prop_max", C.c_double * PCX_MAX_PROPS),
        ("source_q", C.c_void_p),
        ("n_prony", C.c_int32),
        ("prony_moduli", C.c_double * PCX_MAX_PRONY), ("prony_times", C.c_double * PCX_MAX_PRONY),
    ]


class LossDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("energy_weight", C.c_double), ("residual_weight", C.c_double),
        ("reaction_weight", C.c_double),
        ("global_outputs", C.c_void_p),          # DEVICE f64 [nt] (ABI v4)
        ("first_step", C.c_int32),
    ]


class PcxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpancax_b200 error {code}: {msg}")
        self.code = code


_lib = None

# every symbol include/pcx_abi.h declares
SYMBOLS = [
    "pcx_abi_version", "pcx_last_error", "pcx_create", "pcx_set_physics", "pcx_set_field",
    "pcx_destroy", "pcx_num_weights", "pcx_num_trainable_props", "pcx_num_qp", "pcx_nq",
    "pcx_nnpe", "pcx_nd", "pcx_element_geometry", "pcx_field_forward", "pcx_field_backward",
    "pcx_points_forward", "pcx_points_backward", "pcx_energy_force", "pcx_stiffness_action",
    "pcx_element_fields", "pcx_residual_reaction", "pcx_loss_and_grad", "pcx_field_data_loss_and_grad",
    "pcx_num_state_variables", "pcx_get_state", "pcx_set_option", "pcx_set_ownership", "pcx_loss_begin", "pcx_loss_mid", "pcx_loss_finish",
    "pcx_points_jet", "pcx_strong_form_loss_and_grad", "pcx_launch_count", "pcx_adam_step", "pcx_adam_step_sched", "pcx_profile_enable", "pcx_profile_collect",
    "pcx_measure_dmma_peak", "pcx_reserve", "pcx_gather_rows", "pcx_scatter_rows", "pcx_state_step",
]


def load():
    """Load the shared library or raise (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  pancax_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.pcx_abi_version.restype = C.c_int
    lib.pcx_last_error.restype = C.c_char_p
    lib.pcx_create.argtypes = [C.POINTER(vp), C.POINTER(MeshDesc)]
    lib.pcx_set_physics.argtypes = [vp, C.POINTER(PhysicsDesc)]
    lib.pcx_set_field.argtypes = [vp, C.POINTER(FieldDesc)]
    lib.pcx_destroy.argtypes = [vp]
    for name in ("pcx_num_weights", "pcx_num_trainable_props", "pcx_num_qp", "pcx_launch_count"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = i64
    for name in ("pcx_nq", "pcx_nnpe", "pcx_nd"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = i32
    lib.pcx_element_geometry.argtypes = [vp, vp, vp, vp, vp]
    lib.pcx_field_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.pcx_field_backward.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.pcx_points_forward.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    lib.pcx_points_backward.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp]
    lib.pcx_energy_force.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pcx_stiffness_action.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pcx_element_fields.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pcx_residual_reaction.argtypes = [vp, vp, vp, vp]
    lib.pcx_loss_and_grad.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp]
    lib.pcx_points_jet.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.pcx_strong_form_loss_and_grad.argtypes = [vp, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, vp]
    lib.pcx_set_option.argtypes = [vp, i32, i64]
    lib.pcx_num_state_variables.argtypes = [vp]
    lib.pcx_num_state_variables.restype = i64
    lib.pcx_get_state.argtypes = [vp, i32, vp, vp]
    lib.pcx_state_step.argtypes = [vp, vp, vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.pcx_set_ownership.argtypes = [vp, i64, i64]
    lib.pcx_loss_begin.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp]
    lib.pcx_loss_mid.argtypes = [vp, vp, vp, vp]
    lib.pcx_loss_finish.argtypes = [vp, C.POINTER(LossDesc), vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.pcx_field_data_loss_and_grad.argtypes = [vp, vp, vp, vp, vp, vp, i64, dbl, vp, vp, vp]
    lib.pcx_adam_step.argtypes = [vp, vp, vp, vp, i64, i64, dbl, dbl, dbl, dbl, vp]
    lib.pcx_adam_step_sched.argtypes = [vp, vp, vp, vp, i64, vp, i32, dbl, dbl, dbl, dbl, dbl, dbl, vp]
    lib.pcx_profile_enable.argtypes = [vp, C.c_int]
    lib.pcx_profile_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.pcx_measure_dmma_peak.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.pcx_reserve.argtypes = [vp, i32, i64]
    lib.pcx_gather_rows.argtypes = [vp, i64, i64, i32, vp, i64, vp, vp]
    lib.pcx_scatter_rows.argtypes = [vp, i64, i64, i32, vp, i64, vp, i32, vp]
    for name in SYMBOLS:
        if name not in ("pcx_last_error", "pcx_num_weights", "pcx_num_trainable_props",
                        "pcx_num_qp", "pcx_launch_count", "pcx_nq", "pcx_nnpe", "pcx_nd",
                        "pcx_num_state_variables"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PcxError(rc, load().pcx_last_error().decode(errors="replace"))
