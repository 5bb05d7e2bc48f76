"""
ctypes binding of libdiffert2d_b200.so — the C ABI declared in include/differt2d_b200.h.

There is NO fallback: if the shared object is missing or a CUDA call fails, the error is raised.
PyTorch is used by the callers only for device memory and streams; no torch type crosses the ABI.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# D2D_B200_LIB selects another build of the SAME library (diagnostic builds with cull counters); never a fallback
LIB_PATH = os.environ.get("D2D_B200_LIB") or os.path.join(HERE, "_lib", "libdiffert2d_b200.so")

MAX_ORDER = 4
MAX_OBJECTS = 1024

KIND_WALL, KIND_RIS, KIND_VERTEX = 0, 1, 2
GRID_RECEIVERS, GRID_TRANSMITTERS = 0, 1
METHOD_IMAGE, METHOD_FERMAT, METHOD_MINPATH = 0, 1, 2
MODE_HARD, MODE_HARD_SIGMOID, MODE_SIGMOID = 0, 1, 2
GRAD_CLEAN, GRAD_NAN_PARITY = 0, 1
FUN_RECEIVED_POWER, FUN_LENGTH_SQUARED = 0, 1
OPT_ADAM, OPT_SGD, OPT_NEWTON = 0, 1, 2


class D2DError(RuntimeError):
    pass


class D2DProblem(C.Structure):
    _fields_ = [
        ("n_objects", C.c_int32),
        ("objects_xys", C.c_void_p),
        ("object_kinds", C.c_void_p),
        ("object_phis", C.c_void_p),
        ("n_fixed", C.c_int32),
        ("fixed_xy", C.c_void_p),
        ("n_grid", C.c_int64),
        ("grid_xy", C.c_void_p),
        ("grid_role", C.c_int32),
        ("grid_cols", C.c_int32),
        ("min_order", C.c_int32),
        ("max_order", C.c_int32),
        ("filter_nodes", C.c_void_p),
        ("n_filter", C.c_int32),
        ("method", C.c_int32),
        ("steps", C.c_int32),
        ("lr", C.c_float),
        ("x0", C.c_void_p),
        ("mode", C.c_int32),
        ("alpha", C.c_float),
        ("alpha_dev", C.c_void_p),
        ("tol", C.c_float),
        ("patch", C.c_float),
        ("fun", C.c_int32),
        ("r_coef", C.c_double),
        ("height", C.c_double),
        ("reduce_all", C.c_int32),
        ("grad_mode", C.c_int32),
        ("no_cull", C.c_int32),
        ("candidate_slices", C.c_int32),
        ("active_mask", C.c_void_p),
        ("cand_shard_index", C.c_int32),
        ("cand_shard_count", C.c_int32),
        ("optimizer", C.c_int32),
        ("opt_b1", C.c_float),
        ("opt_b2", C.c_float),
        ("opt_eps", C.c_float),
        ("many", C.c_int32),
    ]


class D2DPathRecord(C.Structure):
    _fields_ = [
        ("fixed", C.c_int32),
        ("order", C.c_int32),
        ("grid", C.c_int64),
        ("candidate", C.c_int64),
        ("valid", C.c_float),
        ("loss", C.c_float),
        ("value", C.c_float),
        ("length", C.c_float),
        ("xys", C.c_float * ((MAX_ORDER + 2) * 2)),
    ]


EXPORTS = [
    "d2d_problem_defaults", "d2d_candidates_count", "d2d_candidates_host", "d2d_candidates_device",
    "d2d_problem_num_candidates", "d2d_active_mask_words", "d2d_power_fwd", "d2d_power_bwd", "d2d_paths", "d2d_power_host", "d2d_host_release", "d2d_sanitise_scene", "d2d_affine_points", "d2d_paths_vjp", "d2d_launch_count",
    "d2d_fma_peak_launch", "d2d_last_error", "d2d_abi_version",
]

_lib = None


def lib() -> C.CDLL:
    """Loads the shared object; raises if it has not been built (python -m differt2d_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D2DError(
            f"{LIB_PATH} is missing: build it with `python -m differt2d_b200.build` "
            "(there is no CPU or eager fallback for this path)"
        )
    L = C.CDLL(LIB_PATH)
    P = C.POINTER(D2DProblem)
    vp = C.c_void_p
    L.d2d_problem_defaults.argtypes = [P]
    L.d2d_problem_defaults.restype = None
    L.d2d_paths.argtypes = [P, C.c_float, C.c_int32, vp, C.c_int64, vp, vp]
    L.d2d_paths.restype = C.c_int
    L.d2d_active_mask_words.argtypes = [P]
    L.d2d_active_mask_words.restype = C.c_int64
    L.d2d_candidates_count.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32]
    L.d2d_candidates_count.restype = C.c_int64
    L.d2d_candidates_host.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32, vp]
    L.d2d_candidates_host.restype = C.c_int
    L.d2d_candidates_device.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32, vp, vp]
    L.d2d_candidates_device.restype = C.c_int
    L.d2d_problem_num_candidates.argtypes = [P]
    L.d2d_problem_num_candidates.restype = C.c_int64
    L.d2d_power_fwd.argtypes = [P, vp, vp, vp]
    L.d2d_power_fwd.restype = C.c_int
    L.d2d_power_bwd.argtypes = [P, vp, vp, vp, vp, vp, vp, vp, vp]
    L.d2d_power_bwd.restype = C.c_int
    L.d2d_power_host.argtypes = [P, vp, vp, vp, vp, vp, vp, vp, C.c_int32]
    L.d2d_power_host.restype = C.c_int
    if hasattr(L, "d2d_sanitise_scene") or not os.environ.get("D2D_B200_LIB"):  # (older diagnostic builds lack it)
        L.d2d_sanitise_scene.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp]
        L.d2d_sanitise_scene.restype = C.c_int
        L.d2d_affine_points.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.d2d_affine_points.restype = C.c_int
    if hasattr(L, "d2d_paths_vjp") or not os.environ.get("D2D_B200_LIB"):
        L.d2d_paths_vjp.argtypes = [P, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.d2d_paths_vjp.restype = C.c_int
    L.d2d_host_release.argtypes = []
    L.d2d_host_release.restype = None
    L.d2d_launch_count.argtypes = []
    L.d2d_launch_count.restype = C.c_int64
    L.d2d_fma_peak_launch.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), vp]
    L.d2d_fma_peak_launch.restype = C.c_int
    L.d2d_last_error.argtypes = []
    L.d2d_last_error.restype = C.c_char_p
    L.d2d_abi_version.argtypes = []
    L.d2d_abi_version.restype = C.c_int32
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().d2d_last_error().decode("utf-8", "replace")
        raise D2DError(f"{what} failed (code {rc}): {msg}")


def new_problem() -> D2DProblem:
    p = D2DProblem()
    lib().d2d_problem_defaults(C.byref(p))
    return p
