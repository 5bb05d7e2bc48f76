"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/d2d_oracle.c (the scalar C
restatement of the forward hot path).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libd2d_oracle.so")

MODES = {"hard": 0, "hard_sigmoid": 1, "sigmoid": 2}
FUNS = {"received_power": 0, "length_squared": 1}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "d2d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_candidates.restype = C.c_int64
        L.orc_candidates.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_power_map.restype = C.c_int
        L.orc_power_map.argtypes = [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int,
            C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
            C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
        ]
        L.orc_image_path.restype = C.c_float
        L.orc_image_path.argtypes = [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
            C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def candidates(n: int, order: int, filter_nodes=None) -> np.ndarray:
    f = np.ascontiguousarray(np.asarray(filter_nodes if filter_nodes is not None else [], dtype=np.int32))
    c = lib().orc_candidates(n, order, _p(f), len(f), None)
    out = np.empty((c, order), dtype=np.int32)
    lib().orc_candidates(n, order, _p(f), len(f), _p(out))
    return out


def all_path_candidates(n, min_order=0, max_order=1, *, order=None, filter_nodes=None):
    if order is not None:
        min_order = max_order = order
    out = []
    for k in range(min_order, max_order + 1):
        out.extend(list(candidates(n, k, filter_nodes)))
    return out


def power_map(xys, fixed, grid, *, kinds=None, phis=None, grid_role="receivers", min_order=0, max_order=1,
              filter_nodes=None, mode="hard", alpha=100.0, tol=1e-2, patch=0.0, fun="received_power",
              r_coef=0.5, height=0.1, reduce_all=False, want_valid=False, want_fun=False, nthreads=0):
    """Returns Z [T,R] (or [R]), and optionally valid / fun values [T,R,C]."""
    xys = np.ascontiguousarray(np.asarray(xys, dtype=np.float32).reshape(-1, 2, 2))
    n = xys.shape[0]
    kinds = np.ascontiguousarray(np.zeros(n, np.uint8) if kinds is None else np.asarray(kinds, dtype=np.uint8))
    phis = np.ascontiguousarray(np.zeros(n, np.float32) if phis is None else np.asarray(phis, dtype=np.float32))
    fixed = np.ascontiguousarray(np.asarray(fixed, dtype=np.float32).reshape(-1, 2))
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float32).reshape(-1, 2))
    T, R = fixed.shape[0], grid.shape[0]
    f = np.ascontiguousarray(np.asarray(filter_nodes if filter_nodes is not None else [], dtype=np.int32))
    total = sum(lib().orc_candidates(n, k, _p(f), len(f), None) for k in range(min_order, max_order + 1))
    Z = np.zeros((R,) if reduce_all else (T, R), dtype=np.float32)
    valid = np.empty((T, R, total), dtype=np.float32) if want_valid else None
    vals = np.empty((T, R, total), dtype=np.float32) if want_fun else None
    rc = np.ascontiguousarray(np.array([np.float32(float(r_coef) ** k) for k in range(max_order + 1)], dtype=np.float32))
    h2 = float(np.float32(float(height) * float(height)))
    err = lib().orc_power_map(
        _p(xys), _p(kinds), _p(phis), n, _p(fixed), T, _p(grid), R, 0 if grid_role == "receivers" else 1,
        min_order, max_order, _p(f), len(f), MODES[mode], float(alpha), float(tol), float(patch), FUNS[fun],
        _p(rc), h2, int(reduce_all), _p(Z), _p(valid), _p(vals), int(nthreads),
    )
    if err:
        raise RuntimeError(f"orc_power_map failed with code {err}")
    out = [Z]
    if want_valid:
        out.append(valid)
    if want_fun:
        out.append(vals)
    return out[0] if len(out) == 1 else tuple(out)


def num_threads() -> int:
    return lib().orc_num_threads()
