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
_SO_AD = os.path.join(_HERE, "_build", "libd2d_oracle_ad.so")

MODES = {"hard": 0, "hard_sigmoid": 1, "sigmoid": 2}
FUNS = {"received_power": 0, "length_squared": 1}


def build(force: bool = False) -> str:
    stale = False
    for so, src in ((_SO, "d2d_oracle.c"), (_SO_AD, "d2d_oracle_ad.cpp")):
        srcp = os.path.join(_HERE, src)
        stale = stale or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(srcp)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
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


_lib_ad = None


def lib_ad():
    """oracle/d2d_oracle_ad.cpp — forward-mode AD (dual numbers) over the scalar restatement, fp32 or fp64."""
    global _lib_ad
    if _lib_ad is None:
        if not os.path.exists(_SO_AD):
            build()
        L = C.CDLL(_SO_AD)
        vp = C.c_void_p
        L.orc_ad_power_vjp.restype = C.c_int
        L.orc_ad_power_vjp.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_longlong, C.c_int, C.c_int,
                                       C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, vp,
                                       C.c_float, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        _lib_ad = L
    return _lib_ad


def power_vjp(xys, fixed, grid, Zbar=None, *, kinds=None, phis=None, grid_role="receivers", min_order=0, max_order=1,
              filter_nodes=None, mode="hard", alpha=100.0, tol=1e-2, patch=0.0, fun="received_power", r_coef=0.5,
              height=0.1, real64=False, want_grad=True, nthreads=0) -> dict:
    """
    Value and VJP of the reduce_all map (what jax.vjp of Scene.accumulate_on_*_grid_over_paths(reduce_all=True) gives for
    the cotangent Zbar [R]; None = ones) by forward-mode AD over the literal scalar restatement, in binary32
    (real64=False: the forward value equals power_map(...) bit for bit) or binary64 (the same function of the same fp32
    inputs).  Clean gradients (masked branches are constants).  Returns float64 arrays: Z [R], grid [R,2],
    objects [N,2,2], phis [N], fixed [T,2], alpha [1], and n_valid (paths with a non-zero validity).
    """
    xys = np.ascontiguousarray(np.asarray(xys, dtype=np.float32).reshape(-1, 2, 2))
    n = xys.shape[0]
    kinds = np.ascontiguousarray(np.zeros(n, np.uint8) if kinds is None else np.asarray(kinds, dtype=np.uint8))
    phis = np.ascontiguousarray(np.zeros(n, np.float32) if phis is None else np.asarray(phis, dtype=np.float32))
    fixed = np.ascontiguousarray(np.asarray(fixed, dtype=np.float32).reshape(-1, 2))
    grid = np.ascontiguousarray(np.asarray(grid, dtype=np.float32).reshape(-1, 2))
    T, R = fixed.shape[0], grid.shape[0]
    f = np.ascontiguousarray(np.asarray(filter_nodes if filter_nodes is not None else [], dtype=np.int32))
    zb = None if Zbar is None else np.ascontiguousarray(np.asarray(Zbar, dtype=np.float32).reshape(R))
    rc = np.ascontiguousarray(np.array([np.float32(float(r_coef) ** k) for k in range(max_order + 1)], dtype=np.float32))
    h2 = float(np.float32(float(height) * float(height)))
    out = {"Z": np.zeros(R, np.float64)}
    if want_grad:
        out.update(grid=np.zeros((R, 2), np.float64), objects=np.zeros((n, 2, 2), np.float64), phis=np.zeros(n, np.float64),
                   fixed=np.zeros((T, 2), np.float64), alpha=np.zeros(1, np.float64))
    nv = C.c_longlong(0)
    err = lib_ad().orc_ad_power_vjp(
        int(bool(real64)), _p(xys), _p(kinds), _p(phis), n, _p(fixed), T, _p(grid), R, 0 if grid_role == "receivers" else 1,
        min_order, max_order, _p(f), len(f), MODES[mode], float(alpha), float(tol), float(patch), FUNS[fun], _p(rc), h2,
        _p(zb), _p(out["Z"]), _p(out.get("grid")), _p(out.get("objects")), _p(out.get("phis")), _p(out.get("fixed")),
        _p(out.get("alpha")), C.byref(nv), int(nthreads))
    if err:
        raise RuntimeError(f"orc_ad_power_vjp failed with code {err}" +
                           (" (more tied occluders on one path than the oracle tracks)" if err == 2 else ""))
    out["n_valid"] = int(nv.value)
    return out
