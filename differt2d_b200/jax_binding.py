"""
`jax.ffi` binding of libdiffert2d_b200: the custom calls of jax_ffi/d2d_xla_ffi.cc wrapped in `jax.custom_vjp`
(BASELINE north_star).  Importing this module needs JAX with a CUDA jaxlib AND the shim built by
`python -m differt2d_b200.build --jax-ffi` — neither exists in the image this repository was developed in (no network,
no wheels), where the same C entry points are driven through ctypes + torch (functional.py).  Nothing else in the package
imports this module; without JAX it raises ImportError on import, never a silent fallback.

    from differt2d_b200 import jax_binding as jb
    Z = jb.power_map(cfg, xys, kinds, phis, fixed, grid, alpha, x0)        # differentiable w.r.t. xys, phis, fixed,
    jax.grad(lambda g: jb.power_map(cfg, ..., g, ...).sum())(grid)         # grid and alpha; cfg = functional.TraceConfig

Residuals of the VJP: the inputs plus the activity mask (one bit per fixed point, warp of 32 grid points and candidate)
that the forward call writes — the reference keeps a tape of every intermediate at [n, m] size (scene.py:1920-1952).
"""
from __future__ import annotations

import ctypes as C
import os
from functools import partial

import numpy as np

import jax  # noqa: E402  (ImportError here is the intended failure mode without JAX)
import jax.numpy as jnp

from . import _lib as L
from . import functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libdiffert2d_b200_xla.so")
if not os.path.exists(_SHIM):
    raise ImportError(f"{_SHIM} is missing: build it with `python -m differt2d_b200.build --jax-ffi`")
L.lib()  # the shim links against libdiffert2d_b200.so: load it first
_shim = C.CDLL(_SHIM, mode=C.RTLD_GLOBAL)
jax.ffi.register_ffi_target("d2d_power_fwd", jax.ffi.pycapsule(_shim.D2dPowerFwd), platform="CUDA")
jax.ffi.register_ffi_target("d2d_power_bwd", jax.ffi.pycapsule(_shim.D2dPowerBwd), platform="CUDA")


def _attrs(cfg: F.TraceConfig) -> dict:
    """Static configuration -> FFI attributes (part of the XLA compile-cache key)."""
    return dict(
        filter_nodes=np.asarray(cfg.filter_nodes, dtype=np.int32),
        grid_role=np.int32(F.ROLES[cfg.grid_role]), grid_cols=np.int32(cfg.grid_cols),
        min_order=np.int32(cfg.min_order), max_order=np.int32(cfg.max_order), method=np.int32(F.METHODS[cfg.method]),
        steps=np.int32(cfg.steps), many=np.int32(cfg.many), mode=np.int32(F.MODES[cfg.mode]),
        fun=np.int32(F.FUNS[cfg.fun]), grad_mode=np.int32(F.GRAD_MODES[cfg.grad_mode]),
        candidate_slices=np.int32(cfg.candidate_slices), lr=np.float32(cfg.lr), tol=np.float32(cfg.tol),
        patch=np.float32(cfg.patch), r_coef=np.float64(cfg.r_coef), height=np.float64(cfg.height),
        reduce_all=bool(cfg.reduce_all), cull=bool(cfg.cull))


def _mask_words(cfg: F.TraceConfig, n_objects: int, n_fixed: int, n_grid: int) -> int:
    """Host call on shapes only (d2d_active_mask_words): the size of the VJP residual."""
    dummy = np.zeros(4, np.float32)
    flt = np.ascontiguousarray(np.asarray(cfg.filter_nodes, dtype=np.int32))
    p = L.new_problem()
    p.n_objects, p.objects_xys = n_objects, dummy.ctypes.data
    p.n_fixed, p.fixed_xy = n_fixed, dummy.ctypes.data
    p.n_grid, p.grid_xy = n_grid, dummy.ctypes.data
    p.grid_role, p.grid_cols = F.ROLES[cfg.grid_role], int(cfg.grid_cols)
    p.min_order, p.max_order = int(cfg.min_order), int(cfg.max_order)
    p.filter_nodes, p.n_filter = (flt.ctypes.data if flt.size else None), int(flt.size)
    p.method, p.mode, p.reduce_all = F.METHODS[cfg.method], F.MODES[cfg.mode], int(cfg.reduce_all)
    p.candidate_slices, p.no_cull = int(cfg.candidate_slices), 0 if cfg.cull else 1
    n = L.lib().d2d_active_mask_words(C.byref(p))
    if n < 0:
        raise L.D2DError("d2d_active_mask_words: " + L.lib().d2d_last_error().decode())
    return int(n)


def _z_shape(cfg, fixed, grid):
    return (grid.shape[0],) if cfg.reduce_all else (fixed.shape[0], grid.shape[0])


def _forward(cfg, xys, kinds, phis, fixed, grid, alpha, x0, want_mask: bool):
    n_mask = _mask_words(cfg, xys.shape[0], fixed.shape[0], grid.shape[0]) if want_mask else 0
    out = (jax.ShapeDtypeStruct(_z_shape(cfg, fixed, grid), jnp.float32), jax.ShapeDtypeStruct((n_mask,), jnp.uint32))
    return jax.ffi.ffi_call("d2d_power_fwd", out, vmap_method="sequential")(
        xys, kinds, phis, fixed, grid, jnp.reshape(alpha, (1,)).astype(jnp.float32), x0, **_attrs(cfg))


@partial(jax.custom_vjp, nondiff_argnums=(0,))
def power_map(cfg: F.TraceConfig, xys, kinds, phis, fixed, grid, alpha, x0):
    """facc over all candidates + vmap∘vmap over the grid (scene.py:1892-1937 / :1589-1632) as ONE custom call.
    xys f32[N,2,2], kinds u8[N] (or [0]), phis f32[N] (or [0]), fixed f32[T,2], grid f32[R,2] (= dstack((X, Y)) flattened),
    alpha scalar (may be traced), x0 f32[C,many,max_order] (or [0]).  Returns Z f32[T,R], or [R] with cfg.reduce_all."""
    return _forward(cfg, xys, kinds, phis, fixed, grid, alpha, x0, want_mask=False)[0]


def _power_map_fwd(cfg, xys, kinds, phis, fixed, grid, alpha, x0):
    Z, mask = _forward(cfg, xys, kinds, phis, fixed, grid, alpha, x0, want_mask=True)
    return Z, (xys, kinds, phis, fixed, grid, alpha, x0, mask)


def _power_map_bwd(cfg, res, Zbar):
    xys, kinds, phis, fixed, grid, alpha, x0, mask = res
    T, R, N = fixed.shape[0], grid.shape[0], xys.shape[0]
    shapes = (jax.ShapeDtypeStruct((R, 2) if cfg.reduce_all else (T, R, 2), jnp.float32),
              jax.ShapeDtypeStruct((N, 2, 2), jnp.float32), jax.ShapeDtypeStruct((N,), jnp.float32),
              jax.ShapeDtypeStruct((T, 2), jnp.float32), jax.ShapeDtypeStruct((1,), jnp.float32))
    gbar, obar, pbar, fbar, abar = jax.ffi.ffi_call("d2d_power_bwd", shapes, vmap_method="sequential")(
        xys, kinds, phis, fixed, grid, jnp.reshape(alpha, (1,)).astype(jnp.float32), x0, Zbar.astype(jnp.float32), mask,
        **_attrs(cfg))
    gbar = gbar if cfg.reduce_all else gbar.sum(0)  # every fixed point's map depends on the same grid points
    pbar = pbar if phis.shape[0] else jnp.zeros_like(phis)
    # cotangents in the order of the differentiable arguments: xys, kinds, phis, fixed, grid, alpha, x0
    return obar, None, pbar, fbar, gbar, jnp.reshape(abar, jnp.shape(alpha)), None


power_map.defvjp(_power_map_fwd, _power_map_bwd)
