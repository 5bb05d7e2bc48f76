"""
Array-level entry points over the C ABI (the analogue of the `jax.ffi` custom calls wrapped in
`jax.custom_vjp` described in INTEGRATION.md).  torch supplies device buffers, the current CUDA
stream and the autograd graph node; every numerical step happens in libdiffert2d_b200.so.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .defaults import DEFAULT_ALPHA, DEFAULT_HEIGHT, DEFAULT_PATCH, DEFAULT_R_COEF

MODES = {"hard": L.MODE_HARD, "hard_sigmoid": L.MODE_HARD_SIGMOID, "sigmoid": L.MODE_SIGMOID}
METHODS = {"image": L.METHOD_IMAGE, "fermat": L.METHOD_FERMAT, "minpath": L.METHOD_MINPATH}
FUNS = {"received_power": L.FUN_RECEIVED_POWER, "length_squared": L.FUN_LENGTH_SQUARED}
ROLES = {"receivers": L.GRID_RECEIVERS, "transmitters": L.GRID_TRANSMITTERS}
GRAD_MODES = {"clean": L.GRAD_CLEAN, "nan_parity": L.GRAD_NAN_PARITY}
OPTIMIZERS = {"adam": L.OPT_ADAM, "sgd": L.OPT_SGD, "newton": L.OPT_NEWTON}


@dataclass(frozen=True)
class TraceConfig:
    """Static configuration of one call (what would be FFI attributes / the XLA compile-cache key)."""

    grid_role: str = "receivers"
    min_order: int = 0
    max_order: int = 1
    filter_nodes: tuple = ()
    method: str = "image"
    steps: int = 100
    many: int = 1  # Fermat/MinPath restarts (optimize.py:142-182); x0 is then [C, many, max_order]
    lr: float = 0.1
    # optimize.minimize(..., optimizer=...) (optimize.py:44-97): "adam" = optax.adam(lr, b1, b2, eps) (the reference's
    # default), "sgd" = optax.sgd(lr, momentum = opt_b1), "newton" = damped Newton iterations + implicit reverse mode
    # (a non-parity fast mode, csrc/d2d_newton.cuh); see differt2d_b200.optimizers
    optimizer: str = "adam"
    opt_b1: float = 0.9
    opt_b2: float = 0.999
    opt_eps: float = 1e-8
    mode: str = "hard"
    tol: float = 1e-2
    patch: float = DEFAULT_PATCH
    fun: str = "received_power"
    r_coef: float = DEFAULT_R_COEF
    height: float = DEFAULT_HEIGHT
    reduce_all: bool = False
    grid_cols: int = 0  # row length of a row-major mesh grid (enables 16 x 8 tiles); 0 = unknown
    cull: bool = True   # tile-level candidate culling (identical results)
    candidate_slices: int = 0  # 0 = automatic; > 1 splits each tile's candidate list over that many CTAs
    # (index, count): multi-GPU sharding of the candidate list — this call traces every count-th chunk of 128
    # candidates; Z and all cotangents are then partial sums to be added over the shards (distributed.py)
    cand_shard: tuple = (0, 1)
    # "clean": masked branches have zero cotangent.  "nan_parity": additionally NaN wherever jax.grad over the
    # reference's literal graph yields NaN (geometry.py:1105, :227-230, :163-171) — ImagePath only, diagnostic speed
    grad_mode: str = "clean"


def _dev_f32(x, device) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda":
        raise L.D2DError("differt2d_b200 computes on CUDA devices only (no CPU fallback); got device " + str(device))


class _Packed:
    """Keeps every buffer referenced by a D2DProblem alive for the duration of a call."""

    def __init__(self, cfg: TraceConfig, xys, kinds, phis, fixed, grid, alpha, x0, device):
        _require_cuda(device)
        self.device = device
        self.cfg = cfg
        self.xys = _dev_f32(xys, device).reshape(-1, 2, 2)
        n = self.xys.shape[0]
        self.kinds = None
        if kinds is not None:
            k = kinds if isinstance(kinds, torch.Tensor) else torch.as_tensor(np.asarray(kinds, dtype=np.uint8))
            self.kinds = k.to(device=device, dtype=torch.uint8).contiguous()
            assert self.kinds.numel() == n
        self.phis = None if phis is None else _dev_f32(phis, device).reshape(n)
        self.fixed = _dev_f32(fixed, device).reshape(-1, 2)
        self.grid = _dev_f32(grid, device).reshape(-1, 2)
        self.x0 = None if x0 is None else _dev_f32(x0, device)
        self.mask = None
        self.alpha_t = None
        alpha_f = DEFAULT_ALPHA
        if isinstance(alpha, torch.Tensor) and alpha.device.type == "cuda":
            # a device scalar (a traced value): not readable without a synchronisation; the kernels return NaN
            # everywhere when it is not > 0 (include/differt2d_b200.h, alpha_dev)
            self.alpha_t = alpha.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
        else:
            alpha_f = float(alpha)
            if cfg.mode != "hard" and not alpha_f > 0.0:
                raise L.D2DError(f"alpha must be > 0 (the activations must be non-decreasing), got {alpha_f}")
        self.filter = np.ascontiguousarray(np.asarray(cfg.filter_nodes, dtype=np.int32))
        p = L.new_problem()
        p.n_objects = n
        p.objects_xys = self.xys.data_ptr()
        p.object_kinds = self.kinds.data_ptr() if self.kinds is not None else None
        p.object_phis = self.phis.data_ptr() if self.phis is not None else None
        p.n_fixed = self.fixed.shape[0]
        p.fixed_xy = self.fixed.data_ptr()
        p.n_grid = self.grid.shape[0]
        p.grid_xy = self.grid.data_ptr()
        p.grid_role = ROLES[cfg.grid_role]
        p.min_order, p.max_order = int(cfg.min_order), int(cfg.max_order)
        p.filter_nodes = self.filter.ctypes.data if self.filter.size else None
        p.n_filter = int(self.filter.size)
        p.method = METHODS[cfg.method]
        p.steps = int(cfg.steps)
        p.many = int(cfg.many)
        p.lr = float(cfg.lr)
        p.optimizer, p.opt_b1, p.opt_b2, p.opt_eps = OPTIMIZERS[cfg.optimizer], float(cfg.opt_b1), float(cfg.opt_b2), float(cfg.opt_eps)
        p.x0 = self.x0.data_ptr() if self.x0 is not None else None
        p.mode = MODES[cfg.mode]
        p.alpha = alpha_f
        p.alpha_dev = self.alpha_t.data_ptr() if self.alpha_t is not None else None
        p.tol = float(cfg.tol)
        p.patch = float(cfg.patch)
        p.fun = FUNS[cfg.fun]
        p.r_coef = float(cfg.r_coef)
        p.height = float(cfg.height)
        p.reduce_all = int(cfg.reduce_all)
        p.grid_cols = int(cfg.grid_cols)
        p.no_cull = 0 if cfg.cull else 1
        p.candidate_slices = int(cfg.candidate_slices)
        p.cand_shard_index, p.cand_shard_count = int(cfg.cand_shard[0]), int(cfg.cand_shard[1])
        p.grad_mode = GRAD_MODES[cfg.grad_mode]
        self.p = p
        self.T = self.fixed.shape[0]
        self.R = self.grid.shape[0]
        self.N = n

    @property
    def num_candidates(self) -> int:
        c = L.lib().d2d_problem_num_candidates(C.byref(self.p))
        if c < 0:
            raise L.D2DError("invalid problem: " + L.lib().d2d_last_error().decode())
        return int(c)

    def z_shape(self):
        return (self.R,) if self.cfg.reduce_all else (self.T, self.R)

    def new_mask(self) -> torch.Tensor:
        """Device buffer for the activity mask (the only residual the backward can use), attached to the problem."""
        n = L.lib().d2d_active_mask_words(C.byref(self.p))
        if n < 0:
            raise L.D2DError("invalid problem: " + L.lib().d2d_last_error().decode())
        mask = torch.empty(max(int(n), 1), dtype=torch.int32, device=self.device)
        self.attach_mask(mask)
        return mask

    def attach_mask(self, mask: Optional[torch.Tensor]) -> None:
        if mask is None:
            self.p.active_mask = None
            return
        n = L.lib().d2d_active_mask_words(C.byref(self.p))
        same_dev = mask.device.type == "cuda" and (
            self.device.index is None or mask.device.index is None or mask.device.index == self.device.index)
        if mask.dtype != torch.int32 or not same_dev or mask.numel() < n or not mask.is_contiguous():
            raise L.D2DError(f"activity mask must be a contiguous int32 tensor of >= {n} words on {self.device}")
        self.mask = mask
        self.p.active_mask = mask.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def power_fwd(cfg: TraceConfig, xys, fixed, grid, *, kinds=None, phis=None, alpha=DEFAULT_ALPHA, x0=None,
              want_valid: bool = False, want_mask: bool = False, device=None):
    """Z [T,R] (or [R]); with want_valid also the validity of every (fixed, grid point, candidate); with want_mask
    also the activity mask to hand to power_bwd(mask=...) for the same inputs (returned last)."""
    device = torch.device(device) if device is not None else (
        grid.device if isinstance(grid, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
    pk = _Packed(cfg, xys, kinds, phis, fixed, grid, alpha, x0, device)
    with torch.cuda.device(device):
        Z = torch.empty(pk.z_shape(), dtype=torch.float32, device=device)
        valid = None
        if want_valid:
            valid = torch.empty((pk.T, pk.R, pk.num_candidates), dtype=torch.float32, device=device)
        mask = pk.new_mask() if want_mask else None
        rc = L.lib().d2d_power_fwd(C.byref(pk.p), Z.data_ptr(), valid.data_ptr() if want_valid else None, _stream(device))
        L.check(rc, "d2d_power_fwd")
    out = (Z,) + ((valid,) if want_valid else ()) + ((mask,) if want_mask else ())
    return out if len(out) > 1 else Z


def power_bwd(cfg: TraceConfig, xys, fixed, grid, Zbar=None, *, kinds=None, phis=None, alpha=DEFAULT_ALPHA, x0=None,
              want=("Z", "grid", "objects", "phis", "fixed", "alpha"), mask: Optional[torch.Tensor] = None,
              device=None) -> dict:
    """Recompute-based VJP; returns a dict with the requested cotangents (and Z for value_and_grad).  ``mask``: the
    activity mask a power_fwd(want_mask=True) over the SAME inputs returned; the kernel then re-traces only the
    paths the forward found alive (identical results)."""
    device = torch.device(device) if device is not None else (
        grid.device if isinstance(grid, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
    pk = _Packed(cfg, xys, kinds, phis, fixed, grid, alpha, x0, device)
    pk.attach_mask(mask)
    with torch.cuda.device(device):
        zb = None
        if Zbar is not None:
            zb = _dev_f32(Zbar, device).reshape(pk.z_shape())
        out = {}
        if "Z" in want:
            out["Z"] = torch.empty(pk.z_shape(), dtype=torch.float32, device=device)
        if "grid" in want:
            out["grid"] = torch.empty((*pk.z_shape(), 2), dtype=torch.float32, device=device)
        if "objects" in want:
            out["objects"] = torch.empty((pk.N, 2, 2), dtype=torch.float32, device=device)
        if "phis" in want:
            out["phis"] = torch.empty((pk.N,), dtype=torch.float32, device=device)
        if "fixed" in want:
            out["fixed"] = torch.empty((pk.T, 2), dtype=torch.float32, device=device)
        if "alpha" in want:
            out["alpha"] = torch.empty((1,), dtype=torch.float32, device=device)
        ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        rc = L.lib().d2d_power_bwd(C.byref(pk.p), zb.data_ptr() if zb is not None else None, ptr("Z"), ptr("grid"),
                                   ptr("objects"), ptr("phis"), ptr("fixed"), ptr("alpha"), _stream(device))
        L.check(rc, "d2d_power_bwd")
    return out


def power_host(cfg: TraceConfig, xys, fixed, grid, Zbar=None, *, kinds=None, phis=None, alpha=DEFAULT_ALPHA, x0=None,
               want=("Z", "grid", "objects", "phis", "fixed", "alpha"), device: int = 0) -> dict:
    """
    The host-buffer entry (d2d_power_host): numpy arrays in, numpy arrays out, the library stages them (large 2-D
    grids in row chunks on three streams, copies overlapping the kernels), runs the forward kernel and — when a
    cotangent is wanted — the backward kernel over the activity mask, and synchronises.  What a numpy user of the
    reference calls; `bench.py`'s end-to-end leg.
    """
    if not torch.cuda.is_available():
        raise L.D2DError("differt2d_b200 computes on CUDA devices only (no CPU fallback)")
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))  # noqa: E731
    xys_h, fixed_h, grid_h = f32(xys).reshape(-1, 2, 2), f32(fixed).reshape(-1, 2), f32(grid).reshape(-1, 2)
    N, T, R = xys_h.shape[0], fixed_h.shape[0], grid_h.shape[0]
    Tout = 1 if cfg.reduce_all else T
    kinds_h = None if kinds is None else np.ascontiguousarray(np.asarray(kinds, dtype=np.uint8))
    phis_h = None if phis is None else f32(phis).reshape(N)
    x0_h = None if x0 is None else f32(x0)
    zbar_h = None if Zbar is None else f32(Zbar).reshape(Tout, R)
    flt = np.ascontiguousarray(np.asarray(cfg.filter_nodes, dtype=np.int32))
    p = L.new_problem()
    p.n_objects, p.objects_xys = N, xys_h.ctypes.data
    p.object_kinds = kinds_h.ctypes.data if kinds_h is not None else None
    p.object_phis = phis_h.ctypes.data if phis_h is not None else None
    p.n_fixed, p.fixed_xy = T, fixed_h.ctypes.data
    p.n_grid, p.grid_xy = R, grid_h.ctypes.data
    p.grid_role = ROLES[cfg.grid_role]
    p.min_order, p.max_order = int(cfg.min_order), int(cfg.max_order)
    p.filter_nodes, p.n_filter = (flt.ctypes.data if flt.size else None), int(flt.size)
    p.method, p.steps, p.many, p.lr = METHODS[cfg.method], int(cfg.steps), int(cfg.many), float(cfg.lr)
    p.optimizer, p.opt_b1, p.opt_b2, p.opt_eps = OPTIMIZERS[cfg.optimizer], float(cfg.opt_b1), float(cfg.opt_b2), float(cfg.opt_eps)
    p.x0 = x0_h.ctypes.data if x0_h is not None else None
    p.mode, p.alpha, p.tol, p.patch = MODES[cfg.mode], float(alpha), float(cfg.tol), float(cfg.patch)
    p.fun, p.r_coef, p.height = FUNS[cfg.fun], float(cfg.r_coef), float(cfg.height)
    p.reduce_all, p.grid_cols = int(cfg.reduce_all), int(cfg.grid_cols)
    p.no_cull, p.candidate_slices = (0 if cfg.cull else 1), int(cfg.candidate_slices)
    p.cand_shard_index, p.cand_shard_count = int(cfg.cand_shard[0]), int(cfg.cand_shard[1])
    p.grad_mode = GRAD_MODES[cfg.grad_mode]
    shapes = {"Z": (Tout, R) if not cfg.reduce_all else (R,), "grid": ((Tout, R, 2) if not cfg.reduce_all else (R, 2)),
              "objects": (N, 2, 2), "phis": (N,), "fixed": (T, 2), "alpha": (1,)}
    out = {k: np.empty(shapes[k], dtype=np.float32) for k in want}
    ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
    zsink = out["Z"] if "Z" in out else np.empty(shapes["Z"], dtype=np.float32)
    rc = L.lib().d2d_power_host(C.byref(p), zbar_h.ctypes.data if zbar_h is not None else None, zsink.ctypes.data,
                                ptr("grid"), ptr("objects"), ptr("phis"), ptr("fixed"), ptr("alpha"), int(device))
    L.check(rc, "d2d_power_host")
    return out


_REC_WORDS = C.sizeof(L.D2DPathRecord) // 4  # the record as 32-bit words


def paths(cfg: TraceConfig, xys, fixed, grid, *, kinds=None, phis=None, alpha=DEFAULT_ALPHA, x0=None,
          min_valid: float = 0.5, emit_all: bool = False, device=None) -> dict:
    """
    Materialised paths — Scene.all_paths (emit_all) / all_valid_paths (valid > min_valid), scene.py:1156-1248 —
    and the escape hatch for an arbitrary `fun`: one entry per emitted (fixed point, grid point, candidate),
    sorted by those three indices (= the reference's iteration order).  Returns device tensors
    fixed i32[n], order i32[n], grid i64[n], candidate i64[n], valid/loss/value/length f32[n],
    xys f32[n, MAX_ORDER + 2, 2] (rows beyond order + 2 are zero).
    Two launches of the same kernel: a counting pass sizes the record array, the second fills it.
    """
    device = torch.device(device) if device is not None else (
        grid.device if isinstance(grid, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
    pk = _Packed(cfg, xys, kinds, phis, fixed, grid, alpha, x0, device)
    lib = L.lib()
    with torch.cuda.device(device):
        count = torch.zeros(1, dtype=torch.int64, device=device)
        L.check(lib.d2d_paths(C.byref(pk.p), float(min_valid), int(emit_all), None, 0, count.data_ptr(), _stream(device)),
                "d2d_paths (count)")
        n = int(count.item())
        rec = torch.empty((max(n, 1), _REC_WORDS), dtype=torch.int32, device=device)
        if n:
            L.check(lib.d2d_paths(C.byref(pk.p), float(min_valid), int(emit_all), rec.data_ptr(), n, count.data_ptr(),
                                  _stream(device)), "d2d_paths")
            if int(count.item()) != n:
                raise L.D2DError("d2d_paths: the two passes disagree on the number of records")
        rec = rec[:n]
        i64 = rec[:, 2:6].contiguous().view(torch.int64)
        f32 = rec[:, 6:].contiguous().view(torch.float32)
        out = {"fixed": rec[:, 0].contiguous(), "order": rec[:, 1].contiguous(), "grid": i64[:, 0].contiguous(),
               "candidate": i64[:, 1].contiguous(), "valid": f32[:, 0].contiguous(), "loss": f32[:, 1].contiguous(),
               "value": f32[:, 2].contiguous(), "length": f32[:, 3].contiguous(),
               "xys": f32[:, 4:].reshape(n, L.MAX_ORDER + 2, 2).contiguous()}
        if n:
            c_total = max(pk.num_candidates, 1)
            key = (out["fixed"].to(torch.int64) * max(pk.R, 1) + out["grid"]) * c_total + out["candidate"]
            order = torch.argsort(key)
            out = {k: v[order] for k, v in out.items()}
    out["num_candidates"] = pk.num_candidates
    return out


def paths_vjp(cfg: TraceConfig, xys, fixed, grid, rec: dict, valid_bar: torch.Tensor, xys_bar: torch.Tensor, *, kinds=None,
              phis=None, alpha=DEFAULT_ALPHA, want=("grid", "objects", "phis", "fixed", "alpha"), device=None) -> dict:
    """
    Reverse mode of `paths` (d2d_paths_vjp; ImagePath): pulls the cotangents of the records' validity (valid_bar [n])
    and vertices (xys_bar [n, MAX_ORDER + 2, 2]) back to the grid points (per fixed point: [T, R, 2]), the fixed points,
    the object vertices, the RIS angles and alpha.  `rec` is the dict `paths` returned for the SAME problem.
    """
    device = torch.device(device) if device is not None else (
        grid.device if isinstance(grid, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
    pk = _Packed(cfg, xys, kinds, phis, fixed, grid, alpha, None, device)
    n = int(rec["fixed"].numel())
    with torch.cuda.device(device):
        rf = rec["fixed"].to(device=device, dtype=torch.int32).contiguous()
        rg = rec["grid"].to(device=device, dtype=torch.int64).contiguous()
        rc_ = rec["candidate"].to(device=device, dtype=torch.int64).contiguous()
        vb = valid_bar.detach().to(device=device, dtype=torch.float32).reshape(n).contiguous()
        xb = torch.zeros((max(n, 1), L.MAX_ORDER + 2, 2), dtype=torch.float32, device=device)
        if n:
            xb[:n, : xys_bar.shape[1]] = xys_bar.detach().to(device=device, dtype=torch.float32)
        out = {}
        if "grid" in want:
            out["grid"] = torch.empty((pk.T, pk.R, 2), dtype=torch.float32, device=device)
        if "objects" in want:
            out["objects"] = torch.empty((pk.N, 2, 2), dtype=torch.float32, device=device)
        if "phis" in want:
            out["phis"] = torch.empty((pk.N,), dtype=torch.float32, device=device)
        if "fixed" in want:
            out["fixed"] = torch.empty((pk.T, 2), dtype=torch.float32, device=device)
        if "alpha" in want:
            out["alpha"] = torch.empty((1,), dtype=torch.float32, device=device)
        ptr = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        rc = L.lib().d2d_paths_vjp(C.byref(pk.p), n, rf.data_ptr() if n else None, rg.data_ptr() if n else None,
                                   rc_.data_ptr() if n else None, vb.data_ptr() if n else None, xb.data_ptr() if n else None,
                                   ptr("grid"), ptr("objects"), ptr("phis"), ptr("fixed"), ptr("alpha"), _stream(device))
        L.check(rc, "d2d_paths_vjp")
    return out


def power_value_and_vjp(cfg: TraceConfig, xys, fixed, grid, Zbar=None, *, kinds=None, phis=None, alpha=DEFAULT_ALPHA,
                        x0=None, want=("grid", "objects", "phis", "fixed", "alpha"), device=None) -> dict:
    """Z and the requested cotangents (what jax.value_and_grad / jax.vjp deliver): the forward kernel, then the
    backward kernel over the paths the forward found alive (activity mask).  Returns the dict of power_bwd + "Z"."""
    device = torch.device(device) if device is not None else (
        grid.device if isinstance(grid, torch.Tensor) else torch.device("cuda", torch.cuda.current_device()))
    Z, mask = power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, alpha=alpha, x0=x0, want_mask=True, device=device)
    out = power_bwd(cfg, xys, fixed, grid, Zbar, kinds=kinds, phis=phis, alpha=alpha, x0=x0,
                    want=tuple(w for w in want if w != "Z"), mask=mask, device=device)
    out["Z"] = Z
    return out


class _PowerMap(torch.autograd.Function):
    """custom_vjp analogue: residuals are the inputs only; the backward re-traces (no stored activations)."""

    @staticmethod
    def forward(ctx, xys, phis, fixed, grid, alpha, cfg, kinds, x0):
        ctx.cfg, ctx.kinds, ctx.x0 = cfg, kinds, x0
        need_bwd = any(t.requires_grad for t in (xys, phis, fixed, grid, alpha))
        if not need_bwd:
            ctx.save_for_backward(xys, phis, fixed, grid, alpha)
            ctx.has_mask = False
            return power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, alpha=alpha, x0=x0, device=grid.device)
        # residuals: the inputs and one activity bit per (fixed point, warp, candidate)
        Z, mask = power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, alpha=alpha, x0=x0, want_mask=True,
                            device=grid.device)
        ctx.save_for_backward(xys, phis, fixed, grid, alpha, mask)
        ctx.has_mask = True
        return Z

    @staticmethod
    def backward(ctx, Zbar):
        if ctx.has_mask:
            xys, phis, fixed, grid, alpha, mask = ctx.saved_tensors
        else:
            (xys, phis, fixed, grid, alpha), mask = ctx.saved_tensors, None
        need = ctx.needs_input_grad
        want = [w for w, n in zip(("objects", "phis", "fixed", "grid", "alpha"), need[:5]) if n]
        if ctx.cfg.reduce_all or fixed.shape[0] == 1:
            g = power_bwd(ctx.cfg, xys, fixed, grid, Zbar.contiguous(), kinds=ctx.kinds, phis=phis, alpha=alpha,
                          x0=ctx.x0, want=tuple(want), mask=mask, device=grid.device)
            gg = g.get("grid")
            if gg is not None:
                gg = gg.reshape(-1, 2) if ctx.cfg.reduce_all else gg.sum(dim=0)
        else:
            g = power_bwd(ctx.cfg, xys, fixed, grid, Zbar.contiguous(), kinds=ctx.kinds, phis=phis, alpha=alpha,
                          x0=ctx.x0, want=tuple(want), mask=mask, device=grid.device)
            gg = g.get("grid")
            if gg is not None:
                gg = gg.sum(dim=0)
        return (
            g["objects"].reshape(xys.shape) if "objects" in g else None,
            g["phis"].reshape(phis.shape) if "phis" in g else None,
            g["fixed"].reshape(fixed.shape) if "fixed" in g else None,
            gg.reshape(grid.shape) if gg is not None else None,
            g["alpha"].reshape(alpha.shape) if "alpha" in g else None,
            None, None, None,
        )


def power_map(xys: torch.Tensor, fixed: torch.Tensor, grid: torch.Tensor, *, cfg: TraceConfig,
              phis: Optional[torch.Tensor] = None, alpha=DEFAULT_ALPHA, kinds=None, x0=None) -> torch.Tensor:
    """
    Differentiable power map: Z = f(object vertices, RIS angles, fixed points, grid points, alpha).
    ``torch.autograd`` over it plays the role of ``jax.grad`` over
    ``Scene.accumulate_on_*_grid_over_paths`` in the reference.
    """
    dev = grid.device
    _require_cuda(dev)
    xys = xys.to(dev, torch.float32)
    n = xys.reshape(-1, 2, 2).shape[0]
    phis_t = torch.zeros(n, device=dev) if phis is None else phis.to(dev, torch.float32)
    if not (isinstance(alpha, torch.Tensor) and alpha.device.type == "cuda"):
        if cfg.mode != "hard" and not float(alpha) > 0.0:  # (host values are validated here; device scalars in the kernels)
            raise L.D2DError(f"alpha must be > 0 (the activations must be non-decreasing), got {float(alpha)}")
    alpha_t = alpha if isinstance(alpha, torch.Tensor) else torch.tensor(float(alpha), device=dev)
    alpha_t = alpha_t.to(dev, torch.float32)
    return _PowerMap.apply(xys, phis_t, fixed.to(dev, torch.float32), grid.to(dev, torch.float32), alpha_t, cfg,
                           kinds, x0)


def candidates(n_objects: int, order: int, filter_nodes: Sequence[int] = (), *, device=None) -> np.ndarray:
    """
    [count, order] int32 list of one order — scene.py:122-175.  With ``device`` (a CUDA device) the
    list is produced by the integer decode kernel, otherwise by the host odometer of the same library.
    """
    f = np.ascontiguousarray(np.asarray(filter_nodes, dtype=np.int32))
    fp = f.ctypes.data if f.size else None
    cnt = L.lib().d2d_candidates_count(n_objects, order, fp, int(f.size))
    if cnt < 0:
        raise L.D2DError("d2d_candidates_count: invalid arguments")
    if device is not None:
        dev = torch.device(device)
        _require_cuda(dev)
        with torch.cuda.device(dev):
            out = torch.empty((cnt, order), dtype=torch.int32, device=dev)
            rc = L.lib().d2d_candidates_device(n_objects, order, fp, int(f.size), out.data_ptr() if cnt * order else None,
                                               _stream(dev))
            L.check(rc, "d2d_candidates_device")
        return out.cpu().numpy()
    out = np.empty((cnt, order), dtype=np.int32)
    rc = L.lib().d2d_candidates_host(n_objects, order, fp, int(f.size), out.ctypes.data if out.size else None)
    L.check(rc, "d2d_candidates_host")
    return out


def launch_count() -> int:
    return int(L.lib().d2d_launch_count())
