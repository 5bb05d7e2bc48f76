"""
Multi-GPU plumbing: one process per GPU, the receiver grid sharded by rows: bands of 8 rows
dealt round robin for load balance (row_tiles_cyclic), or contiguous blocks (row_block).  Forward maps and per-receiver cotangents need no communication (disjoint rows);
the scene-parameter cotangents (object vertices, RIS angles, fixed points, alpha — a few KB) are
partial sums and take ONE all-reduce per backward call (NCCL over NVLink on GPUs, gloo in CPU tests).
"""

from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def row_block(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [r0, r1) of `rank`; the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def row_tiles_cyclic(n_rows: int, world: int, rank: int, tile_rows: int = 8) -> "np.ndarray":
    """
    Row indices of `rank` when the grid is dealt out in bands of `tile_rows` rows, round robin (band b goes to
    rank b % world).  The kernels' 16 x 8 tiles stay whole, and every rank sees a statistically identical slice of
    the scene: the culls remove a very different share of the work in different places, so contiguous blocks
    (row_block) leave the ranks unevenly loaded while the step time is the maximum over ranks.
    """
    rows = np.arange(n_rows)
    return rows[(rows // tile_rows) % world == rank]


def init(world: Optional[int] = None, rank: Optional[int] = None, backend: Optional[str] = None):
    """Initialises torch.distributed from the torchrun environment (MASTER_ADDR defaults to 127.0.0.1)."""
    if dist.is_initialized():
        return dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, world_size=world, rank=rank, **kw)
    return dist


def allreduce_sum_(buf: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the packed scene-parameter cotangent buffer (on the current stream)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def pack_param_grads(objects_bar, phis_bar, fixed_bar, alpha_bar) -> torch.Tensor:
    """objects [N,2,2] | phis [N] | fixed [T,2] | alpha [1] -> one flat fp32 buffer (one collective)."""
    return torch.cat([objects_bar.reshape(-1), phis_bar.reshape(-1), fixed_bar.reshape(-1), alpha_bar.reshape(-1)])


def unpack_param_grads(buf: torch.Tensor, n_objects: int, n_fixed: int):
    o = 0
    objects = buf[o:o + 4 * n_objects].reshape(n_objects, 2, 2); o += 4 * n_objects
    phis = buf[o:o + n_objects]; o += n_objects
    fixed = buf[o:o + 2 * n_fixed].reshape(n_fixed, 2); o += 2 * n_fixed
    return objects, phis, fixed, buf[o:o + 1]


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def _current_cuda_device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        from ._lib import D2DError

        raise D2DError("differt2d_b200 computes on CUDA devices only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def sharded_power_vjp(cfg, xys, fixed, X, Y, Zbar=None, *, alpha=100.0, kinds=None, phis=None, device=None,
                      layout: str = "cyclic"):
    """
    Row-sharded forward + VJP: this rank traces its rows of the (n, m) grid (`layout="cyclic"`: bands of 8
    rows dealt round robin, row_tiles_cyclic; `"block"`: one contiguous block, row_block) with the forward kernel and
    the mask-driven backward kernel (functional.power_value_and_vjp) and returns its slice of Z / grid_bar plus the
    ALL-REDUCED scene-parameter cotangents; out["rows"] holds the row indices it owns.
    Zbar: [n, m] (reduce_all) or [T, n, m]; sliced along the row axis.
    """
    import dataclasses

    from . import functional as F

    device = _current_cuda_device(device)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    Xn, Yn = np.asarray(X, dtype=np.float32), np.asarray(Y, dtype=np.float32)
    n, m = Xn.shape
    if layout == "cyclic":
        rows = row_tiles_cyclic(n, world, rank)
    elif layout == "block":
        rows = np.arange(*row_block(n, world, rank))
    else:
        raise ValueError("layout must be 'cyclic' or 'block'")
    grid = torch.as_tensor(np.stack([Xn[rows], Yn[rows]], -1).reshape(-1, 2)).to(device)
    zb = None
    if Zbar is not None:
        zb_all = np.asarray(Zbar, dtype=np.float32)
        if zb_all.shape[-2:] != (n, m):
            raise ValueError(f"Zbar must end with the grid shape {(n, m)}, got {zb_all.shape}")
        zb = zb_all[..., rows, :].reshape(*zb_all.shape[:-2], -1)  # rows are the second-to-last axis, whatever T
    cfg = dataclasses.replace(cfg, grid_cols=m)
    out = F.power_value_and_vjp(cfg, xys, fixed, grid, zb, kinds=kinds, phis=phis, alpha=alpha, device=device)
    n_obj = out["objects"].shape[0]
    buf = pack_param_grads(out["objects"], out["phis"], out["fixed"], out["alpha"])
    allreduce_sum_(buf)
    out["objects"], out["phis"], out["fixed"], out["alpha"] = unpack_param_grads(buf, n_obj, out["fixed"].shape[0])
    out["rows"] = rows
    return out


def candidate_chunks_of_shard(n_candidates: int, slices: int, shard: int, n_shards: int, chunk: int = 128):
    """
    Host-side statement of how the kernels deal ONE order's candidate list to candidate shards (csrc/d2d_driver.cuh,
    for_each_candidate): the list is cut into chunks of 128 candidates; CTA y of shard s walks the chunks q with
    q % (slices * n_shards) == s * slices + y.  Returns the chunk indices of `shard` (all its CTAs), ascending.
    Every chunk belongs to exactly one shard; the shards take turns every `slices` chunks (equal work per GPU on the
    long lists this is meant for).
    """
    n_chunks = (int(n_candidates) + chunk - 1) // chunk
    q = np.arange(n_chunks)
    v = q % (slices * n_shards)
    return q[(v >= shard * slices) & (v < (shard + 1) * slices)]


def sharded_link_power_vjp(cfg, xys, tx, rx, Zbar=None, *, alpha=100.0, kinds=None, phis=None, x0=None, device=None,
                           want=("grid", "objects", "phis", "fixed", "alpha")):
    """
    Point-to-point links with candidate lists too long for one GPU (500 objects at order 3 are 1.2e8
    candidates per link): every rank traces ALL (tx, rx) links over ITS share of the candidate list
    (TraceConfig.cand_shard = (rank, world): chunks of 128 candidates dealt round robin) and ONE all-reduce adds the
    partial Z [T, R], the partial receiver cotangents [T, R, 2] and the partial scene-parameter cotangents.
    Returns the dict of functional.power_value_and_vjp with every entry summed over the ranks.
    """
    import dataclasses

    from . import functional as F

    device = _current_cuda_device(device)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    cfg = dataclasses.replace(cfg, cand_shard=(rank, world))
    grid = torch.as_tensor(np.asarray(rx, dtype=np.float32).reshape(-1, 2)).to(device)
    out = F.power_value_and_vjp(cfg, xys, tx, grid, Zbar, kinds=kinds, phis=phis, alpha=alpha, x0=x0, want=want,
                                device=device)
    keys = [k for k in ("Z", "grid", "objects", "phis", "fixed", "alpha") if k in out]
    buf = torch.cat([out[k].reshape(-1) for k in keys])
    allreduce_sum_(buf)
    o = 0
    for k in keys:
        n = out[k].numel()
        out[k] = buf[o:o + n].reshape(out[k].shape)
        o += n
    return out
