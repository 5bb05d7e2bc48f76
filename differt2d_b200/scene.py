"""
Host-side mirror of differt2d/scene.py for the hot path: the `Scene` container, its canned
builders (used as input generators), `all_path_candidates`, and the three accumulation entry
points with the reference's names, keywords, return structure and iteration order —

    Scene.accumulate_on_receivers_grid_over_paths      scene.py:1803-1953
    Scene.accumulate_on_transmitters_grid_over_paths   scene.py:1489-1648
    Scene.accumulate_over_paths                        scene.py:1272-1334

Arrays in, arrays out: `X`, `Y` may be torch tensors (CUDA: zero-copy; CPU: staged) or numpy
arrays; results come back in the same container type, float32, shaped like `X` (+ a trailing 2 for
gradients), exactly as the reference returns them.  Everything numerical runs in the CUDA library.
"""

from __future__ import annotations

import json
from typing import Any, Callable, Iterator, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import functional as F
from . import optimizers, utils
from .defaults import DEFAULT_ALPHA, DEFAULT_HEIGHT, DEFAULT_PATCH, DEFAULT_R_COEF
from .geometry import RIS, FermatPath, ImagePath, MinPath, ObjectBatch, Path, PathBatch, Point, PointBatch, Vertex, Wall
from .logic import resolve_mode


def _resolve_fun(fun, fun_args, fun_kwargs):
    fun_kwargs = dict(fun_kwargs or {})
    fused = fun in (utils.received_power, utils.length_squared, "received_power", "length_squared")
    if fun_args and fused:
        raise NotImplementedError("positional fun_args are not supported by the fused kernels; use fun_kwargs")
    if fun is utils.received_power or fun == "received_power":
        r_coef = float(fun_kwargs.pop("r_coef", DEFAULT_R_COEF))
        height = float(fun_kwargs.pop("height", DEFAULT_HEIGHT))
        if fun_kwargs:
            raise TypeError(f"unexpected fun_kwargs for received_power: {sorted(fun_kwargs)}")
        return "received_power", r_coef, height
    if fun is utils.length_squared or fun == "length_squared":
        return "length_squared", DEFAULT_R_COEF, DEFAULT_HEIGHT
    if callable(fun):
        # escape hatch: the paths are materialised by the CUDA kernel and `fun` runs in the host framework with the
        # reference's own signature on batched arguments, see Scene._generic_accumulate
        return "generic", DEFAULT_R_COEF, DEFAULT_HEIGHT
    raise NotImplementedError("fun must be utils.received_power, utils.length_squared or a callable "
                              "fun(transmitter, receiver, path, interacting_objects, *fun_args, **fun_kwargs)")


def _default_device():
    if not torch.cuda.is_available():
        raise F.L.D2DError("no CUDA device: differt2d_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Scene:
    """scene.py:178-192 — transmitters / receivers by name (insertion order kept) and a list of objects."""

    def __init__(self, transmitters: Optional[Mapping[str, Point]] = None,
                 receivers: Optional[Mapping[str, Point]] = None, objects: Sequence[Any] = ()):
        self.transmitters = dict(transmitters or {})
        self.receivers = dict(receivers or {})
        self.objects = list(objects)

    # ---- mutators (scene.py:194-426), functional style -------------------------------------------
    def with_transmitters(self, **tx: Point) -> "Scene":
        return Scene(tx, self.receivers, self.objects)

    def with_receivers(self, **rx: Point) -> "Scene":
        return Scene(self.transmitters, rx, self.objects)

    def with_objects(self, *objects) -> "Scene":
        return Scene(self.transmitters, self.receivers, objects)

    def add_objects(self, *objects) -> "Scene":
        return Scene(self.transmitters, self.receivers, [*self.objects, *objects])

    def update_transmitters(self, **tx: Point) -> "Scene":
        return Scene({**self.transmitters, **tx}, self.receivers, self.objects)

    def update_receivers(self, **rx: Point) -> "Scene":
        return Scene(self.transmitters, {**self.receivers, **rx}, self.objects)

    # ---- canned scenes (scene.py:670-935): input generators --------------------------------------
    @classmethod
    def from_walls_array(cls, walls) -> "Scene":  # scene.py:670-693
        walls = np.asarray(walls, dtype=np.float32).reshape(-1, 2, 2)
        return cls(objects=[Wall(xys=w) for w in walls])

    @classmethod
    def square_scene(cls, tx_coords=(0.2, 0.2), rx_coords=(0.5, 0.6)) -> "Scene":  # scene.py:790-836
        walls = [[[0, 0], [1, 0]], [[1, 0], [1, 1]], [[1, 1], [0, 1]], [[0, 1], [0, 0]]]
        return cls({"tx": Point(xy=tx_coords)}, {"rx": Point(xy=rx_coords)}, [Wall(xys=w) for w in walls])

    @classmethod
    def square_scene_with_wall(cls, ratio=0.6, tx_coords=(0.2, 0.5), rx_coords=(0.8, 0.5)) -> "Scene":  # :839-882
        s = cls.square_scene(tx_coords, rx_coords)
        return s.add_objects(Wall(xys=[[0.5, 0.5 * (1 - ratio)], [0.5, 0.5 * (1 + ratio)]]))

    @classmethod
    def square_scene_with_obstacle(cls, ratio=0.1, **kwargs) -> "Scene":  # scene.py:885-935
        s = cls.square_scene(**kwargs)
        hl = 0.5 * ratio
        x0, x1, y0, y1 = 0.5 - hl, 0.5 + hl, 0.5 - hl, 0.5 + hl
        return s.add_objects(Wall(xys=[[x0, y0], [x1, y0]]), Wall(xys=[[x1, y0], [x1, y1]]),
                             Wall(xys=[[x1, y1], [x0, y1]]), Wall(xys=[[x0, y1], [x0, y0]]))

    @classmethod
    def basic_scene(cls, tx_coords=(0.1, 0.1), rx_coords=(0.302, 0.2147)) -> "Scene":  # scene.py:736-787
        s = cls.square_scene(tx_coords, rx_coords)
        return s.add_objects(Wall(xys=[[0.4, 0.0], [0.4, 0.4]]), Wall(xys=[[0.4, 0.4], [0.3, 0.4]]),
                             Wall(xys=[[0.1, 0.4], [0.0, 0.4]]))

    @classmethod
    def random_uniform_scene(cls, *, key, n_transmitters=1, n_walls=1, n_receivers=1) -> "Scene":
        """scene.py:696-733 layout; ``key`` seeds numpy's Generator (the JAX key stream is not reproduced)."""
        rng = np.random.default_rng(key)
        pts = rng.random((n_transmitters + 2 * n_walls + n_receivers, 2), dtype=np.float32)
        tx = {f"tx_{i}": Point(xy=pts[i]) for i in range(n_transmitters)}
        rx = {f"rx_{i}": Point(xy=pts[-(i + 1)]) for i in range(n_receivers)}
        walls = [Wall(xys=pts[2 * i + n_transmitters: 2 * i + 2 + n_transmitters]) for i in range(n_walls)]
        return cls(tx, rx, walls)

    @classmethod
    def from_geojson(cls, s_or_fp, tx_loc="NW", rx_loc="SE") -> "Scene":
        """scene.py:428-668 — one Wall per (coords[i-1], coords[i]) of every Polygon's outer ring."""
        if hasattr(s_or_fp, "read"):
            s_or_fp = s_or_fp.read()
        d = json.loads(s_or_fp)
        walls = []
        for feat in d.get("features", []):
            geom = feat.get("geometry")
            if geom and geom["type"] == "Polygon":
                coords = geom["coordinates"][0]
                for i in range(len(coords)):
                    walls.append(Wall(xys=np.asarray([coords[i - 1], coords[i]], dtype=np.float64).astype(np.float32)))
        sc = cls(objects=walls)
        if walls:
            return sc.with_transmitters(tx=Point(xy=sc.get_location(tx_loc))).with_receivers(
                rx=Point(xy=Scene(objects=walls).get_location(rx_loc)))
        return sc.with_transmitters(tx=Point(xy=[0.0, 0.0])).with_receivers(rx=Point(xy=[1.0, 1.0]))

    def sanitised(self, drop_zero_length: bool = True, normalise: bool = False, return_map: bool = False, device=None):
        """
        SURVEY §8 (f)2 — a scene the fp32 path tracer is well conditioned on (NOT part of the reference: results
        differ from the raw scene's, which stays the parity case).

        * ``drop_zero_length``: removes zero-length walls — `from_geojson` emits one closure wall per polygon ring
          (scene.py:646-652: ``coords[i - 1]`` wraps to the duplicated last vertex).  Every candidate through such a
          wall is invalid (n = 0, loss = 1) but poisons the reference's reverse mode with NaN (geometry.py:1105).
        * ``normalise``: shifts / scales all coordinates to the unit square (computed in float64).  Raw lon/lat
          coordinates (|y| ~ 50) put ~3 % of a wall length of fp32 lattice noise on every parametric coordinate;
          distances change by the scale factor, so `received_power`'s `height` has to be rescaled by the caller
          (the factor is returned with the map).

        ``return_map``: also returns ``{"kept": indices of the kept objects, "origin": [2], "scale": float}`` so that
        cotangents w.r.t. the sanitised vertices can be carried back (d/d raw = d/d sanitised / scale).

        ``device``: a CUDA device — the object table is then built by the library's sanitiser kernel
        (``d2d_sanitise_scene`` / ``d2d_affine_points``, csrc/d2d_sanitise.cu: flags, ordered compaction, bounding box and
        affine map in binary64 on the device); the result is identical to the host path (same binary64 arithmetic).
        """
        if device is not None:
            return self._sanitised_on_device(drop_zero_length, normalise, return_map, torch.device(device))
        kept = [i for i, o in enumerate(self.objects)
                if not (drop_zero_length and isinstance(o, Wall) and not np.any(o.xys[1] != o.xys[0]))]
        objs = [self.objects[i] for i in kept]
        origin, scale = np.zeros(2, np.float64), 1.0
        tx, rx = self.transmitters, self.receivers
        if normalise:
            bb = Scene(tx, rx, objs).bounding_box().astype(np.float64)
            origin = bb[0]
            scale = float(max(bb[1, 0] - bb[0, 0], bb[1, 1] - bb[0, 1])) or 1.0

            def f(a):
                return ((np.asarray(a, np.float64) - origin) / scale).astype(np.float32)

            moved = []
            for o in objs:
                if isinstance(o, Vertex):
                    moved.append(Vertex(xy=f(o.xy)))
                elif isinstance(o, RIS):
                    moved.append(RIS(xys=f(o.xys), phi=o.phi))
                elif isinstance(o, Wall):
                    moved.append(Wall(xys=f(o.xys)))
                else:
                    raise NotImplementedError(f"object type {type(o).__name__}")
            objs = moved
            tx = {k: Point(xy=f(p.xy)) for k, p in tx.items()}
            rx = {k: Point(xy=f(p.xy)) for k, p in rx.items()}
        sc = Scene(tx, rx, objs)
        if return_map:
            return sc, {"kept": kept, "origin": origin.astype(np.float64), "scale": scale}
        return sc

    def _sanitised_on_device(self, drop_zero_length, normalise, return_map, device):
        import ctypes as C

        F._require_cuda(device)
        lib = F.L.lib()
        xys, kinds, phis = self.packed_objects()
        n = xys.shape[0]
        names = list(self.transmitters) + list(self.receivers)
        pts = np.stack([p.xy for p in (*self.transmitters.values(), *self.receivers.values())]) if names \
            else np.zeros((0, 2), np.float32)
        with torch.cuda.device(device):
            d_xys = torch.as_tensor(xys).to(device).contiguous()
            d_kinds, d_phis = torch.as_tensor(kinds).to(device), torch.as_tensor(phis).to(device)
            d_pts = torch.as_tensor(pts).to(device).contiguous()
            o_xys, o_kinds, o_phis = torch.empty_like(d_xys), torch.empty_like(d_kinds), torch.empty_like(d_phis)
            kept = torch.empty(max(n, 1), dtype=torch.int32, device=device)
            n_kept = torch.zeros(1, dtype=torch.int32, device=device)
            affine = torch.zeros(3, dtype=torch.float64, device=device)
            o_pts = torch.empty_like(d_pts)
            st = torch.cuda.current_stream(device).cuda_stream
            ptr = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
            F.L.check(lib.d2d_sanitise_scene(ptr(d_xys), ptr(d_kinds), ptr(d_phis), n, ptr(d_pts), pts.shape[0],
                                             int(bool(drop_zero_length)), int(bool(normalise)), ptr(o_xys), ptr(o_kinds),
                                             ptr(o_phis), kept.data_ptr(), n_kept.data_ptr(), None, affine.data_ptr(), st),
                      "d2d_sanitise_scene")
            F.L.check(lib.d2d_affine_points(ptr(d_pts), pts.shape[0], affine.data_ptr(), ptr(o_pts), st), "d2d_affine_points")
            m = int(n_kept.item())
            kept_h = kept[:m].cpu().numpy().tolist()
            xys_h, kinds_h, phis_h = o_xys[:m].cpu().numpy(), o_kinds[:m].cpu().numpy(), o_phis[:m].cpu().numpy()
            aff = affine.cpu().numpy()
            pts_h = o_pts.cpu().numpy()
        objs = []
        for q in range(m):
            if kinds_h[q] == Vertex.KIND:
                objs.append(Vertex(xy=xys_h[q, 0]))
            elif kinds_h[q] == RIS.KIND:
                objs.append(RIS(xys=xys_h[q], phi=float(phis_h[q])))
            else:
                objs.append(Wall(xys=xys_h[q]))
        nt = len(self.transmitters)
        tx = {k: Point(xy=pts_h[i]) for i, k in enumerate(self.transmitters)}
        rx = {k: Point(xy=pts_h[nt + i]) for i, k in enumerate(self.receivers)}
        sc = Scene(tx, rx, objs)
        if return_map:
            return sc, {"kept": kept_h, "origin": aff[:2].astype(np.float64), "scale": float(aff[2])}
        return sc

    # ---- Plottable pieces needed to build inputs (abc.py:30-126, scene.py:1023-1036) --------------
    def bounding_box(self) -> np.ndarray:
        boxes = [p.bounding_box() for p in self.transmitters.values()]
        boxes += [p.bounding_box() for p in self.receivers.values()]
        boxes += [o.bounding_box() for o in self.objects]
        b = np.stack(boxes)
        return np.vstack([b[:, 0].min(axis=0), b[:, 1].max(axis=0)]).astype(np.float32)

    def grid(self, m: int = 50, n: Optional[int] = None):
        """abc.py:59-81 — (X, Y) of shape (n, m) over the bounding box."""
        bb = self.bounding_box()
        n = m if n is None else n
        x = np.linspace(bb[0, 0], bb[1, 0], m, dtype=np.float32)
        y = np.linspace(bb[0, 1], bb[1, 1], n, dtype=np.float32)
        return np.meshgrid(x, y)

    def center(self) -> np.ndarray:
        bb = self.bounding_box()
        return (np.float32(0.5) * (bb[0] + bb[1])).astype(np.float32)

    def get_location(self, location: str) -> np.ndarray:  # abc.py:95-126
        (xmin, ymin), (xmax, ymax) = self.bounding_box()
        xavg, yavg = np.float32(0.5) * (xmin + xmax), np.float32(0.5) * (ymin + ymax)
        x, y = {"N": (xavg, ymax), "E": (xmax, yavg), "S": (xavg, ymin), "W": (xmin, yavg), "C": (xavg, yavg),
                "NE": (xmax, ymax), "NW": (xmin, ymax), "SE": (xmax, ymin), "SW": (xmin, ymin)}[location]
        return np.array([x, y], dtype=np.float32)

    # ---- packing ----------------------------------------------------------------------------------
    def packed_objects(self):
        """objects as arrays: xys [N,2,2] f32, kinds [N] u8, phis [N] f32."""
        n = len(self.objects)
        xys = np.zeros((n, 2, 2), np.float32)
        kinds = np.zeros(n, np.uint8)
        phis = np.zeros(n, np.float32)
        for i, o in enumerate(self.objects):
            if not isinstance(o, (Wall, Vertex)):
                raise NotImplementedError(f"object type {type(o).__name__}: the kernels support Wall, RIS and Vertex")
            xys[i] = o.packed_xys()
            kinds[i] = o.KIND
            if isinstance(o, RIS):
                phis[i] = o.phi
        return xys, kinds, phis

    def all_transmitter_receiver_pairs(self):  # scene.py:1072-1087
        for tx in self.transmitters.items():
            for rx in self.receivers.items():
                yield tx, rx

    # ---- candidates (scene.py:1089-1134) ------------------------------------------------------------
    def _filter_nodes(self, filter_objects) -> tuple:
        if filter_objects is None:
            return ()
        return tuple(i for i, o in enumerate(self.objects) if not filter_objects(o))

    def all_path_candidates(self, min_order: int = 0, max_order: int = 1, *, order: Optional[int] = None,
                            filter_objects: Optional[Callable[[Any], bool]] = None, device=None) -> list:
        """List with one int32 array (shape (k,)) per candidate, orders ascending, lexicographic inside."""
        if order is not None:
            min_order = max_order = order
        fn = self._filter_nodes(filter_objects)
        if device is None and torch.cuda.is_available():
            device = _default_device()
        out = []
        for k in range(min_order, max_order + 1):
            out.extend(list(F.candidates(len(self.objects), k, fn, device=device)))
        return out

    # ---- accumulation -------------------------------------------------------------------------------
    def _config(self, grid_role, fun, fun_args, fun_kwargs, reduce_all, path_cls, path_cls_kwargs, min_order,
                max_order, order, filter_objects, kwargs):
        kwargs = dict(kwargs)
        if order is not None:
            min_order = max_order = order
        method = getattr(path_cls, "METHOD", None)
        if method is None:
            raise NotImplementedError("path_cls must be ImagePath, FermatPath or MinPath")
        pk = dict(path_cls_kwargs or {})
        steps = int(pk.pop("steps", 100))
        many = int(pk.pop("many", 1))  # geometry.py:1198 / :1282: the path classes default to a single run
        if many < 1:
            raise ValueError("many must be >= 1")
        optimizer = pk.pop("optimizer", None)
        if optimizer is None:
            optimizer = optimizers.adam()  # optimize.py:83
        if not isinstance(optimizer, optimizers.Optimizer):
            raise NotImplementedError("the solver runs inside the CUDA kernels: `optimizer` must be one of "
                                      "differt2d_b200.optimizers.adam / sgd / newton (an optax object cannot be fused)")
        if pk:
            raise NotImplementedError(f"unsupported path_cls_kwargs: {sorted(pk)}")
        mode = resolve_mode(kwargs.pop("approx", None), kwargs.pop("function", None))
        alpha = kwargs.pop("alpha", DEFAULT_ALPHA)
        tol = float(kwargs.pop("tol", 1e-2))
        patch = float(kwargs.pop("patch", DEFAULT_PATCH))
        # extension (not a reference keyword): grad / value_and_grad return NaN wherever jax.grad over the reference's
        # literal graph does (F.TraceConfig.grad_mode); the default is the clean gradient
        nan_parity = bool(kwargs.pop("nan_parity", False))
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        fname, r_coef, height = _resolve_fun(fun, fun_args, fun_kwargs)
        if method == "image" and any(isinstance(o, Vertex) for i, o in enumerate(self.objects)
                                     if i not in self._filter_nodes(filter_objects)):
            raise TypeError("ImagePath cannot interact with Vertex objects (geometry.py:1020 expects walls)")
        cfg = F.TraceConfig(grid_role=grid_role, min_order=min_order, max_order=max_order,
                            filter_nodes=self._filter_nodes(filter_objects), method=method, steps=steps, many=many,
                            lr=optimizer.learning_rate, optimizer=optimizer.kind, opt_b1=optimizer.b1,
                            opt_b2=optimizer.b2, opt_eps=optimizer.eps, mode=mode, tol=tol, patch=patch, fun="received_power" if fname == "generic" else fname,
                            r_coef=r_coef, height=height, reduce_all=bool(reduce_all),
                            grad_mode="nan_parity" if nan_parity else "clean")
        return cfg, alpha, fname == "generic"

    def _generic_accumulate(self, cfg, fun, fun_args, fun_kwargs, xys, kinds, phis, fixed, grid, alpha, x0, device,
                            point_cls=Point, want_grad=False):
        """SURVEY §8 f1 — arbitrary `fun`: Z[t, r] = sum_c valid * fun(transmitter, receiver, path,
        interacting_objects, *fun_args, **fun_kwargs) (scene.py:1909-1916) with the paths materialised by `d2d_paths`
        (every path whose validity is non-zero) and `fun` evaluated ONCE PER ORDER on batched arguments (PointBatch,
        PathBatch, list of ObjectBatch: torch tensors on the device).  `point_cls` is the reference's receiver_cls /
        transmitter_cls: a custom class is constructed as point_cls(xy=<[n, 2] tensor>) for the grid end of the link.
        The summation runs in index order per (t, r) up to the reduction order of index_add_.

        want_grad (ImagePath): also d Z[t, r] / d grid[r] — what jax.grad(facc, argnums=1) gives for ANY fun
        (scene.py:1920-1925).  `fun` is differentiated by torch autograd on the materialised vertices (and on the end
        points it is handed), which yields the cotangents of every record's validity and vertices; `d2d_paths_vjp`
        pulls them back through the path construction and the validity logic.  Returns (Z, dZ [T, R, 2])."""
        rec = F.paths(cfg, xys, fixed, grid, kinds=kinds, phis=phis, alpha=alpha, x0=x0, min_valid=0.0, device=device)
        if want_grad:
            if cfg.method != "image":
                raise NotImplementedError("gradients of a generic `fun` are available for ImagePath only")
            rec["xys"] = rec["xys"].detach().clone().requires_grad_(True)
            rec["valid"] = rec["valid"].detach().clone().requires_grad_(True)
        fixed_t = torch.as_tensor(np.asarray(fixed, dtype=np.float32).reshape(-1, 2), device=device)
        T, R = fixed_t.shape[0], grid.shape[0]
        xys_t = torch.as_tensor(np.asarray(xys, dtype=np.float32).reshape(-1, 2, 2), device=device)
        kinds_t = torch.as_tensor(np.asarray(kinds, dtype=np.uint8), device=device)
        phis_t = torch.as_tensor(np.asarray(phis, dtype=np.float32), device=device)
        cands = {k: torch.as_tensor(F.candidates(len(self.objects), k, cfg.filter_nodes).astype(np.int64), device=device)
                 for k in range(max(cfg.min_order, 1), cfg.max_order + 1)}
        col0, c = {}, 0
        for k in range(cfg.min_order, cfg.max_order + 1):
            col0[k] = c
            c += 1 if k == 0 else int(cands[k].shape[0])
        Z = torch.zeros(T * R, dtype=torch.float32, device=device)
        for k in range(cfg.min_order, cfg.max_order + 1):
            sel = rec["order"] == k
            if not bool(sel.any()):
                continue
            fi, gi = rec["fixed"][sel].to(torch.int64), rec["grid"][sel]
            batch = PathBatch(rec["xys"][sel][:, : k + 2], rec["valid"][sel], rec["loss"][sel], k)
            # the end points a `fun` sees ARE the first / last path vertices (same values; under want_grad the same
            # autograd leaves, so that a `fun` written on transmitter.xy / receiver.xy is differentiated too)
            first, last = batch.xys[:, 0], batch.xys[:, k + 1]
            fixed_xy, grid_xy = (first, last) if cfg.grid_role == "receivers" else (last, first)
            fixed_pts = PointBatch(fixed_xy)
            grid_pts = PointBatch(grid_xy) if point_cls is Point else point_cls(xy=grid_xy)
            tx, rx = (fixed_pts, grid_pts) if cfg.grid_role == "receivers" else (grid_pts, fixed_pts)
            objs = []
            if k > 0:
                idx = cands[k][rec["candidate"][sel] - col0[k]]  # [n, k] object indices of every emitted path
                objs = [ObjectBatch(idx[:, i], xys_t[idx[:, i]], kinds_t[idx[:, i]], phis_t[idx[:, i]]) for i in range(k)]
            val = torch.as_tensor(fun(tx, rx, batch, objs, *fun_args, **dict(fun_kwargs or {})), dtype=torch.float32,
                                  device=device)
            val = val.expand(batch.valid.shape)
            Z = Z.index_add(0, fi * R + gi, batch.valid * val)
        if want_grad:
            # every record belongs to ONE (fixed point, grid point): the cotangent ones of sum(Z) reaches each of them
            if Z.requires_grad:
                vbar, xbar = torch.autograd.grad(Z.sum(), [rec["valid"], rec["xys"]], allow_unused=True)
            else:
                vbar = xbar = None
            vbar = torch.zeros_like(rec["valid"]) if vbar is None else vbar
            xbar = torch.zeros_like(rec["xys"]) if xbar is None else xbar
            # (the vertices handed to `fun` as path.xys include the two end points: rows 0 and order + 1)
            g = F.paths_vjp(cfg, xys, fixed, grid, rec, vbar, xbar, kinds=kinds, phis=phis, alpha=alpha, want=("grid",),
                            device=device)["grid"]
            Z = Z.detach().reshape(T, R)
            return (Z.sum(0), g.sum(0)) if cfg.reduce_all else (Z, g)
        Z = Z.reshape(T, R)
        return Z.sum(0) if cfg.reduce_all else Z

    def _tracked_tensors(self, grid_role, device):
        """torch tensors (objects [N,2,2], phis [N], fixed points [T,2]) that keep the autograd link of every
        coordinate given as a tensor with requires_grad (Point(xy=t), Wall(xys=t), RIS(phi=t)); None when nothing is
        tracked.  The analogue of calling the reference's entry points under jax.grad (scene.py:1272-1334 used by
        examples/plot_power_optimize.py:78-91)."""
        fixed_src = self.transmitters if grid_role == "receivers" else self.receivers
        pts = list(fixed_src.values())
        if not (any(o.tracked is not None or getattr(o, "tracked_phi", None) is not None for o in self.objects)
                or any(p.tracked is not None for p in pts)):
            return None

        def t(tracked, value):
            return (tracked if tracked is not None else torch.as_tensor(value)).to(device=device, dtype=torch.float32)

        if self.objects:
            xys = torch.stack([t(o.tracked, o.packed_xys()).reshape(2, 2) if not isinstance(o, Vertex)
                               else t(o.tracked, o.xy).reshape(1, 2).expand(2, 2) for o in self.objects])
            phis = torch.stack([t(getattr(o, "tracked_phi", None), np.float32(getattr(o, "phi", 0.0))).reshape(())
                                for o in self.objects])
        else:
            xys, phis = torch.zeros((0, 2, 2), device=device), torch.zeros((0,), device=device)
        fixed = torch.stack([t(p.tracked, p.xy).reshape(2) for p in pts]) if pts else torch.zeros((0, 2), device=device)
        return xys, phis, fixed

    def _x0(self, cfg: F.TraceConfig, key, device):
        """Initial guesses per candidate and restart (optimize.py:132, :173-177); `key` is an int seed or an explicit
        [C, max_order] (many == 1) / [C, many, max_order] table."""
        if cfg.method == "image" or cfg.max_order == 0:
            return None
        n_allowed = len(self.objects) - len(cfg.filter_nodes)
        C = sum(F.L.lib().d2d_candidates_count(len(self.objects), k, None, 0) if not cfg.filter_nodes else
                len(F.candidates(len(self.objects), k, cfg.filter_nodes)) for k in range(cfg.min_order, cfg.max_order + 1))
        del n_allowed
        if key is None:
            raise TypeError("FermatPath / MinPath need `key` (an int seed or an explicit x0 table)")
        shape = (C, cfg.max_order) if cfg.many == 1 else (C, cfg.many, cfg.max_order)
        if isinstance(key, (int, np.integer)):
            return np.random.default_rng(int(key)).random(shape, dtype=np.float32)
        x0 = np.asarray(key.detach().cpu() if hasattr(key, "detach") else key, dtype=np.float32)
        if x0.shape != shape:
            raise ValueError(f"x0 table must have shape {shape}, got {x0.shape}")
        return x0

    def _grid_call(self, grid_role, X, Y, fun, fun_args, fun_kwargs, reduce_all, grad, value_and_grad, path_cls,
                   path_cls_kwargs, min_order, max_order, order, filter_objects, key, kwargs, point_cls=Point):
        cfg, alpha, generic = self._config(grid_role, fun, fun_args, fun_kwargs, reduce_all, path_cls, path_cls_kwargs,
                                           min_order, max_order, order, filter_objects, kwargs)
        as_numpy = not isinstance(X, torch.Tensor)
        Xt = torch.as_tensor(np.asarray(X, dtype=np.float32)) if as_numpy else X
        Yt = torch.as_tensor(np.asarray(Y, dtype=np.float32)) if as_numpy else Y
        device = Xt.device if Xt.device.type == "cuda" else _default_device()
        shape = tuple(Xt.shape)
        grid = torch.stack((Xt.to(device, torch.float32), Yt.to(device, torch.float32)), dim=-1).reshape(-1, 2)  # dstack
        fixed_src = self.transmitters if grid_role == "receivers" else self.receivers
        names = list(fixed_src.keys())
        fixed = np.stack([fixed_src[k].xy for k in names]) if names else np.zeros((0, 2), np.float32)
        xys, kinds, phis = self.packed_objects()
        x0 = self._x0(cfg, key, device)
        if len(shape) == 2:
            import dataclasses
            cfg = dataclasses.replace(cfg, grid_cols=int(shape[1]))
        back = (lambda t: t.cpu().numpy()) if as_numpy else (lambda t: t.to(Xt.device))
        want_grad = grad or value_and_grad
        if generic and want_grad:
            Z, dZ = self._generic_accumulate(cfg, fun, fun_args, fun_kwargs, xys, kinds, phis, fixed, grid, alpha, x0,
                                             device, point_cls, want_grad=True)
            if reduce_all:
                Z, dZ = back(Z.reshape(shape)), back(dZ.reshape(*shape, 2))
                return (Z, dZ) if value_and_grad else dZ
            if value_and_grad:
                return ((k, (back(Z[i].reshape(shape)), back(dZ[i].reshape(*shape, 2)))) for i, k in enumerate(names))
            return ((k, back(dZ[i].reshape(*shape, 2))) for i, k in enumerate(names))
        tracked = None if generic else self._tracked_tensors(grid_role, device)
        diff = tracked is not None or (isinstance(alpha, torch.Tensor) and alpha.requires_grad) or \
            (isinstance(X, torch.Tensor) and (X.requires_grad or Y.requires_grad))
        if diff and not generic and not want_grad:
            # called under autograd (the analogue of jax.grad over the entry point): differentiable
            # device tensors come back, in the reference's structure
            if isinstance(X, torch.Tensor):  # keep the link to X / Y
                grid = torch.stack((X.to(device, torch.float32), Y.to(device, torch.float32)), dim=-1).reshape(-1, 2)
            xys_t, phis_t, fixed_t = tracked if tracked is not None else (
                torch.as_tensor(xys, device=device), torch.as_tensor(phis, device=device), torch.as_tensor(fixed, device=device))
            Z = F.power_map(xys_t, fixed_t, grid, cfg=cfg, phis=phis_t, alpha=alpha, kinds=kinds, x0=x0)
            if reduce_all:
                return Z.reshape(shape)
            return ((k, Z[i].reshape(shape)) for i, k in enumerate(names))
        if not want_grad:
            if generic:
                Z = self._generic_accumulate(cfg, fun, fun_args, fun_kwargs, xys, kinds, phis, fixed, grid, alpha, x0,
                                             device, point_cls)
            else:
                Z = F.power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, alpha=alpha, x0=x0, device=device)
            if reduce_all:
                return back(Z.reshape(shape))
            return ((k, back(Z[i].reshape(shape))) for i, k in enumerate(names))
        out = F.power_value_and_vjp(cfg, xys, fixed, grid, None, kinds=kinds, phis=phis, alpha=alpha, x0=x0,
                                    want=("grid",), device=device)
        Z, dZ = out["Z"], out["grid"]
        if reduce_all:
            Z, dZ = back(Z.reshape(shape)), back(dZ.reshape(*shape, 2))
            return (Z, dZ) if value_and_grad else dZ
        if value_and_grad:  # scene.py:1920-1923 value_and_grad overrides grad
            return ((k, (back(Z[i].reshape(shape)), back(dZ[i].reshape(*shape, 2)))) for i, k in enumerate(names))
        return ((k, back(dZ[i].reshape(*shape, 2))) for i, k in enumerate(names))

    def accumulate_on_receivers_grid_over_paths(
        self, X, Y, fun=utils.received_power, fun_args: tuple = (), fun_kwargs: Optional[Mapping[str, Any]] = None, *,
        reduce_all: bool = False, grad: bool = False, value_and_grad: bool = False, path_cls: type = ImagePath,
        path_cls_kwargs: Optional[Mapping[str, Any]] = None, receiver_cls: type = Point, min_order: int = 0,
        max_order: int = 1, order: Optional[int] = None, filter_objects=None, key=None, **kwargs: Any,
    ):
        """scene.py:1803-1953 — per-transmitter maps over a grid of receivers, keyed by transmitter name."""
        return self._grid_call("receivers", X, Y, fun, fun_args, fun_kwargs, reduce_all, grad, value_and_grad,
                               path_cls, path_cls_kwargs, min_order, max_order, order, filter_objects, key, kwargs,
                               receiver_cls)

    def accumulate_on_transmitters_grid_over_paths(
        self, X, Y, fun=utils.received_power, fun_args: tuple = (), fun_kwargs: Optional[Mapping[str, Any]] = None, *,
        reduce_all: bool = False, grad: bool = False, value_and_grad: bool = False, path_cls: type = ImagePath,
        path_cls_kwargs: Optional[Mapping[str, Any]] = None, transmitter_cls: type = Point, min_order: int = 0,
        max_order: int = 1, order: Optional[int] = None, filter_objects=None, key=None, **kwargs: Any,
    ):
        """scene.py:1489-1648 — per-receiver maps over a grid of transmitters, keyed by receiver name."""
        return self._grid_call("transmitters", X, Y, fun, fun_args, fun_kwargs, reduce_all, grad, value_and_grad,
                               path_cls, path_cls_kwargs, min_order, max_order, order, filter_objects, key, kwargs,
                               transmitter_cls)

    def _materialise(self, path_cls, path_cls_kwargs, min_order, max_order, order, filter_objects, key, kwargs,
                     emit_all: bool, min_valid: float = 0.5):
        cfg, alpha, _ = self._config("receivers", utils.received_power, (), None, False, path_cls, path_cls_kwargs,
                                     min_order, max_order, order, filter_objects, kwargs)
        device = _default_device()
        tx_names, rx_names = list(self.transmitters), list(self.receivers)
        if not tx_names or not rx_names:
            return cfg, tx_names, rx_names, None, []
        fixed = np.stack([self.transmitters[k].xy for k in tx_names])
        grid = np.stack([self.receivers[k].xy for k in rx_names])
        xys, kinds, phis = self.packed_objects()
        rec = F.paths(cfg, xys, fixed, torch.as_tensor(grid).to(device), kinds=kinds, phis=phis, alpha=alpha,
                      x0=self._x0(cfg, key, device), min_valid=min_valid, emit_all=emit_all, device=device)
        rec = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in rec.items()}
        cands = []
        for k in range(cfg.min_order, cfg.max_order + 1):
            cands.extend(list(F.candidates(len(self.objects), k, cfg.filter_nodes)))
        return cfg, tx_names, rx_names, rec, cands

    def all_paths(self, path_cls: type = ImagePath, path_cls_kwargs=None, min_order: int = 0, max_order: int = 1,
                  order: Optional[int] = None, filter_objects=None, *, key=None, **kwargs: Any
                  ) -> Iterator[tuple]:
        """scene.py:1156-1217 — (tx name, rx name, valid, path, path_candidate) for EVERY candidate of every
        transmitter-receiver pair, pairs in dictionary order, candidates in list order.  The paths are built and
        validated by the CUDA kernel (`d2d_paths`, emit_all); ``valid`` is a bool (hard logic) or a float32.
        NB the reference draws one PRNG key per (pair, candidate) (:1209-1212); this mirror uses the
        per-candidate x0 table of the grid methods for every pair."""
        cfg, tx_names, rx_names, rec, cands = self._materialise(path_cls, path_cls_kwargs, min_order, max_order, order,
                                                                filter_objects, key, dict(kwargs), emit_all=True)
        if rec is None:
            return
        hard = cfg.mode == "hard"
        for i in range(rec["fixed"].shape[0]):
            k = int(rec["order"][i])
            valid = bool(rec["valid"][i] != 0) if hard else np.float32(rec["valid"][i])
            yield (tx_names[int(rec["fixed"][i])], rx_names[int(rec["grid"][i])], valid,
                   path_cls(xys=rec["xys"][i, : k + 2].copy(), loss=rec["loss"][i]), cands[int(rec["candidate"][i])])

    def all_valid_paths(self, approx: Optional[bool] = None, **kwargs: Any) -> Iterator[tuple]:
        """scene.py:1219-1248 — the paths of all_paths for which logic.is_true(valid) holds (valid > 0.5 with
        approximation, valid itself otherwise, logic.py:542-563), as (tx name, rx name, path, path_candidate).
        Only those paths leave the GPU (compacted by the kernel)."""
        path_cls = kwargs.pop("path_cls", ImagePath)
        kw = dict(approx=approx, **{k: kwargs.pop(k) for k in list(kwargs) if k in ("alpha", "function", "tol", "patch")})
        cfg, tx_names, rx_names, rec, cands = self._materialise(
            path_cls, kwargs.pop("path_cls_kwargs", None), kwargs.pop("min_order", 0), kwargs.pop("max_order", 1),
            kwargs.pop("order", None), kwargs.pop("filter_objects", None), kwargs.pop("key", None), kw, emit_all=False,
            min_valid=0.5)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        if rec is None:
            return
        for i in range(rec["fixed"].shape[0]):
            k = int(rec["order"][i])
            yield (tx_names[int(rec["fixed"][i])], rx_names[int(rec["grid"][i])],
                   path_cls(xys=rec["xys"][i, : k + 2].copy(), loss=rec["loss"][i]), cands[int(rec["candidate"][i])])

    def accumulate_over_paths(self, fun=utils.received_power, fun_args: tuple = (),
                              fun_kwargs: Optional[Mapping[str, Any]] = None, *, reduce_all: bool = False,
                              path_cls: type = ImagePath, path_cls_kwargs=None, min_order: int = 0, max_order: int = 1,
                              order: Optional[int] = None, filter_objects=None, key=None, **kwargs: Any):
        """
        scene.py:1272-1334 — point-to-point: (tx name, rx name, accumulated value) for every pair, or
        their sum.  Runs as a receivers "grid" holding the scene's receivers.
        NB the reference draws one PRNG key per (pair, candidate) here (scene.py:1209-1212); this mirror
        uses the per-candidate table of the grid methods for every pair.
        """
        cfg, alpha, generic = self._config("receivers", fun, fun_args, fun_kwargs, False, path_cls, path_cls_kwargs,
                                           min_order, max_order, order, filter_objects, kwargs)
        device = _default_device()
        tx_names, rx_names = list(self.transmitters), list(self.receivers)
        fixed = np.stack([self.transmitters[k].xy for k in tx_names]) if tx_names else np.zeros((0, 2), np.float32)
        grid = np.stack([self.receivers[k].xy for k in rx_names]) if rx_names else np.zeros((0, 2), np.float32)
        xys, kinds, phis = self.packed_objects()
        x0 = self._x0(cfg, key, device)
        if not tx_names or not rx_names:
            return np.float32(0.0) if reduce_all else iter(())
        tracked = None if generic else self._tracked_tensors("receivers", device)
        rx_tracked = any(p.tracked is not None for p in self.receivers.values())
        if tracked is not None or rx_tracked or (isinstance(alpha, torch.Tensor) and alpha.requires_grad):
            # under autograd (jax.value_and_grad(loss) in examples/plot_power_optimize.py:78-91): differentiable
            # 0-dim device tensors, accumulated like the reference (acc = 0.0; Z = Z + p, scene.py:1305-1334)
            xys_t, phis_t, fixed_t = tracked if tracked is not None else (
                torch.as_tensor(xys, device=device), torch.as_tensor(phis, device=device), torch.as_tensor(fixed, device=device))
            grid_t = torch.stack([(p.tracked if p.tracked is not None else torch.as_tensor(p.xy)).to(device, torch.float32)
                                  for p in self.receivers.values()])
            Zt = F.power_map(xys_t, fixed_t, grid_t, cfg=cfg, phis=phis_t, alpha=alpha, kinds=kinds, x0=x0)
            if reduce_all:
                total = torch.zeros((), device=device)
                for i in range(len(tx_names)):
                    for j in range(len(rx_names)):
                        total = total + Zt[i, j]
                return total
            return ((t, r, Zt[i, j]) for i, t in enumerate(tx_names) for j, r in enumerate(rx_names))
        if generic:
            Z = self._generic_accumulate(cfg, fun, fun_args, fun_kwargs, xys, kinds, phis, fixed,
                                         torch.as_tensor(grid).to(device), alpha, x0, device).cpu().numpy()
        else:
            Z = F.power_fwd(cfg, xys, fixed, torch.as_tensor(grid), kinds=kinds, phis=phis, alpha=alpha, x0=x0,
                            device=device).cpu().numpy()
        if reduce_all:
            total = np.float32(0.0)
            for i in range(len(tx_names)):
                for j in range(len(rx_names)):
                    total = np.float32(total + Z[i, j])
            return total
        return ((t, r, Z[i, j]) for i, t in enumerate(tx_names) for j, r in enumerate(rx_names))
