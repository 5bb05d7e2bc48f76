"""
Host-side mirror of the scene objects and path-class tags of differt2d/geometry.py.

Objects are plain float32 containers: all geometry (normals, images, intersections, residuals —
geometry.py:82-230, 352-721, 811-1288) is evaluated by the CUDA kernels from the packed arrays.
The path classes are *tags* selecting the kernel's construction method.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib as L


def _xy(a, shape):
    arr = np.asarray(a.detach().cpu() if hasattr(a, "detach") else a, dtype=np.float32)
    if arr.shape != shape:
        raise ValueError(f"expected shape {shape}, got {arr.shape}")
    return arr


def _tracked(a):
    """The torch tensor behind a coordinate array when autograd follows it (the analogue of a traced jnp array inside
    jax.grad: `Point(xy=tx_coords)` in examples/plot_power_optimize.py:84), else None."""
    return a if (hasattr(a, "requires_grad") and a.requires_grad) else None


@dataclass(frozen=True)
class Point:
    """geometry.py:270-348"""

    xy: np.ndarray = field(default_factory=lambda: np.zeros(2, np.float32))

    def __post_init__(self):
        object.__setattr__(self, "tracked", _tracked(self.xy))
        object.__setattr__(self, "xy", _xy(self.xy, (2,)))

    def bounding_box(self):
        return np.vstack([self.xy, self.xy])


@dataclass(frozen=True)
class Vertex(Point):
    """geometry.py:352-431 — corner diffraction: no unknown, never occludes, zero residual."""

    KIND = L.KIND_VERTEX

    @staticmethod
    def parameters_count() -> int:
        return 0

    def packed_xys(self):
        return np.stack([self.xy, self.xy])


@dataclass(frozen=True)
class Ray:
    """geometry.py:434-539"""

    xys: np.ndarray = field(default_factory=lambda: np.array([[0.0, 0.0], [1.0, 0.0]], np.float32))

    def __post_init__(self):
        object.__setattr__(self, "tracked", _tracked(self.xys))
        object.__setattr__(self, "xys", _xy(self.xys, (2, 2)))

    def origin(self):
        return self.xys[0]

    def dest(self):
        return self.xys[1]

    def t(self):
        return self.xys[1] - self.xys[0]

    def bounding_box(self):
        return np.vstack([self.xys.min(axis=0), self.xys.max(axis=0)])


@dataclass(frozen=True)
class Wall(Ray):
    """geometry.py:540-680"""

    KIND = L.KIND_WALL

    @staticmethod
    def parameters_count() -> int:
        return 1

    def packed_xys(self):
        return self.xys

    def get_vertices(self):
        return Vertex(xy=self.xys[0]), Vertex(xy=self.xys[1])


@dataclass(frozen=True)
class RIS(Wall):
    """geometry.py:683-721 — reflection angle ``phi`` w.r.t. the normal, default pi/4."""

    phi: float = float(np.pi / 4)
    KIND = L.KIND_RIS

    def __post_init__(self):
        super().__post_init__()
        object.__setattr__(self, "tracked_phi", _tracked(self.phi))
        object.__setattr__(self, "phi", float(self.phi))


class Path:
    """geometry.py:724-973 — a materialised path: ``xys`` [(order + 2), 2] (TX, interaction points, RX) and its
    ``loss``.  The subclasses double as the TAGS that select the kernels' construction method (``path_cls=``);
    the base class itself (every object sampled at t = 0.5 in the reference) is not fused."""

    METHOD = None

    def __init__(self, xys=None, loss=0.0):
        self.xys = np.zeros((2, 2), np.float32) if xys is None else np.asarray(xys, dtype=np.float32)
        self.loss = np.float32(loss)

    def length(self) -> np.float32:
        """geometry.py:811-819 / path_length :176-203 (eps added to every segment vector, fp32)."""
        d = (self.xys[1:] - self.xys[:-1]) + np.float32(1.1920929e-07)
        return np.float32(np.sqrt((d * d).sum(-1, dtype=np.float32)).sum(dtype=np.float32))

    def __repr__(self):
        return f"{type(self).__name__}(xys={self.xys.tolist()}, loss={float(self.loss):.3g})"


class ImagePath(Path):
    """geometry.py:1013-1114"""

    METHOD = "image"


class FermatPath(Path):
    """geometry.py:1117-1204"""

    METHOD = "fermat"


class MinPath(Path):
    """geometry.py:1207-1288"""

    METHOD = "minpath"


class PathBatch:
    """The ``path`` argument of a generic `fun`: all emitted paths of ONE order as device tensors.  The reference
    calls ``fun(transmitter, receiver, path, interacting_objects, *fun_args, **fun_kwargs)`` once per path
    (scene.py:1909-1916, :1318-1325); here the same call is made once per ORDER with every argument batched over the
    n emitted paths: ``path.xys`` f32[n, order + 2, 2], ``path.valid`` / ``path.loss`` f32[n], ``path.order`` int,
    ``path.length()`` f32[n] (geometry.py:811-819)."""

    def __init__(self, xys, valid, loss, order):
        self.xys, self.valid, self.loss, self.order = xys, valid, loss, int(order)

    def length(self):
        d = (self.xys[:, 1:] - self.xys[:, :-1]) + 1.1920929e-07
        return d.square().sum(-1).sqrt().sum(-1)


class PointBatch:
    """The ``transmitter`` / ``receiver`` argument of a generic `fun`: ``xy`` f32[n, 2] (one row per emitted path)."""

    def __init__(self, xy):
        self.xy = xy


class ObjectBatch:
    """One entry of the ``interacting_objects`` list of a generic `fun` (``len(interacting_objects)`` is the order, as
    utils.received_power uses it, utils.py:52): the i-th object of every emitted path — ``index`` i64[n] into
    Scene.objects, ``xys`` f32[n, 2, 2], ``kind`` u8[n] (0 Wall, 1 RIS, 2 Vertex), ``phi`` f32[n]."""

    def __init__(self, index, xys, kind, phi):
        self.index, self.xys, self.kind, self.phi = index, xys, kind, phi
