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


@dataclass(frozen=True)
class Point:
    """geometry.py:270-348"""

    xy: np.ndarray = field(default_factory=lambda: np.zeros(2, np.float32))

    def __post_init__(self):
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


class Path:
    """geometry.py:724-973 — base tag (every object sampled at t = 0.5 in the reference; not fused here)."""

    METHOD = None


class ImagePath(Path):
    """geometry.py:1013-1114"""

    METHOD = "image"


class FermatPath(Path):
    """geometry.py:1117-1204"""

    METHOD = "fermat"


class MinPath(Path):
    """geometry.py:1207-1288"""

    METHOD = "minpath"
