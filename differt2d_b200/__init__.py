"""
differt2d_b200 — B200-native (sm_100a CUDA) implementation of DiffeRT2d's receiver-grid
path-tracing hot path, behind the reference's own Python API names.  See DESIGN.md.
"""

from . import functional, logic, optimizers, utils  # noqa: F401
from ._lib import D2DError  # noqa: F401
from .defaults import DEFAULT_ALPHA, DEFAULT_HEIGHT, DEFAULT_PATCH, DEFAULT_R_COEF  # noqa: F401
from .functional import TraceConfig, power_bwd, power_fwd, power_map  # noqa: F401
from .geometry import RIS, FermatPath, ImagePath, MinPath, Path, Point, Ray, Vertex, Wall  # noqa: F401
from .scene import Scene  # noqa: F401
from .utils import P0, length_squared, received_power  # noqa: F401

__version__ = "0.1.0"
