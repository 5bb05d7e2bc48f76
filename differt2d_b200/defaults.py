"""Default values — differt2d/defaults.py:3-22."""

DEFAULT_ALPHA: float = 100.0
DEFAULT_PATCH: float = 0.0
DEFAULT_R_COEF: float = 0.5
DEFAULT_HEIGHT: float = 0.1
