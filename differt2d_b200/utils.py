"""Host mirror of differt2d/utils.py: the path functions that are fused into the kernels."""

from __future__ import annotations

from .defaults import DEFAULT_HEIGHT, DEFAULT_R_COEF

P0: float = 100.0  # utils.py:12


def received_power(transmitter=None, receiver=None, path=None, interacting_objects=None,
                   r_coef: float = DEFAULT_R_COEF, height: float = DEFAULT_HEIGHT):
    """
    utils.py:16-54 — ``r_coef**n / (height**2 + length**2)``.  Passed as ``fun`` it selects the fused
    D2D_FUN_RECEIVED_POWER epilogue; it is never called per path on the host.
    """
    raise TypeError("received_power is evaluated inside the CUDA kernels; pass it as `fun`")


def length_squared(transmitter=None, receiver=None, path=None, interacting_objects=None):
    """``path.length() ** 2`` (tests/test_scene.py:444-445, :488-489) — D2D_FUN_LENGTH_SQUARED."""
    raise TypeError("length_squared is evaluated inside the CUDA kernels; pass it as `fun`")
