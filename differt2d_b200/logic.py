"""
Host mirror of differt2d/logic.py's process-wide switch (logic.py:58-215).  The numerical part of
logic.py (activation, min/max folds, comparisons) runs inside the CUDA kernels (csrc/d2d_device.cuh);
this module only resolves ``approx=None`` against the global flag, as logic.py:333-334 does.
"""

from __future__ import annotations

import os
import threading
from contextlib import contextmanager
from typing import Optional

_LOCK = threading.RLock()
ENABLE_APPROX: bool = "ENABLE_APPROX" in os.environ  # logic.py:58


def set_approx(enable: bool) -> None:  # logic.py:68-91
    global ENABLE_APPROX
    with _LOCK:
        ENABLE_APPROX = bool(enable)


@contextmanager
def enable_approx(enable: bool = True):  # logic.py:94-196
    global ENABLE_APPROX
    with _LOCK:  # set and restore under the lock, but never hold it across the with-body (other threads must be
        prev = ENABLE_APPROX  # able to call set_approx / enable_approx meanwhile; like the reference's jax.config
        ENABLE_APPROX = bool(enable)  # flag, the switch itself is process-wide)
    try:
        yield
    finally:
        with _LOCK:
            ENABLE_APPROX = prev


def disable_approx(disable: bool = True):  # logic.py:199-215
    return enable_approx(not disable)


def sigmoid(x, alpha):  # logic.py:218-235 — marker; evaluated on the GPU as D2D_MODE_SIGMOID
    raise TypeError("differt2d_b200.logic.sigmoid is a marker selecting the activation inside the CUDA kernels")


def hard_sigmoid(x, alpha):  # logic.py:238-255 — marker; D2D_MODE_HARD_SIGMOID
    raise TypeError("differt2d_b200.logic.hard_sigmoid is a marker selecting the activation inside the CUDA kernels")


def resolve_mode(approx: Optional[bool], function=None) -> str:
    """('hard' | 'hard_sigmoid' | 'sigmoid') from the reference's (approx, function) keywords."""
    if approx is None:
        approx = ENABLE_APPROX
    if not approx:
        return "hard"
    if function is None or function is hard_sigmoid or function == "hard_sigmoid":
        return "hard_sigmoid"  # logic.py:266 default
    if function is sigmoid or function == "sigmoid":
        return "sigmoid"
    raise NotImplementedError(
        "only the two activations shipped by the reference (sigmoid, hard_sigmoid) are fused into the kernels"
    )
