"""
Optimiser descriptors for ``path_cls_kwargs={"optimizer": ...}`` of FermatPath / MinPath.

The reference hands an ``optax.GradientTransformation`` to ``optimize.minimize`` (optimize.py:44-97; default
``optax.adam(learning_rate=0.1)``, optimize.py:83).  The solver here runs inside the CUDA kernels (in registers, per
path), so an arbitrary optax object cannot be executed; the transformations that are fused are described by these
plain records, with optax's own names and defaults:

    adam(learning_rate=0.1, b1=0.9, b2=0.999, eps=1e-8)    optax.adam   (bias-corrected, eps_root = 0)
    sgd(learning_rate, momentum=None)                      optax.sgd    (trace decay = momentum, no Nesterov)
    newton()                                               NOT optax: damped Newton iterations with an implicitly
                                                           differentiated fixed point (csrc/d2d_newton.cuh) — a fast
                                                           mode for callers who want the converged path, not the scan
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Optimizer:
    kind: str
    learning_rate: float = 0.1
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8


def adam(learning_rate: float = 0.1, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> Optimizer:
    return Optimizer("adam", float(learning_rate), float(b1), float(b2), float(eps))


def sgd(learning_rate: float, momentum=None) -> Optimizer:
    return Optimizer("sgd", float(learning_rate), float(momentum or 0.0))


def newton() -> Optimizer:
    return Optimizer("newton")
