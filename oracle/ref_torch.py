"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A literal fp32 restatement, on torch-CPU tensors, of the DiffeRT2d receiver-grid
path-tracing hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.

Why torch-CPU and not the reference itself: JAX / equinox / optax / differt-core
are not installable in this image (no network, no wheels), see DESIGN.md.  Every
function below cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).  The op ORDER is the reference's; each torch op is an IEEE
fp32 op executed separately (eager mode: no FMA contraction), so this file also
defines the *canonical rounding* that the hard-logic masks of the CUDA kernels
are compared against bit-for-bit.

Everything is vectorised over leading "grid" dimensions the way ``jax.vmap``
vectorises the reference (``scene.py:1927-1930``), with a Python loop over the
path candidates exactly as the reference's ``facc`` does (``scene.py:1892-1918``).
Reverse-mode gradients come from torch autograd, whose VJP rules for
``minimum/maximum`` (ties split 0.5/0.5), full ``min`` reductions (ties split
evenly), ``where`` and ``sqrt`` coincide with JAX's; ``norm`` is spelled
``sqrt(x*x+y*y)`` so that the NaN-at-zero behaviour of the reference is kept.

Parity pinning: see tests/test_oracle_kat.py (reference KATs: tests/test_geometry.py,
tests/test_scene.py, tests/test_utils.py, tests/test_logic.py, doctests).  What no
reference test pins (candidate ORDER, non-LOS power maps, gradients through
reflections) is "parity unpinned" and oracle-defined; DESIGN.md lists it.
"""

from __future__ import annotations

import math
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

F32 = torch.float32  # the WORKING dtype (name kept from the fp32-only version); see `precision`
_NP = np.float32
EPS32 = float(np.finfo(np.float32).eps)  # geometry.py:200  jnp.finfo(dtype).eps — the fp32 value in BOTH precisions

KIND_WALL, KIND_RIS, KIND_VERTEX = 0, 1, 2

# defaults.py:3-22, utils.py:12
DEFAULT_ALPHA = 100.0
DEFAULT_PATCH = 0.0
DEFAULT_R_COEF = 0.5
DEFAULT_HEIGHT = 0.1
P0 = 100.0


# Gradient-semantics switch (forward VALUES are identical in both settings).
#   CLEAN = False : the reference's literal op graph; its reverse mode yields NaN wherever a
#                   masked branch divides by zero (geometry.py:1105 single ``where``; ``normalize``
#                   of a zero vector, :227-230; ``alpha * inf`` for parallel segments, :166).
#   CLEAN = True  : the same graph with the standard "double where" guard at those three places,
#                   i.e. masked branches are constants with zero cotangent.  This is the gradient
#                   the CUDA backward kernel defines (DESIGN.md "NaN semantics").
CLEAN = False


class precision:
    """
    ``with precision("f64"):`` evaluates the SAME function of the SAME fp32 inputs in binary64 (inputs are
    rounded to fp32 first, then widened; literal constants 0.005, 1e-2, eps, r_coef**k, height**2 keep their
    fp32 values).  Not a model of the reference (which is fp32 throughout, docs/source/contributing/
    internals.md:7-10): it is the third leg of the parity triangulation — where the fp32 and fp64 evaluations
    of the oracle disagree, the quantity is ill-conditioned in fp32 and no two fp32 implementations (XLA's FMA
    contraction, this oracle, the CUDA kernels) can be expected to agree there (tests/test_gpu_parity.py).
    """

    def __init__(self, name: str = "f64"):
        assert name in ("f32", "f64")
        self.dt, self.np = (torch.float64, np.float64) if name == "f64" else (torch.float32, np.float32)

    def __enter__(self):
        global F32, _NP
        self.prev = (F32, _NP)
        F32, _NP = self.dt, self.np

    def __exit__(self, *a):
        global F32, _NP
        F32, _NP = self.prev


def _c(x: float) -> torch.Tensor:
    """A literal constant of the reference: its fp32 value, in the working dtype."""
    return torch.tensor(float(np.float32(x)), dtype=F32)


class clean_gradients:
    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        global CLEAN
        self.prev = CLEAN
        CLEAN = self.on

    def __exit__(self, *a):
        global CLEAN
        CLEAN = self.prev


def _t(x, requires_grad: bool = False) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(F32)
    return torch.tensor(np.asarray(x, dtype=np.float32), dtype=F32, requires_grad=requires_grad)  # fp32 values, widened


# ----------------------------------------------------------------------------
# logic.py
# ----------------------------------------------------------------------------
class Logic:
    """logic.py:218-617 — hard (bool) or smooth (float in [0,1]) logic, chosen statically."""

    def __init__(self, approx: bool, alpha=DEFAULT_ALPHA, function: str = "hard_sigmoid"):
        assert function in ("hard_sigmoid", "sigmoid")
        self.approx = bool(approx)
        self.alpha = alpha if isinstance(alpha, torch.Tensor) else torch.tensor(float(alpha), dtype=F32)
        self.function = function

    # logic.py:218-255, 260-312
    def activation(self, x: torch.Tensor) -> torch.Tensor:
        z = self.alpha * x
        if self.function == "sigmoid":
            return torch.sigmoid(z)  # jax.nn.sigmoid(alpha * x)
        # jax.nn.hard_sigmoid(z) = relu6(z + 3) / 6 ; relu6 = minimum(maximum(x, 0), 6)
        zero = torch.zeros((), dtype=F32)
        six = torch.full((), 6.0, dtype=F32)
        return torch.minimum(torch.maximum(z + 3.0, zero), six) / 6.0

    def true_value(self):  # logic.py:575-595
        return torch.tensor(1.0, dtype=F32) if self.approx else torch.tensor(True)

    def false_value(self):  # logic.py:598-617
        return torch.tensor(0.0, dtype=F32) if self.approx else torch.tensor(False)

    def lor(self, x, y):  # logic.py:315-335
        return torch.maximum(x, y) if self.approx else torch.logical_or(x, y)

    def land(self, x, y):  # logic.py:338-358
        return torch.minimum(x, y) if self.approx else torch.logical_and(x, y)

    def lnot(self, x):  # logic.py:361-377
        return torch.sub(1.0, x) if self.approx else torch.logical_not(x)

    def gt(self, x, y):  # logic.py:380-404
        return self.activation(x - y) if self.approx else torch.gt(x, y)

    def ge(self, x, y):  # logic.py:407-433
        return self.activation(x - y) if self.approx else torch.ge(x, y)

    def lt(self, x, y):  # logic.py:436-460
        return self.activation(y - x) if self.approx else torch.lt(x, y)

    def le(self, x, y):  # logic.py:463-487
        return self.activation(y - x) if self.approx else torch.le(x, y)

    def lall(self, *xs):  # logic.py:490-513 — jnp.min(jnp.asarray(x)) / jnp.all
        xs = torch.broadcast_tensors(*xs)
        arr = torch.stack(xs, dim=0)
        if self.approx:
            # full reduction over the stacked axis; ties split evenly (== JAX reduce_min rule)
            return _min_even(arr)
        return torch.all(arr, dim=0)

    def lany(self, *xs):  # logic.py:516-539
        xs = torch.broadcast_tensors(*xs)
        arr = torch.stack(xs, dim=0)
        if self.approx:
            return -_min_even(-arr)
        return torch.any(arr, dim=0)

    def is_true(self, x, tol=0.5):  # logic.py:542-556
        return torch.gt(x, 1.0 - tol) if self.approx else x

    def is_false(self, x, tol=0.5):  # logic.py:559-572
        return torch.lt(x, tol) if self.approx else torch.logical_not(x)


class _MinEven(torch.autograd.Function):
    """min over dim 0 with JAX's reduce_min VJP: cotangent split evenly among ties."""

    @staticmethod
    def forward(ctx, arr):
        ans = torch.amin(arr, dim=0)
        # amin returns NaN if any NaN is present (like jnp.min)
        ctx.save_for_backward(arr, ans)
        return ans

    @staticmethod
    def backward(ctx, g):
        arr, ans = ctx.saved_tensors
        loc = (arr == ans.unsqueeze(0)).to(arr.dtype)
        counts = loc.sum(dim=0, keepdim=True)
        return g.unsqueeze(0) * loc / counts


def _min_even(arr):
    return _MinEven.apply(arr)


class _SqrtDiff(torch.autograd.Function):
    """
    Correctly-rounded fp32 sqrt.  torch.sqrt on CPU goes through a vector math library that is
    NOT correctly rounded (1-ulp misses were observed against numpy / C sqrtf), whereas XLA-CPU,
    CUDA ``__fsqrt_rn`` and C ``sqrtf`` all are.  VJP = g * (0.5 / ans), JAX's rule for sqrt
    (inf/NaN at 0 preserved); the backward re-evaluates through ``_sqrt`` so that it is itself
    differentiable (needed to differentiate through the Adam iterations of Fermat/MinPath).
    """

    @staticmethod
    def forward(ctx, x):
        y = torch.from_numpy(np.asarray(np.sqrt(x.detach().contiguous().numpy()), dtype=_NP))
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        with torch.enable_grad():
            y = _sqrt(x)
        return g * (0.5 / y)


def _sqrt(x):
    if x.requires_grad and torch.is_grad_enabled():
        return _SqrtDiff.apply(x)
    return torch.from_numpy(np.asarray(np.sqrt(x.detach().contiguous().numpy()), dtype=_NP))


# ----------------------------------------------------------------------------
# geometry.py free functions
# ----------------------------------------------------------------------------
def dot2(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]


def norm2(v):
    # jnp.linalg.norm == sqrt(sum(x*x)); spelled out so d/dv at 0 is NaN as in JAX
    return _sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1])


def normalize(v):
    """geometry.py:206-230"""
    if CLEAN:
        sq = v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]
        zero = sq == 0.0
        length = _sqrt(torch.where(zero, torch.ones_like(sq), sq))
        length = torch.where(zero | (length == 0.0), torch.ones_like(length), length)
        return v / length.unsqueeze(-1), length
    length = norm2(v)
    length = torch.where(length == 0.0, torch.ones_like(length), length)
    return v / length.unsqueeze(-1), length


def path_length(points):
    """geometry.py:176-203 — points [..., n, 2]"""
    vectors = points[..., 1:, :] - points[..., :-1, :]
    vectors = vectors + EPS32
    lengths = _sqrt(vectors[..., 0] * vectors[..., 0] + vectors[..., 1] * vectors[..., 1])
    total = lengths[..., 0]
    for i in range(1, lengths.shape[-1]):
        total = total + lengths[..., i]
    return total


def segments_intersect(P1, P2, P3, P4, logic: Logic, tol=0.005):
    """geometry.py:82-173 (Graphics Gems III).  All points broadcast over leading dims."""
    hi = _c(np.float32(1.0) + np.float32(tol))  # fl32(1 + tol): a constant folded in fp32 by the reference's trace
    tol = _c(tol)
    A = P2 - P1
    B = P3 - P4
    C = P1 - P3
    a = B[..., 1] * C[..., 0] - B[..., 0] * C[..., 1]
    b = A[..., 0] * C[..., 1] - A[..., 1] * C[..., 0]
    d = A[..., 1] * B[..., 0] - A[..., 0] * B[..., 1]

    def test(num, den):
        den_is_zero = den == 0.0
        den = torch.where(den_is_zero, torch.ones_like(den), den)
        if CLEAN and logic.approx:
            t = num / den
            res = logic.land(logic.ge(t, -tol), logic.le(t, hi))
            # t = +inf gives ge -> 1, le -> 0, and -> 0 : a constant
            return torch.where(den_is_zero, torch.zeros_like(res), res)
        t = torch.where(den_is_zero, torch.full_like(den, float("inf")), num / den)
        return logic.land(logic.ge(t, -tol), logic.le(t, hi))

    return logic.land(test(a, d), test(b, d))


# ----------------------------------------------------------------------------
# Scene objects (Wall / RIS / Vertex) as arrays
# ----------------------------------------------------------------------------
class OScene:
    """
    Array form of a reference ``Scene``: objects ``xys[N,2,2]`` (a Vertex repeats its
    xy in both rows), ``kinds[N]``, ``phis[N]`` (RIS only), named transmitters/receivers.
    """

    def __init__(self, xys, kinds=None, phis=None, transmitters=None, receivers=None):
        self.xys = _t(xys).reshape(-1, 2, 2)
        n = self.xys.shape[0]
        self.kinds = [KIND_WALL] * n if kinds is None else [int(k) for k in kinds]
        self.phis = _t(np.zeros(n, np.float32) if phis is None else phis)
        self.transmitters = {k: _t(v) for k, v in (transmitters or {}).items()}
        self.receivers = {k: _t(v) for k, v in (receivers or {}).items()}

    @property
    def n(self):
        return self.xys.shape[0]

    def cast(self):
        """The same scene with its tensors in the working dtype (see `precision`)."""
        if self.xys.dtype == F32:
            return self
        return OScene(self.xys.detach(), self.kinds, self.phis.detach(), self.transmitters, self.receivers)

    # Ray.origin/dest/t : geometry.py:458-487
    def origin(self, j):
        return self.xys[j, 0]

    def dest(self, j):
        return self.xys[j, 1]

    def t(self, j):
        return self.xys[j, 1] - self.xys[j, 0]

    def normal(self, j):
        """geometry.py:561-573"""
        t = self.t(j)
        n = torch.stack([t[1], -t[0]])
        n, _ = normalize(n)
        return n

    def parameters_count(self, j):
        return 0 if self.kinds[j] == KIND_VERTEX else 1  # geometry.py:381, 578

    def parametric_to_cartesian(self, j, s):
        """geometry.py:581-587 (Wall/RIS), :383-389 (Vertex)"""
        if self.kinds[j] == KIND_VERTEX:
            return self.xys[j, 0]
        return self.origin(j) + s.unsqueeze(-1) * self.t(j)

    def cartesian_to_parametric(self, j, p):
        """geometry.py:589-598"""
        other = p - self.origin(j)
        t = self.t(j)
        sq = dot2(t, t)
        sq = torch.where(sq == 0.0, torch.ones_like(sq), sq)
        return dot2(t, other) / sq

    def contains_parametric(self, j, s, logic: Logic):
        """geometry.py:600-621 ; Vertex :396-403"""
        if self.kinds[j] == KIND_VERTEX:
            return logic.true_value()
        zero = torch.tensor(0.0, dtype=F32)
        one = torch.tensor(1.0, dtype=F32)
        return logic.land(logic.ge(s, zero), logic.le(s, one))

    def intersects_cartesian(self, j, r0, r1, logic: Logic, patch=DEFAULT_PATCH):
        """geometry.py:623-639 ; Vertex :405-414"""
        if self.kinds[j] == KIND_VERTEX:
            return logic.false_value()
        t = self.t(j)
        return segments_intersect(self.origin(j) - patch * t, self.dest(j) + patch * t, r0, r1, logic)

    def evaluate_cartesian(self, j, a, b, c):
        """Wall geometry.py:641-650 ; RIS :698-711 ; Vertex :416-419"""
        kind = self.kinds[j]
        if kind == KIND_VERTEX:
            return torch.zeros(torch.broadcast_shapes(a.shape[:-1], b.shape[:-1], c.shape[:-1]), dtype=F32)
        n = self.normal(j)
        if kind == KIND_WALL:
            i = b - a
            r = c - b
            i, _ = normalize(i)
            r, _ = normalize(r)
            e = r - (i - (2 * dot2(i, n)).unsqueeze(-1) * n)
            return dot2(e, e)
        r = c - b
        r, _ = normalize(r)
        mr = -r
        sin_a = mr[..., 0] * n[1] - mr[..., 1] * n[0]  # jnp.cross(-r, n)
        cos_a = dot2(mr, n)
        sin_p = torch.sin(self.phis[j])
        cos_p = torch.cos(self.phis[j])
        return (sin_a - sin_p) ** 2 + (cos_a - cos_p) ** 2

    def image_of(self, j, p):
        """geometry.py:652-670"""
        i = p - self.origin(j)
        n = self.normal(j)
        return p - (2.0 * dot2(i, n)).unsqueeze(-1) * n

    def bounding_box(self):
        """scene.py:1023-1036"""
        pts = [self.xys.reshape(-1, 2)]
        pts += [v.reshape(1, 2) for v in self.transmitters.values()]
        pts += [v.reshape(1, 2) for v in self.receivers.values()]
        pts = torch.cat(pts, dim=0)
        return torch.stack([pts.min(dim=0).values, pts.max(dim=0).values])

    def grid(self, m=50, n=None):
        """abc.py:59-81 — returns X, Y of shape (n, m), meshgrid 'xy'."""
        bb = self.bounding_box().detach().numpy()
        if n is None:
            n = m
        x = np.linspace(bb[0, 0], bb[1, 0], m, dtype=np.float32)
        y = np.linspace(bb[0, 1], bb[1, 1], n, dtype=np.float32)
        X, Y = np.meshgrid(x, y)
        return torch.from_numpy(X.copy()), torch.from_numpy(Y.copy())


# ----------------------------------------------------------------------------
# scene.py:122-175 — candidates (differt-core 0.0.31 CompleteGraph/DiGraph.all_paths restated)
# ----------------------------------------------------------------------------
def all_path_candidates(num_nodes, min_order=0, max_order=1, *, order=None, filter_nodes=None):
    """
    All sequences of object indices of length k in [min_order, max_order] such that no two
    consecutive indices are equal (complete graph without self loops, from=N, to=N+1,
    depth=k+2, end points stripped — scene.py:154-174), nodes in ``filter_nodes`` never
    visited (scene.py:158-160), in lexicographic (depth-first, ascending-neighbour) order.
    The ORDER is parity-unpinned: differt-core is not vendored; its documented example
    ``generate_all_path_candidates(3, 2)`` is lexicographic.
    """
    if order is not None:
        min_order = max_order = order
    allowed = [i for i in range(num_nodes) if not (filter_nodes and i in filter_nodes)]
    out = []
    for k in range(min_order, max_order + 1):
        if k == 0:
            out.append(np.empty((0,), dtype=np.int32))
            continue
        stack = [[a] for a in reversed(allowed)]
        while stack:
            seq = stack.pop()
            if len(seq) == k:
                out.append(np.asarray(seq, dtype=np.int32))
                continue
            for a in reversed(allowed):
                if a != seq[-1]:
                    stack.append(seq + [a])
    return out


# ----------------------------------------------------------------------------
# Path classes
# ----------------------------------------------------------------------------
def path_loss(scene: OScene, cand, xys):
    """geometry.py:1077-1084 — sum of interaction residuals, sequential adds from 0.0"""
    loss = torch.zeros((), dtype=F32)
    for i, j in enumerate(cand):
        loss = loss + scene.evaluate_cartesian(int(j), xys[i], xys[i + 1], xys[i + 2])
    return loss


def image_path(scene: OScene, tx, cand, rx):
    """
    ImagePath.from_tx_objects_rx — geometry.py:1017-1114.
    tx, rx: [..., 2] broadcastable.  Returns (list of k+2 points, loss).
    """
    k = len(cand)
    if k == 0:
        return [tx, rx], torch.zeros((), dtype=F32)
    images = []
    image = tx
    for j in cand:  # forward scan :1086-1091
        image = scene.image_of(int(j), image)
        images.append(image)
    point = rx
    points = [None] * k
    for i in range(k - 1, -1, -1):  # backward scan (reverse=True) :1093-1107
        j = int(cand[i])
        p = scene.origin(j)
        n = scene.normal(j)
        u = point - images[i]
        v = p - point
        un = dot2(u, n)
        vn = dot2(v, n)
        # single where, exactly as :1105 (NaN cotangent when un == 0 — kept on purpose)
        if CLEAN:
            un_safe = torch.where(un == 0.0, torch.ones_like(un), un)
            inc = torch.where((un == 0.0).unsqueeze(-1), torch.zeros((), dtype=F32),
                              vn.unsqueeze(-1) * u / un_safe.unsqueeze(-1))
        else:
            inc = torch.where((un == 0.0).unsqueeze(-1), torch.zeros((), dtype=F32),
                              vn.unsqueeze(-1) * u / un.unsqueeze(-1))
        point = point + inc
        points[i] = point
    xys = [tx, *points, rx]
    return xys, path_loss(scene, cand, xys)


def parametric_to_cartesian(scene: OScene, cand, theta, tx, rx):
    """geometry.py:976-1010 — theta [..., n_unknowns]"""
    pts = [tx]
    u = 0
    for j in cand:
        j = int(j)
        if scene.parameters_count(j) == 0:
            pts.append(scene.parametric_to_cartesian(j, None))
        else:
            pts.append(scene.parametric_to_cartesian(j, theta[..., u]))
            u += 1
    pts.append(rx)
    return pts


def _stack_pts(pts):
    shape = torch.broadcast_shapes(*[p.shape for p in pts])
    return torch.stack([p.expand(shape) for p in pts], dim=-2)


OPTIMIZER = ("adam", 0.9, 0.999, 1e-8)  # (kind, b1 | momentum, b2, eps): see `optimizer`


class optimizer:
    """``with optimizer("sgd", momentum=0.5):`` — the optax transformation behind optimize.minimize (optimize.py:44-97):
    "adam" = optax.adam(lr, b1, b2, eps), "sgd" = optax.sgd(lr, momentum) (optax.trace(decay=momentum) then scale(-lr))."""

    def __init__(self, kind="adam", b1=0.9, b2=0.999, eps=1e-8, momentum=None):
        self.val = (kind, float(momentum or 0.0) if kind == "sgd" else float(b1), float(b2), float(eps))

    def __enter__(self):
        global OPTIMIZER
        self.prev = OPTIMIZER
        OPTIMIZER = self.val

    def __exit__(self, *a):
        global OPTIMIZER
        OPTIMIZER = self.prev


def minimize_adam(fun, x0, steps=100, lr=0.1, differentiable=False):
    """
    optimize.py:44-97 with optax.adam(0.1) (optax 0.2.4: scale_by_adam b1=.9 b2=.999 eps=1e-8
    eps_root=0, bias correction with count starting at 1, then scale(-lr)).
    ``fun`` maps theta [..., n] -> loss [...]; elements are independent (vmap semantics).
    Returns (x_final, loss at the iterate BEFORE the last update) — optimize.py:96-97.
    """
    kind, b1, b2, eps = OPTIMIZER
    x = x0
    mu = torch.zeros_like(x0)
    nu = torch.zeros_like(x0)
    loss = None
    for count in range(1, steps + 1):
        if differentiable:
            xg = x
            if not xg.requires_grad:
                xg = xg.detach().requires_grad_(True)
            loss = fun(xg)
            (g,) = torch.autograd.grad(loss.sum(), xg, create_graph=True)
            x = xg
        else:
            xg = x.detach().requires_grad_(True)
            with torch.enable_grad():
                loss = fun(xg)
                (g,) = torch.autograd.grad(loss.sum(), xg)
            loss = loss.detach()
            x = xg.detach()
        if kind == "sgd":  # optax.sgd: trace = g + momentum * trace ; update = -lr * trace
            mu = g + b1 * mu
            x = x + (-lr) * mu
            continue
        mu = (1 - b1) * g + b1 * mu
        nu = (1 - b2) * (g * g) + b2 * nu
        bc1 = _NP(1) - _NP(b1) ** _NP(count)
        bc2 = _NP(1) - _NP(b2) ** _NP(count)
        mu_hat = mu / float(bc1)
        nu_hat = nu / float(bc2)
        upd = mu_hat / (torch.sqrt(nu_hat) + eps)
        x = x + (-lr) * upd
    return x, loss


def fermat_or_min_path(scene: OScene, method, tx, cand, rx, x0, steps=100, lr=0.1, differentiable=False):
    """
    FermatPath geometry.py:1121-1204 / MinPath :1211-1288 (many=1).
    ``x0``: initial guess [n_unknowns] (stands in for jax.random.uniform(key, (n,)),
    optimize.py:132; the JAX key stream is not reproducible here — see DESIGN.md).
    """
    k = len(cand)
    if k == 0:
        return [tx, rx], torch.zeros((), dtype=F32)
    n_unknowns = sum(scene.parameters_count(int(j)) for j in cand)
    shape = torch.broadcast_shapes(tx.shape[:-1], rx.shape[:-1])

    def fermat_loss(theta):
        return path_length(_stack_pts(parametric_to_cartesian(scene, cand, theta, tx, rx)))

    def min_loss(theta):
        pts = parametric_to_cartesian(scene, cand, theta, tx, rx)
        out = path_loss(scene, cand, pts)
        return out.expand(shape) if out.dim() == 0 else out

    loss_fun = fermat_loss if method == "fermat" else min_loss
    x0t = _t(x0)
    if x0t.dim() == 2:
        # minimize_many_random_uniform (optimize.py:142-182): x0 [many, n]; vmap over the restarts, then
        # i_min = argmin(losses) (first minimum), xs[i_min], losses[i_min] — per grid element
        thetas, losses = [], []
        for r in range(x0t.shape[0]):
            th0 = x0t[r, :n_unknowns].expand(*shape, n_unknowns).clone()
            if n_unknowns == 0:
                th, ls = th0, loss_fun(th0)
            else:
                th, ls = minimize_adam(loss_fun, th0, steps=steps, lr=lr, differentiable=differentiable)
            thetas.append(th)
            losses.append(ls.expand(shape) if ls.dim() == 0 else ls)
        L = torch.stack(losses)                       # [many, ...]
        i_min = torch.argmin(L.detach(), dim=0, keepdim=True)
        loss = torch.gather(L, 0, i_min).squeeze(0)
        TH = torch.stack(thetas)                      # [many, ..., n]
        theta = torch.gather(TH, 0, i_min.unsqueeze(-1).expand(1, *shape, n_unknowns)).squeeze(0)
        xys = parametric_to_cartesian(scene, cand, theta, tx, rx)
        if method == "fermat":
            return xys, path_loss(scene, cand, xys)
        return xys, loss
    theta0 = x0t[:n_unknowns].expand(*shape, n_unknowns).clone()
    if n_unknowns == 0:
        # scan still runs, on an empty parameter vector; loss is the constant function value
        theta = theta0
        loss = loss_fun(theta)
    else:
        theta, loss = minimize_adam(loss_fun, theta0, steps=steps, lr=lr, differentiable=differentiable)
    xys = parametric_to_cartesian(scene, cand, theta, tx, rx)
    if method == "fermat":
        return xys, path_loss(scene, cand, xys)  # :1202-1204
    return xys, loss  # :1286-1288


def from_tx_objects_rx(scene, method, tx, cand, rx, x0=None, steps=100, lr=0.1, differentiable=False):
    if method == "image":
        return image_path(scene, tx, cand, rx)
    if method == "path":  # Path.from_tx_objects_rx geometry.py:752-809 : s = 0.5 on every object
        pts = [tx] + [scene.parametric_to_cartesian(int(j), torch.tensor(0.5)) for j in cand] + [rx]
        return pts, torch.zeros((), dtype=F32)
    return fermat_or_min_path(scene, method, tx, cand, rx, x0, steps, lr, differentiable)


def on_objects(scene: OScene, cand, xys, logic: Logic):
    """geometry.py:821-854"""
    contains = logic.true_value()
    for i, j in enumerate(cand):
        j = int(j)
        if scene.kinds[j] == KIND_VERTEX:
            c = logic.true_value()
        else:
            s = scene.cartesian_to_parametric(j, xys[i + 1])
            c = scene.contains_parametric(j, s, logic)
        contains = logic.land(contains, c)
    return contains


def intersects_with_objects(scene: OScene, cand, xys, logic: Logic, patch=DEFAULT_PATCH, literal=False):
    """
    geometry.py:856-906 — sequential OR-fold; objects adjacent to a segment are skipped.
    ``literal=True`` calls ``intersects_cartesian`` object by object exactly like the reference;
    the default evaluates the same fp32 expressions for all objects of one segment at once
    (a trailing object axis) and then performs the SAME sequential fold — identical values and
    identical cotangents, ~N times fewer torch dispatches.
    """
    idx = [-1, *[int(c) for c in cand], -1]
    intersects = logic.false_value()
    if literal:
        for i in range(len(xys) - 1):
            for j in range(scene.n):
                if j == idx[i] or j == idx[i + 1]:
                    continue  # jnp.where(ignore, intersects, ...) with a static predicate
                hit = scene.intersects_cartesian(j, xys[i], xys[i + 1], logic, patch)
                intersects = logic.lor(intersects, hit)
        return intersects
    t = scene.xys[:, 1] - scene.xys[:, 0]  # [N,2]
    P1 = scene.xys[:, 0] - patch * t
    P2 = scene.xys[:, 1] + patch * t
    for i in range(len(xys) - 1):
        keep = [j for j in range(scene.n)
                if j != idx[i] and j != idx[i + 1] and scene.kinds[j] != KIND_VERTEX]
        if not keep:
            continue
        r0 = xys[i].unsqueeze(-2)
        r1 = xys[i + 1].unsqueeze(-2)
        hits = segments_intersect(P1[keep], P2[keep], r0, r1, logic)  # [..., len(keep)]
        for q in range(len(keep)):
            intersects = logic.lor(intersects, hits[..., q])
    return intersects


def is_valid(scene: OScene, cand, xys, loss, logic: Logic, tol=1e-2, patch=DEFAULT_PATCH):
    """geometry.py:908-963"""
    v = logic.lall(
        on_objects(scene, cand, xys, logic),
        logic.lnot(intersects_with_objects(scene, cand, xys, logic, patch)),
        logic.lt(loss, _c(tol)),
    )
    if v.dtype.is_floating_point:
        v = torch.nan_to_num(v)
    return v


def received_power(xys, r_coef=DEFAULT_R_COEF, height=DEFAULT_HEIGHT):
    """utils.py:16-54 — python-float constants folded in double, then cast (weak typing)."""
    r = path_length(_stack_pts(xys))
    n = len(xys) - 2
    num = torch.tensor(float(np.float32(r_coef**n)), dtype=F32)  # tensor / tensor: a true IEEE division
    return num / (float(np.float32(height * height)) + r * r)


def length_squared(xys, **_):
    """tests/test_scene.py:444-445, :488-489 — ``fun = path.length() ** 2``"""
    r = path_length(_stack_pts(xys))
    return r * r  # x**2 lowers to x*x in XLA (integer_pow)


FUNS = {"received_power": received_power, "length_squared": length_squared}


# ----------------------------------------------------------------------------
# scene.py accumulation entry points
# ----------------------------------------------------------------------------
def _facc(scene, tx, rx, cands, logic, *, method, fun, fun_kwargs, tol, patch, x0, steps, lr,
          differentiable, collect=None):
    """scene.py:1892-1918 / :1589-1615 — acc = 0.0 ; acc = acc + valid * fun(...) in list order."""
    acc = torch.zeros((), dtype=F32)
    for ci, cand in enumerate(cands):
        xys, loss = from_tx_objects_rx(
            scene, method, tx, cand, rx,
            x0=None if x0 is None else x0[ci], steps=steps, lr=lr, differentiable=differentiable,
        )
        valid = is_valid(scene, cand, xys, loss, logic, tol=tol, patch=patch)
        val = FUNS[fun](xys, **fun_kwargs)
        if collect is not None:
            collect.append((valid, val, xys, loss))
        acc = acc + valid * val
    return acc


def accumulate_on_grid(
    scene: OScene, X, Y, *, grid_role="receivers", fun="received_power", fun_kwargs=None,
    reduce_all=False, grad=False, value_and_grad=False, method="image", min_order=0, max_order=1,
    order=None, filter_nodes=None, x0=None, steps=100, lr=0.1, approx=False, alpha=DEFAULT_ALPHA,
    function="hard_sigmoid", patch=DEFAULT_PATCH, tol=1e-2,
):
    """
    Scene.accumulate_on_receivers_grid_over_paths (scene.py:1803-1953) when grid_role ==
    "receivers"; Scene.accumulate_on_transmitters_grid_over_paths (:1489-1648) when
    "transmitters".  Returns a list of (name, Z | dZ | (Z, dZ)) or the reduced arrays.
    """
    fun_kwargs = fun_kwargs or {}
    scene = scene.cast()
    logic = Logic(approx, alpha, function)
    cands = all_path_candidates(scene.n, min_order, max_order, order=order, filter_nodes=filter_nodes)
    X = _t(X)
    Y = _t(Y)
    want_grad = grad or value_and_grad
    fixed = scene.transmitters if grid_role == "receivers" else scene.receivers
    results = []
    for name, pt in fixed.items():
        grid = torch.stack((X, Y), dim=-1).detach().clone().requires_grad_(want_grad)
        with torch.set_grad_enabled(want_grad):
            if grid_role == "receivers":
                Z = _facc(scene, pt, grid, cands, logic, method=method, fun=fun, fun_kwargs=fun_kwargs,
                          tol=tol, patch=patch, x0=x0, steps=steps, lr=lr, differentiable=want_grad)
            else:
                Z = _facc(scene, grid, pt, cands, logic, method=method, fun=fun, fun_kwargs=fun_kwargs,
                          tol=tol, patch=patch, x0=x0, steps=steps, lr=lr, differentiable=want_grad)
            Z = Z.expand(X.shape) if Z.dim() == 0 else Z
        if want_grad:
            if Z.requires_grad:
                (dZ,) = torch.autograd.grad(Z.sum(), grid, allow_unused=True)
                dZ = torch.zeros_like(grid) if dZ is None else dZ
            else:
                dZ = torch.zeros_like(grid)
            Z = Z.detach()
            results.append((name, (Z, dZ) if value_and_grad else dZ))
        else:
            results.append((name, Z))
    if reduce_all:  # scene.py:1939-1952
        if value_and_grad:
            Zs = torch.zeros((), dtype=F32)
            dZs = torch.zeros((), dtype=F32)
            for _, (p, dp) in results:
                Zs = Zs + p
                dZs = dZs + dp
            return Zs, dZs
        Zs = torch.zeros((), dtype=F32)
        for _, p in results:
            Zs = Zs + p
        return Zs
    return results


def power_map_and_vjp(
    scene: OScene, X, Y, Zbar=None, *, grid_role="receivers", fun="received_power", fun_kwargs=None,
    method="image", min_order=0, max_order=1, filter_nodes=None, x0=None, steps=100, lr=0.1,
    approx=True, alpha=DEFAULT_ALPHA, function="hard_sigmoid", patch=DEFAULT_PATCH, tol=1e-2,
    wrt=("grid", "xys", "phis", "fixed", "alpha"), cand_stride=1,
):
    """
    What ``jax.vjp(lambda scene, tx, grid, alpha: accumulate_…(reduce_all=True, alpha=alpha))``
    produces: Z [n,m] summed over the fixed points, and the cotangents of
    object vertices, RIS angles, fixed points (TX for a receivers grid), grid points and alpha
    for a given Zbar [n,m] (default: ones).
    """
    fun_kwargs = fun_kwargs or {}
    scene = scene.cast()
    X = _t(X)
    Y = _t(Y)
    grid = torch.stack((X, Y), dim=-1).detach().clone().requires_grad_(True)
    xys = scene.xys.detach().clone().requires_grad_(True)
    phis = scene.phis.detach().clone().requires_grad_(True)
    alpha_t = torch.tensor(float(alpha), dtype=F32, requires_grad=True)
    fixed_src = scene.transmitters if grid_role == "receivers" else scene.receivers
    names = list(fixed_src.keys())
    fixed = torch.stack([fixed_src[k] for k in names]).detach().clone().requires_grad_(True)
    s2 = OScene(xys, scene.kinds, phis)
    logic = Logic(approx, alpha_t, function)
    cands = all_path_candidates(scene.n, min_order, max_order, filter_nodes=filter_nodes)
    if cand_stride > 1:  # bounded CPU-baseline samples only: every cand_stride-th candidate
        cands = cands[::cand_stride]
        x0 = None if x0 is None else x0[::cand_stride]
    Z = torch.zeros((), dtype=F32)
    for t in range(fixed.shape[0]):
        if grid_role == "receivers":
            p = _facc(s2, fixed[t], grid, cands, logic, method=method, fun=fun, fun_kwargs=fun_kwargs,
                      tol=tol, patch=patch, x0=x0, steps=steps, lr=lr, differentiable=True)
        else:
            p = _facc(s2, grid, fixed[t], cands, logic, method=method, fun=fun, fun_kwargs=fun_kwargs,
                      tol=tol, patch=patch, x0=x0, steps=steps, lr=lr, differentiable=True)
        Z = Z + p
    Z = Z.expand(X.shape) if Z.dim() == 0 else Z
    Zbar = torch.ones_like(X) if Zbar is None else _t(Zbar)
    inputs = {"grid": grid, "xys": xys, "phis": phis, "fixed": fixed, "alpha": alpha_t}
    outs = {}
    if Z.requires_grad:
        gs = torch.autograd.grad((Z * Zbar).sum(), [inputs[w] for w in wrt], allow_unused=True)
    else:
        gs = [None] * len(wrt)
    for w, g in zip(wrt, gs):
        outs[w] = torch.zeros_like(inputs[w]) if g is None else g
    return Z.detach(), outs


def valid_masks(scene: OScene, tx, rx, *, method="image", min_order=0, max_order=1, filter_nodes=None,
                x0=None, steps=100, lr=0.1, approx=False, alpha=DEFAULT_ALPHA, function="hard_sigmoid",
                patch=DEFAULT_PATCH, tol=1e-2, fun="received_power", fun_kwargs=None):
    """Per-candidate (valid, fun value) for broadcastable tx / rx — the parity probe."""
    scene = scene.cast()
    logic = Logic(approx, alpha, function)
    cands = all_path_candidates(scene.n, min_order, max_order, filter_nodes=filter_nodes)
    collect = []
    with torch.no_grad():
        _facc(scene, _t(tx), _t(rx), cands, logic, method=method, fun=fun, fun_kwargs=fun_kwargs or {},
              tol=tol, patch=patch, x0=x0, steps=steps, lr=lr, differentiable=False, collect=collect)
    shape = torch.broadcast_shapes(_t(tx).shape[:-1], _t(rx).shape[:-1])
    valid = torch.stack([c[0].expand(shape) for c in collect], dim=-1)
    vals = torch.stack([c[1].expand(shape) for c in collect], dim=-1)
    return cands, valid, vals


# ----------------------------------------------------------------------------
# Canned scenes (scene.py builders restated as arrays — Appendix B of SURVEY.md)
# ----------------------------------------------------------------------------
def _square_walls():
    return [[[0, 0], [1, 0]], [[1, 0], [1, 1]], [[1, 1], [0, 1]], [[0, 1], [0, 0]]]


def square_scene():
    """scene.py:790-836"""
    return OScene(_square_walls(), transmitters={"tx": [0.2, 0.2]}, receivers={"rx": [0.5, 0.6]})


def square_scene_with_wall(ratio=0.6):
    """scene.py:839-882"""
    w = _square_walls() + [[[0.5, 0.5 * (1 - ratio)], [0.5, 0.5 * (1 + ratio)]]]
    return OScene(w, transmitters={"tx": [0.2, 0.5]}, receivers={"rx": [0.8, 0.5]})


def square_scene_with_obstacle(ratio=0.1):
    """scene.py:885-935"""
    hl = 0.5 * ratio
    x0, x1, y0, y1 = 0.5 - hl, 0.5 + hl, 0.5 - hl, 0.5 + hl
    w = _square_walls() + [[[x0, y0], [x1, y0]], [[x1, y0], [x1, y1]], [[x1, y1], [x0, y1]], [[x0, y1], [x0, y0]]]
    return OScene(w, transmitters={"tx": [0.2, 0.2]}, receivers={"rx": [0.5, 0.6]})


def basic_scene():
    """scene.py:736-787"""
    w = _square_walls() + [[[0.4, 0.0], [0.4, 0.4]], [[0.4, 0.4], [0.3, 0.4]], [[0.1, 0.4], [0.0, 0.4]]]
    return OScene(w, transmitters={"tx": [0.1, 0.1]}, receivers={"rx": [0.302, 0.2147]})


def scene_from_geojson_rings(rings, tx_loc="NW", rx_loc="SE"):
    """scene.py:628-663 — one Wall per (coords[i-1], coords[i]); i=0 wraps (zero-length closure wall)."""
    walls = []
    for coords in rings:
        n = len(coords)
        for i in range(n):
            walls.append([coords[i - 1], coords[i]])
    sc = OScene(np.asarray(walls, dtype=np.float64).astype(np.float32))
    bb = sc.bounding_box().numpy()
    (xmin, ymin), (xmax, ymax) = bb
    xavg, yavg = np.float32(0.5) * (xmin + xmax), np.float32(0.5) * (ymin + ymax)
    loc = {"N": (xavg, ymax), "E": (xmax, yavg), "S": (xavg, ymin), "W": (xmin, yavg), "C": (xavg, yavg),
           "NE": (xmax, ymax), "NW": (xmin, ymax), "SE": (xmax, ymin), "SW": (xmin, ymin)}
    sc.transmitters = {"tx": _t(loc[tx_loc])}
    sc.receivers = {"rx": _t(loc[rx_loc])}
    return sc
