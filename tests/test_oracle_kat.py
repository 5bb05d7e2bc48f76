"""
Pins the oracle (oracle/ref_torch.py and oracle/d2d_oracle.c) against the known-answer tests the
REFERENCE's own test-suite and doctests hold for this path.  Each test cites the
reference test it restates (paths relative to /root/reference).  CPU only.
"""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import ref_torch as R

APPROX = [False, True]


def T(x):
    return torch.tensor(x, dtype=torch.float32)


@pytest.mark.parametrize("approx", APPROX)
def test_segments_intersect(approx):
    """tests/test_geometry.py:101-120 and doctest geometry.py:139-151"""
    L = R.Logic(approx)
    hit = R.segments_intersect(T([0.0, 0.0]), T([1.0, 0.0]), T([0.5, -1.0]), T([0.5, 1.0]), L)
    assert bool(L.is_true(hit))
    if approx:
        assert float(hit) == 1.0
        assert float(R.segments_intersect(T([0.0, 0.0]), T([1.0, 0.0]), T([0.5, -1.0]), T([0.5, 1.0]),
                                          R.Logic(True, function="sigmoid"))) == 1.0
    miss = R.segments_intersect(T([0.0, 0.0]), T([1.0, 0.0]), T([0.0, 1.0]), T([1.0, 1.0]), L)
    assert bool(L.is_false(miss))


def test_path_length():
    """tests/test_geometry.py:123-135 (== 4.0 exactly despite eps) and doctest geometry.py:193-197"""
    sq = T([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    assert float(R.path_length(sq)) == 4.0
    tri = T([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 0.0]])
    assert np.float32(R.path_length(tri).item()) == np.float32(3.4142137)


def test_normalize():
    """doctest geometry.py:217-225"""
    v, l = R.normalize(T([1.0, 1.0]))
    assert np.allclose(v.numpy(), [0.70710677, 0.70710677]) and np.float32(l.item()) == np.float32(1.4142135)
    v, l = R.normalize(T([0.0, 0.0]))
    assert v.tolist() == [0.0, 0.0] and float(l) == 1.0


@pytest.mark.parametrize("origin,dest", [([0.0, 0.0], [1.0, 0.0]), ([0.0, 0.0], [0.0, 1.0]),
                                         ([0.3, 0.2], [4.0, 2.0]), ([1.0, 1.0], [-2.0, 0.5])])
def test_wall_normal(origin, dest):
    """tests/test_geometry.py:256-261"""
    sc = R.OScene([[origin, dest]])
    n = sc.normal(0)
    v = T(dest) - T(origin)
    assert abs(float(R.dot2(v, n))) < 1e-6
    assert abs(float(R.norm2(n)) - 1.0) < 1e-6


def test_wall_parametric():
    """tests/test_geometry.py:268-322"""
    sc = R.OScene([[[0.0, 0.0], [4.0, 2.0]]])
    assert sc.parametric_to_cartesian(0, T(0.5)).tolist() == [2.0, 1.0]
    for p, want in [([2.0, 1.0], 0.5), ([0.0, 0.0], 0.0), ([4.0, 2.0], 1.0), ([8.0, 4.0], 2.0), ([-4.0, -2.0], -1.0)]:
        assert float(sc.cartesian_to_parametric(0, T(p))) == want
    for approx in APPROX:
        L = R.Logic(approx)
        assert bool(L.is_true(sc.contains_parametric(0, T(0.5), L)))
        assert bool(L.is_false(sc.contains_parametric(0, T(2.0), L)))


@pytest.mark.parametrize("approx", APPROX)
def test_wall_intersects_cartesian(approx):
    """tests/test_geometry.py:324-342 incl. 'intersects on the extremity'"""
    sc = R.OScene([[[0.0, 0.0], [4.0, 2.0]]])
    L = R.Logic(approx)
    assert bool(L.is_true(sc.intersects_cartesian(0, T([0.0, 2.0]), T([4.0, 0.0]), L)))
    assert bool(L.is_false(sc.intersects_cartesian(0, T([0.0, 1.0]), T([4.0, 3.0]), L)))
    assert bool(L.is_false(sc.intersects_cartesian(0, T([0.0, 1.0]), T([2.0, 7.0]), L)))
    got = sc.intersects_cartesian(0, T([0.0, 1.0]), T([0.0, 0.0]), L)
    assert float(got) > 0 if approx else bool(got)


def test_evaluate_cartesian():
    """tests/test_geometry.py:344-355 (Wall, specular = 0) and :365-376 (RIS, phi = 0)"""
    sc = R.OScene([[[0.0, 0.0], [4.0, 0.0]]])
    assert abs(float(sc.evaluate_cartesian(0, T([0.0, 1.0]), T([2.0, 0.0]), T([4.0, 1.0])))) < 1e-6
    assert abs(float(sc.evaluate_cartesian(0, T([0.0, 1.0]), T([2.1, 0.0]), T([4.0, 1.0])))) > 1e-4
    ris = R.OScene([[[0.0, 0.0], [4.0, 0.0]]], kinds=[R.KIND_RIS], phis=[0.0])
    assert abs(float(ris.evaluate_cartesian(0, T([0.0, 1.0]), T([2.0, 0.0]), T([2.0, 1.0])))) < 1e-6
    assert abs(float(ris.evaluate_cartesian(0, T([0.0, 1.0]), T([2.0, 0.0]), T([4.0, 1.0])))) > 1e-4


def test_image_of():
    """doctest geometry.py:663-667"""
    sc = R.OScene([[[0.0, 0.0], [1.0, 0.0]]])
    assert sc.image_of(0, T([0.0, 1.0])).tolist() == [0.0, -1.0]


def test_path_sampled_at_half():
    """tests/test_geometry.py:380-399"""
    sc = R.OScene([[[0.0, 0.0], [2.0, 0.0]]])
    pts, _ = R.from_tx_objects_rx(sc, "path", T([0.0, 1.0]), [0], T([2.0, 1.0]))
    assert abs(float(R.path_length(R._stack_pts(pts))) - 2 * math.sqrt(2)) < 1e-6
    for method in ("path", "image", "fermat", "minpath"):
        pts, _ = R.from_tx_objects_rx(sc, method, T([0.0, 1.0]), [], T([2.0, 1.0]), x0=np.zeros(1, np.float32))
        assert abs(float(R.path_length(R._stack_pts(pts))) - 2.0) < 1e-6


@pytest.mark.parametrize("approx", APPROX)
@pytest.mark.parametrize("method", ["image", "fermat", "minpath"])
def test_is_valid_square_scene(approx, method):
    """tests/test_geometry.py:451-467 — candidate [0,1,2,3] on square_scene is valid for every path class"""
    sc = R.square_scene()
    L = R.Logic(approx)
    x0 = np.random.default_rng(1234).random(4, dtype=np.float32)
    pts, loss = R.from_tx_objects_rx(sc, method, sc.transmitters["tx"], [0, 1, 2, 3], sc.receivers["rx"], x0=x0,
                                     steps=100)
    assert bool(L.is_true(R.is_valid(sc, [0, 1, 2, 3], pts, loss, L)))


def test_image_path_loss_is_zero():
    """tests/test_geometry.py:492-500 (atol 1e-13)"""
    sc = R.square_scene()
    _, loss = R.image_path(sc, sc.transmitters["tx"], [0, 1, 2, 3], sc.receivers["rx"])
    assert abs(float(loss)) <= 1e-13


@pytest.mark.parametrize("method", ["fermat", "minpath"])
def test_solver_simple_reflection(method):
    """tests/test_geometry.py:503-525 — single reflection hits (1, 0), rtol 1e-2; MinPath loss atol 1e-4"""
    sc = R.OScene([[[0.0, 0.0], [2.0, 0.0]]])
    x0 = np.random.default_rng(1234).random(1, dtype=np.float32)
    pts, loss = R.from_tx_objects_rx(sc, method, T([0.0, 1.0]), [0], T([2.0, 1.0]), x0=x0, steps=100)
    got = torch.stack([p.reshape(2) for p in pts]).numpy()
    np.testing.assert_allclose(got, [[0.0, 1.0], [1.0, 0.0], [2.0, 1.0]], rtol=1e-2, atol=1e-2)
    if method == "minpath":
        assert abs(float(loss)) < 1e-4


def test_minimize_quadratic():
    """tests/test_optimize.py:27-74 and doctest optimize.py:67-81"""
    x, y = R.minimize_adam(lambda x: ((x - 1.0) ** 2).sum(-1), torch.zeros(10), steps=1000)
    np.testing.assert_allclose(x.numpy(), np.ones(10), rtol=1e-3)
    x, y = R.minimize_adam(lambda x: ((x - 1.0) ** 2).sum(-1), torch.zeros(10), steps=100)
    np.testing.assert_allclose(x.numpy(), np.ones(10), rtol=1e-2)
    assert abs(float(y)) < 1e-4


def test_received_power():
    """tests/test_utils.py:8-22 — 0.3 / 4"""
    pts = [T([0.0, 0.0]), T([1.0, 0.0]), T([1.0, 1.0])]
    assert abs(float(R.received_power(pts, r_coef=0.3, height=0.0)) - 0.3 / 4.0) < 1e-6


@pytest.mark.parametrize("alpha", [1e-3, 1e-2, 1e-1, 1.0, 10.0])
def test_activation(alpha):
    """tests/test_logic.py:208-218"""
    x = torch.linspace(-5, 5, 200)
    np.testing.assert_allclose(R.Logic(True, alpha, "sigmoid").activation(x).numpy(),
                               (1 / (1 + np.exp(-alpha * x.numpy().astype(np.float64)))), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(R.Logic(True, alpha, "hard_sigmoid").activation(x).numpy(),
                               np.clip(alpha * x.numpy() + 3, 0, 6) / 6, rtol=1e-6, atol=1e-7)


def test_logic_ops():
    """tests/test_logic.py:221-386"""
    rng = np.random.default_rng(0)
    x, y = T(rng.random(200)), T(rng.random(200))
    Ls, Lh = R.Logic(True), R.Logic(False)
    assert torch.equal(Ls.lor(x, y), torch.maximum(x, y)) and torch.equal(Ls.land(x, y), torch.minimum(x, y))
    assert torch.equal(Ls.lnot(x), 1.0 - x)
    assert torch.equal(Lh.ge(x, y), x >= y) and torch.equal(Lh.lt(x, y), x < y) and torch.equal(Lh.le(x, y), x <= y)
    assert torch.equal(Ls.ge(x, y), Ls.activation(x - y)) and torch.equal(Ls.lt(x, y), Ls.activation(y - x))
    assert float(Ls.lall(T(0.2), T(0.7), T(0.5))) == pytest.approx(0.2)
    assert float(Ls.lany(T(0.2), T(0.7), T(0.5))) == pytest.approx(0.7)
    assert float(Ls.true_value()) == 1.0 and float(Ls.false_value()) == 0.0
    assert bool(Lh.true_value()) is True and bool(Lh.false_value()) is False


def test_candidates_kats():
    """tests/test_scene.py:372-399: order 0 -> one empty candidate; filter -> [[], [3]]; plus differt-core's
    documented lexicographic example (3 nodes, order 2)."""
    for impl in (R.all_path_candidates, CO.all_path_candidates):
        got = impl(11, min_order=0, max_order=0)
        assert len(got) == 1 and len(got[0]) == 0
        got = impl(11, order=0)
        assert len(got) == 1 and len(got[0]) == 0
        got = impl(6, min_order=0, max_order=2, filter_nodes=(0, 1, 2, 4, 5))
        assert [c.tolist() for c in got] == [[], [3]]
        assert all(c.dtype == np.int32 for c in got)
        got = impl(3, order=2)
        assert [c.tolist() for c in got] == [[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]]
        for c in impl(5, min_order=0, max_order=3):
            assert all(a != b for a, b in zip(c[:-1], c[1:]))  # no self loops


def test_accumulate_over_paths_los():
    """tests/test_scene.py:443-485 — LOS with fun = length**2 gives 2, 1, 1, 2 and 6 in total"""
    sc = R.OScene(np.zeros((0, 2, 2), np.float32), transmitters={"tx0": [0.0, 0.0], "tx1": [1.0, 0.0]})
    rx = np.array([[1.0, 1.0], [0.0, 1.0]], np.float32)
    res = R.accumulate_on_grid(sc, rx[:, 0], rx[:, 1], fun="length_squared", max_order=1, approx=False)
    vals = [v.tolist() for _, v in res]
    np.testing.assert_allclose(vals, [[2.0, 1.0], [1.0, 2.0]], rtol=1e-6)
    Zc = CO.power_map(np.zeros((0, 2, 2), np.float32), [[0.0, 0.0], [1.0, 0.0]], rx, fun="length_squared", max_order=1)
    np.testing.assert_allclose(Zc, [[2.0, 1.0], [1.0, 2.0]], rtol=1e-6)
    assert float(Zc.sum()) == pytest.approx(6.0)


@pytest.mark.parametrize("role", ["receivers", "transmitters"])
def test_grid_methods_los(role):
    """tests/test_scene.py:487-627 — Z = X^2 + Y^2, grad = [2X, 2Y] on the last axis, value_and_grad tuple,
    results keyed by the fixed points' names in order, reduce_all sums them."""
    fixed = {"a": [0.0, 0.0], "b": [0.0, 0.0]}
    sc = R.OScene(np.zeros((0, 2, 2), np.float32), transmitters=fixed, receivers=fixed)
    x = np.linspace(-2, 2, 9, dtype=np.float32)
    y = np.linspace(-1, 3, 7, dtype=np.float32)
    X, Y = np.meshgrid(x, y)
    res = R.accumulate_on_grid(sc, X, Y, grid_role=role, fun="length_squared", approx=False)
    assert [k for k, _ in res] == ["a", "b"]
    for _, Z in res:
        assert Z.shape == X.shape and Z.dtype == torch.float32
        np.testing.assert_allclose(Z.numpy(), X * X + Y * Y, rtol=1e-5, atol=1e-6)
    Z = R.accumulate_on_grid(sc, X, Y, grid_role=role, fun="length_squared", approx=False, reduce_all=True)
    np.testing.assert_allclose(Z.numpy(), 2 * (X * X + Y * Y), rtol=1e-5, atol=1e-6)
    for _, dZ in R.accumulate_on_grid(sc, X, Y, grid_role=role, fun="length_squared", approx=False, grad=True):
        np.testing.assert_allclose(dZ.numpy(), np.stack([2 * X, 2 * Y], -1), rtol=1e-4, atol=1e-5)
    for _, (Z, dZ) in R.accumulate_on_grid(sc, X, Y, grid_role=role, fun="length_squared", approx=False,
                                            value_and_grad=True):
        np.testing.assert_allclose(Z.numpy(), X * X + Y * Y, rtol=1e-5, atol=1e-6)
        assert dZ.shape == (*X.shape, 2)


def test_geojson_scene(geojson_rings):
    """tests/test_scene.py:217-253 — 28 walls; tx / rx on the NW / SE corners of the bounding box"""
    sc = R.scene_from_geojson_rings(geojson_rings)
    assert sc.n == 28
    bb = sc.bounding_box().numpy()
    assert sc.transmitters["tx"].tolist() == [bb[0, 0], bb[1, 1]]
    assert sc.receivers["rx"].tolist() == [bb[1, 0], bb[0, 1]]
    lens = (sc.xys[:, 1] - sc.xys[:, 0]).abs().sum(-1)
    assert int((lens == 0).sum()) == 2  # one closure wall per ring (scene.py:646-652)


def test_canned_scenes():
    """doctests scene.py:750-759, 804-813, 856-865, 901-910 (object counts, default end points)"""
    assert R.basic_scene().n == 7 and R.basic_scene().transmitters["tx"].tolist() == pytest.approx([0.1, 0.1])
    assert R.square_scene().n == 4 and R.square_scene().receivers["rx"].tolist() == pytest.approx([0.5, 0.6])
    assert R.square_scene_with_wall().n == 5
    assert R.square_scene_with_obstacle().n == 8


@pytest.mark.parametrize("mode,approx,fn", [("hard", False, "hard_sigmoid"), ("hard_sigmoid", True, "hard_sigmoid")])
@pytest.mark.parametrize("scene", ["obstacle", "basic"])
def test_two_oracles_agree_bit_for_bit(mode, approx, fn, scene):
    """The two independent restatements (vectorised torch, scalar C) must produce identical bits for the
    validity of every (receiver, candidate), every path value and the accumulated map."""
    sc = {"obstacle": R.square_scene_with_obstacle, "basic": R.basic_scene}[scene]()
    X, Y = sc.grid(24, 20)
    grid = np.stack([X.numpy(), Y.numpy()], -1)
    tx = sc.transmitters["tx"]
    Zc, vc, fc = CO.power_map(sc.xys.numpy(), tx.numpy()[None], grid, max_order=2, mode=mode, want_valid=True,
                              want_fun=True)
    _, vt, ft = R.valid_masks(sc, tx, torch.from_numpy(grid), max_order=2, approx=approx, function=fn)
    Zt = R.accumulate_on_grid(sc, X, Y, max_order=2, approx=approx, function=fn)[0][1]
    assert np.array_equal(vc[0].reshape(20, 24, -1), vt.float().numpy())
    assert np.array_equal(fc[0].reshape(20, 24, -1), ft.numpy())
    assert np.array_equal(Zc[0].reshape(20, 24), Zt.numpy())


def test_literal_and_vectorised_occlusion_agree():
    sc = R.basic_scene()
    X, Y = sc.grid(9, 7)
    grid = torch.stack((X, Y), -1)
    for approx in APPROX:
        L = R.Logic(approx, 10.0)
        for cand in ([], [0], [4, 1], [2, 6]):
            pts, _ = R.image_path(sc, sc.transmitters["tx"], cand, grid)
            a = R.intersects_with_objects(sc, cand, pts, L)
            b = R.intersects_with_objects(sc, cand, pts, L, literal=True)
            assert torch.equal(a.expand(7, 9), b.expand(7, 9))


def test_clean_gradients_leave_values_unchanged_and_finite():
    sc = R.square_scene_with_obstacle()
    X, Y = sc.grid(8, 8)  # includes receivers ON the walls: the literal gradient is NaN there
    Z0, g0 = R.power_map_and_vjp(sc, X, Y, max_order=1)
    with R.clean_gradients():
        Z1, g1 = R.power_map_and_vjp(sc, X, Y, max_order=1)
    assert torch.equal(Z0, Z1)
    assert torch.isnan(g0["grid"]).any()
    assert all(torch.isfinite(v).all() for v in g1.values())


# ---- committed golden vectors (tests/golden/power_fixtures.npz, made by tests/golden/make_power_fixtures.py) ----------
def _golden():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "power_fixtures.npz"))


@pytest.mark.parametrize("name", ["obstacle", "basic", "geojson", "geojson_norm"])
def test_oracles_reproduce_the_golden_fixtures(name):
    """Both restatements (scalar C, torch) still produce the committed vectors: hard validity of every (receiver,
    candidate) and the hard / hard_sigmoid maps bit for bit, the clean VJP to 1e-6 (torch's threaded reductions)."""
    import torch

    from oracle import c_oracle as CO
    from oracle import ref_torch as R
    from tests import helpers as H
    from tests.golden import make_power_fixtures as M

    g = _golden()
    sc = M.scenes()[name]
    X, Y, xys, fixed = g[f"{name}/X"], g[f"{name}/Y"], g[f"{name}/xys"], g[f"{name}/fixed"]
    Xn, Yn, xys_n, fixed_n = M.inputs(sc)
    assert np.array_equal(X, Xn) and np.array_equal(Y, Yn) and np.array_equal(xys, xys_n) and np.array_equal(fixed, fixed_n)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    shape = tuple(g[f"{name}/hard/valid_shape"])
    want_v = np.unpackbits(g[f"{name}/hard/valid_bits"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)
    Z, v = CO.power_map(xys, fixed, grid, max_order=2, mode="hard", want_valid=True)
    assert np.array_equal(v, want_v) and np.array_equal(Z, g[f"{name}/hard/Z"])
    osc = H.oracle_scene_from_product(sc)
    _, vt, _ = R.valid_masks(osc, osc.transmitters["tx"], torch.from_numpy(np.stack([X, Y], -1)), max_order=2, approx=False)
    assert np.array_equal(vt.numpy().reshape(shape[1:]).astype(np.float32), want_v[0])
    for alpha in (10.0, 100.0):
        Zs, vs = CO.power_map(xys, fixed, grid, max_order=2, mode="hard_sigmoid", alpha=alpha, want_valid=True)
        assert np.array_equal(vs, g[f"{name}/hard_sigmoid_{alpha:g}/valid"])
        assert np.array_equal(Zs, g[f"{name}/hard_sigmoid_{alpha:g}/Z"])
    with R.clean_gradients():
        Zo, gr = R.power_map_and_vjp(osc, X, Y, g[f"{name}/vjp/Zbar"], max_order=2, approx=True, alpha=20.0,
                                     function="hard_sigmoid")
    np.testing.assert_allclose(Zo.numpy(), g[f"{name}/vjp/Z"], rtol=1e-6, atol=1e-7)
    for k in ("grid", "xys", "fixed", "alpha"):
        want = g[f"{name}/vjp/{k}_bar"]
        np.testing.assert_allclose(gr[k].numpy(), want, rtol=1e-5, atol=1e-6 * max(np.abs(want).max(), 1e-30), err_msg=k)


@pytest.mark.parametrize("name", ["obstacle", "basic", "geojson_norm"])
def test_dual_number_oracle_reproduces_the_golden_vjp(name):
    """oracle/d2d_oracle_ad.cpp (forward-mode AD over the scalar port) against the committed vectors, which were
    produced by torch REVERSE-mode autograd over oracle/ref_torch.py: two independent differentiation mechanisms over two
    independent restatements.  Forward value bit for bit with the C oracle; cotangents to 1e-5 (fp32) and the fp64
    evaluation against the committed fp64 leg to 1e-7 (summation order)."""
    from oracle import c_oracle as CO

    g = _golden()
    X, Y, xys, fixed = g[f"{name}/X"], g[f"{name}/Y"], g[f"{name}/xys"], g[f"{name}/fixed"]
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    Zbar = g[f"{name}/vjp/Zbar"]
    a = CO.power_vjp(xys, fixed, grid, Zbar, max_order=2, mode="hard_sigmoid", alpha=20.0)
    Zc = CO.power_map(xys, fixed, grid, max_order=2, mode="hard_sigmoid", alpha=20.0, reduce_all=True)
    assert np.array_equal(a["Z"].astype(np.float32), Zc)
    a64 = CO.power_vjp(xys, fixed, grid, Zbar, max_order=2, mode="hard_sigmoid", alpha=20.0, real64=True)
    for k, ko in (("grid", "grid"), ("objects", "xys"), ("fixed", "fixed"), ("alpha", "alpha")):
        want = g[f"{name}/vjp/{ko}_bar"].astype(np.float64).reshape(-1)
        want64 = g[f"{name}/vjp64/{ko}_bar"].reshape(-1)
        scale = max(np.abs(want64).max(), 1e-30)
        if not (name == "basic" and k == "objects"):  # (axis-aligned scene: structural min/max ties, DESIGN.md "Ties")
            np.testing.assert_allclose(a[k].reshape(-1), want, rtol=1e-5, atol=2e-6 * scale, err_msg=f"{k} fp32")
            np.testing.assert_allclose(a64[k].reshape(-1), want64, rtol=1e-7, atol=1e-9 * scale, err_msg=f"{k} fp64")


def test_dual_number_oracle_other_roles_and_objects():
    """Transmitters-grid role, RIS angle cotangent, sigmoid, hard logic: dual numbers vs torch autograd."""
    import differt2d_b200 as d
    from oracle import c_oracle as CO
    from oracle import ref_torch as R
    from tests import helpers as H

    sc = H.generic_position(d.Scene.square_scene().add_objects(
        d.RIS(xys=[[0.5, 0.3], [0.5, 0.7]], phi=float(np.pi / 4)), d.Wall(xys=[[0.7, 0.15], [0.9, 0.35]])))
    sc = sc.update_receivers(rx2=d.Point(xy=[0.8, 0.3]))
    X, Y = H.jittered_grid(sc, 6, 7, seed=3)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    Zbar = np.random.default_rng(7).standard_normal(X.shape).astype(np.float32)
    osc = H.oracle_scene_from_product(sc)
    for role, mode, alpha in (("receivers", "hard_sigmoid", 2.0), ("transmitters", "sigmoid", 3.0), ("receivers", "hard", 100.0)):
        src = sc.transmitters if role == "receivers" else sc.receivers
        fixed = np.stack([p.xy for p in src.values()])
        a = CO.power_vjp(xys, fixed, grid, Zbar, kinds=kinds, phis=phis, grid_role=role, max_order=2, mode=mode, alpha=alpha)
        with R.clean_gradients():
            Zo, go = R.power_map_and_vjp(osc, X, Y, Zbar, grid_role=role, max_order=2, approx=mode != "hard", alpha=alpha,
                                         function="sigmoid" if mode == "sigmoid" else "hard_sigmoid")
        np.testing.assert_allclose(a["Z"].reshape(X.shape), Zo.numpy(), rtol=1e-6, atol=1e-7)
        for k, ko in (("grid", "grid"), ("objects", "xys"), ("phis", "phis"), ("fixed", "fixed"), ("alpha", "alpha")):
            want = go[ko].numpy().astype(np.float64).reshape(-1)
            scale = max(np.abs(want).max(), 1e-30)
            np.testing.assert_allclose(a[k].reshape(-1), want, rtol=2e-5, atol=2e-6 * scale, err_msg=f"{role} {mode} {k}")
        if mode != "hard" and role == "receivers":
            assert np.abs(a["phis"]).max() > 0


@pytest.mark.parametrize("function,alpha", [("sigmoid", 0.5), ("sigmoid", 1.0), ("sigmoid", 3.0), ("sigmoid", 10.0),
                                            ("hard_sigmoid", 1.0), ("hard_sigmoid", 4.0), ("hard_sigmoid", 100.0)])
@pytest.mark.parametrize("scene", ["obstacle", "basic"])
def test_fold_shortcut_claim_holds_in_the_reference_arithmetic(scene, function, alpha):
    """The kernels skip the occlusion fold of a path whose min(a_on, a_l) is not above 1 - act(0.51) - 4e-6
    (csrc/d2d_device.cuh:fold_skip_bound) and return that min as the validity.  This pins the claim itself on the
    restated reference (geometry.py:908-963 literal, fp32): (i) no test of the fold ever activates above act(0.51), and
    (ii) wherever the bound holds, is_valid — computed WITH the fold — equals min(on_objects, loss < tol) bit for bit."""
    sc = R.square_scene_with_obstacle() if scene == "obstacle" else R.basic_scene()
    logic = R.Logic(True, alpha, function)
    tx = list(sc.transmitters.values())[0]
    bb = sc.bounding_box()
    rng = np.random.default_rng(3)
    n = 20
    xs = torch.tensor(np.sort(rng.uniform(float(bb[0][0]), float(bb[1][0]), n)).astype(np.float32))
    ys = torch.tensor(np.sort(rng.uniform(float(bb[0][1]), float(bb[1][1]), n)).astype(np.float32))
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    rx = torch.stack([X, Y], -1)
    a_max = logic.activation(R._c(0.51))
    bound = (torch.tensor(1.0) - a_max) - 4e-6
    skipped = total = 0
    for cand in R.all_path_candidates(sc.n, 0, 2):
        xys, loss = R.from_tx_objects_rx(sc, "image", tx, cand, rx)
        a_on = R.on_objects(sc, cand, xys, logic)
        a_on = a_on if (torch.is_tensor(a_on) and a_on.ndim) else torch.ones(n, n) * a_on
        a_l = torch.nan_to_num(logic.lt(loss, R._c(1e-2)) * torch.ones(n, n))
        a_in = R.intersects_with_objects(sc, cand, xys, logic)
        a_in = a_in if (torch.is_tensor(a_in) and a_in.ndim) else torch.ones(n, n) * a_in
        assert bool((a_in <= a_max).all()), "a test of the fold activated above act(0.51)"
        valid = R.is_valid(sc, cand, xys, loss, logic)
        valid = valid if valid.ndim else torch.ones(n, n) * valid
        v0 = torch.minimum(a_on, a_l)
        m = v0 <= bound
        skipped += int(m.sum())
        total += m.numel()
        assert torch.equal(valid[m], v0[m]), "the validity depends on the fold where the shortcut says it cannot"
    if function == "sigmoid" and alpha <= 1.0:
        assert skipped > total // 2  # the regime the shortcut exists for (3/4 of the paths at alpha = 1)
    if function == "hard_sigmoid" and alpha >= 100.0:
        assert skipped == 0  # act(0.51) saturates: the bound is negative, the shortcut never fires
