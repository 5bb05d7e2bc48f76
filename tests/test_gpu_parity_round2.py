"""
GPU parity tests, second batch ("Next round" item 1): the configuration bench.py times (raw lon/lat
city scene) against the autograd oracle with an fp64 third leg, the transmitters-grid role against the oracle (smooth
forward and full VJP), RIS objects inside ImagePath, non-default `received_power` parameters.

Bars (BASELINE.json north_star): hard and hard_sigmoid validity bit-exact; maps rtol 1e-5; gradients ELEMENTWISE
rtol 1e-4 (+ 1e-6 of the largest entry), see test_gpu_parity._close.
"""
import numpy as np
import pytest
import torch

import differt2d_b200 as d
from differt2d_b200 import functional as F
from oracle import c_oracle as CO
from oracle import ref_torch as R
from tests import helpers as H
from tests.test_gpu_parity import SCENES, _cfg, _close, _oracle_vjp, _oracle_vjp64

pytestmark = pytest.mark.gpu

FN = {"hard": "hard_sigmoid", "hard_sigmoid": "hard_sigmoid", "sigmoid": "sigmoid"}


def _bench_subgrid(sc, n=13, m=12):
    """n x m points of the 1024 x 1024 bbox grid bench.py traces (every 79th row from 5, every 89th column from 11):
    the same fp32 coordinates the benchmark's launch sees."""
    X, Y = sc.grid(1024, 1024)
    X, Y = X[5::79, 11::89][:n, :m], Y[5::79, 11::89][:n, :m]
    return np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)


@pytest.mark.parametrize("mode,alpha", [("hard_sigmoid", 100.0), ("hard_sigmoid", 20.0), ("sigmoid", 30.0), ("hard", 100.0)])
def test_vjp_raw_geojson_bench_config(mode, alpha):
    """
    The workload of bench.py: example.geojson in RAW lon/lat coordinates, orders 0-2 (785 candidates), forward + VJP
    w.r.t. receivers, object vertices, TX and alpha, on points of the benchmark's own grid.

    1. validity of every (receiver, candidate) bit-identical to the C oracle (hard / hard_sigmoid; sigmoid rtol 1e-5);
    2. Z and every cotangent ELEMENTWISE against torch autograd over the restated graph (fp32, clean gradients);
    3. third leg: the same oracle in binary64.  On raw lon/lat coordinates (|y| ~ 50.7, walls ~ 1e-4 long) fp32 leaves
       ~3 % of a wall length of noise on every parametric coordinate, so the REFERENCE'S OWN fp32 map and gradients are
       far from their fp64 values at most receivers.  Entries where the fp32 and fp64 oracles agree to 1e-3
       must meet the elementwise bar outright; the others must be no farther from the fp32 oracle than the fp32 oracle
       is from fp64 (test_gpu_parity._close).  The excluded counts are printed.
    """
    sc = SCENES["geojson"]
    X, Y = _bench_subgrid(sc)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    Zbar = np.random.default_rng(17).standard_normal(X.shape).astype(np.float32)
    # 1. validity pattern
    Zf, v = F.power_fwd(_cfg(mode, max_order=2, grid_cols=X.shape[1]), xys, fixed, grid, alpha=alpha, want_valid=True,
                        device="cuda")
    Zc, vc = CO.power_map(xys, fixed, grid, max_order=2, mode=mode, alpha=alpha, want_valid=True)
    v = v.cpu().numpy()
    if mode == "sigmoid":
        np.testing.assert_allclose(v, vc, rtol=1e-5, atol=2.5e-7)
    else:
        assert np.array_equal(v, vc), f"{int((v != vc).sum())} of {v.size} validity values differ from the C oracle"
    assert (vc != 0).sum() > 0
    # 2. + 3. value and cotangents: two INDEPENDENT fp32 oracles (reverse mode: torch autograd over the restated graph;
    # forward mode: dual numbers over the scalar C++ port) and the fp64 evaluation of the same function
    Zo, go = _oracle_vjp(sc, X, Y, Zbar, mode, alpha)
    Z64, g64 = _oracle_vjp64(sc, X, Y, Zbar, mode, alpha)
    ad = CO.power_vjp(xys, fixed, grid, Zbar, max_order=2, mode=mode, alpha=alpha)
    out = F.power_value_and_vjp(_cfg(mode, max_order=2, reduce_all=True, grid_cols=X.shape[1]), xys, fixed, grid,
                                Zbar.reshape(-1), alpha=alpha, device="cuda")
    out = {k: t.cpu().numpy() for k, t in out.items()}
    np.testing.assert_allclose(out["Z"].reshape(X.shape), Zo.numpy(), rtol=1e-5, atol=1e-6)
    report = []
    for key, okey in (("grid", "grid"), ("fixed", "fixed"), ("objects", "xys"), ("alpha", "alpha")):
        if key == "alpha" and mode == "hard":
            assert out["alpha"][0] == 0.0
            continue
        got = out[key].reshape(-1).astype(np.float64)
        w32 = go[okey].numpy().reshape(-1).astype(np.float64)
        w64 = g64[okey].numpy().reshape(-1).astype(np.float64)
        a32 = ad[key].reshape(-1)
        scale = max(np.abs(w32).max(), 1e-30)
        # well conditioned in fp32 = BOTH fp32 oracles reproduce the fp64 value to 1e-3 (one alone can agree by chance)
        agree = (np.abs(w32 - w64) <= 1e-3 * np.abs(w64) + 1e-6 * scale) & (np.abs(a32 - w64) <= 1e-3 * np.abs(w64) + 1e-6 * scale)
        tol = 1e-4 * np.abs(w32) + 1e-6 * scale
        bad = (np.abs(got - w32) > tol) & agree
        assert not bad.any(), (f"{key}_bar: {int(bad.sum())} of {int(agree.sum())} well-conditioned entries miss rtol 1e-4; "
                               f"worst |diff| {np.abs(got - w32)[bad].max():.3g} at scale {scale:.3g}")
        # everywhere else: no farther from the fp32 oracle than the fp32 oracles are from fp64 / from each other
        noise = np.maximum(np.abs(w32 - w64), np.abs(a32 - w32))
        still = np.abs(got - w32) > tol + 4.0 * noise + 4.0 * np.percentile(noise, 90)
        assert not still.any(), f"{key}_bar: {int(still.sum())} entries exceed the oracles' own fp32 noise"
        report.append(f"{key}: {int((~agree).sum())}/{agree.size} ill-conditioned in fp32")
    print(f"[parity] raw geojson {mode} alpha={alpha:g}: " + "; ".join(report))


@pytest.mark.parametrize("name", ["basic", "obstacle"])
@pytest.mark.parametrize("mode,alpha", [("hard", 100.0), ("hard_sigmoid", 20.0), ("hard_sigmoid", 100.0), ("sigmoid", 10.0)])
def test_transmitters_grid_forward_and_vjp_vs_oracle(name, mode, alpha):
    """Scene.accumulate_on_transmitters_grid_over_paths (scene.py:1489-1648, the role the reference's own benchmark
    times, tests/benchmarks/test_scene.py:9-29): validity of every (receiver, TX grid point, candidate), the map and
    the full VJP (TX grid points, fixed receivers, vertices, alpha) against the oracle, two receivers, three logics."""
    sc = SCENES[name].update_receivers(rx2=d.Point(xy=[0.8, 0.3]))
    scg = H.generic_position(sc)
    X, Y = H.jittered_grid(sc, 18, 20, seed=5)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    for scene, check_objects in ((sc, mode == "hard"), (scg, True)):
        xys, kinds, phis = scene.packed_objects()
        fixed = np.stack([p.xy for p in scene.receivers.values()])
        cfg = _cfg(mode, max_order=2, grid_role="transmitters", grid_cols=X.shape[1])
        Z, v = F.power_fwd(cfg, xys, fixed, grid, alpha=alpha, want_valid=True, device="cuda")
        Zc, vc = CO.power_map(xys, fixed, grid, grid_role="transmitters", max_order=2, mode=mode, alpha=alpha,
                              want_valid=True)
        if mode == "sigmoid":
            np.testing.assert_allclose(v.cpu().numpy(), vc, rtol=1e-5, atol=2.5e-7)
            np.testing.assert_allclose(Z.cpu().numpy(), Zc, rtol=1e-5, atol=1e-6)
        else:
            assert np.array_equal(v.cpu().numpy(), vc)
            np.testing.assert_allclose(Z.cpu().numpy(), Zc, rtol=1e-6, atol=0)
        Zbar = np.random.default_rng(8).standard_normal(X.shape).astype(np.float32)
        Zo, go = _oracle_vjp(scene, X, Y, Zbar, mode, alpha, role="transmitters")
        cache = {}

        def g64(key, scene=scene, cache=cache):
            def f():
                if "g" not in cache:
                    cache["g"] = _oracle_vjp64(scene, X, Y, Zbar, mode, alpha, role="transmitters")[1]
                return cache["g"][key].numpy()
            return f

        for use_mask in (False, True):
            rcfg = _cfg(mode, max_order=2, grid_role="transmitters", grid_cols=X.shape[1], reduce_all=True)
            if use_mask:
                out = F.power_value_and_vjp(rcfg, xys, fixed, grid, Zbar.reshape(-1), alpha=alpha, device="cuda")
            else:
                out = F.power_bwd(rcfg, xys, fixed, grid, Zbar.reshape(-1), alpha=alpha, device="cuda")
            out = {k: t.cpu().numpy() for k, t in out.items()}
            np.testing.assert_allclose(out["Z"].reshape(X.shape), Zo.numpy(), rtol=1e-5, atol=1e-6)
            _close(out["grid"].reshape(*X.shape, 2), go["grid"].numpy(), 1e-4, "grid_bar (TX grid)", g64("grid"))
            _close(out["fixed"], go["fixed"].numpy(), 1e-4, "fixed_bar (receivers)", g64("fixed"), max_escaped=1.0)
            if check_objects:  # (axis-aligned scenes: structural ties, DESIGN.md "Ties")
                _close(out["objects"], go["xys"].numpy(), 1e-4, "objects_bar", g64("xys"), max_escaped=0.25)
            if mode != "hard":
                _close(out["alpha"], go["alpha"].numpy().reshape(1), 1e-4, "alpha_bar", g64("alpha"), max_escaped=1.0)


def _ris_image_scene():
    """square_scene + a RIS (examples/plot_ris_power_map.py:37-73 geometry) + one slanted wall, in generic position."""
    sc = d.Scene.square_scene().add_objects(d.RIS(xys=[[0.5, 0.3], [0.5, 0.7]], phi=float(np.pi / 4)),
                                            d.Wall(xys=[[0.7, 0.15], [0.9, 0.35]]))
    return H.generic_position(sc)


@pytest.mark.parametrize("mode,alpha", [("hard", 100.0), ("hard_sigmoid", 2.0), ("hard_sigmoid", 50.0), ("sigmoid", 3.0)])
def test_ris_inside_image_path_vs_oracle(mode, alpha):
    """RIS objects visited by ImagePath candidates (RIS is a Wall subclass: image_of / intersects / contains are the
    wall's, the residual is RIS.evaluate_cartesian, geometry.py:698-711): validity, map and VJP including the phi
    cotangent against the oracle.  Small alpha keeps act(tol - loss) alive for the non-specular RIS residuals."""
    sc = _ris_image_scene()
    X, Y = H.jittered_grid(sc, 14, 15, seed=6)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    assert kinds[4] == 1 and phis[4] != 0
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    Z, v = F.power_fwd(_cfg(mode, max_order=2, grid_cols=X.shape[1]), xys, fixed, grid, kinds=kinds, phis=phis,
                       alpha=alpha, want_valid=True, device="cuda")
    Zc, vc = CO.power_map(xys, fixed, grid, kinds=kinds, phis=phis, max_order=2, mode=mode, alpha=alpha, want_valid=True)
    if mode == "hard":
        assert np.array_equal(v.cpu().numpy(), vc) and np.array_equal(Z.cpu().numpy(), Zc)
    else:
        # sinf / cosf of phi: CUDA libm vs glibc may differ in the last bit -> the RIS residual is not bit-pinned
        np.testing.assert_allclose(v.cpu().numpy(), vc, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(Z.cpu().numpy(), Zc, rtol=2e-5, atol=1e-6)
    Zbar = np.random.default_rng(9).standard_normal(X.shape).astype(np.float32)
    Zo, go = _oracle_vjp(sc, X, Y, Zbar, mode, alpha)
    g64c = {}

    def g64(key):
        def f():
            if "g" not in g64c:
                g64c["g"] = _oracle_vjp64(sc, X, Y, Zbar, mode, alpha)[1]
            return g64c["g"][key].numpy()
        return f

    out = F.power_value_and_vjp(_cfg(mode, max_order=2, reduce_all=True, grid_cols=X.shape[1]), xys, fixed, grid,
                                Zbar.reshape(-1), kinds=kinds, phis=phis, alpha=alpha, device="cuda")
    out = {k: t.cpu().numpy() for k, t in out.items()}
    np.testing.assert_allclose(out["Z"].reshape(X.shape), Zo.numpy(), rtol=2e-5, atol=1e-6)
    _close(out["grid"].reshape(*X.shape, 2), go["grid"].numpy(), 1e-4, "grid_bar", g64("grid"))
    _close(out["fixed"], go["fixed"].numpy(), 1e-4, "fixed_bar", g64("fixed"), max_escaped=1.0)
    _close(out["objects"], go["xys"].numpy(), 1e-4, "objects_bar", g64("xys"), max_escaped=0.25)
    _close(out["phis"], go["phis"].numpy(), 1e-4, "phis_bar", g64("phis"), max_escaped=1.0)
    if mode != "hard":
        _close(out["alpha"], go["alpha"].numpy().reshape(1), 1e-4, "alpha_bar", g64("alpha"), max_escaped=1.0)
        assert np.abs(out["phis"]).max() > 0, "the RIS angle carries no gradient in this configuration"


@pytest.mark.parametrize("mode", ["hard", "hard_sigmoid"])
def test_non_default_received_power_parameters(mode):
    """utils.received_power(r_coef=0.7, height=0.25) (utils.py:16-54; Python-float constants folded in double, then
    cast): hard map bit-exact against the C oracle, smooth map and VJP against the autograd oracle."""
    sc = H.generic_position(SCENES["obstacle"])
    X, Y = H.jittered_grid(sc, 16, 18, seed=12)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    kw = dict(r_coef=0.7, height=0.25)
    cfg = _cfg(mode, max_order=2, grid_cols=X.shape[1], reduce_all=True, **kw)
    Z = F.power_fwd(cfg, xys, fixed, grid, alpha=40.0, device="cuda")
    Zc = CO.power_map(xys, fixed, grid, max_order=2, mode=mode, alpha=40.0, reduce_all=True, **kw)
    if mode == "hard":
        assert np.array_equal(Z.cpu().numpy(), Zc)
    else:
        np.testing.assert_allclose(Z.cpu().numpy(), Zc, rtol=1e-6, atol=0)
    Zd = CO.power_map(xys, fixed, grid, max_order=2, mode=mode, alpha=40.0, reduce_all=True)
    assert not np.allclose(Zc, Zd, rtol=1e-3), "the parameters must matter"
    Zbar = np.random.default_rng(10).standard_normal(X.shape).astype(np.float32)
    osc = H.oracle_scene_from_product(sc)
    with R.clean_gradients():
        Zo, go = R.power_map_and_vjp(osc, X, Y, Zbar, max_order=2, approx=mode != "hard", alpha=40.0,
                                     function=FN[mode], fun_kwargs=kw)
    out = F.power_value_and_vjp(cfg, xys, fixed, grid, Zbar.reshape(-1), alpha=40.0, device="cuda")
    out = {k: t.cpu().numpy() for k, t in out.items()}
    np.testing.assert_allclose(out["Z"].reshape(X.shape), Zo.numpy(), rtol=1e-5, atol=1e-6)
    _close(out["grid"].reshape(*X.shape, 2), go["grid"].numpy(), 1e-4, "grid_bar")
    _close(out["fixed"], go["fixed"].numpy(), 1e-4, "fixed_bar")
    _close(out["objects"], go["xys"].numpy(), 1e-4, "objects_bar")
    # the Scene API forwards fun_kwargs (scene.py:1909 fun(..., **fun_kwargs))
    Zs = sc.accumulate_on_receivers_grid_over_paths(X, Y, fun=d.received_power, fun_kwargs=kw, reduce_all=True,
                                                    max_order=2, approx=mode != "hard", alpha=40.0)
    assert np.array_equal(Zs.reshape(-1), Z.cpu().numpy())


def test_golden_fixture_vjp_on_raw_geojson():
    """tests/golden/power_fixtures.npz holds the VJP of the raw lon/lat scene in fp32 AND fp64: the CUDA path against
    the committed fp32 vectors, with the committed fp64 leg deciding which entries are meaningful in fp32."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "power_fixtures.npz"))
    name = "geojson"
    X, Y, xys, fixed = g[f"{name}/X"], g[f"{name}/Y"], g[f"{name}/xys"], g[f"{name}/fixed"]
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    out = F.power_bwd(_cfg("hard_sigmoid", max_order=2, reduce_all=True, grid_cols=X.shape[1]), xys, fixed, grid,
                      g[f"{name}/vjp/Zbar"].reshape(-1), alpha=20.0, device="cuda")
    out = {k: t.cpu().numpy() for k, t in out.items()}
    np.testing.assert_allclose(out["Z"].reshape(X.shape), g[f"{name}/vjp/Z"], rtol=1e-5, atol=1e-6)
    for key, okey in (("grid", "grid"), ("fixed", "fixed"), ("objects", "xys"), ("alpha", "alpha")):
        _close(out[key].reshape(-1), g[f"{name}/vjp/{okey}_bar"].reshape(-1), 1e-4, f"{key}_bar",
               g[f"{name}/vjp64/{okey}_bar"].reshape(-1), max_escaped=1.0)


# ---- BASELINE config 5: optimisation loop over accumulate_over_paths, candidate-sharded long lists ---------------------
def _oracle_link_powers(osc, tx, alpha, max_order, approx):
    """scene.accumulate_over_paths for one transmitter (scene.py:1272-1334) on the oracle: list of per-receiver powers."""
    logic = R.Logic(approx, alpha, "hard_sigmoid")
    cands = R.all_path_candidates(osc.n, 0, max_order)
    return [R._facc(osc, tx, rx, cands, logic, method="image", fun="received_power", fun_kwargs={}, tol=1e-2,
                    patch=0.0, x0=None, steps=100, lr=0.1, differentiable=True) for rx in osc.receivers.values()]


@pytest.mark.parametrize("max_order", [0, 1])
def test_power_optimize_loop_matches_the_oracle(max_order):
    """
    examples/plot_power_optimize.py:63-91,168-229 of the reference, restated on torch: loss(tx) = -min over the two
    receivers of accumulate_over_paths(received_power)/P0, jax.value_and_grad -> optax.adam(0.01) + zero_nans,
    alpha = logspace(0, 2, 101) with approximation.  The product side calls Scene.accumulate_over_paths with
    Point(xy=<tensor requiring grad>) under torch autograd (the CUDA forward + VJP kernels through functional.power_map);
    the oracle side is the restated graph + autograd (clean gradients == zero_nans on the un == 0 / d == 0 branches).
    The transmitter trajectories must agree to rtol 1e-3 over 20 iterations.
    """
    sc0 = d.Scene.square_scene_with_obstacle().with_receivers(rx_0=d.Point(xy=[0.3, 0.1]), rx_1=d.Point(xy=[0.5, 0.1]))
    osc = H.oracle_scene_from_product(sc0)
    alphas = np.logspace(0, 2, 101).astype(np.float32)[:20]
    start = [0.5, 0.7]

    def run(loss_fn):
        tx = torch.tensor(start, dtype=torch.float32, requires_grad=True)
        opt = torch.optim.Adam([tx], lr=0.01)  # optax.adam(0.01): b1 .9, b2 .999, eps 1e-8, bias-corrected
        traj, losses = [], []
        for a in alphas:
            opt.zero_grad()
            loss = loss_fn(tx, float(a))
            loss.backward()
            tx.grad.nan_to_num_(nan=0.0)  # optax.zero_nans()
            opt.step()
            traj.append(tx.detach().clone().numpy())
            losses.append(float(loss))
        return np.stack(traj), np.asarray(losses)

    def loss_product(tx, alpha):
        scene = sc0.with_transmitters(tx=d.Point(xy=tx))
        acc = None
        for _, _, power in scene.accumulate_over_paths(fun=d.received_power, max_order=max_order, approx=True, alpha=alpha):
            pw = power / d.P0
            acc = pw if acc is None else torch.minimum(acc, pw)
        return -acc.cpu() if acc.device.type != "cpu" else -acc

    def loss_oracle(tx, alpha):
        with R.clean_gradients():
            acc = None
            for power in _oracle_link_powers(osc, tx, alpha, max_order, True):
                pw = power / R.P0
                acc = pw if acc is None else torch.minimum(acc, pw)
            return -acc

    tp, lp = run(loss_product)
    to, lo = run(loss_oracle)
    np.testing.assert_allclose(lp, lo, rtol=1e-3, err_msg="loss trajectory")
    np.testing.assert_allclose(tp, to, rtol=1e-3, err_msg="transmitter trajectory")
    assert np.abs(tp[-1] - np.asarray(start)).max() > 0.05, "the transmitter did not move"


def test_scene_entry_points_are_differentiable_under_autograd():
    """The analogue of jax.grad over Scene.accumulate_on_receivers_grid_over_paths(reduce_all=True) w.r.t. a wall
    vertex, the transmitter and alpha, through the Scene API itself, against the direct VJP."""
    base = H.generic_position(SCENES["obstacle"])
    X, Y = H.jittered_grid(base, 12, 14, seed=4)
    w3 = torch.tensor(base.objects[3].xys, requires_grad=True)
    txt = torch.tensor([0.2, 0.2], requires_grad=True)
    alpha = torch.tensor(30.0, requires_grad=True)
    objs = list(base.objects)
    objs[3] = d.Wall(xys=w3)
    sc = d.Scene({"tx": d.Point(xy=txt)}, base.receivers, objs)
    Z = sc.accumulate_on_receivers_grid_over_paths(X, Y, reduce_all=True, max_order=2, approx=True, alpha=alpha)
    assert isinstance(Z, torch.Tensor) and Z.shape == X.shape and Z.requires_grad
    Zbar = torch.as_tensor(np.random.default_rng(2).standard_normal(X.shape).astype(np.float32), device=Z.device)
    (Z * Zbar).sum().backward()
    xys, _, _ = base.packed_objects()
    ref = F.power_bwd(_cfg("hard_sigmoid", max_order=2, reduce_all=True, grid_cols=X.shape[1]), xys,
                      np.array([[0.2, 0.2]], np.float32), np.stack([X, Y], -1).reshape(-1, 2), Zbar.reshape(-1),
                      alpha=30.0, device="cuda")
    sc_ = max(float(ref["objects"].abs().max()), 1e-30)
    assert torch.allclose(w3.grad, ref["objects"][3].cpu(), rtol=1e-3, atol=1e-4 * sc_)
    assert torch.allclose(txt.grad, ref["fixed"][0].cpu(), rtol=1e-3, atol=1e-4 * float(ref["fixed"].abs().max()))
    assert torch.allclose(alpha.grad.reshape(1), ref["alpha"].cpu(), rtol=1e-3)


@pytest.mark.parametrize("mode", ["hard", "hard_sigmoid"])
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_candidate_shards_sum_to_the_unsharded_link(mode, shards):
    """point-to-point links whose candidate list is sharded over GPUs.  All shards traced one after the
    other on this GPU: validity flags are disjoint over the shards and OR to the unsharded flags bit for bit; the
    partial Z and cotangents add up to the unsharded ones (fp32 summation order aside)."""
    from tests.test_gpu_parity import _random_walls

    sc = _random_walls(60, length=0.15)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    grid = np.stack([p.xy for p in sc.receivers.values()])
    kw = dict(min_order=0, max_order=3)
    Zf, vf = F.power_fwd(_cfg(mode, **kw), xys, fixed, grid, alpha=30.0, want_valid=True, device="cuda")
    full = F.power_value_and_vjp(_cfg(mode, **kw), xys, fixed, grid, None, alpha=30.0, device="cuda")
    assert float(vf.sum()) > 0
    vsum = torch.zeros_like(vf)
    acc = None
    for s in range(shards):
        cfg = _cfg(mode, cand_shard=(s, shards), **kw)
        _, v = F.power_fwd(cfg, xys, fixed, grid, alpha=30.0, want_valid=True, device="cuda")
        assert float(((v != 0) & (vsum != 0)).sum()) == 0, "two shards traced the same candidate"
        vsum += v
        part = F.power_value_and_vjp(cfg, xys, fixed, grid, None, alpha=30.0, device="cuda")
        acc = part if acc is None else {k: acc[k] + part[k] for k in acc}
    assert torch.equal(vsum, vf)
    for k in full:
        scale = max(float(full[k].abs().max()), 1e-30)
        assert torch.allclose(acc[k], full[k], rtol=1e-4, atol=1e-5 * scale), k
    # the single-process form of the multi-GPU entry point (world size 1 = the whole list)
    from differt2d_b200 import distributed as D

    one = D.sharded_link_power_vjp(_cfg(mode, **kw), xys, fixed, grid, None, alpha=30.0, device="cuda")
    for k in full:
        scale = max(float(full[k].abs().max()), 1e-30)
        assert torch.allclose(one[k], full[k], rtol=1e-4, atol=1e-5 * scale), k


def test_row_sharded_entry_point_single_process():
    """distributed.sharded_power_vjp without a process group (world size 1): defaults to the current CUDA device,
    slices Zbar along the row axis for [T, n, m] cotangents, equals the unsharded call."""
    from differt2d_b200 import distributed as D

    sc = d.Scene.basic_scene().update_transmitters(tx2=d.Point(xy=[0.7, 0.6]))
    X, Y = sc.grid(40, 24)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    Zbar = np.random.default_rng(1).standard_normal((2, *X.shape)).astype(np.float32)
    cfg = _cfg("hard_sigmoid", max_order=2)
    out = D.sharded_power_vjp(cfg, xys, fixed, X, Y, Zbar, alpha=40.0)
    assert out["rows"].tolist() == list(range(24))
    grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
    ref = F.power_value_and_vjp(_cfg("hard_sigmoid", max_order=2, grid_cols=40), xys, fixed, grid,
                                Zbar.reshape(2, -1), alpha=40.0, device="cuda")
    assert torch.equal(out["Z"], ref["Z"]) and torch.equal(out["grid"], ref["grid"])
    for k in ("objects", "fixed", "alpha"):
        assert torch.allclose(out[k], ref[k], rtol=1e-4, atol=1e-5 * float(ref[k].abs().max()))
    # rows of another rank's share: band layout of a 2-rank job, evaluated here by hand
    rows = D.row_tiles_cyclic(24, 2, 1)
    sub = F.power_value_and_vjp(_cfg("hard_sigmoid", max_order=2, grid_cols=40), xys, fixed,
                                np.stack([X[rows], Y[rows]], -1).reshape(-1, 2).astype(np.float32),
                                Zbar[:, rows].reshape(2, -1), alpha=40.0, device="cuda")
    assert torch.equal(sub["Z"].reshape(2, len(rows), 40), ref["Z"].reshape(2, 24, 40)[:, rows])


def test_traced_alpha_not_positive_poisons_the_outputs():
    """A device-resident (traced) alpha cannot be validated on the host: alpha <= 0 must not return plausible numbers
    (ADVICE r1): every output is NaN; host scalars are rejected with an error."""
    sc = SCENES["obstacle"]
    X, Y = sc.grid(16, 16)
    grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
    xys, _, _ = sc.packed_objects()
    fixed = np.array([[0.2, 0.2]], np.float32)
    cfg = _cfg("hard_sigmoid", max_order=1, reduce_all=True, grid_cols=16)
    bad = torch.tensor(-5.0, device="cuda")
    Z = F.power_fwd(cfg, xys, fixed, grid, alpha=bad, device="cuda")
    assert bool(torch.isnan(Z).all())
    out = F.power_bwd(cfg, xys, fixed, grid, None, alpha=bad, device="cuda")
    assert all(bool(torch.isnan(t).all()) for t in out.values())
    ok = F.power_fwd(cfg, xys, fixed, grid, alpha=torch.tensor(5.0, device="cuda"), device="cuda")
    assert bool(torch.isfinite(ok).all())
    with pytest.raises(d.D2DError):
        F.power_fwd(cfg, xys, fixed, grid, alpha=-5.0, device="cuda")
    with pytest.raises(d.D2DError):
        F.power_map(torch.as_tensor(xys).cuda(), torch.as_tensor(fixed).cuda(), torch.as_tensor(grid).cuda(), cfg=cfg, alpha=0.0)
    # hard logic does not use alpha at all
    assert bool(torch.isfinite(F.power_fwd(_cfg("hard", max_order=1), xys, fixed, grid, alpha=-5.0, device="cuda")).all())


def test_host_entry_is_reentrant_across_threads():
    """d2d_power_host keeps its staging arenas per calling thread (include/differt2d_b200.h): concurrent calls from
    several threads, different problem sizes, each equal to the single-threaded result."""
    import threading

    sc = d.Scene.basic_scene()
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    cfgs, refs, grids = [], [], []
    for n, m in [(40, 32), (64, 48), (24, 80), (512, 520)]:
        X, Y = sc.grid(m, n)
        grids.append(np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32))
        cfgs.append(_cfg("hard_sigmoid", max_order=2, reduce_all=True, grid_cols=m))
        refs.append(F.power_host(cfgs[-1], xys, fixed, grids[-1], None, alpha=25.0))
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got = F.power_host(cfgs[i], xys, fixed, grids[i], None, alpha=25.0)
                assert np.array_equal(got["Z"], refs[i]["Z"]) and np.array_equal(got["grid"], refs[i]["grid"])
            F.L.lib().d2d_host_release()
        except Exception as e:  # noqa: BLE001
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4) for _ in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


# ---- transmitters-grid cull ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["basic", "obstacle", "wall", "geojson_norm", "geojson"])
@pytest.mark.parametrize("alpha", [1.0, 10.0, 100.0, 1000.0])
def test_transmitters_grid_cull_equals_nocull(name, alpha):
    """The tile / macro-tile cull of the transmitters-grid role (csrc/d2d_driver.cuh: tile_may_be_valid_tx, the receiver
    unfolded through the candidate, s-range rule on every interaction with the error recursion of the thread's own image
    chain) must not change a single bit of the map or of the per-point cotangents: every scene, every logic, the alpha
    sweep, bbox grids (points ON the walls), two receivers, orders 0-2."""
    sc = SCENES[name]
    sc = sc.update_receivers(rx2=d.Point(xy=sc.center() + np.float32(0.31) * (sc.bounding_box()[1] - sc.center())))
    n = 512 if name.startswith("geojson") else 384
    X, Y = sc.grid(n, n)
    grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.receivers.values()])
    for mode in ("hard", "hard_sigmoid", "sigmoid"):
        if mode == "hard" and alpha != 100.0:
            continue
        kw = dict(max_order=2, grid_cols=n, grid_role="transmitters")
        a = F.power_fwd(_cfg(mode, **kw), xys, fixed, grid, alpha=alpha, device="cuda")
        b = F.power_fwd(_cfg(mode, cull=False, **kw), xys, fixed, grid, alpha=alpha, device="cuda")
        assert torch.equal(a, b), (mode, int((a != b).sum()))
        assert float(a.abs().max()) > 0
        ga = F.power_bwd(_cfg(mode, reduce_all=True, **kw), xys, fixed, grid, None, alpha=alpha, device="cuda")
        gb = F.power_bwd(_cfg(mode, reduce_all=True, cull=False, **kw), xys, fixed, grid, None, alpha=alpha, device="cuda")
        assert torch.equal(ga["Z"], gb["Z"]) and torch.equal(ga["grid"], gb["grid"]), mode
        for k in ("objects", "fixed", "alpha"):  # fp32 atomics: order not fixed
            scale = max(gb[k].abs().max().item(), 1e-30)
            assert torch.allclose(ga[k], gb[k], rtol=1e-3, atol=1e-4 * scale), (mode, k)


@pytest.mark.parametrize("name", ["basic", "geojson_norm"])
def test_transmitters_grid_cull_order3_and_flat_tiles(name):
    """Three-interaction chains (every stage of the unfolded test), 1-D tiles (grid_cols = 0) and a jittered grid."""
    sc = SCENES[name]
    n = 96 if name.startswith("geojson") else 192
    X, Y = H.jittered_grid(sc, n, n, seed=5)
    grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.receivers.values()])
    for mode in ("hard", "hard_sigmoid", "sigmoid"):
        for cols in (n, 0):
            kw = dict(min_order=3, max_order=3, grid_cols=cols, candidate_slices=1, grid_role="transmitters")
            a = F.power_fwd(_cfg(mode, **kw), xys, fixed, grid, alpha=100.0, device="cuda")
            b = F.power_fwd(_cfg(mode, cull=False, **kw), xys, fixed, grid, alpha=100.0, device="cuda")
            assert torch.equal(a, b), (mode, cols, int((a != b).sum()))


def test_transmitters_grid_cull_random_scenes():
    """Random short walls in generic position (many orientations, many near-grazing incidences), hard logic masks bit
    for bit against the C oracle with the cull on."""
    from tests.test_gpu_parity import _random_walls

    for seed in (1, 2, 3):
        sc = _random_walls(24, seed=seed, length=0.25)
        X, Y = sc.grid(64, 48)
        grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
        xys, _, _ = sc.packed_objects()
        fixed = np.stack([p.xy for p in sc.receivers.values()])
        Z, v = F.power_fwd(_cfg("hard", max_order=2, grid_cols=64, grid_role="transmitters"), xys, fixed, grid,
                           want_valid=True, device="cuda")
        Zc, vc = CO.power_map(xys, fixed, grid, grid_role="transmitters", max_order=2, mode="hard", want_valid=True)
        assert np.array_equal(v.cpu().numpy(), vc) and np.array_equal(Z.cpu().numpy(), Zc)
        assert vc.sum() > 0


# ---- Fermat / MinPath at the default steps = 100: how far can two fp32 implementations agree? --------
@pytest.mark.parametrize("method", ["fermat", "minpath"])
@pytest.mark.parametrize("mode", ["hard", "hard_sigmoid"])
def test_solver_steps100_distance_is_within_the_oracles_own_fp32_noise(method, mode):
    """
    optimize.minimize (optimize.py:83-97) with the reference's default steps = 100, forward AND VJP through the scan.
    Near convergence Adam's update m / (sqrt(v) + eps) is a ratio of vanishing quantities: the 100th iterate and, much
    more so, its derivative amplify last-bit differences.  That was a claim in round 1; here it is measured: the same
    oracle graph evaluated in fp64 (ref_torch.precision) is the yardstick.  For every output the distance between the
    CUDA kernels and the fp32 oracle must not exceed 4 x the distance between the fp32 oracle and its own fp64
    evaluation (+ the elementwise bar), i.e. the kernels are as close to the reference's fp32 graph as that graph is to
    the function it computes.  Both distances are printed.
    """
    from tests.test_gpu_parity import _vertex_scene

    sc = H.generic_position(_vertex_scene())
    osc = H.oracle_scene_from_product(sc)
    X, Y = H.jittered_grid(sc, 4, 5, seed=3)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    C = 1 + 7 + 42
    x0 = np.random.default_rng(1234).random((C, 2), dtype=np.float32)
    Zbar = (0.5 + np.random.default_rng(7).random(X.shape)).astype(np.float32)
    cfg = _cfg(mode, max_order=2, method=method, steps=100, grid_cols=X.shape[1], reduce_all=True)
    got = F.power_bwd(cfg, xys, fixed, grid, Zbar.reshape(-1), kinds=kinds, phis=phis, x0=x0, alpha=100.0, device="cuda")
    kw = dict(method=method, max_order=2, x0=x0, steps=100, approx=mode != "hard", alpha=100.0, function="hard_sigmoid")
    Z32, g32 = R.power_map_and_vjp(osc, X, Y, Zbar, **kw)
    with R.precision("f64"):
        Z64, g64 = R.power_map_and_vjp(osc, X, Y, Zbar, **kw)
    legs = {"Z": (got["Z"], Z32, Z64), "grid": (got["grid"], g32["grid"], g64["grid"]),
            "objects": (got["objects"], g32["xys"], g64["xys"]), "fixed": (got["fixed"], g32["fixed"], g64["fixed"])}
    if mode != "hard":
        legs["alpha"] = (got["alpha"], g32["alpha"], g64["alpha"])
    report = []
    for k, (a, w32, w64) in legs.items():
        a = a.cpu().numpy().reshape(-1).astype(np.float64)
        w32 = w32.detach().numpy().reshape(-1).astype(np.float64)
        w64 = w64.detach().numpy().reshape(-1).astype(np.float64)
        scale = max(np.abs(w64).max(), 1e-30)
        d_kernel = np.abs(a - w32).max() / scale
        d_oracle = np.abs(w32 - w64).max() / scale
        report.append(f"{k}: kernel-vs-fp32 {d_kernel:.2e}, fp32-vs-fp64 {d_oracle:.2e}")
        assert d_kernel <= 4.0 * d_oracle + 1e-4, f"{method} {mode} {k}: {report[-1]}"
    print(f"[parity] {method} {mode} steps=100: " + "; ".join(report))


def test_scene_sanitiser_on_the_device():
    """d2d_sanitise_scene / d2d_affine_points (csrc/d2d_sanitise.cu) against the host path of
    Scene.sanitised (same binary64 arithmetic: identical tables), the parity switch (nothing dropped, nothing moved:
    the scene comes back unchanged), and a mixed scene (a Vertex is never flagged, a RIS keeps its angle)."""
    raw = SCENES["geojson"]
    for drop, norm in ((True, False), (False, True), (True, True), (False, False)):
        host, hm = raw.sanitised(drop_zero_length=drop, normalise=norm, return_map=True)
        dev, dm = raw.sanitised(drop_zero_length=drop, normalise=norm, return_map=True, device="cuda")
        assert dm["kept"] == hm["kept"] and len(dev.objects) == (26 if drop else 28)
        assert np.array_equal(dm["origin"], hm["origin"]) and dm["scale"] == hm["scale"]
        assert np.array_equal(dev.packed_objects()[0], host.packed_objects()[0])
        for k in raw.transmitters:
            assert np.array_equal(dev.transmitters[k].xy, host.transmitters[k].xy)
        for k in raw.receivers:
            assert np.array_equal(dev.receivers[k].xy, host.receivers[k].xy)
        if not drop and not norm:
            assert np.array_equal(dev.packed_objects()[0], raw.packed_objects()[0])
    mixed = d.Scene.square_scene().add_objects(d.Vertex(xy=[0.3, 0.4]), d.RIS(xys=[[0.5, 0.3], [0.5, 0.7]], phi=0.6),
                                                d.Wall(xys=[[0.2, 0.2], [0.2, 0.2]]))
    dev, dm = mixed.sanitised(return_map=True, device="cuda")
    assert dm["kept"] == [0, 1, 2, 3, 4, 5] and isinstance(dev.objects[4], d.Vertex) and isinstance(dev.objects[5], d.RIS)
    assert abs(dev.objects[5].phi - 0.6) < 1e-6
    # the sanitised scene traces without the NaN-poisoning closure walls and gives the same map: they never carry a path
    # (residual |i_hat|^2 = 1 > tol) unless the receiver coincides with the transmitter — the bbox grid's NW corner —
    # where both directions vanish; hence a grid strictly inside the box
    X, Y = H.jittered_grid(raw, 32, 40, seed=4)
    Za = raw.accumulate_on_receivers_grid_over_paths(X, Y, reduce_all=True, max_order=1, approx=False)
    Zb = raw.sanitised(device="cuda").accumulate_on_receivers_grid_over_paths(X, Y, reduce_all=True, max_order=1, approx=False)
    assert np.array_equal(Za, Zb)


# ---- optimisers of FermatPath / MinPath: optax.adam hyper-parameters, optax.sgd, Newton fast mode -------
def _solver_case_with_optimizer(method, mode, steps, opt, oracle_opt, *, lr, n=5, m=6):
    from tests.test_gpu_parity import _vertex_scene

    sc = H.generic_position(_vertex_scene())
    osc = H.oracle_scene_from_product(sc)
    X, Y = H.jittered_grid(sc, n, m, seed=3)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    x0 = np.random.default_rng(1234).random((50, 2), dtype=np.float32)
    Zbar = (0.5 + np.random.default_rng(7).random(X.shape)).astype(np.float32)
    cfg = _cfg(mode, max_order=2, method=method, steps=steps, grid_cols=m, reduce_all=True, lr=lr, **opt)
    got = F.power_bwd(cfg, xys, fixed, grid, Zbar.reshape(-1), kinds=kinds, phis=phis, x0=x0, alpha=50.0, device="cuda")
    with R.optimizer(**oracle_opt):
        Zo, go = R.power_map_and_vjp(osc, X, Y, Zbar, method=method, max_order=2, x0=x0, steps=steps, lr=lr,
                                     approx=mode != "hard", alpha=50.0, function="hard_sigmoid")
    want = {"Z": Zo, "grid": go["grid"], "objects": go["xys"], "fixed": go["fixed"], "alpha": go["alpha"]}
    err = {}
    for k, w in want.items():
        a = got[k].cpu().numpy().reshape(-1).astype(np.float64)
        b = w.detach().numpy().reshape(-1).astype(np.float64)
        err[k] = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) if np.abs(b).max() > 0 else np.abs(a).max()
    return err


@pytest.mark.parametrize("method", ["fermat", "minpath"])
@pytest.mark.parametrize("name,opt,oracle_opt,lr", [
    ("adam_hyper", dict(optimizer="adam", opt_b1=0.8, opt_b2=0.95, opt_eps=1e-6), dict(kind="adam", b1=0.8, b2=0.95, eps=1e-6), 0.05),
    ("sgd", dict(optimizer="sgd", opt_b1=0.0), dict(kind="sgd", momentum=0.0), 0.02),
    ("sgd_momentum", dict(optimizer="sgd", opt_b1=0.6), dict(kind="sgd", momentum=0.6), 0.01),
])
def test_solver_other_optimizers_vs_oracle(method, name, opt, oracle_opt, lr):
    """optimize.minimize(..., optimizer=...) (optimize.py:44-97) with optax.adam's hyper-parameters changed and with
    optax.sgd (plain and with momentum): forward map and the VJP through the scan against torch autograd over the
    restated loop (short scans: 6 steps pin the algebra of every term, see test_solver_vjp_through_adam_scan)."""
    err = _solver_case_with_optimizer(method, "hard_sigmoid", 6, opt, oracle_opt, lr=lr)
    for k, e in err.items():
        assert e < 2e-4, (name, k, e, err)


@pytest.mark.parametrize("method", ["fermat", "minpath"])
def test_newton_fast_mode_converges_to_the_image_path(method):
    """D2D_OPT_NEWTON (csrc/d2d_newton.cuh; not a reference optimiser).  In a convex room (square_scene, generic
    position) and with hard logic the stationary point of both losses that a valid path sits at is the specular path,
    i.e. what ImagePath constructs in closed form (geometry.py:1017-1114).  (With smooth logic the iterative path
    classes legitimately differ from ImagePath: a candidate that crosses a wall instead of reflecting has residual 4
    under the image method and anything in [0, 4] at Fermat's crossing point; measured: 100 Adam steps reproduce 38-43 %
    of the image-method map at alpha = 30, 72-84 % with hard logic.)  30 damped Newton iterations must reproduce the
    image-method map on more receivers than 100 Adam steps (the reference's default) do, and put the vertices of the
    valid paths closer to the image method's."""
    sc = H.generic_position(SCENES["square"])
    X, Y = H.jittered_grid(sc, 20, 22, seed=3)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    x0 = np.random.default_rng(5).random((17, 2), dtype=np.float32)
    kw = dict(max_order=2, grid_cols=X.shape[1], reduce_all=True)
    zi = F.power_fwd(_cfg("hard", **kw), xys, fixed, grid, device="cuda").cpu().numpy()
    zn = F.power_fwd(_cfg("hard", method=method, optimizer="newton", steps=30, **kw), xys, fixed, grid, x0=x0,
                     device="cuda").cpu().numpy()
    za = F.power_fwd(_cfg("hard", method=method, steps=100, **kw), xys, fixed, grid, x0=x0, device="cuda").cpu().numpy()
    close = lambda a: float(np.isclose(a, zi, rtol=1e-3, atol=1e-5 * np.abs(zi).max()).mean())  # noqa: E731
    assert close(zn) > 0.93 and close(zn) > close(za), (close(zn), close(za))
    sub = grid[::37]
    pk = dict(max_order=2)
    ri = F.paths(_cfg("hard", **pk), xys, fixed, sub, emit_all=True, device="cuda")
    rn = F.paths(_cfg("hard", method=method, optimizer="newton", steps=30, **pk), xys, fixed, sub, x0=x0, emit_all=True,
                 device="cuda")
    ra = F.paths(_cfg("hard", method=method, steps=100, **pk), xys, fixed, sub, x0=x0, emit_all=True, device="cuda")
    ok = (ri["valid"] > 0) & (rn["valid"] > 0) & (ra["valid"] > 0)
    dn = (rn["xys"] - ri["xys"]).abs().reshape(ok.shape[0], -1).max(-1).values[ok]
    da = (ra["xys"] - ri["xys"]).abs().reshape(ok.shape[0], -1).max(-1).values[ok]
    assert float(dn.median()) < 1e-5 and float(dn.median()) < float(da.median()), (float(dn.median()), float(da.median()))


def test_newton_implicit_reverse_mode_matches_a_converged_scan():
    """Reverse mode of the Newton mode = implicit differentiation of the fixed point (d theta* = -H^-1 dg/dq dq): one
    solve + one dual evaluation.  Yardstick: plain gradient descent (optax.sgd, whose unrolled reverse sweep is pinned
    against the autograd oracle by test_solver_other_optimizers_vs_oracle) run to convergence on FermatPath's convex loss
    — the derivative of a contracting iteration converges to the implicit derivative.  Smooth logic, so that the
    validity depends on the interaction points themselves and theta_bar is not zero (with hard logic the envelope
    theorem removes the implicit term: the power depends on theta* only through the minimised length)."""
    sc = H.generic_position(SCENES["square"])
    X, Y = H.jittered_grid(sc, 14, 16, seed=8)
    X, Y = np.ascontiguousarray(X, np.float32), np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    x0 = np.full((5, 1), 0.5, np.float32)
    Zbar = (0.5 + np.random.default_rng(7).random(X.shape)).astype(np.float32).reshape(-1)
    kw = dict(min_order=1, max_order=1, method="fermat", mode="hard_sigmoid", grid_cols=X.shape[1], reduce_all=True)
    new = F.power_bwd(F.TraceConfig(optimizer="newton", steps=30, **kw), xys, fixed, grid, Zbar, x0=x0, alpha=8.0, device="cuda")
    sgd = F.power_bwd(F.TraceConfig(optimizer="sgd", opt_b1=0.0, lr=0.02, steps=4000, **kw), xys, fixed, grid, Zbar, x0=x0,
                      alpha=8.0, device="cuda")
    zn, zs = new["Z"].cpu().numpy(), sgd["Z"].cpu().numpy()
    same = np.isclose(zn, zs, rtol=1e-4, atol=1e-6 * np.abs(zs).max())
    assert same.mean() > 0.95, same.mean()
    gn, gs = new["grid"].cpu().numpy().reshape(-1, 2)[same], sgd["grid"].cpu().numpy().reshape(-1, 2)[same]
    rel = np.abs(gn - gs).max() / np.abs(gs).max()
    assert rel < 5e-3, f"receiver cotangents: Newton (implicit) vs converged gradient descent differ by {rel:.2e}"
    assert np.abs(gs).max() > 0
    if same.all():
        for k in ("objects", "fixed", "alpha"):
            a, b = new[k].cpu().numpy().reshape(-1), sgd[k].cpu().numpy().reshape(-1)
            assert np.abs(a - b).max() <= 5e-3 * max(np.abs(b).max(), 1e-30), k


def test_scene_api_accepts_optimizer_descriptors():
    from tests.test_gpu_parity import _vertex_scene

    sc = _vertex_scene()
    X, Y = sc.grid(24, 20)
    isv = lambda o: isinstance(o, d.Vertex)  # noqa: E731
    base = dict(path_cls=d.FermatPath, reduce_all=True, max_order=1, key=1234, approx=False, filter_objects=isv)
    Za = sc.accumulate_on_receivers_grid_over_paths(X, Y, **base)
    Zb = sc.accumulate_on_receivers_grid_over_paths(X, Y, path_cls_kwargs={"optimizer": d.optimizers.adam(0.1)}, **base)
    assert np.array_equal(Za, Zb)  # optimize.py:83: adam(0.1) IS the default
    Zc = sc.accumulate_on_receivers_grid_over_paths(
        X, Y, path_cls_kwargs={"optimizer": d.optimizers.newton(), "steps": 10}, **base)
    assert np.isfinite(Zc).all() and (Zc > 0).any()
    with pytest.raises(NotImplementedError):
        sc.accumulate_on_receivers_grid_over_paths(X, Y, path_cls_kwargs={"optimizer": object()}, **base)


# ---- gradients of a generic `fun` -------------------------------------------------------------------------
@pytest.mark.parametrize("approx", [False, True])
@pytest.mark.parametrize("role", ["receivers", "transmitters"])
def test_generic_fun_gradients_match_the_fused_kernels(approx, role):
    """scene.py:1920-1925 differentiates acc = sum valid * fun(tx, rx, path, objects) for ANY fun.  A python restatement of
    utils.received_power, evaluated and differentiated by torch autograd on the materialised vertices and pulled back by
    d2d_paths_vjp, must give the fused backward kernel's grad / value_and_grad (per fixed point and reduced)."""
    sc = H.generic_position(SCENES["obstacle"]).update_receivers(rx2=d.Point(xy=[0.8, 0.3]))
    X, Y = H.jittered_grid(sc, 14, 16, seed=4)

    def my_power(transmitter, receiver, path, interacting_objects, r_coef=0.5, height=0.1):
        r = path.length()
        return (r_coef ** len(interacting_objects)) / (height * height + r * r)

    call = sc.accumulate_on_receivers_grid_over_paths if role == "receivers" else sc.accumulate_on_transmitters_grid_over_paths
    kw = dict(max_order=2, approx=approx, alpha=30.0)
    Zg, dZg = call(X, Y, fun=my_power, reduce_all=True, value_and_grad=True, **kw)
    Zf, dZf = call(X, Y, reduce_all=True, value_and_grad=True, **kw)
    np.testing.assert_allclose(Zg, Zf, rtol=2e-6, atol=1e-7)
    scale = np.abs(dZf).max()
    np.testing.assert_allclose(dZg, dZf, rtol=1e-4, atol=1e-5 * scale)
    assert scale > 0
    per_g = dict(call(X, Y, fun=my_power, grad=True, **kw))
    per_f = dict(call(X, Y, grad=True, **kw))
    assert list(per_g) == list(per_f)
    for k in per_f:
        assert per_g[k].shape == (*X.shape, 2)
        np.testing.assert_allclose(per_g[k], per_f[k], rtol=1e-4, atol=1e-5 * np.abs(per_f[k]).max())


def test_generic_fun_gradient_through_the_end_points():
    """A fun written on the END POINTS it is handed (transmitter.xy, receiver.xy), reference LOS KAT style
    (tests/test_scene.py:487-627: maps X^2 + Y^2, gradients [2X, 2Y])."""
    sc = d.Scene(transmitters={"tx": d.Point(xy=[0.25, -0.5])}, receivers={"rx": d.Point(xy=[0.0, 0.0])}, objects=[])
    x = np.linspace(-2, 2, 12, dtype=np.float32)
    y = np.linspace(-1, 3, 9, dtype=np.float32)
    X, Y = np.meshgrid(x, y)

    def dist2(transmitter, receiver, path, interacting_objects):
        dlt = receiver.xy - transmitter.xy
        return (dlt * dlt).sum(-1)

    Z, dZ = sc.accumulate_on_receivers_grid_over_paths(X, Y, fun=dist2, reduce_all=True, value_and_grad=True,
                                                       max_order=0, approx=False)
    np.testing.assert_allclose(Z, (X - 0.25) ** 2 + (Y + 0.5) ** 2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dZ, np.stack([2 * (X - 0.25), 2 * (Y + 0.5)], -1), rtol=1e-5, atol=1e-5)
