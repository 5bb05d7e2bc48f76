#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (parity-test cases, not the headline bench line).
One JSON line per config; CUDA-event timing, 3 warm-up + `--steps` timed calls, inputs resident in HBM.
    python scripts/bench_configs.py [--steps 5] [--only cfg5b]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import differt2d_b200 as d  # noqa: E402
from differt2d_b200 import functional as F  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def pack(sc, n, m=None):
    X, Y = sc.grid(m or n, n)
    grid = torch.from_numpy(np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)).cuda()
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    return xys, kinds, phis, fixed, grid


def ncand(n, lo, hi):
    return sum(1 if k == 0 else n * (n - 1) ** (k - 1) for k in range(lo, hi + 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    out = []

    def emit(name, ms, paths, **kw):
        line = {"config": name, "ms": ms, "paths": paths, "paths_per_s": paths / (ms * 1e-3), **kw}
        print(json.dumps(line), flush=True)
        out.append(line)

    want = lambda k: not a.only or k in a.only.split(",")  # noqa: E731

    if want("cfg1"):
        sc = d.Scene.square_scene_with_obstacle()
        xys, kinds, phis, fixed, grid = pack(sc, 100)
        cfg = F.TraceConfig(mode="hard", max_order=1, grid_cols=100)
        ms = timed(lambda: F.power_fwd(cfg, xys, fixed, grid, device="cuda"), a.steps)
        emit("cfg1: obstacle scene, ImagePath orders 0-1, 100x100, hard, forward", ms, 1e4 * ncand(8, 0, 1))
    if want("txgrid"):
        # the reference's own benchmark workload (tests/benchmarks/test_scene.py:9-29): accumulate_on_transmitters_grid_
        # over_paths on basic_scene, beside the receivers role on the same scene and grid
        sc = d.Scene.basic_scene()
        xys, kinds, phis = sc.packed_objects()
        for n in (50, 1024):
            X, Y = sc.grid(n, n)
            grid = torch.from_numpy(np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)).cuda()
            for approx, mode in ((False, "hard"), (True, "hard_sigmoid")):
                for order in (1, 2):
                    res = {}
                    for role, src in (("receivers", sc.transmitters), ("transmitters", sc.receivers)):
                        fixed = np.stack([p.xy for p in src.values()])
                        for cull in (True, False):
                            cfg = F.TraceConfig(mode=mode, max_order=order, grid_cols=n, grid_role=role, cull=cull)
                            res[(role, cull)] = timed(lambda: F.power_fwd(cfg, xys, fixed, grid, device="cuda"), a.steps)
                    emit(f"txgrid: basic_scene, ImagePath orders 0-{order}, {n}x{n}, approx={approx}, forward",
                         res[("transmitters", True)], n * n * ncand(7, 0, order),
                         rx_grid_ms=res[("receivers", True)], tx_grid_nocull_ms=res[("transmitters", False)],
                         rx_grid_nocull_ms=res[("receivers", False)],
                         tx_over_rx=res[("transmitters", True)] / res[("receivers", True)])
    if want("cfg2"):
        sc = d.Scene.square_scene_with_obstacle()
        xys, kinds, phis, fixed, grid = pack(sc, 1024)
        for mode in ("hard_sigmoid", "sigmoid"):
            for alpha in (1.0, 10.0, 100.0, 1000.0):
                cfg = F.TraceConfig(mode=mode, max_order=2, grid_cols=1024, reduce_all=True)
                ms = timed(lambda: F.power_bwd(cfg, xys, fixed, grid, None, alpha=alpha, want=("Z", "fixed"),
                                               device="cuda"), a.steps)
                emit(f"cfg2: obstacle scene, ImagePath orders 0-2, 1024x1024, {mode} alpha={alpha:g}, value + d/dTX",
                     ms, 1024 * 1024 * ncand(8, 0, 2))
    if want("cfg3"):
        from tests import helpers as H
        sc = d.Scene.from_geojson(H.geojson_text())
        xys, kinds, phis, fixed, grid = pack(sc, 2048)
        cfg = F.TraceConfig(mode="hard_sigmoid", max_order=2, grid_cols=2048, reduce_all=True)
        ms = timed(lambda: F.power_bwd(cfg, xys, fixed, grid, None, alpha=100.0, device="cuda"), a.steps)
        emit("cfg3: geojson city scene (raw), ImagePath orders 0-2, 2048x2048, hard_sigmoid, fused value + full VJP, 1 GPU",
             ms, 2048 * 2048 * ncand(28, 0, 2))
    if want("cfg4"):
        sc = d.Scene.basic_scene()
        objs = list(sc.objects)
        v = objs[5].get_vertices()[1]
        del objs[5]
        sc = d.Scene(sc.transmitters, sc.receivers, [*objs, v])
        xys, kinds, phis, fixed, grid = pack(sc, 1024)
        C = ncand(7, 0, 2)
        x0 = np.random.default_rng(1234).random((C, 2), dtype=np.float32)
        for method in ("fermat", "minpath"):
            cfg = F.TraceConfig(mode="hard", max_order=2, method=method, steps=100, grid_cols=1024)
            ms = timed(lambda: F.power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, x0=x0, device="cuda"), a.steps)
            emit(f"cfg4: basic scene + vertex, {method} orders 0-2, 100 Adam steps, 1024x1024, hard, forward", ms,
                 1024 * 1024 * C, adam_steps_per_s=1024 * 1024 * (C - 1) * 100 / (ms * 1e-3))
            cfgn = F.TraceConfig(mode="hard", max_order=2, method=method, steps=30, optimizer="newton", grid_cols=1024)
            ms = timed(lambda: F.power_fwd(cfgn, xys, fixed, grid, kinds=kinds, phis=phis, x0=x0, device="cuda"), a.steps)
            emit(f"cfg4: same, {method}, 30 damped Newton iterations (non-parity fast mode), forward", ms, 1024 * 1024 * C)
            cfgs = F.TraceConfig(mode="hard_sigmoid", max_order=2, method=method, steps=30, optimizer="newton", grid_cols=1024,
                                 reduce_all=True)
            ms = timed(lambda: F.power_value_and_vjp(cfgs, xys, fixed, grid, None, kinds=kinds, phis=phis, x0=x0, alpha=100.0,
                                                     device="cuda"), a.steps)
            emit(f"cfg4: same, {method}, Newton, hard_sigmoid, forward + implicit VJP", ms, 1024 * 1024 * C)
            cfga = F.TraceConfig(mode="hard_sigmoid", max_order=2, method=method, steps=100, grid_cols=1024, reduce_all=True)
            ms = timed(lambda: F.power_value_and_vjp(cfga, xys, fixed, grid, None, kinds=kinds, phis=phis, x0=x0, alpha=100.0,
                                                     device="cuda"), a.steps)
            emit(f"cfg4: same, {method}, 100 Adam steps, hard_sigmoid, forward + VJP through the scan", ms, 1024 * 1024 * C)
    if want("cfg5a"):
        sc = d.Scene.square_scene().add_objects(d.RIS(xys=[[0.5, 0.3], [0.5, 0.7]], phi=float(np.pi / 4)))
        xys, kinds, phis, fixed, grid = pack(sc, 300)
        x0 = np.random.default_rng(1234).random((5, 1), dtype=np.float32)
        cfg = F.TraceConfig(mode="hard", min_order=1, max_order=1, method="minpath", steps=1000, grid_cols=300)
        ms = timed(lambda: F.power_fwd(cfg, xys, fixed, grid, kinds=kinds, phis=phis, x0=x0, device="cuda"), a.steps)
        emit("cfg5a: square scene + RIS, MinPath order 1, 1000 Adam steps, 300x300, hard, forward", ms, 300 * 300 * 5,
             adam_steps_per_s=300 * 300 * 5 * 1000 / (ms * 1e-3))
    if want("cfg5b"):
        rng = np.random.default_rng(1234)
        pts = rng.random((1 + 2 * 500 + 2, 2), dtype=np.float32)  # scene.py:718-733 layout
        fixed = pts[:1]
        xys = pts[1:1001].reshape(500, 2, 2)
        grid = torch.from_numpy(pts[-2:][::-1].copy()).cuda()
        for mode in ("hard", "hard_sigmoid"):
            cfg = F.TraceConfig(mode=mode, min_order=3, max_order=3)
            ms = timed(lambda: F.power_fwd(cfg, xys, fixed, grid, alpha=100.0, device="cuda"), a.steps)
            emit(f"cfg5b: 500 random walls, ImagePath order 3, 1 TX x 2 RX point-to-point, {mode}, forward", ms,
                 2 * 500 * 499 * 499)
            ms = timed(lambda: F.power_bwd(cfg, xys, fixed, grid, None, alpha=100.0, device="cuda"), a.steps)
            emit(f"cfg5b: same, {mode}, fused value + full VJP", ms, 2 * 500 * 499 * 499)


if __name__ == "__main__":
    main()
