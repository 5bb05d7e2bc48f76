"""GPU diagnostic: Fermat/MinPath reverse mode vs the autograd oracle, error table over steps/modes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from oracle import ref_torch as R
from tests import helpers as H
from tests.test_gpu_parity import _vertex_scene, _ris_scene

def run(sc, method, mode, steps, max_order, min_order=0, n=6, m=7, generic=True, alpha=100.0):
    if generic:
        sc = H.generic_position(sc)
    osc = H.oracle_scene_from_product(sc)
    X, Y = H.jittered_grid(sc, n, m, seed=3)
    X = np.ascontiguousarray(X, np.float32); Y = np.ascontiguousarray(Y, np.float32)
    grid = np.stack([X, Y], -1).reshape(-1, 2)
    xys, kinds, phis = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    N = xys.shape[0]
    C = sum((1 if k == 0 else N * (N - 1) ** (k - 1)) for k in range(min_order, max_order + 1))
    x0 = np.random.default_rng(1234).random((C, max(max_order, 1)), dtype=np.float32)
    rng = np.random.default_rng(7)
    Zbar = (0.5 + rng.random(X.shape)).astype(np.float32)
    cfg = F.TraceConfig(mode=mode, min_order=min_order, max_order=max_order, method=method, steps=steps, grid_cols=m, reduce_all=True)
    t0 = time.time()
    out = F.power_bwd(cfg, xys, fixed, grid, Zbar.reshape(-1), kinds=kinds, phis=phis, x0=x0, alpha=alpha, device="cuda")
    torch.cuda.synchronize()
    t1 = time.time()
    Zo, go = R.power_map_and_vjp(osc, X, Y, Zbar, method=method, min_order=min_order, max_order=max_order, x0=x0, steps=steps,
                                 approx=mode != "hard", alpha=alpha, function=mode if mode != "hard" else "hard_sigmoid")
    t2 = time.time()
    def err(a, b):
        a = a.detach().cpu().numpy().reshape(-1).astype(np.float64); b = b.detach().cpu().numpy().reshape(-1).astype(np.float64)
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30), np.abs(b).max()
    row = {"Z": err(out["Z"], Zo), "grid": err(out["grid"], go["grid"]), "objects": err(out["objects"], go["xys"]),
           "phis": err(out["phis"], go["phis"]), "fixed": err(out["fixed"], go["fixed"]), "alpha": err(out["alpha"], go["alpha"])}
    print(f"{method:8s} {mode:13s} steps={steps:4d} orders {min_order}-{max_order} gpu {t1-t0:.2f}s cpu {t2-t1:.1f}s  " +
          "  ".join(f"{k}:{v[0]:.1e}(|{v[1]:.1e}|)" for k, v in row.items()), flush=True)

for method in ("fermat", "minpath"):
    for mode in ("hard", "hard_sigmoid", "sigmoid"):
        for steps in (1, 3, 10, 30, 100):
            run(_vertex_scene(), method, mode, steps, 2)
for steps in (5, 50, 300, 1000, 1100):
    run(_ris_scene(), "minpath", "hard_sigmoid", steps, 1, min_order=1, alpha=10.0)
    run(_ris_scene(), "minpath", "hard", steps, 1, min_order=1)
run(_vertex_scene(), "fermat", "hard_sigmoid", 10, 3, n=3, m=3)
run(_vertex_scene(), "minpath", "hard_sigmoid", 10, 3, n=3, m=3)
