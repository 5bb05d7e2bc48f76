"""GPU diagnostic: Newton fast mode vs ImagePath vs Adam on one receiver (path vertices per candidate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from tests import helpers as H
sc = H.generic_position(d.Scene.square_scene_with_obstacle())
xys, _, _ = sc.packed_objects()
fixed = np.stack([p.xy for p in sc.transmitters.values()])
grid = np.array([[0.62, 0.31]], np.float32)
x0 = np.random.default_rng(5).random((65, 2), dtype=np.float32)
def rec(**kw):
    r = F.paths(F.TraceConfig(mode="hard", max_order=2, **kw), xys, fixed, grid, x0=x0, emit_all=True, device="cuda")
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in r.items()}
img = rec()
for name, kw in (("adam100", dict(method="fermat", steps=100)), ("newton12", dict(method="fermat", optimizer="newton", steps=12)),
                 ("newton40", dict(method="fermat", optimizer="newton", steps=40)), ("min_newton12", dict(method="minpath", optimizer="newton", steps=12))):
    r = rec(**kw)
    dx = np.abs(r["xys"] - img["xys"]).reshape(65, -1).max(-1)
    print(name, "valid equal:", int((r["valid"] == img["valid"]).sum()), "/65; max |X - X_image| over VALID image paths:",
          dx[img["valid"] > 0].max() if (img["valid"] > 0).any() else None, " median over all:", np.median(dx))
    for c in (1, 9, 10, 30):
        print("   cand", c, "order", r["order"][c], "image X", img["xys"][c, 1:3].round(4).tolist(), name, r["xys"][c, 1:3].round(4).tolist(),
              "loss", float(r["loss"][c]), "img loss", float(img["loss"][c]), "valid", r["valid"][c], img["valid"][c])

# the failing test's configuration: maps through power_fwd / power_bwd
X, Y = H.jittered_grid(sc, 20, 22, seed=3)
G = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
Zbar = (0.5 + np.random.default_rng(7).random(X.shape)).astype(np.float32).reshape(-1)
for mode in ("hard", "hard_sigmoid"):
    kw = dict(max_order=2, grid_cols=X.shape[1], reduce_all=True)
    zi = F.power_fwd(F.TraceConfig(mode=mode, **kw), xys, fixed, G, alpha=30.0, device="cuda").cpu().numpy()
    zib = F.power_bwd(F.TraceConfig(mode=mode, **kw), xys, fixed, G, Zbar, alpha=30.0, device="cuda")["Z"].cpu().numpy()
    for method in ("fermat", "minpath"):
        cfg = F.TraceConfig(mode=mode, method=method, optimizer="newton", steps=12, **kw)
        zf = F.power_fwd(cfg, xys, fixed, G, x0=x0, alpha=30.0, device="cuda").cpu().numpy()
        zb = F.power_bwd(cfg, xys, fixed, G, Zbar, x0=x0, alpha=30.0, device="cuda")["Z"].cpu().numpy()
        za = F.power_fwd(F.TraceConfig(mode=mode, method=method, steps=100, **kw), xys, fixed, G, x0=x0, alpha=30.0, device="cuda").cpu().numpy()
        c = lambda a, b: float(np.isclose(a, b, rtol=2e-3, atol=1e-4 * np.abs(b).max()).mean())
        print(mode, method, "image fwd==bwd", np.array_equal(zi, zib), "| newton fwd~image", c(zf, zi), "newton bwd~image", c(zb, zi),
              "newton fwd==bwd", np.array_equal(zf, zb), "| adam100~image", c(za, zi), " max", float(np.abs(zi).max()))
