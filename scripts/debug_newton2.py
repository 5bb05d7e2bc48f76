import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from tests import helpers as H
sc = H.generic_position(d.Scene.square_scene())
xys, _, _ = sc.packed_objects()
fixed = np.stack([p.xy for p in sc.transmitters.values()])
X, Y = H.jittered_grid(sc, 20, 22, seed=3)
G = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
x0 = np.random.default_rng(5).random((17, 2), dtype=np.float32)
kw = dict(max_order=2, grid_cols=22, reduce_all=True)
zi = F.power_fwd(F.TraceConfig(mode="hard", **kw), xys, fixed, G, device="cuda").cpu().numpy()
zn = F.power_fwd(F.TraceConfig(mode="hard", method="fermat", optimizer="newton", steps=12, **kw), xys, fixed, G, x0=x0, device="cuda").cpu().numpy()
za = F.power_fwd(F.TraceConfig(mode="hard", method="fermat", steps=100, **kw), xys, fixed, G, x0=x0, device="cuda").cpu().numpy()
print("nan", np.isnan(zn).sum(), "zi[:6]", zi[:6], "zn[:6]", zn[:6], "za[:6]", za[:6])
print("rel newton", np.median(np.abs(zn - zi) / np.abs(zi)), "rel adam", np.median(np.abs(za - zi) / np.abs(zi)))
c = lambda a, b: float(np.isclose(a, b, rtol=1e-3, atol=1e-5 * np.abs(b).max()).mean())
print("close adam100", c(za, zi))
for method in ("fermat", "minpath"):
    for mode in ("hard", "hard_sigmoid"):
        zi_ = F.power_fwd(F.TraceConfig(mode=mode, **kw), xys, fixed, G, alpha=30.0, device="cuda").cpu().numpy()
        for steps in (8, 12, 20, 30):
            z = F.power_fwd(F.TraceConfig(mode=mode, method=method, optimizer="newton", steps=steps, **kw), xys, fixed, G, x0=x0, alpha=30.0, device="cuda").cpu().numpy()
            print(method, mode, "newton", steps, "close", c(z, zi_))
        z = F.power_fwd(F.TraceConfig(mode=mode, method=method, steps=100, **kw), xys, fixed, G, x0=x0, alpha=30.0, device="cuda").cpu().numpy()
        print(method, mode, "adam 100 close", c(z, zi_))
