"""GPU diagnostic: find (point, candidate) pairs whose validity differs between cull and no-cull."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from tests import helpers as H

sc = H.normalised(d.Scene.from_geojson(H.geojson_text()))
n = 96
X, Y = H.jittered_grid(sc, n, n, seed=5)
grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
xys, _, _ = sc.packed_objects()
fixed = np.stack([p.xy for p in sc.transmitters.values()])
mode = sys.argv[1] if len(sys.argv) > 1 else "sigmoid"
res = {}
for cull in (True, False):
    cfg = F.TraceConfig(mode=mode, min_order=3, max_order=3, grid_cols=n, cull=cull)
    Z, v = F.power_fwd(cfg, xys, fixed, grid, alpha=100.0, want_valid=True, device="cuda")
    res[cull] = (Z, v)
Za, va = res[True]; Zb, vb = res[False]
print("Z differ:", int((Za != Zb).sum()), "valid differ:", int((va != vb).sum()))
idx = torch.nonzero(va != vb)
cands = F.candidates(xys.shape[0], 3)
deg = np.where(np.abs(xys[:, 1] - xys[:, 0]).sum(-1) == 0)[0]
print("degenerate walls:", deg)
rows = []
for t, r, c in idx[:4000].cpu().numpy():
    rows.append((r, c, *cands[c], float(va[t, r, c]), float(vb[t, r, c])))
rows = np.array(rows)
if len(rows):
    print("first 30:")
    for row in rows[:30]:
        r = int(row[0])
        print(f"r={r} ({grid[r,0]:.6f},{grid[r,1]:.6f}) tile=({(r % n)//16},{(r // n)//8}) cand={row[2:5].astype(int)} cull={row[5]:.3e} nocull={row[6]:.3e}")
    print("max nocull value among differing:", rows[:, 6].max(), "min:", rows[:, 6].min())
    has_deg = np.isin(rows[:, 2:5].astype(int), deg).any(-1)
    print("fraction with degenerate wall:", has_deg.mean())
    np.save("gpurun_out/cull_diff.npy", rows)
np.save("gpurun_out/cull_xys.npy", xys); np.save("gpurun_out/cull_grid.npy", grid); np.save("gpurun_out/cull_fixed.npy", fixed)
