"""GPU diagnostic: (grid point, candidate) pairs whose validity differs between the culled and the unculled launch
of the transmitters-grid role.   python scripts/debug_txcull.py [scene] [alpha] [mode] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from tests.test_gpu_parity import SCENES
name = sys.argv[1] if len(sys.argv) > 1 else "basic"
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
mode = sys.argv[3] if len(sys.argv) > 3 else "hard_sigmoid"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 384
sc = SCENES[name]
sc = sc.update_receivers(rx2=d.Point(xy=sc.center() + np.float32(0.31) * (sc.bounding_box()[1] - sc.center())))
X, Y = sc.grid(n, n)
grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
xys, _, _ = sc.packed_objects()
fixed = np.stack([p.xy for p in sc.receivers.values()])
kw = dict(max_order=2, grid_cols=n, grid_role="transmitters")
for t in range(fixed.shape[0]):
    fx = fixed[t:t + 1]
    a, va = F.power_fwd(F.TraceConfig(mode=mode, **kw), xys, fx, grid, alpha=alpha, want_valid=True, device="cuda")
    b, vb = F.power_fwd(F.TraceConfig(mode=mode, cull=False, **kw), xys, fx, grid, alpha=alpha, want_valid=True, device="cuda")
    diff = (va != vb).nonzero().cpu().numpy()
    print("fixed", t, fx.tolist(), "differing (t, r, c):", len(diff))
    cands = sc.all_path_candidates(0, 2)
    for (_, r, c) in diff[:10]:
        print("  r", int(r), "row", int(r) // n, "col", int(r) % n, "tx", grid[r].tolist(), "cand", cands[c].tolist(),
              "culled", float(va[0, r, c]), "unculled", float(vb[0, r, c]))
print("xys", xys.tolist())
