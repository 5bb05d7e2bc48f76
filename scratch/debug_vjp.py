import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import differt2d_b200 as d
from differt2d_b200 import functional as F
from oracle import ref_torch as R
from tests import helpers as H

name, mode, alpha = sys.argv[1], sys.argv[2], float(sys.argv[3])
sc = {"basic": d.Scene.basic_scene(), "obstacle": d.Scene.square_scene_with_obstacle()}[name]
X, Y = H.jittered_grid(sc, 20, 22, seed=3)
rng = np.random.default_rng(7)
Zbar = rng.standard_normal(X.shape).astype(np.float32)
osc = H.oracle_scene_from_product(sc)
xys, kinds, phis = sc.packed_objects()
fixed = np.stack([p.xy for p in sc.transmitters.values()])
grid = np.stack([X, Y], -1).reshape(-1, 2)
fn = "sigmoid" if mode == "sigmoid" else "hard_sigmoid"

def both(zb, mn, mx, filt=()):
    with R.clean_gradients():
        Zo, go = R.power_map_and_vjp(osc, X, Y, zb, max_order=mx, min_order=mn, approx=True, alpha=alpha, function=fn,
                                     filter_nodes=filt, wrt=("xys",))
    out = F.power_bwd(F.TraceConfig(mode=mode, min_order=mn, max_order=mx, reduce_all=True, filter_nodes=tuple(filt)),
                      xys, fixed, grid, zb.reshape(-1), alpha=alpha, device="cuda", want=("objects",))
    a = out["objects"].cpu().numpy(); b = go["xys"].numpy()
    return a, b, np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

for k in range(3):
    a, b, e = both(Zbar, k, k)
    print("order", k, "rel err", e)
k = int(sys.argv[4]) if len(sys.argv) > 4 else 2
# bisect receivers
idx = np.arange(X.size)
while len(idx) > 1:
    half = idx[: len(idx) // 2]
    zb = np.zeros(X.size, np.float32); zb[half] = Zbar.reshape(-1)[half]
    a, b, e = both(zb.reshape(X.shape), k, k)
    absd = np.abs(a - b).max()
    print(len(idx), "first half abs diff", absd, "rel", e)
    zb2 = np.zeros(X.size, np.float32); zb2[idx[len(idx) // 2:]] = Zbar.reshape(-1)[idx[len(idx) // 2:]]
    a2, b2, e2 = both(zb2.reshape(X.shape), k, k)
    absd2 = np.abs(a2 - b2).max()
    idx = half if absd >= absd2 else idx[len(idx) // 2:]
r = int(idx[0])
print("receiver", r, grid[r], "zbar", Zbar.reshape(-1)[r])
zb = np.zeros(X.size, np.float32); zb[r] = 1.0
# per candidate for this receiver
cands = R.all_path_candidates(osc.n, k, k)
N = osc.n
for c in cands:
    filt = tuple(j for j in range(N) if j not in set(c.tolist()))
    a, b, e = both(zb.reshape(X.shape), k, k, filt)
    if np.abs(a - b).max() > 1e-3 * max(1.0, np.abs(b).max()):
        print("cand-set", c.tolist(), "abs diff", np.abs(a - b).max(), "scale", np.abs(b).max())
        print("got\n", a.reshape(N, 4)); print("want\n", b.reshape(N, 4))
        logic = R.Logic(True, alpha, fn)
        for cc in R.all_path_candidates(N, k, k, filter_nodes=filt):
            with torch.no_grad():
                pts, loss = R.image_path(osc, osc.transmitters["tx"], cc, torch.tensor(grid[r]))
                on = R.on_objects(osc, cc, pts, logic); it = R.intersects_with_objects(osc, cc, pts, logic)
                lt = logic.lt(loss, torch.tensor(1e-2))
            print("  cand", cc.tolist(), "on", float(on), "inter", float(it), "lt", float(lt), "loss", float(loss))
        break
