"""float64 census of why (tile, candidate) pairs survive / die, normalised or raw city scene, order 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import differt2d_b200 as d
from tests import helpers as H
coords = sys.argv[1] if len(sys.argv) > 1 else "normalised"
sc = d.Scene.from_geojson(H.geojson_text())
if coords == "normalised":
    sc = H.normalised(sc)
xys, _, _ = sc.packed_objects(); xys = xys.astype(np.float64)
tx = np.stack([p.xy for p in sc.transmitters.values()])[0].astype(np.float64)
N = xys.shape[0]
P1 = xys[:, 0]; t = xys[:, 1] - xys[:, 0]
m = np.stack([t[:, 1], -t[:, 0]], -1); L = np.linalg.norm(m, axis=-1); deg = L == 0; L[deg] = 1; n = m / L[:, None]
tt = (t * t).sum(-1); tt[tt == 0] = 1
X, Y = sc.grid(1024, 1024); X = X.astype(np.float64); Y = Y.astype(np.float64)
xz = -0.03
rng = np.random.default_rng(0)
def mirror(p, j): return p - 2 * ((p - P1[j]) @ n[j])[..., None] * n[j]
def cross_s(p, A, j):
    u = p - A; v = P1[j] - p
    un = u @ n[j]; vn = v @ n[j]
    with np.errstate(all="ignore"):
        g = vn / un
    Xp = p + g[..., None] * u
    return ((Xp - P1[j]) @ t[j]) / tt[j], Xp, g, un
cands = [(a, b) for a in range(N) for b in range(N) if a != b]
tiles = [(rng.integers(64), rng.integers(128)) for _ in range(150)]
tot = 0; sA = 0; sAB = 0; sABdeg = 0
pt_visits = 0; pt_last = 0; pt_on = 0; pt_right = 0
for (bx, by) in tiles:
    xs = X[0, bx * 16:(bx + 1) * 16]; ys = Y[by * 8:(by + 1) * 8, 0]
    box = np.array([[xs.min(), ys.min()], [xs.max(), ys.min()], [xs.min(), ys.max()], [xs.max(), ys.max()]])
    pts = np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2)
    for (a, b) in cands:
        tot += 1
        I1 = mirror(tx, a); I2 = mirror(I1, b)
        if deg[b]:
            keepA = True
        else:
            s2, _, g2, un2 = cross_s(box, I2, b)
            keepA = (np.sign(un2).min() != np.sign(un2).max()) or not (s2.max() < xz or s2.min() > 1 - xz)
        if not keepA: continue
        sA += 1
        keepB = True
        if not deg[a] and not deg[b]:
            R = mirror(mirror(box, b), a)
            u = tx - R; un = u @ n[a]; vn = (P1[a] - tx) @ n[a]
            g = vn / un; Xp = tx + g[:, None] * u; s1 = ((Xp - P1[a]) @ t[a]) / tt[a]
            keepB = (np.sign(un).min() != np.sign(un).max()) or not (s1.max() < xz or s1.min() > 1 - xz)
        if not keepB: continue
        sAB += 1
        if deg[a] or deg[b]: sABdeg += 1
        # per point
        s2p, X2, g2p, _ = cross_s(pts, I2, b) if not deg[b] else (np.zeros(len(pts)), pts, np.full(len(pts), -0.5), None)
        ok2 = (np.minimum(s2p, 1 - s2p) > xz)
        s1p, X1, g1p, _ = cross_s(X2, I1, a) if not deg[a] else (np.zeros(len(pts)), X2, np.full(len(pts), -0.5), None)
        ok1 = (np.minimum(s1p, 1 - s1p) > xz)
        pt_visits += len(pts); pt_last += ok2.sum(); pt_on += (ok1 & ok2).sum()
        right = (g2p > -1) & (g2p < 0) & (g1p > -1) & (g1p < 0)
        pt_right += (ok1 & ok2 & right).sum()
print(coords, "tile-cands", tot, "survive A (last s-range) %.3f" % (sA / tot), "survive A&B %.3f" % (sAB / tot), "of which with degenerate wall %.3f" % (sABdeg / max(sAB, 1)))
print("per point among A&B survivors: pass last %.3f, pass on_objects %.3f, and right side %.3f" % (pt_last / pt_visits, pt_on / pt_visits, pt_right / pt_visits))
