"""CPU (float64) estimate of candidate survival under tile-cull variants on the geojson scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import differt2d_b200 as d
from tests import helpers as H

coords = sys.argv[1] if len(sys.argv) > 1 else "raw"
TW, TH = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 8)
sc = d.Scene.from_geojson(H.geojson_text())
if coords == "normalised":
    sc = H.normalised(sc)
xys, _, _ = sc.packed_objects()
xys = xys.astype(np.float64)
tx = np.stack([p.xy for p in sc.transmitters.values()])[0].astype(np.float64)
N = xys.shape[0]
P1 = xys[:, 0]; t = xys[:, 1] - xys[:, 0]
m = np.stack([t[:, 1], -t[:, 0]], -1); L = np.linalg.norm(m, axis=-1); L[L == 0] = 1; n = m / L[:, None]
tt = (t * t).sum(-1); tt[tt == 0] = 1
X, Y = sc.grid(1024, 1024)
X = X.astype(np.float64); Y = Y.astype(np.float64)
S = max(np.abs(X).max(), np.abs(Y).max())
eps = 2.0 ** -24
xz = -0.03
rng = np.random.default_rng(0)

def mirror(p, j):
    return p - 2 * ((p - P1[j]) @ n[j]) * n[j]

def backproj(q, A, j):
    u = q - A; v = P1[j] - q
    un = u @ n[j]; vn = v @ n[j]
    with np.errstate(all="ignore"):
        g = vn / un
    return q + g[..., None] * u, un, g

def spar(Xp, j):
    return ((Xp - P1[j]) @ t[j]) / tt[j]

cands = [(a,) for a in range(N)] + [(a, b) for a in range(N) for b in range(N) if a != b]
ntx, nty = 1024 // TW, 1024 // TH
tiles = [(rng.integers(ntx), rng.integers(nty)) for _ in range(200)]
deg = (np.abs(t).sum(-1) == 0)
tot = 0; pt_spec = 0; surv3 = 0; surv1 = 0; surv1tight = 0; surv2 = 0; pt_on_last = 0; pt_on_all = 0; npts = 0
for (bx, by) in tiles:
    xs = X[0, bx * TW:(bx + 1) * TW]; ys = Y[by * TH:(by + 1) * TH, 0]
    box = np.array([[xs.min(), ys.min()], [xs.max(), ys.min()], [xs.min(), ys.max()], [xs.max(), ys.max()]])
    pts = np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2)
    for c in cands:
        tot += 1
        A = tx
        Is = [A]
        for j in c:
            A = mirror(A, j); Is.append(A)
        j = c[-1]
        # exact per point
        Xp, unp, _ = backproj(pts, Is[-1], j)
        sp = spar(Xp, j)
        on_last = (np.minimum(sp, 1 - sp) > xz)
        on_all = on_last.copy()
        if len(c) == 2:
            X1, _, _ = backproj(Xp, Is[1], c[0])
            s1 = spar(X1, c[0])
            on_all &= np.minimum(s1, 1 - s1) > xz
        # specular: g in (-1,0) at every interaction (per point)
        _, unl, gl = backproj(pts, Is[-1], j)
        spec = (gl > -1) & (gl < 0)
        if len(c) == 2:
            _, _, g0 = backproj(Xp, Is[1], c[0])
            spec &= (g0 > -1) & (g0 < 0)
        anydeg = any(deg[k] for k in c)
        pt_spec += (on_all & spec & (not anydeg)).sum()
        pt_on_last += on_last.sum(); pt_on_all += on_all.sum(); npts += len(pts)
        # stage 1 at corners
        Xc, unc, gc = backproj(box, Is[-1], j)
        if not (np.all(unc > 0) or np.all(unc < 0)):
            surv1 += 1; surv1tight += 1; surv2 += 1; continue
        scn = spar(Xc, j)
        tlen = np.abs(t[j]).sum() / tt[j]
        E = eps * S * tlen * 1.0
        tol_t = 2.5 * E + 1e-6
        lo, hi = xz - tol_t, 1 - xz + tol_t
        keep_t = not (scn.max() < lo or scn.min() > hi)
        surv1tight += keep_t
        # old tol (approx)
        u = box - Is[-1]; ul = np.linalg.norm(u, axis=-1)
        gmax = np.abs(gc).max(); umax = ul.max(); unmin = np.abs(unc).min(); Gmax = (np.abs(gc) * ul).max()
        tolX = 16 * eps * S * (gmax + umax / unmin + Gmax / unmin) + 4 * eps * (S + Gmax)
        tol_o = 1e-4 + 4 * tolX / np.sqrt(tt[j])
        keep_o = not (scn.max() < xz - tol_o or scn.min() > 1 - xz + tol_o)
        surv1 += keep_o
        if not keep_t:
            continue
        if len(c) == 1:
            surv2 += 1; continue
        a = max(scn.min() - tol_t, lo); b = min(scn.max() + tol_t, hi)
        ends = P1[j] + np.array([[a], [b]]) * t[j]
        X1e, un1, g1 = backproj(ends, Is[1], c[0])
        if not (np.all(un1 > 0) or np.all(un1 < 0)):
            surv2 += 1; continue
        s1e = spar(X1e, c[0])
        j0 = c[0]
        tlen0 = np.abs(t[j0]).sum() / tt[j0]
        u1 = ends - Is[1]
        Lsens = (np.abs(1 + g1) * (1 + np.linalg.norm(u1, axis=-1) / np.abs(un1))).max()
        E1 = eps * S * tlen0 * (1 + Lsens)
        tol1 = 2.5 * E1 + 1e-6
        keep2 = not (s1e.max() < xz - tol1 or s1e.min() > 1 - xz + tol1)
        surv2 += keep2
print(f"per-point on_all & specular & non-degenerate: {pt_spec/npts:.4f}")
print(f"{coords} tiles {TW}x{TH}: candidates/tile {len(cands)}; survive old stage1 {surv1/tot:.4f}; tight stage1 {surv1tight/tot:.4f}; "
      f"tight stage1+2 {surv2/tot:.4f}; per-point on_last {pt_on_last/npts:.4f} on_all {pt_on_all/npts:.4f}")
