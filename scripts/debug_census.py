"""GPU diagnostic (needs the --debug-counters build): census of the forward kernel's culls on the bench workload.
    D2D_B200_LIB=differt2d_b200/_lib/libdiffert2d_b200_dbg.so python scripts/debug_census.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as B
from differt2d_b200 import _lib as L, functional as F
NAMES = {0: "tile-cull evaluations", 1: "rule 1 last interaction", 2: "rule 1 earlier stage", 3: "rule 1' unfolded first",
         4: "rule 3 zero-length", 5: "rule 2 wrong side", 6: "tile survivors", 7: "  survivors with zero-length wall, not provably dead",
         8: "  kept: u.n changes sign", 9: "  kept: NaN", 10: "  kept: u.n near zero", 11: "rule 1' evaluated, kept", 12: "rule 1' not applicable",
         20: "macro-tile tests (cluster stage)", 21: "  macro survivors", 13: "warp-cull tests", 14: "warp-cull kept (= warp visits)", 16: "thread visits", 17: "thread passes on_objects",
         18: "  of which with zero-length wall", 22: "thread evaluates the loss", 23: "  killed by the loss",
         24: "thread enters the occlusion fold", 26: "  killed by the fold", 25: "    at the first object tested (the hint)",
         27: "fold: (segment, object) tests", 28: "fold:   of which reach the exact divisions", 19: "thread valid != 0"}
lib = L.lib()
lib.d2d_debug_counters.argtypes = [C.c_int32, C.c_void_p, C.c_int32]
for coords in ("raw", "normalised"):
    sc = B.load_scene(coords)
    xys, _, _ = sc.packed_objects()
    fixed = np.stack([p.xy for p in sc.transmitters.values()])
    X, Y = sc.grid(1024, 1024)
    grid = np.stack([X, Y], -1).reshape(-1, 2).astype(np.float32)
    out = (C.c_uint64 * 32)()
    lib.d2d_debug_counters(1, None, 1)
    Z = F.power_fwd(F.TraceConfig(mode="hard_sigmoid", max_order=2, grid_cols=1024), xys, fixed, grid, alpha=100.0, device="cuda")
    torch.cuda.synchronize()
    lib.d2d_debug_counters(1, C.addressof(out), 0)
    print("==", coords)
    for i, nm in NAMES.items():
        print(f"  {out[i]:12d}  {nm}")
