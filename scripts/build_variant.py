#!/usr/bin/env python
"""A/B timing builds: recompiles the ImagePath forward / backward units with extra nvcc flags and links them with the
default build's other objects into differt2d_b200/_lib/variants/lib_<name>.so (select at run time with D2D_B200_LIB).
    python scripts/build_variant.py <name> <flags...>      e.g.  nowarp -DD2D_WARP_FULL_CULL=0
"""
import concurrent.futures as cf, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differt2d_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.OUT_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
hot = [u for u in B.UNITS if u[1].startswith(("d2d_forward_m", "d2d_backward_m"))]
rest = [os.path.join(B.OUT_DIR, u[1]) for u in B.UNITS if u not in hot]

def one(item):
    src, objname, extra = item
    obj = os.path.join(out_dir, f"{name}_{objname}")
    cmd = [B._nvcc(), *B.ARCH, *B.COMMON, *extra, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return obj

with cf.ThreadPoolExecutor(max_workers=6) as ex:
    objs = list(ex.map(one, hot))
lib = os.path.join(out_dir, f"lib_{name}.so")
r = subprocess.run([B._nvcc(), *B.ARCH, "-shared", "-o", lib, *objs, *rest, "--cudart", "static"], capture_output=True, text=True)
if r.returncode: raise RuntimeError(r.stderr)
for o in objs: os.remove(o)
print(lib)
