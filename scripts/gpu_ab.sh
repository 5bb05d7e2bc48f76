#!/bin/bash
# A/B timing of library variants (scripts/build_variant.py) + a parity subset on the default build.
# Usage: bash scripts/gpu_ab.sh <tag> "<pytest -k expr or empty>" <variant> [<variant> ...]     ("default" = the shipped .so)
TAG=${1:-ab}; shift
K="$1"; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$K" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
fi
for v in "$@"; do
  if [ $v = default ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/differt2d_b200/_lib/variants/lib_$v.so; fi
  timeout 300 python bench.py --only-dense --steps 5 > $OUT/dense_$v.json 2> $OUT/dense_$v.err
  CS="raw"; if [ $v = default ] || [ $v = base ]; then CS="raw normalised"; fi
  for c in $CS; do
    timeout 300 python bench.py --coords $c --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $OUT/city_${c}_$v.json 2> $OUT/city_${c}_$v.err
  done
  python - <<PY
import json
def ld(p):
    try: return json.load(open(p))
    except Exception as e: return None
d=ld("$OUT/dense_$v.json"); r=ld("$OUT/city_raw_$v.json"); n=ld("$OUT/city_normalised_$v.json")
s="%-10s" % "$v"
if d: s+=" dense fwd %.3f bwd %.3f |" % (d["fwd_ms"], d["bwd_ms"])
for nm,l in (("raw",r),("norm",n)):
    if l:
        k=l.get("kernel_split_eager_pass") or {}
        s+=" %s step %.3f fwd %.3f bwd %.3f e2e %.3f |" % (nm, l["ms_per_step"], k.get("fwd_ms",-1), k.get("bwd_ms",-1), l["e2e"]["ms_per_step"])
print(s)
PY
done
unset D2D_B200_LIB
