mkdir -p gpurun_out/r02z
for v in default base default base; do
  if [ $v = default ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/differt2d_b200/_lib/variants/lib_$v.so; fi
  timeout 300 python scripts/bench_configs.py --only cfg3,cfg5b --steps 5 > gpurun_out/r02z/cfg_$v.jsonl 2> gpurun_out/r02z/cfg_$v.err
  python -c "
import json
for l in open('gpurun_out/r02z/cfg_$v.jsonl'):
    d=json.loads(l); print('$v %-100s %8.3f ms' % (d['config'][:100], d['ms']))"
  timeout 120 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02z/city_raw_$v.json 2> gpurun_out/r02z/city_raw_$v.err
  python -c "
import json; l=json.load(open('gpurun_out/r02z/city_raw_$v.json')); k=l['kernel_split_eager_pass']; print('$v raw', 'step %.3f fwd %.3f bwd %.3f' % (l['ms_per_step'], k['fwd_ms'], k['bwd_ms']))"
done
unset D2D_B200_LIB
timeout 120 python bench.py --only-dense --steps 5 > gpurun_out/r02z/dense_default.json; python -c "
import json; d=json.load(open('gpurun_out/r02z/dense_default.json')); print('dense fwd %.3f bwd %.3f' % (d['fwd_ms'], d['bwd_ms']))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:power_ -s 6 -c 2 -f -o gpurun_out/r02z/prof_dense python bench.py --only-dense --steps 1 > gpurun_out/r02z/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
