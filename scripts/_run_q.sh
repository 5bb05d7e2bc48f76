mkdir -p gpurun_out/r02x
for v in default h8 h32 base default h32; do
  if [ $v = default ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/differt2d_b200/_lib/variants/lib_$v.so; fi
  for c in raw normalised; do
  timeout 120 python bench.py --coords $c --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02x/city_${c}_$v.json 2> gpurun_out/r02x/city_${c}_$v.err
  python -c "
import json; l=json.load(open('gpurun_out/r02x/city_${c}_$v.json')); k=l['kernel_split_eager_pass']; print('$v $c', 'step %.3f fwd %.3f bwd %.3f' % (l['ms_per_step'], k['fwd_ms'], k['bwd_ms']))"
  done
done
