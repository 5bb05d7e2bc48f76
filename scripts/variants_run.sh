mkdir -p gpurun_out/r02g
python -m pytest tests/test_gpu_parity_round2.py -m gpu -q -p no:cacheprovider -k "sanitiser" 2>&1 | tail -5 > gpurun_out/r02g/pytest_san.log
for v in default b4 b3 f6 f10; do
  if [ $v = default ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=differt2d_b200/_lib/variants/lib_$v.so; fi
  python bench.py --only-dense --steps 5 > gpurun_out/r02g/dense_$v.json 2> gpurun_out/r02g/dense_$v.err
  python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02g/city_$v.json 2> gpurun_out/r02g/city_$v.err
done
unset D2D_B200_LIB
tail -3 gpurun_out/r02g/pytest_san.log
