# register-cap sweep: bash scripts/variants_run.sh  (variants built with D2D_NVCC_EXTRA into differt2d_b200/_lib/variants/)
mkdir -p gpurun_out/r02j
python scripts/debug_newton.py > gpurun_out/r02j/newton.txt 2>&1
for v in default b6 b8; do
  if [ $v = default ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=differt2d_b200/_lib/variants/lib_$v.so; fi
  python bench.py --only-dense --steps 5 > gpurun_out/r02j/dense_$v.json 2> gpurun_out/r02j/dense_$v.err
  python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02j/city_$v.json 2> gpurun_out/r02j/city_$v.err
done
unset D2D_B200_LIB
cat gpurun_out/r02j/newton.txt
