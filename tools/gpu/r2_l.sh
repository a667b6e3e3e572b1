set -x
cd /root/repo; mkdir -p gpurun_out
for v in "" _m1 _m2a _m2b _m2c; do
  SVAE_LIB_VARIANT=$v python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2l_kernel_only$v.json 2> gpurun_out/r2l_kernel_only$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2l_kernel_only$v.json'))
print('variant "$v":', {k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
done
