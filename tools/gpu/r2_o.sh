set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/run_bwd_once.py > gpurun_out/r2o_bwd_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_bwd_once.log
grep -E "one_pass|rc=|B=" gpurun_out/r2o_bwd_once.log | cut -c1-110 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9}' | sort | uniq -c | sort -rn | head -5
for v in "" _late; do
  SVAE_LIB_VARIANT=$v python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2o_kernel_only$v.json 2> gpurun_out/r2o_kernel_only$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2o_kernel_only$v.json'))
print('variant "$v":', {k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
done
