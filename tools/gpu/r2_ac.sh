set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/run_bwd_once.py > gpurun_out/r2ac_bwd_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2ac_bwd_once.log
grep -E "one_pass|rc=|B=|Error" gpurun_out/r2ac_bwd_once.log | cut -c1-110 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9}' | sort | uniq -c | sort -rn | head -6
python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2ac_kernel_only.json 2> gpurun_out/r2ac_kernel_only.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ac_kernel_only.json'))
print({k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -3
