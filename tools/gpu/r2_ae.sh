set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do timeout 200 python tests/run_bwd_once.py 16 8 4096 4 2>&1 | grep -E "dq:|Error" | cut -c1-120; done
python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2ae_kernel_only.json 2> gpurun_out/r2ae_kernel_only.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ae_kernel_only.json'))
print({k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
