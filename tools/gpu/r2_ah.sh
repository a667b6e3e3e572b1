cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/dbg/bwd_watch.py 2>&1 | tail -30
timeout 300 python tools/dbg/bwd_watch.py 5 4096 8 64 2>&1 | tail -30
python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2ah_kernel_only.json 2> gpurun_out/r2ah_kernel_only.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2ah_kernel_only.json'))
print({k: round(x['us_per_launch'],1) for k,x in d['kernels'].items()})
PY
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -3
