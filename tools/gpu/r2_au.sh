cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/dbg/profile_small_ops.py 2>&1 | grep -v Warn | tee gpurun_out/r2au_small_ops.log | tail -50
timeout 300 python -m pytest tests/test_gpu_cross_attention.py -m gpu -x -q 2>&1 | tail -2
python bench.py --config c5 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', d['ms_per_token'], d['value'], d['e2e']['value'])"
