cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -3
python tools/dbg/sample_time.py 2>&1 | tail -4
python bench.py --config c5 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', d['ms_per_token'], d['value'], d['e2e']['value'])"
