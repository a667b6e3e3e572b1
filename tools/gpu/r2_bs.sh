cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bottleneck-leg 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'])"
