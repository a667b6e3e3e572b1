cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -2
python tools/dbg/bwd_flush_repro.py noflush 2>&1 | tail -3
python bench.py --kernel-only --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({n:round(v['us_per_launch'],1) for n,v in d['kernels'].items()})"
