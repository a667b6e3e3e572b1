cd /root/repo
timeout 300 python -m pytest tests/test_gpu_rotary.py -m gpu -x -q 2>&1 | tail -2
python tools/dbg/rotary_pair_bench.py 2>&1 | tail -3
