set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_graph_step.py tests/test_gpu_bottleneck.py tests/test_gpu_linear.py tests/test_gpu_optim.py -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2u_tests.log
tail -30 gpurun_out/r2u_tests.log
python bench.py --steps 20 --warmup 3 --no-bottleneck-leg > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -5 gpurun_out/r2u_bench.err
python bench.py --steps 20 --warmup 3 --no-graph --no-bottleneck-leg --no-cpu-baseline > gpurun_out/r2u_bench_nograph.json 2> gpurun_out/r2u_bench_nograph.err
python - <<PY
import json
for f in ('r2u_bench','r2u_bench_nograph'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','final_loss')}, d['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
