set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2be_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2be_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 40 --warmup 3 > gpurun_out/r2be_bench.json 2> gpurun_out/r2be_bench.err; echo "bench rc=$?"
python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/r2be_bench_c5.json 2> gpurun_out/r2be_bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2be_bench_reference.json 2> gpurun_out/r2be_bench_reference.err; echo "ref rc=$?"
python - <<PY
import json
for f in ('r2be_bench','r2be_bench_c5','r2be_bench_reference'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','ms_per_token')}, d.get('e2e'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
