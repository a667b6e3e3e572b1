set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2bm_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2bm_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 40 --warmup 3 > gpurun_out/r2bm_bench.json 2> gpurun_out/r2bm_bench.err; echo "bench rc=$?"
python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2bm_kernel_only.json 2> gpurun_out/r2bm_kernel_only.err
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2bm_bench_c4.json 2> gpurun_out/r2bm_bench_c4.err
python bench.py --config c5 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2bm_bench_c5.json 2> gpurun_out/r2bm_bench_c5.err
python - <<PY
import json
for f in ('r2bm_bench','r2bm_kernel_only','r2bm_bench_c4','r2bm_bench_c5'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','ms_per_token')}, d.get('e2e'), d.get('clocks'))
        if f == 'r2bm_kernel_only': print({n:round(v['us_per_launch'],1) for n,v in d['kernels'].items()})
        if f == 'r2bm_bench': print({k: (v['launches'], round(v['us_per_launch'],1)) for k,v in d['kernels'].items()}); print(d.get('roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2bm_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-bottleneck-leg --profile-steps 0 > gpurun_out/r2bm_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
