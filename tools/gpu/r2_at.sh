set -x
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2at_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2at_gpu_suite.log
python bench.py --steps 40 --warmup 3 > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err; echo "bench rc=$?"
python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2at_kernel_only.json 2> gpurun_out/r2at_kernel_only.err
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2at_bench_c4.json 2> gpurun_out/r2at_bench_c4.err
python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/r2at_bench_c5.json 2> gpurun_out/r2at_bench_c5.err
python - <<PY
import json
for f in ('r2at_bench','r2at_kernel_only','r2at_bench_c4','r2at_bench_c5'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','ms_per_token')}, d.get('e2e'))
    except Exception as e: print(f, 'ERR', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2at_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-bottleneck-leg --profile-steps 0 > gpurun_out/r2at_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd1|attn_fwd_persist|gelu_bwd|vocab_ce16s" -s 8 -c 4 -o gpurun_out/r2at_top python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-bottleneck-leg --profile-steps 0 > gpurun_out/r2at_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2at_*
