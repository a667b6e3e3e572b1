set -x
cd /root/repo; mkdir -p gpurun_out
python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2ab_kernel_only.json 2> gpurun_out/r2ab_kernel_only.err &&
ncu --set full --clock-control none --import-source on -k regex:"attn_bwd1|xattn|attn_fwd_persist" -s 6 -c 3 -o gpurun_out/r2ab_attn python bench.py --kernel-only --steps 3 --warmup 2 > gpurun_out/r2ab_ncu.log 2>&1
python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2ab_bench_c4.json 2> gpurun_out/r2ab_bench_c4.err
python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/r2ab_bench_c5.json 2> gpurun_out/r2ab_bench_c5.err; tail -3 gpurun_out/r2ab_bench_c5.err
python - <<PY
import json
for f in ('r2ab_kernel_only','r2ab_bench_c4','r2ab_bench_c5'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','ms_per_token')}, d.get('e2e'))
        print('  ', {k: round(v['us_per_launch'],1) for k,v in (d.get('kernels') or {}).items() if 'attn' in k})
    except Exception as e: print(f, 'ERR', e)
PY
