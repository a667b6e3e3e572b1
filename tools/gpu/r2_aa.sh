set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2aa_tests.log
tail -6 gpurun_out/r2aa_tests.log
python bench.py --steps 30 --warmup 3 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; tail -3 gpurun_out/r2aa_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aa_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','final_loss')}, d['e2e'], d['clocks'])
print({k: round(v['us_per_launch'],1) for k,v in d['kernels'].items()})
PY
