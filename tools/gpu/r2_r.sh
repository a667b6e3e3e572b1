set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
tail -5 gpurun_out/r2r_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e'])
print(d['bottleneck'])
print(d['attention'])
for r in d['rooflines']: print(r['kernel'], round(r['us_per_launch'],1), round(r['frac'],3), round(r['tensor_frac_burst'],3))
PY
