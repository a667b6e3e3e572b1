cd /root/repo; mkdir -p gpurun_out
for rc in 16384 32768 65536; do
SVAE_CE_ROW_CHUNK=$rc python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-bottleneck-leg --profile-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $rc', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
