cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_data_parallel_nccl.py -m gpu -x -q 2>&1 | tail -5
for ov in 1 0; do
SVAE_DP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2aj_bench_n2_ov$ov.json 2> gpurun_out/r2aj_bench_n2_ov$ov.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2aj_bench_n2_ov$ov.json').read().strip().splitlines()[-1])
print('overlap=$ov', d['ms_per_step'], d['value'], d['e2e'])
PY
done
