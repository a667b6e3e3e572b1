set -x
cd /root/repo; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_data_parallel_nccl.py -m gpu -x -q > gpurun_out/r2x_nccl_test.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_nccl_test.log; tail -12 gpurun_out/r2x_nccl_test.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2x_bench_n2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-graph > gpurun_out/r2x_bench_n2_nograph.json 2> gpurun_out/r2x_bench_n2_nograph.err; echo "bench n2 nograph rc=$?"
python - <<PY
import json
for f in ('r2x_bench_n2','r2x_bench_n2_nograph'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','final_loss','n_gpus')}, d['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
