set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
python tests/mma_bench.py > gpurun_out/r2a_mma_bench.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-bottleneck-leg > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
tail -c 600 gpurun_out/r2a_tests.log
