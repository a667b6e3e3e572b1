set -x
cd /root/repo; mkdir -p gpurun_out
python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c_kernel_only.json 2> gpurun_out/r2c_kernel_only.err
SVAE_ATTN_BWD_TWO_PASS=1 python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2c_kernel_only_twopass.json 2>> gpurun_out/r2c_kernel_only.err
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err
tail -c 400 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_kernel_only.json | head -c 3000
