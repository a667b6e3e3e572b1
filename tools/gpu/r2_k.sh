set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/run_bwd_once.py 5 8 4096 4 > gpurun_out/r2k_bwd_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_bwd_once.log
timeout 300 python tests/timeline_bwd1.py > gpurun_out/r2k_timeline_bwd1.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_timeline_bwd1.log
python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2k_kernel_only.json 2> gpurun_out/r2k_kernel_only.err
grep -E "one_pass|rc=" gpurun_out/r2k_bwd_once.log | cut -c1-120; grep -E "^knock|rc=" gpurun_out/r2k_timeline_bwd1.log; head -c 300 gpurun_out/r2k_kernel_only.json
