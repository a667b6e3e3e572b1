set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/run_bwd_once.py 5 8 4096 4 > gpurun_out/r2e_bwd_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_bwd_once.log
timeout 300 python tests/timeline_bwd1.py > gpurun_out/r2e_timeline_bwd1.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_timeline_bwd1.log
python bench.py --kernel-only --steps 4 --warmup 2 > gpurun_out/r2e_kernel_only.json 2> gpurun_out/r2e_kernel_only.err &&
ncu --set full --clock-control none --import-source on -k regex:attn_bwd1 -s 2 -c 1 -o gpurun_out/r2e_bwd1 python bench.py --kernel-only --steps 4 --warmup 2 > gpurun_out/r2e_ncu.log 2>&1
cat gpurun_out/r2e_bwd_once.log gpurun_out/r2e_timeline_bwd1.log; head -c 600 gpurun_out/r2e_kernel_only.json
