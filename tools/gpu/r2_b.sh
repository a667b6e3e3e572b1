set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/run_bwd_once.py > gpurun_out/r2b_bwd_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_bwd_once.log
tail -c 3000 gpurun_out/r2b_bwd_once.log
