set -x
cd /root/repo; mkdir -p gpurun_out
timeout 120 python tests/run_xattn_once.py > gpurun_out/r2z_xattn_once.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_xattn_once.log; grep -v "^  File\|^    " gpurun_out/r2z_xattn_once.log | tail -8
timeout 600 python -m pytest tests/test_gpu_cross_attention.py -m gpu -x -q > gpurun_out/r2z_xattn.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_xattn.log; tail -12 gpurun_out/r2z_xattn.log
