set -x
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cross_attention.py -m gpu -x -q > gpurun_out/r2y_xattn.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_xattn.log; tail -25 gpurun_out/r2y_xattn.log
