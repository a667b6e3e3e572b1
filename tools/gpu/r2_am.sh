cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_data_parallel_nccl.py -m gpu -x -q -k single_process 2>&1 | grep -E "^E  |assert|Error" | head -20
