cd /root/repo
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k checkpointing 2>&1 | grep -E "^E |Error|assert" | head -20
