cd /root/repo; mkdir -p gpurun_out
timeout 600 python tests/profile_decode.py 2>&1 | tail -40
nproc; cat /proc/cpuinfo | grep "model name" | head -1; uptime
