cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/dbg/dp_equiv.py 2>&1 | tail -32
