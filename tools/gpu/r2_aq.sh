cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/dbg/profile_gemms.py 2>&1 | grep -v Warning | tee gpurun_out/r2aq_gemms.log | tail -45
