cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/dbg/profile_copies_by_site.py 2>&1 | grep -v Warn | tee gpurun_out/r2ba_copy_sites.log | tail -32
