cd /root/repo; python tools/dbg/gemm_qkv.py 2>&1 | tail -9
