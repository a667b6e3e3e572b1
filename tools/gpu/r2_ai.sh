cd /root/repo; mkdir -p gpurun_out
python tools/dbg/gemm_shapes.py 2>&1 | tee gpurun_out/r2ai_gemm_shapes.log | tail -20
python tools/dbg/ext_event.py 2>&1 | tail -8
