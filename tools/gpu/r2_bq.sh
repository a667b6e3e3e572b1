cd /root/repo; mkdir -p gpurun_out
timeout 120 python tests/sanitize_small.py 2>&1 | tail -2
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tests/sanitize_small.py > gpurun_out/r2bq_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|sanitize_small" gpurun_out/r2bq_memcheck.log | head -10
