set -x
cd /root/repo; mkdir -p gpurun_out
for cfg in "4 1 4096 4" "1 1 32768 4" "10 8 4096 4" "16 8 4096 4"; do
  timeout 120 python tests/run_bwd_once.py $cfg > gpurun_out/r2ad_once.log 2>&1; echo "cfg $cfg rc=$?"; grep -E "dq:|Error" gpurun_out/r2ad_once.log | cut -c1-120 | head -3
done
