set -x
cd /root/repo; mkdir -p gpurun_out
timeout 300 python tests/timeline_bwd1.py > gpurun_out/r2m_timeline_bwd1.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_timeline_bwd1.log
grep -E "^knock|rc=|Error|error" gpurun_out/r2m_timeline_bwd1.log
