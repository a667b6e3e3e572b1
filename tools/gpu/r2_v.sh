set -x
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log
tail -4 gpurun_out/r2v_tests.log
python tests/profile_step.py > gpurun_out/r2v_profile_step.log 2>&1
sed -n '/KERNELS/,$p' gpurun_out/r2v_profile_step.log | head -60
