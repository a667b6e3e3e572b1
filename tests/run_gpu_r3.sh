#!/bin/bash
OUT=${1:-gpurun_out/r3}
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n 12 $OUT/$name.log; }
run t4_sm100 python -m pytest tests/test_gpu_attention.py -q --no-header -rfE -p no:cacheprovider -k "sm100 or masked or validation"
run t5_full python -m pytest tests/test_gpu_attention.py -q --no-header -rfE -p no:cacheprovider -k "full_size"
run t6_model python -m pytest tests/test_gpu_model.py -q --no-header -rfE -p no:cacheprovider
run t7_smoke python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --kernel-only --steps 10 --warmup 3 > $OUT/kernel_only.json 2> $OUT/kernel_only.err; echo "kernel-only exit $?"; cat $OUT/kernel_only.json
timeout 600 python tests/profile_step.py > $OUT/profile_step.log 2>&1; echo "profile exit $?"; tail -60 $OUT/profile_step.log | cut -c1-200
