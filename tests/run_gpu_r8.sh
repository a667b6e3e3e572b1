#!/bin/bash
OUT=${1:-gpurun_out/r8}
mkdir -p $OUT
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > $OUT/$name.log 2>&1; echo "exit $?" | tee -a $OUT/$name.log; tail -n 8 $OUT/$name.log; }
run t4_sm100 python -m pytest tests/test_gpu_attention.py -q --no-header -rfE -p no:cacheprovider -x -k "sm100 and not raw_scores or masked or validation"
run t5_full python -m pytest tests/test_gpu_attention.py -q --no-header -rfE -p no:cacheprovider -k "full_size"
run t6_model python -m pytest tests/test_gpu_model.py -q --no-header -rfE -p no:cacheprovider
timeout 600 python bench.py --kernel-only --steps 10 --warmup 3 > $OUT/kernel_only.json 2> $OUT/kernel_only.err; echo "kernel-only exit $?"; cat $OUT/kernel_only.json; tail -3 $OUT/kernel_only.err
