#!/bin/bash
# Full GPU pass: test suite, kernel-only bench, full bench.  Usage: bash tests/run_gpu_all.sh <outdir>
OUT=${1:-gpurun_out/all}
mkdir -p $OUT
bash tests/run_gpu_suite.sh $OUT > $OUT/suite.log 2>&1; grep -E "^===|passed|failed|error|exit" $OUT/suite.log
timeout 600 python bench.py --kernel-only --steps 10 --warmup 3 > $OUT/kernel_only.json 2> $OUT/kernel_only.err; echo "kernel-only exit $?"; cat $OUT/kernel_only.json; tail -3 $OUT/kernel_only.err
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
