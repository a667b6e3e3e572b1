#!/bin/bash
# Quick GPU pass over the newest tests + model tests + short bench.  Usage: bash tests/run_gpu_new.sh <outdir> [pytest files...]
OUT=${1:-gpurun_out/new}; shift
mkdir -p $OUT
timeout 900 python -m pytest "$@" -q --no-header -rfE -p no:cacheprovider -x > $OUT/new_tests.log 2>&1; echo "new tests exit $?"; tail -n 30 $OUT/new_tests.log
timeout 600 python -m pytest tests/test_gpu_model.py -q --no-header -rfE -p no:cacheprovider > $OUT/t6_model.log 2>&1; echo "model exit $?"; tail -n 15 $OUT/t6_model.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
