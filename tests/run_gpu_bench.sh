#!/bin/bash
# bench + ncu evidence on the GPU box. Usage: bash tests/run_gpu_bench.sh <outdir>
OUT=${1:-gpurun_out/bench}
mkdir -p $OUT
python -m pytest tests/test_gpu_attention.py -q --no-header -p no:cacheprovider -k "layouts" > $OUT/t_layouts.log 2>&1; tail -3 $OUT/t_layouts.log
timeout 600 python bench.py --kernel-only --steps 10 --warmup 3 > $OUT/kernel_only.json 2> $OUT/kernel_only.err; echo "kernel-only exit $?"; cat $OUT/kernel_only.json
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python bench.py --kernel-only --steps 2 --warmup 1 > $OUT/plain_kernel_only.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_kernel_only.csv python bench.py --kernel-only --steps 2 --warmup 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 300 python bench.py --kernel-only --steps 2 --warmup 1 > $OUT/plain_kernel_only2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 6 -o $OUT/prof_attn python bench.py --kernel-only --steps 2 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
