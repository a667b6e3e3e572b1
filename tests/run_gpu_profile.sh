#!/bin/bash
# ncu evidence on the GPU box (one gpurun call): launch list of the bench command + one full capture of the
# attention kernels.  Usage: bash tests/run_gpu_profile.sh <outdir>
OUT=${1:-gpurun_out/prof}
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > $OUT/plain_bench.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
KCMD="python bench.py --kernel-only --steps 2 --warmup 1"
timeout 300 $KCMD > $OUT/plain_kernel_only.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -o $OUT/prof_attn $KCMD > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
