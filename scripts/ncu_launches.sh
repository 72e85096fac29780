#!/bin/bash
# ncu launch list (every kernel with its device time) of one short bench run; shares, not absolutes.
# usage (under gpurun): bash scripts/ncu_launches.sh <tag> [extra bench args]
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_bench_$tag.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launches_${tag}_summary.txt
python scripts/summarize_launches.py gpurun_out/launches_$tag.csv td_fused_kernel > gpurun_out/launches_${tag}_one_step.txt
head -40 gpurun_out/launches_${tag}_one_step.txt
