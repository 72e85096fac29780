#!/bin/bash
# one `ncu --set full` capture of the kernels matching <regex> in a short bench run (1 GPU). usage: ncu_kernel.sh <tag> <regex> [skip] [count]
tag=$1; regex=$2; skip=${3:-2}; count=${4:-1}
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $count -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --tile-rows 12500 > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
