#!/bin/bash
# ncu --set full captures of the round's new kernels (one launch each) in a short bench run
ncu --set full --clock-control none --import-source on -k "regex:solve_colorless_kernel|solve_bwd_replay_kernel|render_mix_tiled_kernel|render_groups_cluster_kernel" \
    -s 4 -c 4 -o gpurun_out/prof_extra -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_extra.log 2>&1
ls -la gpurun_out/prof_extra.ncu-rep
