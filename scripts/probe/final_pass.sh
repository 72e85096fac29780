set -x
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
bash scripts/ncu_launches.sh r02g --no-render --no-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:td_fused_kernel -s 3 -c 1 -o gpurun_out/prof_t3b -f python scripts/time_td_kernels.py 12500 47360 cluster 3 > gpurun_out/ncu_t3b.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:solve_fwd_mixed_kernel|solve_bwd_replay_kernel" -s 2 -c 2 -o gpurun_out/prof_k1b -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-render --no-configs --no-graph > gpurun_out/ncu_k1b.log 2>&1
python scripts/timeline_step.py > gpurun_out/timeline_r02h.txt 2> gpurun_out/timeline_r02h.err
ls -la gpurun_out/*.ncu-rep | tail -3
tail -c 300 gpurun_out/r02g_bench.err
