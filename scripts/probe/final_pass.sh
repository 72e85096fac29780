set -x
python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
bash scripts/ncu_launches.sh r02s --no-render --no-configs > /dev/null 2>&1
python scripts/timeline_step.py > gpurun_out/timeline_r02s.txt 2> gpurun_out/timeline_r02s.err
ncu --set full --clock-control none --import-source on -k "regex:solve_fwd_mixed_kernel|solve_bwd_replay_kernel" -s 2 -c 2 -o gpurun_out/prof_k1c -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-render --no-configs --no-graph > gpurun_out/ncu_k1c.log 2>&1
tail -c 200 gpurun_out/r02s_bench.err
