set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
bash scripts/ncu_launches.sh r02m --no-render --no-configs > /dev/null 2>&1
python scripts/timeline_step.py > gpurun_out/timeline_r02m.txt 2> gpurun_out/timeline_r02m.err
tail -c 300 gpurun_out/r02m_bench.err
