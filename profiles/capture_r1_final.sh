#!/bin/bash
# Final round-1 capture after the agent-record change (the dense pass and the reference arm are those of capture_r1.sh / v4):
# parity tests, smoke, bench, ncu launch list, ncu --set full of the kernels of a mid-epidemic day.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-dense > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'edge_pass|states_pre|post_prepare|nab_count|trace_|infect_kernel|test_prob' -s 640 -c 9 -f -o gpurun_out/prof_day python bench.py --steps 1 --warmup 1 --no-cpu --no-dense > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
