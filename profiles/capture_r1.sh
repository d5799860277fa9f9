#!/bin/bash
# Round-1 GPU capture (run under gpurun from the repo root): parity tests, bench (both arms), ncu launch list,
# ncu --set full of every kernel of a mid-epidemic day and of the dense edge pass.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-dense > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_pass|states_pre|post_prepare|nab_count|trace_|infect_kernel|test_prob' -s 640 -c 16 -f -o gpurun_out/prof_day python bench.py --steps 1 --warmup 1 --no-cpu --no-dense > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_pass_kernel' -s 5 -c 2 -f -o gpurun_out/prof_dense python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_dense.log 2>&1
ls -la gpurun_out
