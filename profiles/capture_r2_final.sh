#!/bin/bash
# Round-2 closing capture on one B200 (one gpurun call): GPU test suite, smoke, the default bench, the reference arm, and the ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-dense --no-ensemble > gpurun_out/final_launches_bench.log 2>&1
gzip -f gpurun_out/final_launches.csv
