#!/bin/bash
# quick GPU loop for the dense edge pass: parity tests that cover it, the isolated timing of bench.py, one ncu --set full capture
timeout 600 python -m pytest tests/test_gpu_sim.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_quick.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"]); print(d["edge_pass_dense"])
PY
if [ "$1" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_pass_kernel' -s 5 -c 2 -f -o gpurun_out/prof_dense_q python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_dense_q.log 2>&1
fi
