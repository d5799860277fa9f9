#!/bin/bash
# quick GPU loop: parity tests (single-GPU sim, ops, partition), then the bench without the CPU leg
timeout 900 python -m pytest tests/test_gpu_sim.py tests/test_gpu_ops.py tests/test_gpu_partition.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu $BENCH_ARGS > gpurun_out/bench_quick.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value %.4g  ms/step %.2f  us/day %.1f  e2e %.4g" % (d["value"], d["ms_per_step"], d["us_per_day"], d["e2e"]["value"]))
for k,v in d["kernels"].items(): print("  %-24s %6.1f us  share %.3f  frac %s" % (k, v["us_per_launch"], v["share_of_step"], v.get("frac")))
print(d["edge_pass_dense"])
PY
