#!/bin/bash
# bench without the CPU / dense legs + one ncu --set full capture of the per-agent kernels of a mid-epidemic day
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-dense > gpurun_out/bench_quick.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value %.4g  ms/step %.2f  us/day %.1f  e2e %.4g" % (d["value"], d["ms_per_step"], d["us_per_day"], d["e2e"]["value"]))
for k,v in d["kernels"].items(): print("  %-24s %6.1f us  share %.3f  frac %s" % (k, v["us_per_launch"], v["share_of_step"], v.get("frac")))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'states_pre|post_prepare|nab_count|infect_kernel|test_prob' -s 450 -c 5 -f -o gpurun_out/prof_agent python bench.py --steps 1 --warmup 1 --no-cpu --no-dense > gpurun_out/ncu_agent.log 2>&1
