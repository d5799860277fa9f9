#!/bin/bash
# time the dense edge pass variants (CVB_DENSE_VARIANT) with bench.py's isolated measurement
for v in "$@"; do
  CVB_DENSE_VARIANT=$v timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_var$v.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_var$v.json").read().strip().splitlines()[-1])
print("variant $v", d["edge_pass_dense"]["us_per_launch"], d["edge_pass_dense"]["frac"])
PY
done
