#!/bin/bash
# compute-sanitizer memcheck over small runs of every path: single-GPU sim with all interventions (smoke), dense / dynamic layers,
# agent partition with in-process ranks, device population generation
export PYTHONPATH=$PWD:$PWD/tests
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python - <<'PY' > gpurun_out/sanitize.log 2>&1
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
import covasim_b200 as cv, scenarios
from covasim_b200 import partition as cvpart
for name in ('dynamic2k', 'variants4k'):
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name])); sim.run(); print(name, sim.summary['cum_infections'])
sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS['hybrid3k']), use_adjacency=False); sim.run(); print('dense', sim.summary['cum_infections'])
spec = scenarios.SCENARIOS['hybrid3k']
comms = cvpart.LocalComm.make(3)
sims = [cv.Sim(**scenarios.build(cv, spec), partition=comms[r], pop_gen='device') for r in range(3)]
cvpart.run_local(sims, lambda s: s.initialize()); cvpart.run_local(sims, lambda s: s.run())
print('partitioned', sims[0].summary['cum_infections'])
torch.cuda.synchronize()
print('SANITIZE DONE')
PY
echo "sanitizer exit $?" >> gpurun_out/sanitize.log
tail -15 gpurun_out/sanitize.log
