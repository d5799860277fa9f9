'''
The dense edge-streaming pass alone (cvb_edge_pass with no adjacency bound) on the day-60 state of the C2 sim, as bench.py's
"edge_pass_dense" leg times it; the launch shape comes from CVB_DENSE_VARIANT (0 default, 1/2 register-staged shapes, 3 bulk-copy staging).
    CVB_DENSE_VARIANT=3 python profiles/dense_variants.py
Prints one JSON line with the time per launch.  (That every variant gives the same run is checked by
tests/test_gpu_fused.py::test_bulk_copy_staged_dense_pass_gives_the_same_run.)
'''
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import covasim_b200 as cv  # noqa: E402

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pars = dict(pop_size=pop, pop_type='hybrid', n_days=180, pop_infected=max(1, int(0.005 * pop)), rand_seed=1, verbose=0)
sim = cv.Sim(pars, interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)], pop_exact=False)
sim.initialize()
sim.run(until=60)
call = cv._capi.call
h, st, t = sim._handle, sim._stream_ptr, sim.t
E = sum(len(l) for l in sim.people.contacts.values())
call('cvb_bind_adjacency', h, None, None, 0, 0)
call('cvb_update_states_pre', h, t, st)
call('cvb_post_and_prepare', h, t, st)
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device=sim.device)
times = []
for r in range(23):
    flush.max()
    call('cvb_prepare_transmission', h, t, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    call('cvb_edge_pass', h, t, st)
    b.record()
    torch.cuda.synchronize()
    if r >= 3:
        times.append(a.elapsed_time(b))
us = 1e3 * float(np.mean(times))
algo = 12 * E + 8 * pop
print(json.dumps(dict(variant=os.environ.get('CVB_DENSE_VARIANT', '0'), pop_size=pop, us_per_launch=round(us, 2), min_us=round(1e3 * min(times), 2),
                      gbs=round(algo / us / 1e3, 1))))
