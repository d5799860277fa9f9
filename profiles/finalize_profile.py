''' Where Sim.finalize() spends its host time (the last ~1 ms of bench.py's e2e step): cProfile of finalize() after a C2 run. '''
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import covasim_b200 as cv  # noqa: E402

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sim = cv.Sim(pop_size=pop, pop_type='hybrid', n_days=180, pop_infected=pop // 200, rand_seed=1, verbose=0, pop_exact=False,
             interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=20), cv.contact_tracing(trace_probs=0.3, start_day=30)])
sim.initialize()
snap = sim.snapshot()
for rep in range(3):
    sim.restore(snap)
    sim.set_seed()
    sim._advance(sim.npts)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    sim.finalize()
    pr.disable()
    el = time.perf_counter() - t0
print(f'finalize: {1e3 * el:.3f} ms')
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22)
print(s.getvalue()[:4000])
