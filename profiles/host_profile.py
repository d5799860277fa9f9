import os, sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import covasim_b200 as cv
n = 2_000_000
pars = dict(pop_size=n, pop_type='hybrid', n_days=60, pop_infected=n // 200, rand_seed=1, verbose=0, use_waning=True)
variants = [cv.variant('alpha', days=5, n_imports=100), cv.variant('delta', days=15, n_imports=100)]
ivs = [cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=10), cv.contact_tracing(trace_probs=0.3, start_day=15),
       cv.vaccinate_prob('pfizer', days=list(range(10, 30)), prob=0.01), cv.vaccinate_prob('pfizer', days=[40], prob=0.05, booster=True, label='booster')]
sim = cv.Sim(pars, variants=variants, interventions=ivs, pop_exact=False)
sim.initialize()
snap = sim.snapshot(pinned=False)
for rep in range(2):
    sim.restore(snap); sim.set_seed(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr = cProfile.Profile() if rep == 1 else None
    if pr: pr.enable()
    while not sim.complete:
        sim.step()
    if pr: pr.disable()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'rep {rep}: host loop {1e3*(t1-t0):.1f} ms, + sync {1e3*(t2-t1):.1f} ms')
    sim.finalize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(18); print(s.getvalue()[:3500])
