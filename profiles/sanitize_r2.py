'''
Small runs of every round-2 path for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_r2.py
fused day pipeline with testing + tracing + vaccination days (hybrid), dynamic layer (dense pass inside the fused day), three variants
with bed limits and importations, an ensemble in lockstep (cvb_run_days_multi), snapshot / restore / copy, the taped infect entry point.
'''
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import covasim_b200 as cv  # noqa: E402
import scenarios  # noqa: E402

for name in ('hybrid3k', 'dynamic2k', 'variants4k', 'clip3k'):
    spec = scenarios.SCENARIOS[name]
    sim = cv.Sim(**scenarios.build(cv, spec, n_days=min(spec['pars']['n_days'], 25)))
    sim.initialize()
    snap = sim.snapshot()
    sim.run(until=12)
    twin = sim.copy()
    sim.run(reset_seed=False)
    twin.run(reset_seed=False)
    assert np.array_equal(sim.results['cum_infections'].values, twin.results['cum_infections'].values)
    sim.restore(snap)
    sim.run()
    print(name, 'fused days', sim.fused_days, 'of', sim.npts, 'cum_infections', sim.summary['cum_infections'], flush=True)
base = cv.Sim(pop_size=4000, pop_type='hybrid', n_days=20, pop_infected=80, rand_seed=3, verbose=0, beta=0.03,
              interventions=[cv.test_prob(symp_prob=0.2, asymp_prob=0.01, start_day=3), cv.contact_tracing(trace_probs=0.4, start_day=5)])
msim = cv.MultiSim(base, n_runs=4).run()
print('multisim', [float(r['cum_infections'][-1]) for r in msim.member_results], flush=True)
torch.cuda.synchronize()
print('ok')
