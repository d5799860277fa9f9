'''
TEST INFRASTRUCTURE -- distribution of epidemic curves of the UNMODIFIED reference over many seeds, for the
statistical comparison of the native-RNG mode (north star: "epidemic curves over 200 seeds must be statistically
indistinguishable from the reference MultiSim").  Run in the build container:  python -m oracle.gen_stats
Writes tests/golden/stats_ref.npz: new_infections, new_deaths, cum_infections, new_diagnoses, new_quarantined as
float32[n_seeds, npts] plus the configuration.
'''
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refenv  # noqa: E402

N_SEEDS = 200
KEYS = ('new_infections', 'new_deaths', 'cum_infections', 'new_diagnoses', 'new_quarantined', 'n_exposed')
PARS = dict(pop_size=10000, pop_type='hybrid', n_days=60, pop_infected=50, verbose=0)
INTERVENTIONS = [('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=15)), ('contact_tracing', dict(trace_probs=0.3, start_day=20))]


def main():
    cv = refenv.import_reference()
    out = {k: [] for k in KEYS}
    t0 = time.time()
    for i in range(N_SEEDS):
        ivs = [getattr(cv, name)(**kw) for name, kw in INTERVENTIONS]
        sim = cv.Sim(dict(PARS, rand_seed=1000 + i), interventions=ivs)
        sim.run()
        for k in KEYS:
            out[k].append(sim.results[k].values.astype(np.float32))
        if i % 20 == 0:
            print(f'  seed {i}: cum_infections={sim.summary["cum_infections"]:.0f} ({time.time() - t0:.0f} s)')
    path = os.path.join(ROOT, 'tests', 'golden', 'stats_ref.npz')
    np.savez_compressed(path, config=np.array(json.dumps(dict(pars=PARS, interventions=INTERVENTIONS, first_seed=1000))),
                        **{k: np.stack(v) for k, v in out.items()})
    print(f'wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)')


if __name__ == '__main__':
    main()
