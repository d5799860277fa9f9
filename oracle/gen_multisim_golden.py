'''
TEST INFRASTRUCTURE -- golden vectors for the ensemble reductions (reference run.py:220-374 MultiSim.reduce / mean / combine):
runs the UNMODIFIED reference's MultiSim on four small members and stores every member's result series next to what the
reference's reduce(), mean() and combine() make of them.  Run from the repo root:  python -m oracle.gen_multisim_golden
'''
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refenv  # noqa: E402
cv = refenv.import_reference()


def main():
    base = cv.Sim(pop_size=600, n_days=14, verbose=0, pop_infected=30, rand_seed=5, beta=0.03,
                  variants=[cv.variant('alpha', days=3, n_imports=10)])
    out = {}
    msim = cv.MultiSim(base, n_runs=4)
    msim.run(keep_people=False, verbose=0)
    keys = msim.sims[0].result_keys()
    vkeys = list(msim.sims[0].results['variant'].keys())
    out['keys'] = np.array(keys)
    out['vkeys'] = np.array(vkeys)
    for i, sim in enumerate(msim.sims):
        for k in keys:
            out[f'member{i}/{k}'] = np.array(sim.results[k].values)
        for k in vkeys:
            out[f'member{i}/variant/{k}'] = np.array(sim.results['variant'][k].values)
    for name, call in (('median', lambda m: m.reduce()), ('mean', lambda m: m.mean()), ('quant', lambda m: m.reduce(quantiles=dict(low=0.25, high=0.75)))):
        call(msim)
        for k in keys:
            r = msim.results[k]
            out[f'{name}/{k}/values'], out[f'{name}/{k}/low'], out[f'{name}/{k}/high'] = np.array(r.values), np.array(r.low), np.array(r.high)
        for k in vkeys:
            r = msim.results['variant'][k]
            out[f'{name}/variant/{k}/values'], out[f'{name}/variant/{k}/low'], out[f'{name}/variant/{k}/high'] = np.array(r.values), np.array(r.low), np.array(r.high)
    msim.combine()
    for k in keys:
        out[f'combine/{k}'] = np.array(msim.results[k].values)
    for k in vkeys:
        out[f'combine/variant/{k}'] = np.array(msim.results['variant'][k].values)
    path = os.path.join(ROOT, 'tests', 'golden', 'multisim_ref.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB, {len(out)} arrays)')


if __name__ == '__main__':
    main()
