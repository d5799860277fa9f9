'''
TEST INFRASTRUCTURE -- records what the UNMODIFIED reference's TransTree (analysis.py:1772-1925) computes for two golden scenarios,
as tests/golden/transtree_ref.npz: count_targets() over the whole run and over a window of days, and the transmissions list.
Run from the repo root:  python -m oracle.gen_transtree_golden
The infection logs of the same runs are in tests/golden/<scenario>.npz, so tests/test_post_cpu.py can rebuild the tree with
covasim_b200.analysis.TransTree from plain arrays and compare.
'''
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refenv  # noqa: E402
cv = refenv.import_reference()
import scenarios  # noqa: E402

# the 'detailed' dataframe (analysis.py:1928-2001) needs sciris features the stand-in does not have and is not recorded here
cv.TransTree.make_detailed = lambda self, people, reset=False: None

out = {}
for name in ('hybrid3k', 'variants4k'):
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
    sim.run()
    tt = cv.TransTree(sim)
    out[f'{name}/n_targets'] = np.array(tt.count_targets())
    out[f'{name}/n_targets_10_30'] = np.array(tt.count_targets(start_day=10, end_day=30))
    out[f'{name}/transmissions'] = np.array(tt.count_transmissions(), dtype=np.int64).reshape(-1, 2)
    out[f'{name}/sources'] = np.array([-1 if s is None else s for s in tt.sources], dtype=np.int64)
    out[f'{name}/shape'] = np.array([len(tt), tt.pop_size, tt.n_days], dtype=np.int64)
    print(name, len(tt), out[f'{name}/n_targets'].sum(), len(out[f'{name}/transmissions']))
path = os.path.join(ROOT, 'tests', 'golden', 'transtree_ref.npz')
np.savez_compressed(path, **out)
print('wrote', path, os.path.getsize(path))
