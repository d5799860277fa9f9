'''
Replay mode on the GPU (cv.Sim(rng='mt')): the device path driven by the reference's two MT19937 streams must
reproduce the UNMODIFIED REFERENCE -- not just the oracle -- bit for bit: every result series, the final People arrays
and the infection log recorded in tests/golden by oracle/gen_golden.py, including the 58 values of the reference's own
tests/baseline.json.  (The three float32-mean results are compared at 1e-6: summation order.)
'''
import hashlib
import json

import numpy as np
import pytest

import scenarios

pytestmark = pytest.mark.gpu
LOOSE = ('pop_nabs', 'pop_protection', 'pop_symp_protection')


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


_runs = {}


def run_replay(cv, name):
    if name not in _runs:
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]), rng='mt')
        sim.run()
        _runs[name] = sim
    return _runs[name]


@pytest.mark.parametrize('name', list(scenarios.SCENARIOS.keys()))
def test_replay_reproduces_reference(cv, golden, name):
    g = golden(name)
    sim = run_replay(cv, name)
    for key in g.files:
        if key.startswith('results/'):
            k = key.split('/', 1)[1]
            np.testing.assert_allclose(sim.results[k].values, g[key], rtol=1e-6 if k in LOOSE else 1e-12, atol=0, equal_nan=True, err_msg=k)
        elif key.startswith('vresults/'):
            k = key.split('/', 1)[1]
            np.testing.assert_allclose(sim.results['variant'][k].values, g[key], rtol=1e-12, atol=0, equal_nan=True, err_msg=k)
    for k in cv.defaults.all_states:
        got = sim.people.to_numpy(k)
        if f'people/{k}' in g.files:
            assert np.array_equal(got, g[f'people/{k}'], equal_nan=True), k
        else:
            assert digest(got) == str(g[f'people_digest/{k}']), k
    log = sim.infection_log
    # the reference's log is in call order; compare as sorted multisets of (date, variant, layer, target, source)
    ref = np.stack([g['log/date'], g['log/variant'], g['log/layer'], g['log/target'], g['log/source']], axis=1)
    mine = np.stack([log['date'], log['variant'], log['layer'].astype(np.int32), log['target'], log['source']], axis=1)
    ref = ref[np.lexsort(ref.T[::-1])]
    mine = mine[np.lexsort(mine.T[::-1])]
    assert np.array_equal(ref, mine)


def test_replay_reproduces_baseline_json(cv, golden):
    ''' reference tests/test_baselines.py:81-95 against tests/baseline.json, run on the GPU '''
    base = json.loads(str(golden('baseline20k')['baseline_json']))
    assert len(base) == 58
    sim = run_replay(cv, 'baseline20k')
    for k, v in base.items():
        assert np.isclose(sim.summary[k], v, rtol=1e-6 if k in LOOSE else 1e-12, atol=0), (k, sim.summary[k], v)
