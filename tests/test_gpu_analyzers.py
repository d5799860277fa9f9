'''
Analyzers on a device sim (reference analysis.py:23-425, sim.py:677-678): called once per day after the day's transmission and
counts, reading the People device tensors -- the Analyzer classes and a plain callable.
'''
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_analyzers_in_a_device_sim():
    import covasim_b200 as cv
    seen = []
    snap = cv.snapshot(5, 10)
    hist = cv.age_histogram(days=[10])
    sim = cv.Sim(pop_size=2000, pop_type='hybrid', n_days=20, pop_infected=50, rand_seed=3, verbose=0, beta=0.03,
                 analyzers=[snap, hist, lambda s: seen.append((s.t, int(s.people.exposed.sum())))])
    sim.run()
    assert [t for t, _ in seen] == list(range(sim.npts))
    # the callable runs at the end of the step: its count is the day's n_exposed result
    assert [c for _, c in seen] == [int(v) for v in sim.results['n_exposed'].values]
    assert len(snap.snapshots) == 2 and snap.finalized and hist.finalized
    day10 = snap.get(10)
    assert int(day10['exposed'].sum()) == int(sim.results['n_exposed'].values[10])
    for state in ('exposed', 'severe', 'dead', 'tested', 'diagnosed'):
        want = np.histogram(day10['age'][~np.isnan(day10[f'date_{state}'])], bins=np.linspace(0, 100, 11))[0]
        assert np.array_equal(hist.get(10)[state], want), state
    assert hist.get(10)['exposed'].sum() > 50
    assert sim.get_analyzer(cv.snapshot) is snap and sim.get_analyzers(cv.age_histogram) == [hist]
