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


def test_transtree_and_story_of_a_device_sim():
    ''' sim.make_transtree() (a host view of the device infection log) and people.story() after a run on the GPU '''
    import numpy as np
    import covasim_b200 as cv
    sim = cv.Sim(pop_size=3000, pop_type='hybrid', n_days=40, pop_infected=60, rand_seed=7, verbose=0, beta=0.03,
                 interventions=[cv.test_prob(symp_prob=0.3, asymp_prob=0.02, start_day=5), cv.contact_tracing(trace_probs=0.5, start_day=8)])
    sim.run()
    tt = sim.make_transtree()
    log = sim.infection_log
    assert len(tt) == len(log['target']) == int(sim.summary['cum_infections'])
    assert len(tt.transmissions) == int((log['source'] >= 0).sum())
    assert tt.n_targets.sum() <= len(tt.transmissions) and tt.r0() > 0.5
    spreader = int(np.bincount(log['source'][log['source'] >= 0]).argmax())
    lines = sim.people.story(spreader, quiet=True)
    assert lines[0].startswith(f'This is the story of {spreader}, a ') and 'COVID' in lines[0]
    assert sum('gave COVID to' in l for l in lines) == int((log['source'] == spreader).sum())
    assert any('became infectious' in l for l in lines)
    never = int(np.nonzero(sim.people.to_numpy('naive'))[0][0])
    assert any('did not contract COVID' in l for l in sim.people.story(never, quiet=True))
