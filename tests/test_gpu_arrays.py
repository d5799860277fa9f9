'''
The upward face of the boundary: ``sim.people.<field>`` is device memory, but the reference's plug-ins
treat People arrays as NumPy arrays.  The idioms below are taken from the reference's examples and tests
(examples/t05_custom_intervention.py, examples/t08_boosters.py:52, tests/test_immunity.py:137, 263, 283, covasim/utils.py:487-669)
and must work unchanged on the device arrays (covasim_b200/devarray.py); plus Sim.copy() of a running sim (base.py:444-446).
'''
import numpy as np
import pytest

import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


def test_numpy_idioms_on_people_arrays(cv):
    sim = cv.Sim(pop_size=4000, pop_type='hybrid', n_days=30, pop_infected=80, rand_seed=2, verbose=0, beta=0.03,
                 interventions=[cv.vaccinate_prob('pfizer', days=5, prob=0.3)])
    sim.run(until=20)
    P = sim.people
    doses = P.doses.copy()                                            # examples/t08_boosters.py:52
    assert type(doses).__name__ == 'DeviceArray' and doses.data_ptr() != P.doses.data_ptr()
    one_dose = cv.true(P.doses != 2)                                  # ... cv.true(sim.people.doses != 2)
    assert len(one_dose) == int((np.asarray(P.doses) != 2).sum())
    nab = P.nab.copy()                                                # tests/test_immunity.py:137
    assert np.array_equal(np.asarray(nab), P.nab.get())
    exposed_ever = np.isfinite(P.date_exposed)                        # tests/test_immunity.py:263, 283
    assert int(exposed_ever.sum()) == int(np.isfinite(P.to_numpy('date_exposed')).sum()) > 80
    assert np.array_equal(P.exposed.nonzero()[0].cpu().numpy(), np.nonzero(P.to_numpy('exposed'))[0])       # utils.py:506
    both = np.intersect1d(cv.true(P.exposed), cv.true(P.vaccinated))  # interventions.py:956-961 style set algebra
    assert np.array_equal(both, np.nonzero(P.to_numpy('exposed') & P.to_numpy('vaccinated'))[0])
    assert np.digitize(P.age, [0, 18, 65]).max() == 3                 # people.py:153
    assert float(P.age.mean()) == pytest.approx(float(np.mean(P.to_numpy('age'))), rel=1e-6)
    assert int(np.count_nonzero(P.dead)) == sim.people.count('dead')
    import pandas as pd                                               # base.py:1190-1193 People.to_df style
    df = pd.DataFrame({'age': np.asarray(P.age), 'exposed': np.asarray(P.exposed)})
    assert len(df) == 4000 and df.exposed.sum() == int(P.exposed.sum())


class protect_elderly:
    ''' reference examples/t05_custom_intervention.py: on a given day, people over 70 become non-susceptible '''
    def __init__(self, day):
        self.day, self.initialized, self.done = day, False, False

    def initialize(self, sim):
        self.initialized = True

    def __call__(self, sim):
        if sim.t == self.day:
            elderly = sim.people.age > 70
            sim.people.rel_sus[elderly] = 0.0
            self.n_protected = int(elderly.sum())
            self.done = True


def test_custom_intervention_writes_device_arrays(cv):
    iv = protect_elderly(8)
    sim = cv.Sim(pop_size=5000, pop_type='hybrid', n_days=40, pop_infected=100, rand_seed=4, verbose=0, beta=0.03, interventions=[iv])
    sim.run()
    assert iv.done and iv.n_protected > 100
    age, d_exp = sim.people.to_numpy('age'), sim.people.to_numpy('date_exposed')
    assert not np.any((age > 70) & (d_exp > 8)), 'nobody over 70 may be infected after the intervention (rel_sus = 0)'
    assert np.any((age <= 70) & (d_exp > 8))


def test_copy_of_a_running_sim(cv):
    ''' sim.copy() in the middle of a run: the copy continues to the same end as the original, and is independent of it '''
    spec = scenarios.SCENARIOS['hybrid3k']
    sim = cv.Sim(**scenarios.build(cv, spec))
    sim.run(until=18)
    twin = sim.copy()
    assert twin.t == 18 and twin.people.exposed.data_ptr() != sim.people.exposed.data_ptr()
    sim.run(reset_seed=False)
    before = twin.people.to_numpy('exposed').copy()
    assert np.array_equal(before, twin.people.to_numpy('exposed'))   # the original's run did not touch the copy
    twin.run(reset_seed=False)
    for k in sim.result_keys():
        assert np.array_equal(sim.results[k].values, twin.results[k].values, equal_nan=True), k
    for k in ('exposed', 'date_exposed', 'nab', 'quarantined', 'doses'):
        x, y = sim.people.to_numpy(k), twin.people.to_numpy(k)
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == 'f')), k
    a, b = sim.infection_log, twin.infection_log
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_vaccinate_num_reference_examples(cv):
    '''
    The reference's own uses of vaccinate_num, unchanged: examples/t08_boosters.py:15-45 (doses per day and the people to boost
    as functions of the sim, booster=True), tests/test_immunity.py:170-181 (fluctuating doses + subtarget, doses recorded by an
    analyzer) and tests/test_immunity.py:214-232 (two people, two vaccines, 1000 days, no infections).
    '''
    def num_doses(sim):
        return sim.t * 10 if sim.t < 50 else 500

    def num_boosters(sim):
        return 0 if sim.t < 50 else 50
    pfizer = cv.vaccinate_num(vaccine='pfizer', sequence='age', num_doses=num_doses)
    booster_target = {'inds': lambda sim: cv.true(sim.people.doses != 2), 'vals': 0}
    booster = cv.vaccinate_num(vaccine='pfizer', sequence='age', subtarget=booster_target, booster=True, num_doses=num_boosters)
    n_doses = []
    sim = cv.Sim(beta=0.015, n_days=90, interventions=[pfizer, booster], analyzers=lambda sim: n_doses.append(sim.people.doses.copy()), verbose=0)
    sim.run()
    doses = np.array([np.asarray(d) for d in n_doses])
    per_day = np.diff(doses.sum(axis=1), prepend=0)
    assert np.array_equal(per_day, sim.results['new_doses'].values)
    assert per_day[:50].max() <= 490 and np.all(per_day[50:] <= 550)               # never more than the day's doses + boosters
    assert np.array_equal(per_day[1:20], 10 * np.arange(1, 20))                       # first weeks: every dose is a first dose
    assert doses[-1].max() >= 3 and (doses[-1] >= 3).sum() > 100                      # boosters reached people with two doses
    first = np.argmax(doses > 0, axis=0)
    got = doses[-1] > 0
    age = sim.people.to_numpy('age')
    assert np.corrcoef(age[got], first[got])[0, 1] < -0.8                             # oldest first

    n_days = 60
    nd = {i: (i ** 2) * (i % 2 == 0) for i in np.arange(n_days)}
    sub = dict(inds=np.arange(10000), vals=0.1)
    seq = cv.vaccinate_num(vaccine='pfizer', sequence='age', num_doses=nd, subtarget=sub)
    sim2 = cv.Sim(pop_size=20000, n_days=n_days, rescale=False, use_waning=True, variants=cv.variant('beta', days=20, n_imports=20), interventions=seq, verbose=0)
    sim2.run()
    assert sim2.results['new_doses'].values[1::2].sum() == 0 and sim2.summary['cum_doses'] > 10000
    assert sim2.results['new_doses'].values[10] == 100 and sim2.results['cum_vaccinated'][-1] <= 20000

    vac1 = cv.vaccinate_num(vaccine='pfizer', sequence=[0], num_doses=1)
    vac2 = cv.vaccinate_num(vaccine='jj', sequence=[1], num_doses=1)
    sim3 = cv.Sim(n_days=200, pop_size=4, pop_infected=0, variants=cv.variant('beta', days=20, n_imports=0), interventions=[vac1, vac2], verbose=0)
    sim3.run()
    assert list(sim3.people.to_numpy('doses')) == [2, 1, 0, 0] and sim3.summary['cum_infections'] == 0
    assert np.all(sim3.people.to_numpy('nab')[:2] > 0) and np.all(sim3.people.to_numpy('nab')[2:] == 0)
