'''
Whole-path GPU parity: covasim_b200 (native-RNG mode, fused edge pass) against the oracle in Philox mode,
same scenario, same population, stepped in lockstep with every People array compared every day, then
results and the infection log.  Bit-exact for flags / dates / counters / infection log; 1e-6 relative for
NAb and immunity floats.  Plus the reference's RNG-free known-answer tests and its state-diagram invariants.
'''
import numpy as np
import pytest

import parity
import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


@pytest.mark.parametrize('name', ['random2k_nowaning', 'dynamic2k', 'hybrid3k', 'variants4k', 'dynpars3k', 'clip3k', 'rescale3k', 'fracsus2k', 'sequence3k', 'testnum3k', 'testnum_rescale2k', 'subtarget3k', 'ili3k', 'capacity3k', 'vaccnum3k', 'testnum_sub3k', 'targeteff3k', 'swab3k', 'quarpol3k'])
@pytest.mark.parametrize('fused', [True, False])
def test_lockstep_parity(cv, name, fused):
    ''' fused=True: days without a host decision go through cvb_run_days (day_fused.cu); False: the per-step entry points every day '''
    sim, orc = parity.build_pair(cv, name, sim_kwargs=dict(fused=fused))
    parity.run_lockstep(sim, orc)
    assert sim.summary['cum_infections'] > scenarios.SCENARIOS[name]['pars']['pop_infected']     # the epidemic actually ran
    if not fused:
        assert sim.fused_days == 0


ODD_SPEC = dict(
    # population size not a multiple of 4 (scalar tail path of the 4-agents-per-thread kernels), 2 variants, all interventions
    pars=dict(pop_size=2501, pop_infected=45, pop_type='hybrid', n_days=35, verbose=0, rand_seed=21, beta=0.025, n_imports=0.8),
    variants=[dict(variant='delta', days=6, n_imports=15)],
    interventions=[('test_prob', dict(start_day=3, symp_prob=0.3, asymp_prob=0.02, quar_policy='both', test_delay=1)),
                   ('contact_tracing', dict(trace_probs=dict(h=0.9, s=0.4, w=0.4, c=0.1), trace_time=dict(h=0, s=1, w=1, c=3), start_day=5)),
                   ('vaccinate_prob', dict(vaccine='moderna', days=7, prob=0.2))])


def test_dense_and_adjacency_paths_agree(cv):
    ''' use_adjacency=False streams every layer densely (the reference's access pattern); both must match the oracle '''
    spec = scenarios.SCENARIOS['hybrid3k']
    sim = cv.Sim(**scenarios.build(cv, spec), use_adjacency=False)
    sim.initialize()
    import numpy as np
    from oracle import cvoracle as cvo
    pop = dict(age=sim.people.to_numpy('age').astype(np.float64), sex=sim.people.to_numpy('sex'),
               contacts={lk: l.to_numpy() for lk, l in sim.people.contacts.items()})
    orc = cvo.OracleSim(**scenarios.build(cvo, spec), rng='philox', popdict=pop)
    orc.initialize()
    parity.run_lockstep(sim, orc, every=5)


class clip_layer:
    ''' Test-only intervention: on `day`, drop the first `frac` of a layer's edges (what cv.clip_edges does, reference
    interventions.py:589-666 + base.py:1742-1757 Layer.pop_inds); works on both the device sim and the oracle '''
    def __init__(self, day, lkey, frac):
        self.day, self.lkey, self.frac = day, lkey, frac
        self.initialized = False

    def initialize(self, sim):
        self.initialized = True

    def __call__(self, sim):
        if sim.t != self.day:
            return
        import numpy as np
        if hasattr(sim, 'P'):                                   # oracle
            layer = sim.contacts[self.lkey]
            k = int(len(layer['p1']) * self.frac)
            for c in ('p1', 'p2', 'beta'):
                layer[c] = layer[c][k:].copy()
        else:
            layer = sim.people.contacts[self.lkey]
            k = int(len(layer) * self.frac)
            layer.pop_inds(np.arange(k))


def test_layer_edit_rebuilds_adjacency(cv):
    import numpy as np
    from oracle import cvoracle as cvo
    pars = dict(pop_size=3000, pop_infected=60, pop_type='hybrid', n_days=30, verbose=0, rand_seed=13, beta=0.03)
    sim = cv.Sim(pars, interventions=[clip_layer(8, 'c', 0.5), cv.test_prob(symp_prob=0.3, start_day=3), cv.contact_tracing(trace_probs=0.5, start_day=4)])
    sim.initialize()
    pop = dict(age=sim.people.to_numpy('age').astype(np.float64), sex=sim.people.to_numpy('sex'),
               contacts={lk: l.to_numpy() for lk, l in sim.people.contacts.items()})
    orc = cvo.OracleSim(pars, interventions=[clip_layer(8, 'c', 0.5), cvo.test_prob(symp_prob=0.3, start_day=3), cvo.contact_tracing(trace_probs=0.5, start_day=4)],
                        rng='philox', popdict=pop)
    orc.initialize()
    parity.run_lockstep(sim, orc)
    assert len(sim.people.contacts['c']) == len(orc.contacts['c']['p1']) < len(pop['contacts']['c']['p1'])


def test_lockstep_parity_odd_population(cv):
    sim, orc = parity.build_pair(cv, spec=ODD_SPEC)
    parity.run_lockstep(sim, orc)


@pytest.mark.parametrize('name', ['default20k', 'baseline20k'])
def test_endpoint_parity_20k(cv, name):
    ''' The two 20k-agent configurations (BASELINE.json config 1 and the reference's own baseline test sim) '''
    sim, orc = parity.build_pair(cv, name)
    sim.run()
    orc.run()
    parity.compare_people(sim, orc, 'at the end')
    parity.compare_results(sim, orc)
    parity.compare_log(sim, orc)


def test_step_by_step_equals_run(cv):
    ''' reference tests/test_resume.py:120-138: a step() loop and run(until=..) + run() are the same simulation '''
    spec = scenarios.SCENARIOS['hybrid3k']
    a = cv.Sim(**scenarios.build(cv, spec)).run()
    b = cv.Sim(**scenarios.build(cv, spec))
    b.run(until=20)
    b.run()
    c = cv.Sim(**scenarios.build(cv, spec))
    c.initialize()
    while not c.complete:
        c.step()
    c.finalize()
    for k in a.result_keys():
        assert np.array_equal(a.results[k].values, b.results[k].values, equal_nan=True), k
        assert np.array_equal(a.results[k].values, c.results[k].values, equal_nan=True), k


def test_no_transmission_when_beta_is_zero(cv):
    ''' reference tests/unittests/test_transmission.py:16-35 '''
    sim = cv.Sim(pop_size=5000, pop_infected=50, n_days=30, beta=0.0, rand_seed=4).run()
    assert sim.results['new_infections'].values.sum() == 0
    assert sim.summary['cum_infections'] == 50


def test_everyone_dies(cv):
    ''' reference tests/unittests/test_mortality.py:13-24 '''
    big = 1e6
    sim = cv.Sim(pop_size=500, pop_infected=500, n_days=120, rand_seed=4, use_waning=False, rel_symp_prob=big, rel_severe_prob=big,
                 rel_crit_prob=big, rel_death_prob=big).run()
    assert sim.summary['cum_deaths'] == 500


def test_exact_progression_days(cv):
    ''' reference tests/unittests/test_progression.py:17-108: zero-variance durations put every transition on the exact day '''
    dur = {k: dict(dist='normal_int', par1=v, par2=0.0) for k, v in dict(exp2inf=3, inf2sym=2, sym2sev=4, sev2crit=3, asym2rec=7, mild2rec=7,
                                                                          sev2rec=9, crit2rec=9, crit2die=5).items()}
    big = 1e6
    sim = cv.Sim(pop_size=300, pop_infected=300, n_days=40, rand_seed=1, use_waning=False, dur=dur, rel_symp_prob=big, rel_severe_prob=big,
                 rel_crit_prob=big, rel_death_prob=big).run()
    r = sim.results
    assert r['new_infectious'].values[3] == 300 and r['new_infectious'].values.sum() == 300
    assert r['new_symptomatic'].values[5] == 300
    assert r['new_severe'].values[9] == 300
    assert r['new_critical'].values[12] == 300
    assert r['new_deaths'].values[17] == 300


STATE_MATRIX = '''
susceptible   1  0 -1 -1 -1 -1 -1  0  0 -1 -1  0  0  0
naive         1  1 -1 -1 -1 -1 -1  0 -1 -1 -1  0  0  0
exposed      -1 -1  1  0  0  0  0  0  0 -1 -1  0  0  0
infectious   -1 -1  1  1  0  0  0  0  0 -1 -1  0  0  0
symptomatic  -1 -1  1  1  1  0  0  0  0 -1 -1  0  0  0
severe       -1 -1  1  1  1  1  0  0  0 -1 -1  0  0  0
critical     -1 -1  1  1  1  1  1  0  0 -1 -1  0  0  0
tested        0  0  0  0  0  0  0  1  0  0  0  0  0  0
diagnosed     0  0  0  0  0  0  0  1  1  0  0  0  0  0
recovered    -1 -1 -1 -1 -1 -1 -1  0  0  1 -1  0  0  0
dead         -1 -1 -1 -1 -1 -1 -1  0  0 -1  1 -1 -1  0
known_contact 0  0  0  0  0  0  0  0  0  0 -1  1  0  0
quarantined   0  0  0  0  0  0  0  0  0  0 -1  1  1  0
vaccinated    0  0  0  0  0  0  0  0  0  0  0  0  0  1
'''


@pytest.mark.parametrize('use_waning', [False, True])
def test_state_diagram(cv, use_waning):
    ''' reference tests/test_immunity.py:25-85 + tests/state_diagram.xlsx (transcribed in SURVEY.md Appendix B) '''
    rows = [l.split() for l in STATE_MATRIX.strip().splitlines()]
    names = [r[0] for r in rows]
    M = np.array([[int(x) for x in r[1:]] for r in rows])
    if use_waning:
        M[names.index('susceptible'), names.index('recovered')] = 0
        M[names.index('recovered'), names.index('susceptible')] = 1
    sim = cv.Sim(pop_size=2000, pop_infected=40, n_days=70, rand_seed=1, use_waning=use_waning, beta=0.03, pop_type='hybrid',
                 interventions=[cv.test_prob(symp_prob=0.4, asymp_prob=0.02, start_day=5),
                                cv.contact_tracing(trace_probs=0.5, start_day=8),
                                cv.vaccinate_prob('pfizer', days=10, prob=0.3)] if use_waning else
                 [cv.test_prob(symp_prob=0.4, asymp_prob=0.02, start_day=5), cv.contact_tracing(trace_probs=0.5, start_day=8)])
    sim.run()
    P = {k: sim.people.to_numpy(k) for k in names}
    for i, s1 in enumerate(names):
        on = P[s1]
        assert on.any() or s1 in ('critical', 'vaccinated'), f'nobody is {s1}: the check would be vacuous'
        for j, s2 in enumerate(names):
            if M[i, j] == 1:
                assert P[s2][on].all(), f'{s1} must imply {s2}'
            elif M[i, j] == -1:
                assert not P[s2][on].any(), f'{s1} must exclude {s2}'


def test_multisim_members_equal_individual_runs(cv):
    ''' reference tests/test_run.py: MultiSim members are the base sim with rand_seed + i (run.py:1363-1365) '''
    pars = dict(pop_size=3000, pop_infected=40, pop_type='hybrid', n_days=25, rand_seed=5, beta=0.025)
    make = lambda seed: cv.Sim(dict(pars, rand_seed=seed), interventions=[cv.test_prob(symp_prob=0.2, start_day=5)])
    msim = cv.MultiSim(make(5), n_runs=4)
    msim.run()
    for i in range(4):
        solo = make(5 + i).run()
        for k in ('new_infections', 'cum_infections', 'n_exposed', 'new_diagnoses', 'cum_deaths'):
            assert np.array_equal(msim.member_results[i][k], solo.results[k].values), (i, k)
    msim.reduce()
    lo, mid, hi = msim.results['cum_infections'].low, msim.results['cum_infections'].values, msim.results['cum_infections'].high
    assert np.all(lo <= mid) and np.all(mid <= hi)
    finals = msim.summarize('cum_infections')
    assert len(set(finals.tolist())) > 1                      # different seeds, different epidemics
