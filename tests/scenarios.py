'''
Scenario specifications shared by the golden-vector generator (oracle/gen_golden.py, which runs the
unmodified reference), the oracle tests and the GPU parity tests.  A scenario is plain data:
``pars`` (simulation parameters), ``interventions`` (list of (name, kwargs)), ``variants`` (list of
kwargs); ``build(module, spec)`` instantiates it against any module exposing the reference's
constructor names (the reference itself, oracle.cvoracle, or covasim_b200).
'''
import copy

import numpy as np

SCENARIOS = {
    # reference tests/test_baselines.py:18-56 -- the sim behind tests/baseline.json
    'baseline20k': dict(
        pars=dict(use_waning=True, pop_size=20000, pop_infected=100, pop_type='hybrid', n_days=60, verbose=0, rand_seed=2),
        interventions=[('change_beta', dict(days=40, changes=0.5)),
                       ('test_prob', dict(start_day=20, symp_prob=0.1, asymp_prob=0.01)),
                       ('contact_tracing', dict(trace_probs=0.3, start_day=50)),
                       ('vaccinate_prob', dict(vaccine='pfizer', days=30, prob=0.1))],
    ),
    # BASELINE.json config 1: cv.Sim() defaults
    'default20k': dict(pars=dict(verbose=0), interventions=[]),
    # small hybrid with everything on, early start so that tracing/quarantine/vaccination all fire
    'hybrid3k': dict(
        pars=dict(pop_size=3000, pop_infected=60, pop_type='hybrid', n_days=45, verbose=0, rand_seed=11, beta=0.02),
        interventions=[('test_prob', dict(start_day=5, symp_prob=0.3, asymp_prob=0.02, symp_quar_prob=0.6, asymp_quar_prob=0.1, test_delay=1, sensitivity=0.9, loss_prob=0.1)),
                       ('contact_tracing', dict(trace_probs=dict(h=1.0, s=0.5, w=0.5, c=0.2), trace_time=dict(h=0, s=1, w=1, c=2), start_day=8)),
                       ('vaccinate_prob', dict(vaccine='pfizer', days=[10, 12], prob=0.15)),
                       ('change_beta', dict(days=[15, 30], changes=[0.6, 0.9], layers='c'))],
    ),
    # random population, no waning, no interventions
    'random2k_nowaning': dict(pars=dict(pop_size=2000, pop_infected=30, n_days=40, verbose=0, rand_seed=5, use_waning=False, beta=0.025), interventions=[]),
    # two extra variants, imports, bed limits, vaccination + booster
    'variants4k': dict(
        pars=dict(pop_size=4000, pop_infected=50, pop_type='hybrid', n_days=50, verbose=0, rand_seed=3, beta=0.02,
                  n_imports=1.5, n_beds_hosp=5, n_beds_icu=1),
        variants=[dict(variant='alpha', days=8, n_imports=20), dict(variant='delta', days=15, n_imports=25)],
        interventions=[('test_prob', dict(start_day=3, symp_prob=0.2, asymp_prob=0.01)),
                       ('contact_tracing', dict(trace_probs=0.4, start_day=6)),
                       ('vaccinate_prob', dict(vaccine='pfizer', days=5, prob=0.3)),
                       ('vaccinate_prob', dict(vaccine='jj', days=32, prob=0.2, booster=True, label='jj_boost'))],
    ),
    # parameters edited during the run (dynamic_pars): transmissibility down and up again, deadlier disease, daily importations on and off
    'dynpars3k': dict(
        pars=dict(pop_size=3000, pop_infected=40, pop_type='hybrid', n_days=40, verbose=0, rand_seed=21, beta=0.02),
        interventions=[('dynamic_pars', dict(pars={'beta': dict(days=[10, 25], vals=[0.008, 0.025]), 'rel_death_prob': dict(days=15, vals=3.0),
                                                   'n_imports': dict(days=[5, 30], vals=[3, 0])})),
                       ('test_prob', dict(start_day=5, symp_prob=0.2, asymp_prob=0.01))],
    ),
    # contacts removed and restored during the run (clip_edges), with tracing over the clipped network
    'clip3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=40, verbose=0, rand_seed=31, beta=0.022),
        interventions=[('clip_edges', dict(days=[8, 25], changes=[0.3, 1.0], layers=['s', 'w'])),
                       ('clip_edges', dict(days=12, changes=0.5)),
                       ('test_prob', dict(start_day=4, symp_prob=0.3, asymp_prob=0.02)),
                       ('contact_tracing', dict(trace_probs=0.5, start_day=6))],
    ),
    # dynamic rescaling: 3000 agents standing for up to 24000 people (reference sim.py:535-555)
    'rescale3k': dict(
        pars=dict(pop_size=3000, pop_scale=8, rescale=True, pop_infected=60, pop_type='hybrid', n_days=40, verbose=0, rand_seed=41, beta=0.025),
        interventions=[('test_prob', dict(start_day=5, symp_prob=0.2, asymp_prob=0.01)),
                       ('vaccinate_prob', dict(vaccine='pfizer', days=12, prob=0.2))],
    ),
    # a third of the population immune from the start (frac_susceptible, reference sim.py:519-521)
    'fracsus2k': dict(pars=dict(pop_size=2000, pop_infected=40, n_days=30, verbose=0, rand_seed=51, beta=0.03, frac_susceptible=0.67), interventions=[]),
    # a sequence of two testing regimes (nested interventions), traced
    'sequence3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=35, verbose=0, rand_seed=61, beta=0.022),
        interventions=[('sequence', dict(days=[5, 20], interventions=[('test_prob', dict(symp_prob=0.4, asymp_prob=0.0)),
                                                                      ('test_prob', dict(symp_prob=0.1, asymp_prob=0.02, test_delay=1))])),
                       ('contact_tracing', dict(trace_probs=0.4, start_day=8))],
    ),
    # number-based testing (test_num) with quarantine testing, traced; and with dynamic rescaling (share of tests inside the sample)
    'testnum3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=35, verbose=0, rand_seed=71, beta=0.022),
        interventions=[('test_num', dict(daily_tests=90, symp_test=50.0, quar_test=3.0, quar_policy='both', start_day=4, sensitivity=0.9, loss_prob=0.05, test_delay=1)),
                       ('contact_tracing', dict(trace_probs=0.5, start_day=6))],
    ),
    'testnum_rescale2k': dict(
        pars=dict(pop_size=2000, pop_scale=6, rescale=True, pop_infected=40, pop_type='hybrid', n_days=30, verbose=0, rand_seed=81, beta=0.025),
        interventions=[('test_num', dict(daily_tests=[200] * 31, symp_test=80.0, start_day=3))],
    ),
    # number-based testing with the two options that re-weight agents: influenza-like illness (3 % of the population every day test like
    # symptomatic people) and a subtarget (every 5th agent three times as likely, a block of agents never), daily quarantine testing
    'testnum_sub3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=30, verbose=0, rand_seed=131, beta=0.022),
        interventions=[('test_num', dict(daily_tests=120, symp_test=20.0, quar_test=2.0, quar_policy='daily', start_day=3, sensitivity=0.95, ili_prev=0.03,
                                         subtarget=dict(inds=np.concatenate([np.arange(0, 3000, 5), np.arange(2001, 2400, 5)]),
                                                        vals=np.concatenate([np.full(600, 3.0), np.zeros(80)])))),
                       ('contact_tracing', dict(trace_probs=0.5, start_day=5))],
    ),
    # a two-dose vaccine given by its target efficacies (reference tests/test_immunity.py:238-290): NAb level and boost are derived from them
    'targeteff3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=40, verbose=0, rand_seed=141, beta=0.022),
        interventions=[('vaccinate_prob', dict(vaccine=dict(nab_init=None, nab_boost=3, doses=2, interval=14, target_eff=[0.7, 0.95]), label='trial', days=[3, 5], prob=0.3)),
                       ('test_prob', dict(start_day=5, symp_prob=0.2, asymp_prob=0.01))],
    ),
    # symptom-onset-to-swab delay (swab_delay): symptomatic people test with a probability / weight that follows the time since onset
    'swab3k': dict(
        pars=dict(pop_size=3000, pop_infected=80, pop_type='hybrid', n_days=30, verbose=0, rand_seed=151, beta=0.025),
        interventions=[('test_prob', dict(start_day=3, end_day=16, symp_prob=0.25, asymp_prob=0.005, symp_quar_prob=0.5, quar_policy='daily',
                                          swab_delay=dict(dist='lognormal', par1=3, par2=4))),
                       ('test_num', dict(daily_tests=100, symp_test=30.0, quar_test=2.0, start_day=17, swab_delay=dict(dist='lognormal', par1=2, par2=3))),
                       ('contact_tracing', dict(trace_probs=0.5, start_day=5))],
    ),
    # quarantine testing on given days after the start of quarantine (quar_policy as a list / a number instead of a keyword)
    'quarpol3k': dict(
        pars=dict(pop_size=3000, pop_infected=80, pop_type='hybrid', n_days=30, verbose=0, rand_seed=161, beta=0.025),
        interventions=[('test_prob', dict(start_day=3, end_day=17, symp_prob=0.3, asymp_prob=0.01, symp_quar_prob=0.9, asymp_quar_prob=0.5, quar_policy=[1, 4])),
                       ('test_num', dict(daily_tests=80, symp_test=30.0, quar_test=40.0, quar_policy=2, start_day=18)),
                       ('contact_tracing', dict(trace_probs=0.6, start_day=4))],
    ),
    # subtargeting: explicit testing / vaccination probabilities for given agents (a scalar for every 4th agent; a ramp over a block)
    'subtarget3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=35, verbose=0, rand_seed=91, beta=0.022),
        interventions=[('test_prob', dict(start_day=3, symp_prob=0.2, asymp_prob=0.01, subtarget=dict(inds=np.arange(0, 3000, 4), vals=0.15))),
                       ('vaccinate_prob', dict(vaccine='pfizer', days=[6, 9], prob=0.02,
                                               subtarget=dict(inds=np.arange(1000, 2500), vals=np.linspace(0.0, 0.6, 1500))))],
    ),
    # influenza-like illness: a random 2 % of the population tests like symptomatic people every day, with quarantine testing
    'ili3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=30, verbose=0, rand_seed=101, beta=0.022),
        interventions=[('test_prob', dict(start_day=3, symp_prob=0.3, asymp_prob=0.005, symp_quar_prob=0.8, asymp_quar_prob=0.2, ili_prev=0.02)),
                       ('contact_tracing', dict(trace_probs=0.5, start_day=5))],
    ),
    # tracing capacity: at most 6 of the day's cases are traced
    'capacity3k': dict(
        pars=dict(pop_size=3000, pop_infected=80, pop_type='hybrid', n_days=30, verbose=0, rand_seed=111, beta=0.03),
        interventions=[('test_prob', dict(start_day=3, symp_prob=0.5, asymp_prob=0.05)),
                       ('contact_tracing', dict(trace_probs=0.6, trace_time=dict(h=0, s=1, w=1, c=2), start_day=4, capacity=6))],
    ),
    # doses per day along a priority sequence (vaccinate_num): oldest first with subtarget weights, days without doses (second doses
    # deferred) and days with fewer doses than second doses due (who waits is drawn); a one-dose booster in random order
    'vaccnum3k': dict(
        pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=45, verbose=0, rand_seed=121, beta=0.022),
        interventions=[('vaccinate_num', dict(vaccine='pfizer', sequence='age', subtarget=dict(inds=np.arange(0, 3000, 3), vals=0.4),
                                              num_doses={2: 80, 3: 120, 4: 0, 5: 150, 6: 90, 8: 60, 23: 30, 24: 0, 25: 100, 26: 40, 27: 200, 28: 35, 30: 400, 33: 90})),
                       ('vaccinate_num', dict(vaccine='jj', booster=True, num_doses=25, label='jj_boost')),
                       ('test_prob', dict(start_day=5, symp_prob=0.2, asymp_prob=0.01))],
    ),
    # dynamic layer (BASELINE.json config 5 member shape, scaled down)
    'dynamic2k': dict(pars=dict(pop_size=2000, pop_infected=40, n_days=30, verbose=0, rand_seed=8, beta=0.02,
                                dynam_layer=dict(a=1)), interventions=[]),
}


def build(mod, spec, **extra):
    ''' Instantiate a scenario against module ``mod`` (reference covasim, oracle.cvoracle or covasim_b200) '''
    spec = copy.deepcopy(spec)
    pars = spec['pars']
    def make(name, kw):
        if name == 'sequence':                         # nested interventions are given as (name, kwargs) pairs too
            kw = dict(kw, interventions=[make(n, k) for n, k in kw['interventions']])
        return getattr(mod, name)(**kw)
    ivs = [make(name, kw) for name, kw in spec.get('interventions', [])]
    vs = [mod.variant(**kw) for kw in spec.get('variants', [])]
    kwargs = dict(pars)
    kwargs['interventions'] = ivs
    if vs:
        kwargs['variants'] = vs
    kwargs.update(extra)
    return kwargs
