'''
Parity at the sizes the benchmark runs at (BASELINE config 2: 1M agents, ~17.8M edges): the paths small scenarios never reach --
64-bit adjacency offsets over 35M entries, transmitter lists of tens of thousands of agents, the dense streaming pass with a transmit
bitmap larger than shared memory -- checked against the oracle (Philox mode) and against each other.
'''
import numpy as np
import pytest

import parity
import scenarios

pytestmark = pytest.mark.gpu

# the C2 recipe with the epidemic and the interventions moved to the first days, so that eight days exercise testing, tracing,
# quarantine, isolation and ~10^4 transmissions per day
C2_EARLY = dict(
    pars=dict(pop_size=1_000_000, pop_type='hybrid', n_days=8, pop_infected=50_000, rand_seed=1, verbose=0),
    interventions=[('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=2)), ('contact_tracing', dict(trace_probs=0.3, start_day=3))])


@pytest.mark.parametrize('fused', [True, False])
def test_c2_shape_lockstep_against_oracle(fused):
    ''' 1M agents, every People array compared with the oracle after every day (fused day kernels / per-step entry points) '''
    import covasim_b200 as cv
    sim, orc = parity.build_pair(cv, spec=C2_EARLY, sim_kwargs=dict(fused=fused, pop_exact=False))
    orc.keep_log = True
    parity.run_lockstep(sim, orc)
    assert sim.summary['cum_infections'] > 100_000
    assert sim.summary['cum_quarantined'] > 1_000 and sim.summary['cum_diagnoses'] > 1_000
    assert (sim.fused_days == sim.npts) == fused


def test_dense_pass_large_population_equals_adjacency_form():
    '''
    2M agents with use_adjacency=False: the transmit bitmap (250 KB) no longer fits shared memory, which selects the global-bitmap
    shape of the dense streaming pass (edge_pass.cu: CVB_DENSE(256, 1, false, 4, 1)); the run must equal the adjacency form.
    '''
    import covasim_b200 as cv
    spec = dict(pars=dict(pop_size=2_000_000, pop_type='hybrid', n_days=6, pop_infected=80_000, rand_seed=3, verbose=0),
                interventions=[('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=1)), ('contact_tracing', dict(trace_probs=0.3, start_day=2))])
    runs = []
    for use_adj in (True, False):
        sim = cv.Sim(**scenarios.build(cv, spec), use_adjacency=use_adj, pop_exact=False, pop_gen='device')
        sim.run()
        runs.append(sim)
    a, b = runs
    assert a._adj is not None and b._adj is None
    for k in a.people.keys():
        x, y = a.people.to_numpy(k), b.people.to_numpy(k)
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == 'f')), k
    for k in a.result_keys():
        assert np.array_equal(a.results[k].values, b.results[k].values, equal_nan=True), k
    la, lb = a.infection_log, b.infection_log
    for k in la:
        assert np.array_equal(la[k], lb[k]), k
    assert a.summary['cum_infections'] > 100_000


def test_dynamic_layer_large_population_fused_equals_per_step():
    ''' 2M agents, one dynamic layer (regenerated on the device every day, streamed densely with the bitmap in global memory): the
    fused day kernels and the per-step entry points give the same simulation '''
    import covasim_b200 as cv
    runs = []
    for fused in (True, False):
        sim = cv.Sim(pop_size=2_000_000, pop_type='random', n_days=6, pop_infected=60_000, rand_seed=5, verbose=0, dynam_layer=dict(a=1),
                     interventions=[cv.test_prob(symp_prob=0.1, asymp_prob=0.01, start_day=1)], pop_gen='device', fused=fused)
        sim.run()
        runs.append(sim)
    a, b = runs
    assert a.fused_days == a.npts and b.fused_days == 0
    for k in a.people.keys():
        x, y = a.people.to_numpy(k), b.people.to_numpy(k)
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == 'f')), k
    for k in a.result_keys():
        assert np.array_equal(a.results[k].values, b.results[k].values, equal_nan=True), k
    assert a.summary['cum_infections'] > 100_000
