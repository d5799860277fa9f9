'''
Device-side population generation on the GPU: the library's keyed uniforms against the oracle's
NumPy Philox, the generated population against the oracle's plain-loop restatement (bit for bit), a simulation on a
device-generated population in lockstep with the oracle, and the partitioned adjacency built from streamed chunks.
'''
import numpy as np
import pytest
import torch

import parity
import scenarios
from oracle import cvoracle as cvo, philox as ph

pytestmark = pytest.mark.gpu


def test_keyed_uniform_kernel_matches_numpy_philox():
    from covasim_b200 import population as cvpop
    seed = 0x1234_5678_9ABC
    u = cvpop.device_uniforms(seed, torch.device('cuda', 0))
    for sub, i0, n, slot in [(0, 0, 1000, 0), (4, 2 ** 32 - 5, 64, 1), (2, 123_456_789_012, 333, 2), (1, 7, 0, 0)]:
        got = u(sub, i0, n, slot).cpu().numpy()
        want = ph.keyed_uniform(seed, 10, sub, 0, np.arange(i0, i0 + n, dtype=np.int64), slot)
        assert np.array_equal(got, want)


@pytest.mark.parametrize('pop_type,n,seed', [('hybrid', 6000, 3), ('random', 5000, 9)])
def test_device_population_matches_oracle(pop_type, n, seed):
    from covasim_b200 import population as cvpop, parameters as cvpar
    pars = cvpar.make_pars(pop_size=n, pop_type=pop_type)
    dev = torch.device('cuda', 0)
    pop = cvpop.KeyedPop(pars, seed, dev, cvpop.device_uniforms(seed, dev)).materialize()
    ref = cvo.make_keyed_pop(pars, seed)
    assert np.array_equal(pop['age'].cpu().numpy(), ref['age']) and np.array_equal(pop['sex'].cpu().numpy(), ref['sex'])
    for lk, want in ref['contacts'].items():
        assert np.array_equal(pop['contacts'][lk]['p1'].cpu().numpy(), want['p1']), lk
        assert np.array_equal(pop['contacts'][lk]['p2'].cpu().numpy(), want['p2']), lk


def test_sim_on_device_population_in_lockstep_with_oracle():
    import covasim_b200 as cv
    spec = scenarios.SCENARIOS['hybrid3k']
    sim, orc = parity.build_pair(cv, spec=spec, sim_kwargs=dict(pop_gen='device'))
    # the population the sim generated on the device is the oracle's restatement of the same keyed draws
    ref = cvo.make_keyed_pop(sim.pars, sim.pars['rand_seed'])
    for lk, want in ref['contacts'].items():
        assert np.array_equal(sim.people.contacts[lk]['p1'].cpu().numpy(), want['p1'])
    parity.run_lockstep(sim, orc)
    assert sim.summary['cum_infections'] > 300


@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_device_population_equals_single(world):
    import covasim_b200 as cv
    from covasim_b200 import partition as cvpart
    spec = scenarios.SCENARIOS['hybrid3k']
    ref = cv.Sim(**scenarios.build(cv, spec), pop_gen='device')
    ref.run()
    comms = cvpart.LocalComm.make(world)
    sims = [cv.Sim(**scenarios.build(cv, spec), partition=comms[r], pop_gen='device') for r in range(world)]
    cvpart.run_local(sims, lambda s: s.initialize())
    cvpart.run_local(sims, lambda s: s.run())
    for k in ref.people.keys():
        whole = ref.people.to_numpy(k)
        parts = np.concatenate([s.people.to_numpy(k) for s in sims], axis=-1)
        assert np.array_equal(whole, parts, equal_nan=(whole.dtype.kind == 'f')), k
    for k in ('new_infections', 'new_diagnoses', 'new_quarantined', 'n_exposed', 'cum_deaths'):
        assert np.array_equal(sims[0].results[k].values, ref.results[k].values), k
    assert ref.summary['cum_infections'] > 300
