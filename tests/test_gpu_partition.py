'''
Agent-partitioned simulation (BASELINE config 4) against the single-GPU run of the
SAME simulation: every random draw is keyed on global agent / edge ids, so the People arrays (concatenated over
ranks), every result series and the infection log must be IDENTICAL for any number of ranks -- and the single-GPU
run is itself checked bit for bit against the oracle in test_gpu_sim.py.

The ranks run as threads of this process on one GPU (partition.LocalComm: the same kernels and the same exchange
points as under torchrun, with device copies instead of ncclAllGather), so the test runs on the one-GPU box;
tests/multi_gpu_partition.py runs the same comparison under torchrun with NCCL.
'''
import numpy as np
import pytest
import torch

import scenarios

pytestmark = pytest.mark.gpu

PART_SCENARIOS = {
    # testing + tracing with delays + vaccination + per-layer beta changes (hybrid3k of scenarios.py)
    'hybrid3k': scenarios.SCENARIOS['hybrid3k'],
    # three variants, importations, vaccine + booster (variants4k without the bed limits, which need a daily global count)
    'variants4k': dict(pars={k: v for k, v in scenarios.SCENARIOS['variants4k']['pars'].items() if not k.startswith('n_beds')},
                       variants=scenarios.SCENARIOS['variants4k']['variants'], interventions=scenarios.SCENARIOS['variants4k']['interventions']),
    # ... and with them: the day's severe / critical counts are summed over the ranks before anybody is infected
    'variants4k_beds': scenarios.SCENARIOS['variants4k'],
    'random2k_nowaning': scenarios.SCENARIOS['random2k_nowaning'],
    # dynamic rescaling: the count of non-naive agents is summed over the ranks, the agents made naive again are picked by global position
    'rescale3k': scenarios.SCENARIOS['rescale3k'],
    # number-based testing: the n smallest keys of the whole population, found from every rank's own n smallest; with tracing, and with rescaling
    'testnum3k': scenarios.SCENARIOS['testnum3k'],
    'testnum_rescale2k': scenarios.SCENARIOS['testnum_rescale2k'],
    # doses per day along a population-wide priority sequence (oldest first; a random order for the booster), second doses deferred and drawn
    'vaccnum3k': scenarios.SCENARIOS['vaccnum3k'],
    'testnum_sub3k': scenarios.SCENARIOS['testnum_sub3k'],
    # every candidate on ONE rank (a priority list of the first 600 agents), fewer doses than candidates: that rank's own k smallest decide
    'vaccnum_front': dict(pars=dict(pop_size=3000, pop_infected=50, pop_type='hybrid', n_days=30, verbose=0, rand_seed=5, beta=0.022),
                          interventions=[('vaccinate_num', dict(vaccine='pfizer', sequence=np.arange(600), num_doses=35))]),
    # population size not divisible by 32 * world
    'odd5003': dict(pars=dict(pop_size=5003, pop_infected=80, pop_type='hybrid', n_days=35, verbose=0, rand_seed=4, beta=0.02),
                    interventions=[('test_prob', dict(start_day=4, symp_prob=0.3, asymp_prob=0.02)),
                                   ('contact_tracing', dict(trace_probs=0.5, trace_time=1, start_day=6))]),
}


def run_partitioned(cv, spec, world, fused=True):
    from covasim_b200 import partition as cvpart
    comms = cvpart.LocalComm.make(world)
    sims = [cv.Sim(**scenarios.build(cv, spec), partition=comms[r], fused=fused) for r in range(world)]
    cvpart.run_local(sims, lambda s: s.initialize())
    cvpart.run_local(sims, lambda s: s.run())
    logs = cvpart.run_local(sims, lambda s: s.infection_log)
    return sims, logs


# world 1: every row is whole (~36 entries per agent), which selects the 32-lanes-per-transmitter form of edge_pass_partition_kernel;
# 2-4 ranks select the 16- and 8-lane forms
@pytest.mark.parametrize('name,world', [('hybrid3k', 1), ('variants4k', 1), ('hybrid3k', 2), ('hybrid3k', 3), ('variants4k', 2), ('variants4k_beds', 3), ('random2k_nowaning', 4), ('odd5003', 3), ('rescale3k', 2), ('rescale3k', 3), ('testnum3k', 3), ('testnum_rescale2k', 2), ('vaccnum3k', 2), ('vaccnum3k', 3), ('testnum_sub3k', 2), ('vaccnum_front', 3)])
@pytest.mark.parametrize('fused', [True, False])
def test_partitioned_equals_single(name, world, fused):
    ''' fused=True: the days without a host decision go through the fused kernels phase by phase (cvb_fused_phase) with the exchanges in between '''
    import covasim_b200 as cv
    spec = PART_SCENARIOS[name]
    ref = cv.Sim(**scenarios.build(cv, spec))
    ref.run()
    sims, logs = run_partitioned(cv, spec, world, fused)
    if fused and not spec['pars'].get('n_imports') and name not in ('rescale3k', 'testnum3k', 'testnum_rescale2k', 'vaccnum3k', 'testnum_sub3k', 'vaccnum_front'):        # (days on which the population may still be rescaled need the host, too)                      # (daily importations are drawn on the host: those runs keep the per-step path)
        # (5003 agents over 3 ranks: the last rank's share is not a multiple of 4, so ALL ranks keep the per-step path)
        assert all(s.fused_days > 0.5 * s.npts for s in sims) or (name == 'odd5003' and all(s.fused_days == 0 for s in sims)), [s.fused_days for s in sims]
    if not fused:
        assert all(s.fused_days == 0 for s in sims)
    # ranges tile the population
    assert [s.id0 for s in sims] == list(np.cumsum([0] + [s.n_local for s in sims[:-1]]))
    assert sum(s.n_local for s in sims) == ref.n
    # every People array, concatenated over ranks, is identical (floats included: same arithmetic, same keys)
    for k in ref.people.keys():
        whole = ref.people.to_numpy(k)
        parts = np.concatenate([s.people.to_numpy(k) for s in sims], axis=-1)
        assert np.array_equal(whole, parts, equal_nan=(whole.dtype.kind == 'f')), f'{name} x{world}: People.{k} differs'
    # results: counters are integers (exact); the three float64 population sums differ only by summation order
    for s in sims:
        for k in ref.result_keys():
            a, b = s.results[k].values, ref.results[k].values
            # r_eff: the single-GPU run takes float32 means of the date arrays like the reference (sim.py:915-923), the
            # partitioned run combines per-rank float64 sums
            assert np.allclose(a, b, rtol=1e-6 if k == 'r_eff' else 1e-12, atol=0, equal_nan=True), f'{name} x{world}: result {k} differs'
        for k in ref.result_keys('variant'):
            assert np.array_equal(s.results['variant'][k].values, ref.results['variant'][k].values), f'{name} x{world}: variant/{k} differs'
        assert s.summary['cum_infections'] == ref.summary['cum_infections']
    # infection log: same transmissions, same sources, same layers
    want = ref.infection_log
    for log in logs:
        for k in ('source', 'target', 'date', 'layer', 'variant'):
            assert np.array_equal(log[k], want[k]), f'{name} x{world}: infection log {k} differs'
    assert ref.summary['cum_infections'] > 4 * spec['pars']['pop_infected']         # the epidemic actually spread


def test_partitioned_rejects_what_it_cannot_do():
    import covasim_b200 as cv
    from covasim_b200 import partition as cvpart
    comm = cvpart.LocalComm.make(1)[0]
    with pytest.raises(NotImplementedError):
        cv.Sim(pop_size=2000, n_days=5, dynam_layer=dict(a=1), partition=comm).initialize()
    with pytest.raises(NotImplementedError):
        cv.Sim(pop_size=2000, n_days=5, rng='mt', partition=comm)
    with pytest.raises(ValueError):
        cvpart.plan(40, 4)                      # chunks of 32: the last ranks would own nobody


def test_hit_capacity_overflow_is_reported():
    ''' A day with more successful transmissions than hit_capacity must fail loudly, never drop infections silently '''
    import covasim_b200 as cv
    from covasim_b200 import partition as cvpart
    comms = cvpart.LocalComm.make(2)
    # very infectious and a deliberately tiny capacity: more successful transmissions per day than the 1024 slots
    sims = [cv.Sim(pop_size=20_000, pop_infected=4_000, n_days=10, beta=0.5, verbose=0, partition=comms[r], hit_capacity=1024) for r in range(2)]
    cvpart.run_local(sims, lambda s: s.initialize())
    with pytest.raises(RuntimeError, match='hit_capacity'):
        cvpart.run_local(sims, lambda s: s.run())


@pytest.mark.parametrize('n_bytes,offset', [(4096, 0), (4096, 8192), (1000, 4), (52, 20)])
def test_peer_push_places_the_chunk_in_every_buffer(n_bytes, offset):
    '''
    cvb_peer_push (the data half of the exchange over peer memory, partition.PeerExchange): the chunk lands at the given offset of EVERY
    listed buffer and nowhere else, through the 16-byte path and the 4-byte path (sizes / offsets that are not multiples of 16).  Here the
    "peers" are three buffers of one GPU; under torchrun they are the ranks' symmetric-memory mappings (tests/multi_gpu_partition.py).
    '''
    import ctypes as C
    import torch
    from covasim_b200 import _capi
    dev = torch.device('cuda', 0)
    src = torch.randint(0, 255, (n_bytes,), dtype=torch.uint8, device=dev)
    bufs = [torch.full((3 * 8192,), 7, dtype=torch.uint8, device=dev) for _ in range(3)]
    ptrs = (C.c_uint64 * 3)(*[b.data_ptr() for b in bufs])
    _capi.call('cvb_peer_push', src.data_ptr(), n_bytes, ptrs, 3, offset, None)
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b[offset:offset + n_bytes], src)
        assert bool((b[:offset] == 7).all()) and bool((b[offset + n_bytes:] == 7).all())
    with pytest.raises(_capi.CvbError):
        _capi.call('cvb_peer_push', src.data_ptr(), 6, ptrs, 3, 0, None)             # sizes are multiples of 4 bytes
