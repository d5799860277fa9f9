'''
CPU tests of the ensemble plumbing (covasim_b200/run.py): sharding, packing and the one results gather, with a
world_size-2 gloo process group.  The payload is the oracle (the device sims need a GPU); the plumbing under test is
device independent.
'''
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def member_vector(i):
    ''' One ensemble member: a small oracle sim with seed 10 + i, flattened like MultiSim does '''
    from oracle import cvoracle as cvo
    sim = cvo.OracleSim(pop_size=400, pop_infected=20, n_days=12, rand_seed=10 + i, beta=0.03).run()
    return np.concatenate([np.asarray(sim.results[k], dtype=np.float64) for k in ('new_infections', 'cum_infections', 'n_exposed', 'pop_nabs')])


def _worker(rank, world, port, n_items, q):
    for p in (ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from covasim_b200 import run as cvrun
    mine = cvrun.shard_indices(n_items, rank, world)
    local = {i: member_vector(i) for i in mine}
    full = cvrun.gather_results(local, n_items)
    q.put((rank, mine, np.stack(full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_items', [5, 2])
def test_gloo_gather_matches_serial(n_items):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n_items
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = np.stack([member_vector(i) for i in range(n_items)])
    owned = []
    for rank, mine, full in got:
        assert np.array_equal(full, serial)             # every rank ends with the whole ensemble, in member order
        owned += mine
    assert sorted(owned) == list(range(n_items))        # every member ran exactly once


def test_shard_indices_cover():
    from covasim_b200 import run as cvrun
    for n in (0, 1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            parts = [cvrun.shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_pack_unpack_roundtrip_and_reduce():
    from covasim_b200 import run as cvrun, Sim
    sim = Sim(pop_size=100, n_days=5)
    sim.pars['n_variants'] = 2
    sim._init_results()
    rng = np.random.RandomState(0)
    for k in sim.result_keys():
        sim.results[k].values[:] = rng.random_sample(sim.npts)
    for k in sim.result_keys('variant'):
        sim.results['variant'][k].values[:] = rng.random_sample((2, sim.npts))
    vec, keys, vkeys = cvrun.pack_results(sim.results)
    back = cvrun.unpack_results(vec, keys, vkeys, sim.npts, 2)
    for k in keys:
        assert np.array_equal(back[k], sim.results[k].values)
    for k in vkeys:
        assert np.array_equal(back['variant'][k], sim.results['variant'][k].values)
    ms = cvrun.MultiSim(sim, n_runs=3)
    ms.member_results = [cvrun.unpack_results(vec * (1 + 0.1 * i), keys, vkeys, sim.npts, 2) for i in range(3)]
    ms.reduce()
    assert np.allclose(ms.results['cum_infections'].values, back['cum_infections'] * 1.1)
    assert np.all(ms.results['cum_infections'].low <= ms.results['cum_infections'].values)
    ms.combine()
    assert np.allclose(ms.results['cum_infections'].values, back['cum_infections'] * 3.3)
    assert np.allclose(ms.results['prevalence'].values, back['prevalence'] * 1.1)


# ---- reductions against the reference's MultiSim (run.py:220-374), from golden member results ------------------------------------------
def _members_from_golden(g):
    keys, vkeys = [str(k) for k in g['keys']], [str(k) for k in g['vkeys']]
    members = []
    for i in range(4):
        m = {k: g[f'member{i}/{k}'] for k in keys}
        m['variant'] = {k: g[f'member{i}/variant/{k}'] for k in vkeys}
        members.append(m)
    return keys, vkeys, members


def test_reduce_mean_combine_match_reference(golden):
    ''' Same member result series -> the reference's reduce() (median + 10/90 % band), mean() (+- 2 std), custom quantiles and combine() '''
    from covasim_b200 import run as cvrun
    g = golden('multisim_ref')
    keys, vkeys, members = _members_from_golden(g)
    msim = cvrun.MultiSim(base_sim=object(), n_runs=4)
    msim.member_results = members
    for name, call in (('median', lambda m: m.reduce()), ('mean', lambda m: m.mean()), ('quant', lambda m: m.reduce(quantiles=dict(low=0.25, high=0.75)))):
        call(msim)
        for k in keys:
            r = msim.results[k]
            for part, got in (('values', r.values), ('low', r.low), ('high', r.high)):
                np.testing.assert_allclose(got, g[f'{name}/{k}/{part}'], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=f'{name} {k} {part}')
        for k in vkeys:
            r = msim.results['variant'][k]
            for part, got in (('values', r.values), ('low', r.low), ('high', r.high)):
                np.testing.assert_allclose(got, g[f'{name}/variant/{k}/{part}'], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=f'{name} variant/{k} {part}')
    msim.combine()
    for k in keys:
        np.testing.assert_allclose(msim.results[k].values, g[f'combine/{k}'], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=f'combine {k}')
    for k in vkeys:
        np.testing.assert_allclose(msim.results['variant'][k].values, g[f'combine/variant/{k}'], rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=f'combine variant/{k}')
