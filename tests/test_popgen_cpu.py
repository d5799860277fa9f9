'''
CPU tests of the keyed (device-side) population generator's host logic (covasim_b200/population.py:KeyedPop): the torch
index arithmetic, fed with uniforms from the oracle's NumPy Philox, must reproduce the oracle's plain-loop restatement
(oracle/cvoracle.py:make_keyed_pop) bit for bit -- whole layers and arbitrary chunks -- and the populations it draws must
have the statistics of the reference's own generators (population.py:143-364, here through the exact host port).
The same comparison with the uniforms coming from the CUDA kernel is tests/test_gpu_popgen.py.
'''
import numpy as np
import pytest
import torch

from oracle import cvoracle as cvo, philox as ph


def oracle_uniforms(seed):
    return lambda sub, i0, n, slot: torch.as_tensor(ph.keyed_uniform(seed, 10, sub, 0, np.arange(i0, i0 + n, dtype=np.int64), slot))


def make(pop_type, n, seed, **kw):
    from covasim_b200 import population as cvpop, parameters as cvpar
    pars = cvpar.make_pars(pop_size=n, pop_type=pop_type, **kw)
    return pars, cvpop.KeyedPop(pars, seed, 'cpu', oracle_uniforms(seed))


@pytest.mark.parametrize('pop_type,n,seed', [('hybrid', 2500, 3), ('random', 1800, 9), ('hybrid', 37, 1)])
def test_keyed_pop_matches_plain_loops(pop_type, n, seed):
    pars, gen = make(pop_type, n, seed)
    pop, ref = gen.materialize(), cvo.make_keyed_pop(pars, seed)
    assert np.array_equal(pop['age'].numpy(), ref['age']) and np.array_equal(pop['sex'].numpy(), ref['sex'])
    assert list(pop['contacts'].keys()) == list(ref['contacts'].keys())
    for lk, want in ref['contacts'].items():
        got = pop['contacts'][lk]
        assert np.array_equal(got['p1'].numpy(), want['p1']) and np.array_equal(got['p2'].numpy(), want['p2']), lk
        assert got['p1'].dtype == torch.int32 and np.all(got['beta'].numpy() == 1)


def test_chunks_tile_the_layers():
    pars, gen = make('hybrid', 3000, 5)
    for lk in gen.layer_keys():
        p1, p2, _ = gen.layer_edges(lk)
        m, off = gen.plans[lk]['m'], gen.plans[lk]['offsets']
        parts, e_expected = [], 0
        for a in range(0, m, 700):
            c1, c2, e0 = gen.layer_edges(lk, a, a + 700)
            assert e0 == e_expected == int(off[a])
            e_expected += c1.numel()
            parts.append((c1, c2))
        assert torch.equal(torch.cat([p[0] for p in parts]), p1) and torch.equal(torch.cat([p[1] for p in parts]), p2)
        assert e_expected == gen.plans[lk]['n_edges']


def test_statistics_match_the_reference_generators():
    ''' Same distributions as the reference's generators (host port, reference-exact for exact=True): layer sizes, degrees, ages '''
    from covasim_b200 import population as cvpop, parameters as cvpar, utils as cvu
    n = 40_000
    pars, gen = make('hybrid', n, 11)
    pop = gen.materialize()
    rng = cvu.HostStreams(11)
    ref = cvpop.make_randpop(cvpar.make_pars(pop_size=n, pop_type='hybrid'), rng, exact=False)
    for lk in ('h', 's', 'w', 'c'):
        a, b = pop['contacts'][lk]['p1'].numel(), len(ref['contacts'][lk]['p1'])
        assert abs(a - b) / b < 0.03, (lk, a, b)
    # ages: two-sample KS distance
    x, y = np.sort(pop['age'].numpy()), np.sort(ref['age'])
    grid = np.linspace(0, 100, 401)
    ks = np.max(np.abs(np.searchsorted(x, grid) / n - np.searchsorted(y, grid) / n))
    assert ks < 0.02, ks
    assert abs(pop['sex'].numpy().mean() - 0.5) < 0.01
    # school / work layers only connect agents in their age bands; households are cliques of consecutive agents
    age = pop['age'].numpy().astype(np.float32)
    for lk, (a0, a1) in (('s', (6, 22)), ('w', (22, 65))):
        for col in ('p1', 'p2'):
            ids = pop['contacts'][lk][col].numpy()
            assert np.all((age[ids] >= a0) & (age[ids] < a1))
    h1, h2 = pop['contacts']['h']['p1'].numpy(), pop['contacts']['h']['p2'].numpy()
    assert np.all(h2 > h1) and np.max(h2 - h1) < 20
    # degree distribution of the community layer: mean ~ contacts, as in the reference
    deg = np.bincount(pop['contacts']['c']['p1'].numpy(), minlength=n) + np.bincount(pop['contacts']['c']['p2'].numpy(), minlength=n)
    deg_ref = np.bincount(ref['contacts']['c']['p1'], minlength=n) + np.bincount(ref['contacts']['c']['p2'], minlength=n)
    assert abs(deg.mean() - deg_ref.mean()) < 0.3 and abs(deg.std() - deg_ref.std()) < 0.3


def test_poisson_cdf():
    from covasim_b200 import population as cvpop
    for lam in (0.7, 2.0, 16, 20):
        cdf = cvpop.poisson_cdf(lam)
        assert np.array_equal(cdf, cvo.poisson_cdf(lam))
        assert abs(cdf[-1] - 1) < 1e-12 and np.all(np.diff(cdf) >= 0)
        k = np.arange(len(cdf))
        pmf = np.diff(np.concatenate([[0], cdf]))
        assert abs((k * pmf).sum() - lam) < 1e-9


def test_choose_distinct_properties():
    ''' The O(k) sampler of native-RNG mode: k distinct values in range, deterministic given the stream, same function in the oracle '''
    from covasim_b200 import utils as cvu
    for n, k in [(10, 0), (10, 1), (1000, 7), (1000, 124), (1000, 126), (50, 50), (2_000_000, 5000)]:
        a = cvu.choose_distinct(np.random.RandomState(3), n, k)
        b = cvo.choose_distinct(np.random.RandomState(3), n, k)
        assert len(a) == k and len(np.unique(a)) == k and (k == 0 or (a.min() >= 0 and a.max() < n))
        assert np.array_equal(a, b)
    # uniform: every value about equally likely
    counts = np.zeros(200)
    rs = np.random.RandomState(0)
    for _ in range(4000):
        counts[cvu.choose_distinct(rs, 200, 10)] += 1
    assert abs(counts.mean() - 200) < 1e-9 and counts.std() < 3 * np.sqrt(200 * 0.95)
