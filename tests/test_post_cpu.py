'''
CPU tests of the post-processing that consumes the hot path's outputs -- the windowed r_eff methods and the generation time
(reference sim.py:888-1025) -- as functions of plain arrays, against the values the unmodified reference computed for the
golden scenarios (oracle/gen_golden.py stores r_eff/infectious, r_eff/outcome and gen_time next to the People arrays and the
infection log they are computed from).
'''
import numpy as np
import pytest

import scenarios

FULL = ['hybrid3k', 'random2k_nowaning', 'variants4k', 'dynamic2k', 'dynpars3k', 'clip3k', 'rescale3k', 'fracsus2k', 'sequence3k',
        'testnum3k', 'subtarget3k', 'capacity3k']


@pytest.mark.parametrize('name', FULL)
def test_r_eff_methods_and_gen_time(name, golden):
    from covasim_b200.sim import r_eff_windowed, gen_time
    g = golden(name)
    npts = scenarios.SCENARIOS[name]['pars'].get('n_days', 60) + 1
    for method in ('infectious', 'outcome'):
        got = r_eff_windowed(method, g['people/date_infectious'], g['people/date_recovered'], g['people/date_dead'], g['log/source'], npts)
        np.testing.assert_allclose(got, g[f'r_eff/{method}'], rtol=1e-12, atol=0, equal_nan=True, err_msg=f'{name} {method}')
        assert np.isfinite(got).sum() > npts // 2
    gt = gen_time(g['people/date_exposed'], g['people/date_symptomatic'], g['log/source'], g['log/target'])
    np.testing.assert_allclose([gt['true'], gt['true_std'], gt['clinical'], gt['clinical_std']], g['gen_time'], rtol=1e-12)
    assert np.isnan(gt['true']) or 2 < gt['true'] < 15          # NaN when sources were made naive again (rescaling), as in the reference


# ---- goodness of fit (reference analysis.py:991-1222 Fit, misc.py:707-795 compute_gof) ---------------------------------------------
FIT_SCENARIOS = ['hybrid3k', 'variants4k', 'rescale3k', 'testnum3k']


def _fit_inputs(g):
    import json
    results = {k[len('results/'):]: g[k] for k in g.files if k.startswith('results/')}
    data = {k[len('fit/data/'):]: g[k] for k in g.files if k.startswith('fit/data/')}
    data['day'] = g['fit/days']
    return results, data, json.loads(str(g['fit/keys']))


@pytest.mark.parametrize('name', FIT_SCENARIOS)
def test_fit_matches_reference(name, golden):
    ''' Same result series, same (synthetic, partly missing) data -> the reference's keys, gofs, losses and mismatch '''
    from covasim_b200 import analysis as cva
    g = golden(name)
    results, data, ref_keys = _fit_inputs(g)
    npts = len(results['cum_infections'])
    fit = cva.Fit(data=data, results=results, npts=npts)
    assert fit.keys == ref_keys
    for k in ref_keys:
        np.testing.assert_allclose(fit.gofs[k], g[f'fit/gofs/{k}'], rtol=1e-12, atol=0)
        np.testing.assert_allclose(fit.losses[k], g[f'fit/losses/{k}'], rtol=1e-12, atol=0)
    assert fit.mismatch == pytest.approx(float(g['fit/mismatch']), rel=1e-12)
    fit2 = cva.Fit(data=data, results=results, npts=npts, keys=['cum_diagnoses', 'new_infections'], weights=dict(new_infections=2.5),
                   use_squared=True, as_scalar='mean')
    assert fit2.mismatch == pytest.approx(float(g['fit/mismatch_custom']), rel=1e-12)
    # dates instead of day numbers, and the ensemble form
    import datetime as dt
    dated = {k: v for k, v in data.items() if k != 'day'}
    dated['date'] = [(dt.date(2020, 3, 1) + dt.timedelta(days=int(d))).strftime('%Y-%m-%d') for d in data['day']]
    assert cva.Fit(data=dated, results=results, npts=npts, start_date=dt.date(2020, 3, 1)).mismatch == pytest.approx(fit.mismatch, rel=1e-14)
    worse = {k: (v * 1.5 if k.startswith('cum_') else v) for k, v in results.items()}
    mm = cva.fit_members([results, worse], data, npts)
    assert mm[0] == pytest.approx(fit.mismatch, rel=1e-14) and mm[1] > mm[0]


def test_compute_gof_options():
    from covasim_b200.analysis import compute_gof
    a, p = np.array([1.0, 2.0, 4.0]), np.array([1.5, 2.0, 3.0])
    np.testing.assert_allclose(compute_gof(a, p), np.abs(a - p) / 4.0)
    np.testing.assert_allclose(compute_gof(a, p, normalize=False), np.abs(a - p))
    np.testing.assert_allclose(compute_gof(a, p, use_frac=True), np.abs(a - p) / (np.maximum(a, p) + 1e-9))
    assert compute_gof(a, p, normalize=False, use_squared=True, as_scalar='sum') == pytest.approx(0.25 + 1.0)
    assert compute_gof(a, p, estimator=lambda x, y: float(np.max(np.abs(x - y)))) == 1.0


# ---- analyzers (reference analysis.py:23-425) on a stand-in sim with CPU tensors ------------------------------------------------------
class _FakePeople:
    def __init__(self, arrays):
        self._a = arrays

    def __getattr__(self, k):
        return self._a[k]

    def __getitem__(self, k):
        return self._a[k]

    def keys(self):
        return list(self._a.keys())

    def to_numpy(self, k):
        return self._a[k].numpy()


class _FakeSim:
    def __init__(self, n=5000, npts=20, seed=0):
        import torch
        rng = np.random.RandomState(seed)
        nan = np.where(rng.random_sample(n) < 0.6, np.nan, rng.randint(0, npts, n)).astype(np.float32)
        self.people = _FakePeople(dict(age=torch.as_tensor((rng.random_sample(n) * 100).astype(np.float32)),
                                       date_exposed=torch.as_tensor(nan), date_dead=torch.as_tensor(np.where(rng.random_sample(n) < 0.97, np.nan, 5).astype(np.float32)),
                                       exposed=torch.as_tensor(rng.random_sample(n) < 0.1)))
        self.npts, self.t = npts, 0
        self.rescale_vec = np.linspace(1, 2, npts)

    def day(self, d):
        return int(d)

    def date(self, d):
        return f'day{int(d):03d}'


def test_analyzers_on_a_stand_in_sim():
    import covasim_b200 as cv
    sim = _FakeSim()
    snap = cv.snapshot(3, 7)
    hist = cv.age_histogram(days=[7, 19], states=['exposed', 'date_dead'])
    with pytest.raises(RuntimeError):
        snap(sim)                                              # not initialised yet
    for an in (snap, hist):
        an.initialize(sim)
    for t in range(sim.npts):
        sim.t = t
        if t == 5:
            sim.people._a['exposed'] = ~sim.people._a['exposed']      # the snapshots must be copies of the day's state
        for an in (snap, hist):
            an(sim)
    for an in (snap, hist):
        an.finalize(sim)
    assert list(snap.snapshots.keys()) == ['day003', 'day007']
    assert not np.array_equal(snap.get(3)['exposed'], snap.get(7)['exposed'])
    age = sim.people.age.numpy()
    for date, t in (('day007', 7), ('day019', 19)):
        for state in ('exposed', 'dead'):
            want = np.histogram(age[~np.isnan(sim.people[f'date_{state}'].numpy())], bins=np.linspace(0, 100, 11))[0] * sim.rescale_vec[t]
            np.testing.assert_allclose(hist.hists[date][state], want)
    with pytest.raises(RuntimeError):
        snap.finalize(sim)                                     # finalising twice is an error, as in the reference
    assert isinstance(snap, cv.Analyzer) and snap.label == 'snapshot'


@pytest.mark.parametrize('name', ['hybrid3k', 'variants4k'])
def test_transtree_matches_reference(name):
    ''' The array TransTree built from a golden infection log against what the reference's TransTree computed for the same run
    (tests/golden/transtree_ref.npz, oracle/gen_transtree_golden.py): count_targets with and without a day window, the sources
    of every person, and the transmissions '''
    import os
    from covasim_b200.analysis import TransTree
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, 'golden', f'{name}.npz'))
    ref = np.load(os.path.join(here, 'golden', 'transtree_ref.npz'))
    n_log, pop_size, n_days = (int(x) for x in ref[f'{name}/shape'])
    log = {k: g['log/' + k] for k in ('source', 'target', 'date', 'layer', 'variant')}
    tt = TransTree(log=log, pop_size=pop_size, n_days=n_days)
    assert len(tt) == n_log
    assert np.array_equal(tt.sources, ref[f'{name}/sources'])
    assert np.array_equal(tt.count_targets(), ref[f'{name}/n_targets'])
    assert np.array_equal(tt.count_targets(start_day=10, end_day=30), ref[f'{name}/n_targets_10_30'])
    mine = tt.count_transmissions()
    want = ref[f'{name}/transmissions']
    assert np.array_equal(mine[np.lexsort(mine.T[::-1])], want[np.lexsort(want.T[::-1])])


def test_compact_arena_encoding_round_trip():
    '''
    Sim._compact_arena (the compact form Sim.restore sends to the device: per array one fill value + exceptions, or dense) decoded in
    NumPy must give back every array byte for byte: constant arrays, sparse exceptions, dense arrays, bool arrays whose length is not a
    multiple of 4 (the last word shares its bytes with alignment padding), 2-D by-variant arrays, an all-NaN date array.
    '''
    import types
    import torch
    from covasim_b200.sim import Sim
    rng = np.random.default_rng(3)
    n, nv = 4003, 3
    fields = [('uid', np.int32, (n,)), ('age', np.float32, (n,)), ('sex', np.bool_, (n,)), ('susceptible', np.bool_, (n,)), ('date_exposed', np.float32, (n,)),
              ('date_dead', np.float32, (n,)), ('sus_imm', np.float32, (nv, n)), ('exposed_by_variant', np.bool_, (nv, n)), ('n_infections', np.int32, (n,))]
    layout, total = [], 0
    for name, dt, shape in fields:
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        layout.append((name, dt, shape, total, nbytes))
        total += (nbytes + 255) // 256 * 256
    arena = rng.integers(0, 255, total, dtype=np.uint8)                 # padding holds garbage, as on the device
    vals = dict(uid=np.arange(n, dtype=np.int32), age=rng.random(n, dtype=np.float32) * 90, sex=rng.random(n) < 0.5, susceptible=np.ones(n, dtype=bool),
                date_exposed=np.full(n, np.nan, dtype=np.float32), date_dead=np.full(n, np.nan, dtype=np.float32),
                sus_imm=np.zeros((nv, n), dtype=np.float32), exposed_by_variant=np.zeros((nv, n), dtype=bool), n_infections=np.zeros(n, dtype=np.int32))
    sick = rng.choice(n, 17, replace=False)
    vals['susceptible'][sick] = False
    vals['date_exposed'][sick] = rng.integers(0, 9, 17)
    vals['exposed_by_variant'][1, sick[:5]] = True
    vals['n_infections'][sick] = 1
    vals['n_infections'][-1] = 2                                        # the very last agent: the word shared with the padding
    vals['susceptible'][-1] = False
    for name, dt, shape, off, nbytes in layout:
        arena[off:off + nbytes] = np.ascontiguousarray(vals[name]).view(np.uint8).reshape(-1)
    host = torch.from_numpy(arena.copy())
    stub = types.SimpleNamespace(people=types.SimpleNamespace(_layout=layout))
    c = Sim._compact_arena(stub, host, pinned=False)
    assert c['n_seg'] >= 6 and c['arena_bytes'] == total and c['h2d_bytes'] < total
    table = c['table'].numpy()
    out = np.full(total, 0xEE, dtype=np.uint8)
    words = out.view(np.uint32)
    seg = table[:3 * c['n_seg']].reshape(-1, 3)
    for begin, count, value in seg:
        words[begin:begin + count] = np.uint32(value)
    idx, val = table[3 * c['n_seg']:3 * c['n_seg'] + c['n_exc']], table[3 * c['n_seg'] + c['n_exc']:3 * c['n_seg'] + 2 * c['n_exc']]
    words[idx] = val.astype(np.uint32)
    for a, b in c['dense']:
        out[a:b] = arena[a:b]
    dense_names = set()
    for name, dt, shape, off, nbytes in layout:
        got = out[off:off + nbytes].view(dt).reshape(shape)
        assert np.array_equal(got, vals[name], equal_nan=np.dtype(dt).kind == 'f'), name
        if any(a <= off and off + nbytes <= b for a, b in c['dense']):
            dense_names.add(name)
    assert dense_names == {'uid', 'age', 'sex'}                        # what really is dense at day 0 -- everything else is one value + exceptions
