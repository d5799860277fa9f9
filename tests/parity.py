'''
Helpers for the GPU parity tests: build a covasim_b200 sim and an oracle sim from the same scenario and
the same population, step them in lockstep and compare every People array each day.
'''
import numpy as np

import scenarios
from oracle import cvoracle as cvo

# float fields compared with a relative tolerance (float64 pow / exp / log / cos on the device differ from
# glibc's by <= 2 ulp before the result is rounded to float32); everything else must be bit-identical
TOL_FIELDS = {'sus_imm': 1e-6, 'symp_imm': 1e-6, 'sev_imm': 1e-6, 'peak_nab': 1e-6, 'nab': 1e-6}
# r_eff: the mean infectious duration is a float64 device sum here and a float32 NumPy mean in the oracle / reference (whose own
# cv.diff_sims compares with np.isclose, rtol 1e-5)
LOOSE_RESULTS = {'pop_nabs': 1e-6, 'pop_protection': 1e-6, 'pop_symp_protection': 1e-6, 'r_eff': 1e-6}


def build_pair(cv, name=None, spec=None, sim_kwargs=None, **extra):
    ''' (device sim, oracle sim), both initialised, sharing one population; ``sim_kwargs`` go to the device sim only '''
    spec = spec or scenarios.SCENARIOS[name]
    sim = cv.Sim(**scenarios.build(cv, spec, **extra), **(sim_kwargs or {}))
    sim.initialize()
    pop = dict(age=sim.people.to_numpy('age').astype(np.float64), sex=sim.people.to_numpy('sex'),
               contacts={lk: l.to_numpy() for lk, l in sim.people.contacts.items()})
    orc = cvo.OracleSim(**scenarios.build(cvo, spec, **extra), rng='philox', popdict=pop)
    orc.initialize()
    return sim, orc


def compare_people(sim, orc, where=''):
    bad = []
    for k in cvo.cvd.all_states:
        a = sim.people.to_numpy(k)
        b = orc.P[k]
        if k in TOL_FIELDS:
            ok = np.allclose(a, b, rtol=TOL_FIELDS[k], atol=0, equal_nan=True)
        else:
            ok = np.array_equal(a, b, equal_nan=(a.dtype.kind == 'f'))
        if not ok:
            if a.dtype.kind == 'f':
                diff = ~(np.isclose(a, b, rtol=TOL_FIELDS.get(k, 0), atol=0, equal_nan=True))
            else:
                diff = a != b
            idx = np.argwhere(diff)[:5]
            bad.append(f'{k}: {int(diff.sum())} differ, e.g. at {idx.tolist()} got {a[tuple(idx[0])]} want {b[tuple(idx[0])]}')
    assert not bad, f'People mismatch {where}:\n  ' + '\n  '.join(bad)


def compare_layers(sim, orc, where=''):
    for lk, layer in sim.people.contacts.items():
        got = layer.to_numpy()
        for c in ('p1', 'p2', 'beta'):
            assert np.array_equal(got[c], orc.contacts[lk][c]), f'layer {lk}.{c} differs {where}'


def compare_results(sim, orc):
    bad = []
    for k in sim.result_keys():
        a, b = sim.results[k].values, orc.results[k]
        rtol = LOOSE_RESULTS.get(k, 1e-9)
        if not np.allclose(a, b, rtol=rtol, atol=0, equal_nan=True):
            i = int(np.argwhere(~np.isclose(a, b, rtol=rtol, atol=0, equal_nan=True))[0][0])
            bad.append(f'{k}: first differs on day {i}: got {a[i]} want {b[i]}')
    for k in sim.result_keys('variant'):
        a, b = sim.results['variant'][k].values, orc.results['variant'][k]
        if not np.allclose(a, b, rtol=1e-9, atol=0, equal_nan=True):
            bad.append(f'variant/{k} differs')
    assert not bad, 'Result mismatch:\n  ' + '\n  '.join(bad)


def compare_log(sim, orc):
    log = sim.infection_log
    tgt, src, date, var, lay = [], [], [], [], []
    lmap = {lk: i for i, lk in enumerate(orc.contacts.keys())}
    lmap.update(seed_infection=-1, importation=-2)
    for e in orc.infection_log:
        n = len(e['target'])
        tgt.append(e['target'])
        src.append(np.full(n, -1, dtype=np.int32) if e['source'] is None else e['source'])
        date.append(np.full(n, e['date'], dtype=np.int32))
        var.append(np.full(n, e['variant'], dtype=np.int32))
        lay.append(np.full(n, lmap[e['layer']], dtype=np.int32))
    tgt, src, date, var, lay = (np.concatenate(x) for x in (tgt, src, date, var, lay))
    order = np.lexsort((tgt, lay, var, date))
    assert np.array_equal(log['target'], tgt[order])
    assert np.array_equal(log['date'], date[order])
    assert np.array_equal(log['variant'], var[order])
    assert np.array_equal(log['layer'], lay[order])
    assert np.array_equal(log['source'], src[order])


def check_packed_state(sim, where=''):
    ''' After a day that went through the fused kernels: the library's packed state words must equal what the People arrays say '''
    bad, diff, examples = sim.check_packed_state()
    assert bad == 0 and diff == 0, (f'packed state words differ from the People arrays {where}: {diff} agents (inexpressible: {bad}); '
                                    f'(agent, stored, recomputed) = {[(int(a), hex(int(b) & 0xFFFFFFFF), hex(int(c) & 0xFFFFFFFF)) for a, b, c in examples[:min(diff, 8)]]}')


def run_lockstep(sim, orc, every=1):
    compare_people(sim, orc, 'after initialize')
    sim.set_seed()
    orc.rng.set_seed(orc.pars['rand_seed'])
    while not sim.complete:
        t = sim.t
        fused_before = getattr(sim, 'fused_days', 0)
        sim.step()
        orc.step()
        if t % every == 0 or sim.complete:
            compare_people(sim, orc, f'after day {t}')
            compare_layers(sim, orc, f'after day {t}')
            if getattr(sim, 'fused_days', 0) > fused_before and not sim.pars['analyzers']:
                check_packed_state(sim, f'after day {t}')
    sim.finalize()
    orc.finalize()
    compare_results(sim, orc)
    compare_log(sim, orc)
