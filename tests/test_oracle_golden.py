'''
Pins the oracle (oracle/cvoracle.py, 'mt' RNG mode) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py): every result series, the final People arrays, the
infection log, and -- for 'baseline20k' -- the 58 values of the reference's own tests/baseline.json.
Also checks the oracle's kernel restatements against the recorded inputs/outputs of the reference's
Numba kernels.  CPU only.
'''
import hashlib
import json
import os

import numpy as np
import pytest

import scenarios
from oracle import cvoracle as cvo

LOOSE = ('pop_nabs', 'pop_protection', 'pop_symp_protection')   # float32 reductions: summation order is platform-dependent
_cache = {}


def run_oracle(name):
    if name not in _cache:
        sim = cvo.OracleSim(**scenarios.build(cvo, scenarios.SCENARIOS[name]), rng='mt')
        sim.run()
        _cache[name] = sim
    return _cache[name]


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


@pytest.mark.parametrize('name', list(scenarios.SCENARIOS.keys()))
def test_results_match_reference(name, golden):
    g = golden(name)
    sim = run_oracle(name)
    for key in g.files:
        if key.startswith('results/'):
            k = key.split('/', 1)[1]
            rtol = 1e-6 if k in LOOSE else 1e-12
            np.testing.assert_allclose(sim.results[k], g[key], rtol=rtol, atol=0, equal_nan=True, err_msg=k)
        elif key.startswith('vresults/'):
            k = key.split('/', 1)[1]
            np.testing.assert_allclose(sim.results['variant'][k], g[key], rtol=1e-12, atol=0, equal_nan=True, err_msg=k)


@pytest.mark.parametrize('name', list(scenarios.SCENARIOS.keys()))
def test_people_and_log_match_reference(name, golden):
    g = golden(name)
    sim = run_oracle(name)
    for lk, layer in sim.contacts.items():
        if f'final_contacts_len/{lk}' in g.files:             # layers as the run left them (clip_edges moves edges): same edges, same order
            assert len(layer['p1']) == int(g[f'final_contacts_len/{lk}'])
            assert digest(layer['p1']) + digest(layer['p2']) == str(g[f'final_contacts_digest/{lk}']), lk
        else:
            assert len(layer['p1']) == int(g[f'contacts_len/{lk}'])
    for k in cvo.cvd.all_states:
        if f'people/{k}' in g.files:
            assert np.array_equal(sim.P[k], g[f'people/{k}'], equal_nan=True), k
        else:
            assert digest(sim.P[k]) == str(g[f'people_digest/{k}']), k
    log = sim.infection_log
    tgt = np.concatenate([e['target'] for e in log])
    src = np.concatenate([np.full(len(e['target']), -1, dtype=np.int32) if e['source'] is None else e['source'] for e in log])
    date = np.concatenate([np.full(len(e['target']), e['date'], dtype=np.int32) for e in log])
    var = np.concatenate([np.full(len(e['target']), e['variant'], dtype=np.int32) for e in log])
    assert np.array_equal(tgt, g['log/target'])
    assert np.array_equal(src, g['log/source'])
    assert np.array_equal(date, g['log/date'])
    assert np.array_equal(var, g['log/variant'])


def test_baseline_json(golden):
    ''' The 58 golden values of the reference's tests/baseline.json (reference tests/test_baselines.py:81-95) '''
    g = golden('baseline20k')
    base = json.loads(str(g['baseline_json']))
    assert len(base) == 58
    sim = run_oracle('baseline20k')
    for k, v in base.items():
        assert np.isclose(sim.summary[k], v, rtol=1e-6 if k in LOOSE else 1e-12, atol=0), (k, sim.summary[k], v)


@pytest.mark.parametrize('name,day', [('hybrid3k', 12), ('hybrid3k', 25), ('variants4k', 20)])
def test_kernel_vectors(name, day, golden):
    ''' Recorded calls of the reference's Numba kernels on one simulated day '''
    g = golden(name)
    pre = f'k{day}/'
    vl = cvo.compute_viral_load(int(g[pre + 'vl/t']), g[pre + 'vl/date_inf'], g[pre + 'vl/date_rec'], g[pre + 'vl/date_dead'],
                                g[pre + 'vl/frac_time'], g[pre + 'vl/load_ratio'], g[pre + 'vl/high_cap'])
    assert np.array_equal(vl, g[pre + 'vl/out'])
    n_calls = int(g[pre + 'n_calls'])
    assert n_calls > 0
    sim = run_oracle(name)     # only for the (static) layer arrays, regenerated from the seed
    by_len = {len(l['p1']): l for l in sim.contacts.values()}
    for j in range(n_calls):
        a = {k.split('/')[-1]: g[k] for k in g.files if k.startswith(f'{pre}ts{j}/')}
        rt, rs = cvo.compute_trans_sus(a['rel_trans'], a['rel_sus'], a['inf'], a['sus'], a['beta_layer'], a['viral_load'],
                                       a['symp'], a['iso'], a['quar'], a['asymp_factor'], a['iso_factor'], a['quar_factor'],
                                       a['immunity_factors'])
        assert np.array_equal(rt, a['out_trans'])
        assert np.array_equal(rs, a['out_sus'])
        c = {k.split('/')[-1]: g[k] for k in g.files if k.startswith(f'{pre}ci{j}/')}
        layer = by_len[int(c['n_edges'])]
        assert digest(layer['p1']) == str(c['p1_digest'])
        us = [c['u_dir1'], c['u_dir2']]
        src, tgt = cvo.compute_infections(c['beta'], layer['p1'], layer['p2'], layer['beta'], c['rel_trans'], c['rel_sus'],
                                          lambda d, e: us[d])
        assert np.array_equal(src, c['out_src']) and np.array_equal(tgt, c['out_tgt'])
    for lk, layer in sim.contacts.items():
        out = cvo.find_contacts(layer['p1'], layer['p2'], g[f'{pre}fc/{lk}/inds'])
        assert np.array_equal(out, g[f'{pre}fc/{lk}/out'])


def test_known_answers():
    ''' RNG-free known answers (reference tests/unittests/test_transmission.py:16-35, test_mortality.py:13-24) '''
    sim = cvo.OracleSim(pop_size=1500, pop_infected=30, n_days=30, beta=0.0, rand_seed=4).run()
    assert sim.results['new_infections'].sum() == 0
    big = 1e6
    sim = cvo.OracleSim(pop_size=400, pop_infected=400, n_days=120, rand_seed=4, use_waning=False, rel_symp_prob=big,
                        rel_severe_prob=big, rel_crit_prob=big, rel_death_prob=big).run()
    assert sim.summary['cum_deaths'] == 400


@pytest.mark.parametrize('name', ['hybrid3k', 'default20k', 'dynamic2k', 'baseline20k'])
def test_initial_population_matches_reference(name, golden):
    ''' Ages, initial transmissibility and every layer's edge list (order included) as the reference built them (population.py:143-364) '''
    g = golden(name)
    sim = cvo.OracleSim(**scenarios.build(cvo, scenarios.SCENARIOS[name]), rng='mt')
    sim.initialize()
    assert np.array_equal(sim.P['age'], g['pop/age'].astype(sim.P['age'].dtype))
    assert np.array_equal(sim.P['rel_trans'], g['init/rel_trans'])
    for lk, layer in sim.contacts.items():
        assert len(layer['p1']) == int(g[f'contacts_len/{lk}'])
        assert digest(layer['p1']) + digest(layer['p2']) == str(g[f'contacts_digest/{lk}']), lk


def test_config_tables():
    '''
    The parameter tables of the product (covasim_b200/{defaults,parameters}.py) and of the oracle's own copy
    (oracle/{ref_defaults,ref_parameters}.py) against the tables recorded from the unmodified reference
    (tests/golden/ref_config.json, written by oracle/gen_config_golden.py): a constant cannot be wrong on both sides.
    The product modules are loaded from their files so that this CPU test does not load the CUDA library.
    '''
    import importlib.util
    import sys
    import types
    from oracle import gen_config_golden as gen, ref_defaults as od, ref_parameters as op
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = json.load(open(os.path.join(root, 'tests', 'golden', 'ref_config.json')))
    pkg = types.ModuleType('_cvb_cfg')                      # a stand-in package holding only the two config modules
    pkg.__path__ = [os.path.join(root, 'covasim_b200')]
    sys.modules['_cvb_cfg'] = pkg
    mods = {}
    for name in ('defaults', 'parameters'):
        spec = importlib.util.spec_from_file_location(f'_cvb_cfg.{name}', os.path.join(root, 'covasim_b200', f'{name}.py'))
        mods[name] = importlib.util.module_from_spec(spec)
        sys.modules[f'_cvb_cfg.{name}'] = mods[name]
        spec.loader.exec_module(mods[name])
    for label, (par, dfl) in dict(oracle=(op, od), product=(mods['parameters'], mods['defaults'])).items():
        got = json.loads(json.dumps(gen.collect(par, dfl), sort_keys=True))
        for key, val in got.items():
            assert val == ref[key], f'{label} copy of "{key}" differs from the reference'
    fields = ref['people_fields']
    for dfl in (od, mods['defaults']):
        assert list(dfl.person_fields) == fields['person'] and list(dfl.states) == fields['states'] and list(dfl.dates) == fields['dates']
        assert list(dfl.variant_states) == fields['variant_states'] and list(dfl.by_variant_states) == fields['by_variant_states']
        assert list(dfl.imm_states) == fields['imm_states'] and list(dfl.nab_states) == fields['nab_states']
        assert list(dfl.vacc_states) == fields['vacc_states'] and list(dfl.durs) == fields['durs']


def test_oracle_does_not_import_the_product():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = 'import sys; from oracle import cvoracle, philox; assert not any(m.startswith("covasim_b200") for m in sys.modules), "oracle imported the product"'
    subprocess.run([sys.executable, '-c', code], check=True, cwd=root)
