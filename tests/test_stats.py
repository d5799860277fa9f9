'''
Native-RNG mode vs the reference, statistically (north star: KS test and 95 % envelopes over many seeds, mirroring
the regression checks of the reference's tests/baseline.json).  The reference distribution (200 seeds of the
unmodified reference, MT19937 streams) is committed in tests/golden/stats_ref.npz by oracle/gen_stats.py; here the
same configuration is run with Philox keyed draws -- 200 seeds on the GPU through cv.MultiSim, and 60 seeds of the
oracle on the CPU -- and the distributions of the epidemic summaries must be indistinguishable.
'''
import json

import numpy as np
import pytest
from scipy import stats

KS_ALPHA = 1e-3          # the seeds are fixed, so the test is deterministic; alpha only sets how strict it is


def summaries(curves):
    ''' Scalar summaries per seed from dict key -> [n_seeds, npts] '''
    ci, ni = curves['cum_infections'], curves['new_infections']
    return {
        'final cum_infections': ci[:, -1],
        'cum_infections day 30': ci[:, 30],
        'peak new_infections': ni.max(axis=1),
        'day of peak': ni.argmax(axis=1).astype(float),
        'total deaths': curves['new_deaths'].sum(axis=1),
        'total diagnoses': curves['new_diagnoses'].sum(axis=1),
        'total quarantined': curves['new_quarantined'].sum(axis=1),
        'final n_exposed': curves['n_exposed'][:, -1],
    }


def compare(ref, got, label):
    sr, sg = summaries(ref), summaries(got)
    bad = []
    for k in sr:
        res = stats.ks_2samp(sr[k], sg[k])
        if res.pvalue < KS_ALPHA:
            bad.append(f'{k}: KS D={res.statistic:.3f} p={res.pvalue:.2e} (ref median {np.median(sr[k]):.1f}, {label} median {np.median(sg[k]):.1f})')
    assert not bad, f'{label} differs from the reference distribution:\n  ' + '\n  '.join(bad)
    # 95 % envelopes: the ensemble median of each daily curve stays inside the reference's central 95 % band
    for key in ('new_infections', 'cum_infections', 'new_deaths'):
        lo, hi = np.quantile(ref[key], [0.025, 0.975], axis=0)
        med = np.median(got[key], axis=0)
        inside = np.mean((med >= lo) & (med <= hi))
        assert inside >= 0.95, f'{label}: median {key} curve is inside the reference 95 % envelope on only {inside:.0%} of days'
        # and the two medians track each other within the band width
        ref_med = np.median(ref[key], axis=0)
        assert np.all(np.abs(med - ref_med) <= np.maximum(hi - lo, 1.0)), f'{label}: median {key} curve leaves the band width'


def load_ref(golden):
    g = golden('stats_ref')
    cfg = json.loads(str(g['config']))
    keys = [k for k in g.files if k != 'config']
    return cfg, {k: g[k].astype(np.float64) for k in keys}


def test_oracle_philox_matches_reference_distribution(golden):
    ''' CPU: 60 seeds of the oracle in Philox mode vs 200 seeds of the reference '''
    from oracle import cvoracle as cvo
    cfg, ref = load_ref(golden)
    got = {k: [] for k in ref}
    for i in range(60):
        ivs = [getattr(cvo, name)(**kw) for name, kw in cfg['interventions']]
        sim = cvo.OracleSim(dict(cfg['pars'], rand_seed=5000 + i), interventions=ivs, rng='philox')
        sim.keep_log = False
        sim.run()
        for k in got:
            got[k].append(np.asarray(sim.results[k], dtype=np.float64))
    compare(ref, {k: np.stack(v) for k, v in got.items()}, 'oracle (philox)')


@pytest.mark.gpu
def test_gpu_native_rng_matches_reference_distribution(golden):
    ''' GPU: 200 seeds through cv.MultiSim vs 200 seeds of the reference '''
    import covasim_b200 as cv
    cfg, ref = load_ref(golden)
    n_seeds = ref['cum_infections'].shape[0]
    ivs = [getattr(cv, name)(**kw) for name, kw in cfg['interventions']]
    base = cv.Sim(dict(cfg['pars'], rand_seed=7000), interventions=ivs)
    msim = cv.MultiSim(base, n_runs=n_seeds)
    msim.run()
    got = {k: np.stack([m[k] for m in msim.member_results]) for k in ref}
    compare(ref, got, 'covasim_b200 (philox)')
