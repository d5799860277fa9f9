'''
TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, Covasim 3.1.7) in the build container through the stand-in sciris/pylab modules
of oracle/shim.  Run from the repo root:  python -m oracle.gen_golden

For every scenario of tests/scenarios.py it records
  * every result time series of the finished sim (and the 58-value summary for 'baseline20k', which
    it also checks against the reference's own tests/baseline.json);
  * the final People arrays (full arrays for the small scenarios, SHA-256 digests for the 20k ones);
  * the infection log as flat arrays (source, target, date, layer index, variant index);
  * kernel-level vectors for selected days: the exact inputs and outputs of the reference's Numba
    kernels compute_viral_load / compute_trans_sus / compute_infections / find_contacts, together
    with the uniform draws compute_infections consumed.  Those draws are not observable from outside
    Numba, so the script keeps a NumPy RandomState mirror of the Numba stream in lockstep (every Numba-
    stream consumer is wrapped and its output asserted equal to the mirror's) -- which is also the
    check that the two MT19937 streams are the same algorithm.
'''
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refenv  # noqa: E402
cv = refenv.import_reference()
import covasim.utils as cvu  # noqa: E402
from oracle import cvoracle as cvo  # noqa: E402
import scenarios  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
FULL_PEOPLE = {'hybrid3k', 'random2k_nowaning', 'variants4k', 'dynamic2k', 'dynpars3k', 'clip3k', 'rescale3k', 'fracsus2k', 'sequence3k', 'testnum3k', 'testnum_rescale2k', 'subtarget3k', 'ili3k', 'capacity3k', 'vaccnum3k', 'testnum_sub3k', 'targeteff3k', 'swab3k', 'quarpol3k'}
KERNEL_DAYS = {'hybrid3k': [12, 25], 'variants4k': [20], 'baseline20k': []}


class NumbaMirror:
    ''' Keeps np.random.RandomState in lockstep with Numba's MT19937 stream by wrapping its consumers '''

    def __init__(self):
        self.rs = np.random.RandomState()
        self.capture = None            # dict to fill with kernel vectors, or None
        self.orig = {k: getattr(cvu, k) for k in ('set_seed', 'choose', 'choose_r', 'poisson', 'n_poisson',
                                                  'compute_infections', 'compute_trans_sus', 'compute_viral_load')}
        cvu.set_seed = self.set_seed
        cvu.choose = self.choose
        cvu.choose_r = self.choose_r
        cvu.poisson = self.poisson
        cvu.n_poisson = self.n_poisson
        cvu.compute_infections = self.compute_infections
        cvu.compute_trans_sus = self.compute_trans_sus
        cvu.compute_viral_load = self.compute_viral_load
        self.n_checked = 0

    def restore(self):
        for k, f in self.orig.items():
            setattr(cvu, k, f)

    def set_seed(self, seed=None):
        self.orig['set_seed'](seed)
        self.rs.seed(int(seed))

    def choose(self, max_n, n):
        out = self.orig['choose'](max_n, n)
        assert np.array_equal(out, self.rs.choice(int(max_n), int(n), replace=False))
        return out

    def choose_r(self, max_n, n):
        out = self.orig['choose_r'](max_n, n)
        assert np.array_equal(out, self.rs.choice(int(max_n), int(n), replace=True))
        return out

    def poisson(self, rate):
        out = self.orig['poisson'](rate)
        assert out == self.rs.poisson(np.float32(rate), 1)[0]
        return out

    def n_poisson(self, rate, n):
        out = self.orig['n_poisson'](rate, n)
        assert np.array_equal(out, self.rs.poisson(np.float32(rate), int(n)))
        return out

    def compute_viral_load(self, t, *args):
        out = self.orig['compute_viral_load'](t, *args)
        if self.capture is not None:
            self.capture['vl'] = dict(t=np.int32(t), date_inf=args[0].copy(), date_rec=args[1].copy(), date_dead=args[2].copy(),
                                      frac_time=args[3], load_ratio=args[4], high_cap=args[5], out=out.copy())
        return out

    def compute_trans_sus(self, *args):
        rt, rs = self.orig['compute_trans_sus'](*args)
        if self.capture is not None:
            names = ('rel_trans', 'rel_sus', 'inf', 'sus', 'beta_layer', 'viral_load', 'symp', 'iso', 'quar',
                     'asymp_factor', 'iso_factor', 'quar_factor', 'immunity_factors')
            rec = {n: np.copy(a) for n, a in zip(names, args)}
            rec['out_trans'], rec['out_sus'] = rt.copy(), rs.copy()
            self.capture.setdefault('ts', []).append(rec)
        return rt, rs

    def compute_infections(self, beta, p1, p2, betas, rel_trans, rel_sus, legacy=False):
        src, tgt = self.orig['compute_infections'](beta, p1, p2, betas, rel_trans, rel_sus, legacy)
        drawn = []

        def draw(direction, edges):
            u = self.rs.random_sample(len(edges))
            drawn.append(u)
            return u
        s2, t2 = cvo.compute_infections(beta, p1, p2, betas, rel_trans, rel_sus, draw)
        assert np.array_equal(src, s2) and np.array_equal(tgt, t2), 'Numba-stream mirror lost lockstep'
        self.n_checked += 1
        if self.capture is not None:
            self.capture.setdefault('ci', []).append(dict(
                beta=np.float32(beta), p1=p1.copy(), p2=p2.copy(), betas=betas.copy(), rel_trans=rel_trans.copy(),
                rel_sus=rel_sus.copy(), u_dir1=drawn[0], u_dir2=drawn[1], out_src=src.copy(), out_tgt=tgt.copy(),
                n_edges=np.int64(len(p1)), p1_digest=np.array(digest(p1))))
        return src, tgt


def digest(arr):
    a = np.ascontiguousarray(arr)
    return hashlib.sha256(a.tobytes()).hexdigest()


def run_scenario(name, spec):
    mirror = NumbaMirror()
    out = {}
    try:
        sim = cv.Sim(**scenarios.build(cv, spec))
        sim.initialize()
        out['pop/age'] = np.array(sim.people.age)
        for lk, layer in sim.people.contacts.items():
            out[f'contacts_digest/{lk}'] = np.array(digest(layer['p1']) + digest(layer['p2']))
            out[f'contacts_len/{lk}'] = np.int64(len(layer))
        out['init/rel_trans'] = sim.people.rel_trans.copy()
        kdays = KERNEL_DAYS.get(name, [])
        sim.set_seed()          # what Sim.run() does before stepping (reference sim.py:713-716)
        while not sim.complete:
            t = sim.t
            if t in kdays:
                mirror.capture = {}
                # find_contacts vector: contacts of everyone currently infectious, per layer
                inds = cvu.true(sim.people.infectious)
                for lk, layer in sim.people.contacts.items():
                    out[f'k{t}/fc/{lk}/inds'] = inds.astype(np.int64)
                    out[f'k{t}/fc/{lk}/out'] = layer.find_contacts(inds)
            sim.step()
            if t in kdays:
                cap, mirror.capture = mirror.capture, None
                for k, v in cap['vl'].items():
                    out[f'k{t}/vl/{k}'] = np.asarray(v)
                for j, rec in enumerate(cap.get('ts', [])):
                    for k, v in rec.items():
                        out[f'k{t}/ts{j}/{k}'] = np.asarray(v)
                for j, rec in enumerate(cap.get('ci', [])):
                    for k, v in rec.items():
                        if k in ('p1', 'p2', 'betas'):
                            continue        # the layer arrays are regenerated from the seed by the tests
                        out[f'k{t}/ci{j}/{k}'] = np.asarray(v)
                out[f'k{t}/n_calls'] = np.int64(len(cap.get('ci', [])))
        for lk, layer in sim.people.contacts.items():          # layers as the run left them (clip_edges moves edges)
            out[f'final_contacts_digest/{lk}'] = np.array(digest(layer['p1']) + digest(layer['p2']))
            out[f'final_contacts_len/{lk}'] = np.int64(len(layer))
        sim.finalize()
    finally:
        mirror.restore()
    print(f'  {name}: numba-stream lockstep verified on {mirror.n_checked} compute_infections calls')

    for k in sim.result_keys():
        out[f'results/{k}'] = np.array(sim.results[k].values)
    # the other r_eff methods and the generation time (post-processing of the dates and the infection log, sim.py:888-1025)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for method in ('infectious', 'outcome'):
            out[f'r_eff/{method}'] = np.array(sim.compute_r_eff(method=method))
        gt = sim.compute_gen_time()
        out['gen_time'] = np.array([gt['true'], gt['true_std'], gt['clinical'], gt['clinical_std']], dtype=np.float64)
        sim.compute_r_eff()                                    # back to the default
    # goodness of fit against a synthetic data set (analysis.py:991-1222 Fit): the data and the reference's answers
    import datetime as _dt
    import pandas as pd
    days = np.arange(3, sim.npts - 2)
    fit_data = {'cum_diagnoses': 0.8 * sim.results['cum_diagnoses'].values[days] + 3 * np.sin(days),
                'cum_deaths': sim.results['cum_deaths'].values[days] + 1.0,
                'cum_tests': 1.1 * sim.results['cum_tests'].values[days],
                'new_infections': sim.results['new_infections'].values[days] * 0.9 + 2}
    fit_data['cum_diagnoses'][[4, 9]] = np.nan               # missing observations are skipped
    start = sim['start_day']
    df = pd.DataFrame(fit_data, index=[start + _dt.timedelta(days=int(d)) for d in days])
    sim.data = df
    fit = cv.Fit(sim)
    out['fit/days'] = days
    for k, v in fit_data.items():
        out[f'fit/data/{k}'] = v
    out['fit/mismatch'] = np.float64(fit.mismatch)
    out['fit/keys'] = np.array(json.dumps(list(fit.keys)))
    for k in fit.keys:
        out[f'fit/gofs/{k}'] = np.array(fit.gofs[k])
        out[f'fit/losses/{k}'] = np.array(fit.losses[k])
    fit2 = cv.Fit(sim, keys=['cum_diagnoses', 'new_infections'], weights=dict(new_infections=2.5), use_squared=True, as_scalar='mean')
    out['fit/mismatch_custom'] = np.float64(fit2.mismatch)
    for k in sim.result_keys('variant'):
        out[f'vresults/{k}'] = np.array(sim.results['variant'][k].values)
    for k in cvo.cvd.all_states:
        arr = np.asarray(sim.people[k])
        if name in FULL_PEOPLE:
            out[f'people/{k}'] = arr
        else:
            out[f'people_digest/{k}'] = np.array(digest(arr))
    log = sim.people.infection_log
    lkeys = list(sim.people.contacts.keys())
    lmap = {lk: i for i, lk in enumerate(lkeys)}
    lmap.update(seed_infection=-1, importation=-2)
    vmap = {v: k for k, v in sim['variant_map'].items()}
    out['log/source'] = np.array([-1 if e['source'] is None else e['source'] for e in log], dtype=np.int32)
    out['log/target'] = np.array([e['target'] for e in log], dtype=np.int32)
    out['log/date'] = np.array([e['date'] for e in log], dtype=np.int32)
    out['log/layer'] = np.array([lmap[e['layer']] for e in log], dtype=np.int32)
    out['log/variant'] = np.array([vmap[e['variant']] for e in log], dtype=np.int32)
    out['summary_json'] = np.array(json.dumps({k: float(v) for k, v in sim.summary.items()}))
    return sim, out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])                          # python -m oracle.gen_golden [scenario ...]: regenerate a subset
    for name, spec in scenarios.SCENARIOS.items():
        if only and name not in only:
            continue
        sim, out = run_scenario(name, spec)
        if name == 'baseline20k':
            with open(os.path.join(refenv.REFERENCE, 'tests', 'baseline.json')) as f:
                base = json.load(f)['summary']
            # pop_nabs / pop_protection are float32 np.sum / np.nanmean reductions whose last bits depend on the
            # platform's SIMD summation order (seen here: 2e-8 relative); the reference's own check
            # (cv.diff_sims -> np.isclose, rtol=1e-5) tolerates that.  Everything else must be exact.
            loose = ('pop_nabs', 'pop_protection', 'pop_symp_protection')
            bad = [k for k, v in base.items() if not np.isclose(sim.summary[k], v, rtol=1e-6 if k in loose else 1e-12, atol=0)]
            assert not bad, f'reference run does not reproduce its own baseline.json: {bad}'
            out['baseline_json'] = np.array(json.dumps(base))
            print(f'  baseline20k: all {len(base)} values of the reference tests/baseline.json reproduced')
        path = os.path.join(GOLDEN, f'{name}.npz')
        np.savez_compressed(path, **out)
        print(f'wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB, {len(out)} arrays)')


if __name__ == '__main__':
    main()
