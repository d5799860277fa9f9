'''
TEST INFRASTRUCTURE -- records People.infect calls of the UNMODIFIED reference (/root/reference, Covasim 3.1.7) with the random
draws they consumed, as tests/golden/infect_tape.npz.  Run from the repo root:  python -m oracle.gen_infect_golden

For a handful of People.infect calls of two scenarios (three variants with bed limits; waning with vaccination) it stores
  * the arguments (inds, variant, hosp_max, icu_max, t) and the People arrays of the touched agents BEFORE the call;
  * every array the call's samplers returned, in call order (covasim.utils.sample / binomial_arr are wrapped while infect runs;
    binomial_arr is re-stated as ``u = np.random.random(n); return u < p`` -- the reference's own one-liner, utils.py:302-310 --
    so that the uniforms themselves are on the tape; the stream consumption is unchanged and the scenario's final summary is
    asserted equal to the committed golden run);
  * the same draws re-indexed per agent and per prognosis step, the layout cvb_infect_list_taped takes (16 slots per agent);
  * the People arrays of the touched agents AFTER the call.
tests/test_gpu_ops.py feeds the tape to the CUDA infect kernel and compares the result with the reference's arrays, bit for bit
(NAb levels at 1e-6): a DIRECT check of the prognosis tree against the reference, without the oracle in between.
'''
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import refenv  # noqa: E402
cv = refenv.import_reference()
import covasim.utils as cvu  # noqa: E402
import covasim.people as cvppl  # noqa: E402
import scenarios  # noqa: E402

PRE = ('susceptible', 'naive', 'recovered', 'diagnosed', 'exposed', 'peak_nab', 'nab', 't_nab_event', 'n_breakthroughs', 'n_infections',
       'rel_trans', 'date_recovered', 'date_diagnosed', 'symp_prob', 'severe_prob', 'crit_prob', 'death_prob')
POST = ('susceptible', 'naive', 'recovered', 'diagnosed', 'exposed', 'peak_nab', 't_nab_event', 'n_breakthroughs', 'n_infections', 'rel_trans',
        'exposed_variant', 'date_exposed', 'date_infectious', 'date_symptomatic', 'date_severe', 'date_critical', 'date_recovered', 'date_dead',
        'date_diagnosed', 'dur_exp2inf', 'dur_inf2sym', 'dur_sym2sev', 'dur_sev2crit', 'dur_disease')
PRE = PRE + tuple(k for k in POST if k not in PRE)     # (fields a reinfection leaves as they were, e.g. dur_inf2sym of an asymptomatic case)
WANT = {'variants4k': [(9, None), (20, None), (33, None)], 'baseline20k': [(35, None), (52, None)]}     # (day, -): the call of that day that infected most agents


def record(name, spec, days):
    calls, tape = [], []
    orig_infect, orig_sample, orig_binom = cvppl.People.infect, cvu.sample, cvu.binomial_arr
    state = dict(active=False)

    def sample(*args, **kwargs):
        out = orig_sample(*args, **kwargs)
        if state['active']:
            tape.append(('sample', np.array(out, dtype=np.float64, copy=True)))
        return out

    def binomial_arr(prob_arr):
        u = np.random.random(len(prob_arr))            # the reference's implementation (utils.py:302-310), with the uniforms kept
        if state['active']:
            tape.append(('uniform', u.copy()))
        return u < prob_arr

    def infect(self, inds, hosp_max=None, icu_max=None, source=None, layer=None, variant=0):
        want = self.t in days and len(inds) >= 2 and not state['active']
        if not want:
            return orig_infect(self, inds, hosp_max=hosp_max, icu_max=icu_max, source=source, layer=layer, variant=variant)
        uniq = np.unique(inds)
        pre = {k: np.array(self[k][uniq]) for k in PRE}
        pre['symp_imm'] = np.array(self.symp_imm[variant, uniq])
        pre['sev_imm'] = np.array(self.sev_imm[variant, uniq])
        del tape[:]
        state['active'] = True
        try:
            out = orig_infect(self, inds, hosp_max=hosp_max, icu_max=icu_max, source=source, layer=layer, variant=variant)
        finally:
            state['active'] = False
        post = {k: np.array(self[k][uniq]) for k in POST}
        post['exposed_by_variant'] = np.array(self.exposed_by_variant[variant, uniq])
        calls.append(dict(t=int(self.t), inds=np.array(inds), uniq=uniq, variant=int(variant), hosp_max=bool(hosp_max), icu_max=bool(icu_max),
                          infected=np.array(out), pre=pre, post=post, tape=[(k, a.copy()) for k, a in tape]))
        return out

    cvppl.People.infect, cvu.sample, cvu.binomial_arr = infect, sample, binomial_arr
    try:
        sim = cv.Sim(**scenarios.build(cv, spec))
        sim.run()
    finally:
        cvppl.People.infect, cvu.sample, cvu.binomial_arr = orig_infect, orig_sample, orig_binom
    best = {}
    rank = lambda c: (c['variant'] > 0 and c['t'] >= 20, len(c['infected']))     # later days: a call of an imported variant if there is one
    for c in calls:                                    # per wanted day, the call that infected most agents
        if c['t'] not in best or rank(c) > rank(best[c['t']]):
            best[c['t']] = c
    calls = [best[t] for t in sorted(best)]
    golden = np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz'), allow_pickle=False)
    for k in ('cum_infections', 'cum_deaths', 'cum_severe'):
        assert np.array_equal(sim.results[k].values, golden[f'results/{k}']), f'{name}: wrapping changed the run ({k})'
    return sim, calls


def per_agent_draws(call, use_waning):
    '''
    The tape as float64[n, 16]: row = infected agent (ascending id, the order of People.infect after np.unique), column = prognosis
    step (covasim_b200/csrc/infect.cu): 0 exp2inf, 1 symptomatic?, 2 asym2rec | inf2sym, 3 severe?, 4 mild2rec | sym2sev,
    5 critical?, 6 sev2rec | sev2crit, 7 dies?, 8 crit2rec | crit2die, 9 initial NAb level.  Who took which branch is read off the
    reference's own results (a date is set or it is not); people.py:513-584 gives the order of the draws.
    '''
    inf = call['infected']
    pos = {a: j for j, a in enumerate(call['uniq'])}
    rows = np.array([pos[a] for a in inf])
    P = call['post']
    symp = ~np.isnan(P['date_symptomatic'][rows])
    sev = ~np.isnan(P['date_severe'][rows])
    crit = ~np.isnan(P['date_critical'][rows])
    dead = ~np.isnan(P['date_dead'][rows])
    n = len(inf)
    D = np.full((n, 16), 0.5)
    tape = list(call['tape'])

    def take(kind, mask, col):
        k, arr = tape.pop(0)
        assert k == kind and len(arr) == int(mask.sum()), (k, kind, len(arr), int(mask.sum()))
        D[mask, col] = arr

    every = np.ones(n, dtype=bool)
    take('sample', every, 0)
    take('uniform', every, 1)
    take('sample', ~symp, 2)
    take('sample', symp, 2)
    take('uniform', symp, 3)
    take('sample', symp & ~sev, 4)
    take('sample', sev, 4)
    take('uniform', sev, 5)
    take('sample', sev & ~crit, 6)
    take('sample', crit, 6)
    take('uniform', crit, 7)
    take('sample', crit & ~dead, 8)
    take('sample', dead, 8)
    if use_waning:
        no_prior = ~(call['pre']['nab'][rows] > 0)
        if no_prior.any():
            take('sample', no_prior, 9)
    assert not tape, f'{len(tape)} draws left on the tape'
    return D


def main():
    out = {}
    n_calls = 0
    for name, want in WANT.items():
        spec = scenarios.SCENARIOS[name]
        sim, calls = record(name, spec, {d for d, _ in want})
        assert len(calls) == len(want), (name, [c['t'] for c in calls])
        for c in calls:
            pre = f'{name}/t{c["t"]}/'
            D = per_agent_draws(c, bool(sim['use_waning']))
            out[pre + 'inds'] = c['inds'].astype(np.int32)
            out[pre + 'uniq'] = c['uniq'].astype(np.int32)
            out[pre + 'infected'] = c['infected'].astype(np.int32)
            out[pre + 'draws'] = D
            out[pre + 'args'] = np.array([c['t'], c['variant'], int(c['hosp_max']), int(c['icu_max'])], dtype=np.int32)
            for k, v in c['pre'].items():
                out[pre + 'pre/' + k] = v
            for k, v in c['post'].items():
                out[pre + 'post/' + k] = v
            n_calls += 1
            print(f'{name} day {c["t"]}: {len(c["inds"])} targets, {len(c["infected"])} infected, variant {c["variant"]}, hosp_max {c["hosp_max"]}, '
                  f'{int((~np.isnan(c["post"]["date_dead"])).sum())} will die')
    path = os.path.join(ROOT, 'tests', 'golden', 'infect_tape.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes,', n_calls, 'calls')


if __name__ == '__main__':
    main()
