'''
TEST INFRASTRUCTURE -- records test_prob.apply / People.test calls of the UNMODIFIED reference (/root/reference) with the uniforms they
consumed, as tests/golden/test_tape.npz.  Run from the repo root:  python -m oracle.gen_test_golden

For two days of the 'hybrid3k' scenario (test_prob with quarantine-specific probabilities, sensitivity 0.9, loss 0.1, one day of delay)
it stores the People arrays the intervention reads BEFORE the call, the three uniform arrays the call consumed -- binomial_arr over the
whole population (interventions.py:975), n_binomial(sensitivity) over the infectious who tested and n_binomial(1 - loss_prob) over the
undiagnosed positives (people.py:603-611); covasim.utils.binomial_arr / n_binomial are wrapped with their own one-line definitions so that
the uniforms are kept -- re-indexed per agent as float64[N][3], and the arrays the call writes AFTER it.
tests/test_gpu_ops.py feeds the tape to the CUDA test_prob kernel (cvb_test_prob_taped) and compares with the reference's arrays bit for bit.
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
import covasim.interventions as cvi  # noqa: E402
import scenarios  # noqa: E402

PRE = ('symptomatic', 'diagnosed', 'infectious', 'quarantined', 'tested', 'date_quarantined', 'date_end_quarantine', 'date_diagnosed', 'date_tested',
       'date_pos_test')
POST = ('tested', 'date_tested', 'date_diagnosed', 'date_pos_test')
DAYS = {12, 25}


def main():
    name = 'hybrid3k'
    calls, tape, state = [], [], dict(active=False)
    orig_apply, orig_binom, orig_nbinom = cvi.test_prob.apply, cvu.binomial_arr, cvu.n_binomial

    def binomial_arr(prob_arr):
        u = np.random.random(len(prob_arr))                      # utils.py:302-310
        if state['active']:
            tape.append(u.copy())
        return u < prob_arr

    def n_binomial(prob, n):
        u = np.random.random(n)                                  # utils.py:313-324
        if state['active']:
            tape.append(u.copy())
        return u < prob

    def apply(self, sim):
        if sim.t not in DAYS:
            return orig_apply(self, sim)
        P = sim.people
        pre = {k: np.array(P[k]) for k in PRE}
        del tape[:]
        state['active'] = True
        try:
            out = orig_apply(self, sim)
        finally:
            state['active'] = False
        calls.append(dict(t=int(sim.t), pre=pre, post={k: np.array(P[k]) for k in POST}, tape=[a.copy() for a in tape], test_inds=np.array(out)))
        return out

    cvi.test_prob.apply, cvu.binomial_arr, cvu.n_binomial = apply, binomial_arr, n_binomial
    try:
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
        sim.run()
    finally:
        cvi.test_prob.apply, cvu.binomial_arr, cvu.n_binomial = orig_apply, orig_binom, orig_nbinom
    golden = np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz'))
    for k in ('cum_infections', 'cum_diagnoses', 'cum_tests'):
        assert np.array_equal(sim.results[k].values, golden[f'results/{k}']), f'wrapping changed the run ({k})'
    out = {}
    n = sim['pop_size']
    for c in calls:
        u_test, u_sens, u_loss = c['tape']
        assert len(u_test) == n
        tested = np.nonzero(u_test < 2)[0][np.isin(np.arange(n), c['test_inds'])]          # = test_inds (ascending)
        assert np.array_equal(tested, np.unique(c['test_inds']))
        inf_tested = tested[c['pre']['infectious'][tested]]
        assert len(u_sens) == len(inf_tested)
        pos = inf_tested[u_sens < 0.9]
        undiag = pos[np.isnan(c['pre']['date_diagnosed'][pos])]
        assert len(u_loss) == len(undiag)
        T = np.full((n, 3), 0.5)
        T[:, 0] = u_test
        T[inf_tested, 1] = u_sens
        T[undiag, 2] = u_loss
        pre = f'{name}/t{c["t"]}/'
        out[pre + 'tape'] = T
        out[pre + 'n_tests'] = np.int64(len(tested))
        for k, v in c['pre'].items():
            out[pre + 'pre/' + k] = v
        for k, v in c['post'].items():
            out[pre + 'post/' + k] = v
        print(f'day {c["t"]}: {len(tested)} tested, {len(inf_tested)} of them infectious, {len(undiag)} undiagnosed positives, '
              f'{int(np.sum(c["post"]["date_pos_test"] == c["t"]))} new positive results')
    path = os.path.join(ROOT, 'tests', 'golden', 'test_tape.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
