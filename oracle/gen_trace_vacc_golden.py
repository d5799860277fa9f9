'''
TEST INFRASTRUCTURE -- records contact_tracing.apply and BaseVaccination.vaccinate calls of the UNMODIFIED reference (/root/reference)
with the random draws they consumed, as tests/golden/trace_tape.npz and tests/golden/vacc_tape.npz.
Run from the repo root:  python -m oracle.gen_trace_vacc_golden

Tracing ('hybrid3k': per-layer trace probabilities 1 / 0.5 / 0.5 / 0.2 and trace times 0 / 1 / 1 / 2; days 12 and 25): the People arrays
the intervention reads BEFORE the call, the quarantine requests pending before it, the uniforms binomial_filter consumed per traced layer
(covasim.utils.binomial_filter is wrapped with its own one-line definition, utils.py:383-396, so that the uniforms are kept; they are
stored per (layer, contact) as float64[n_layers][N], -1 where the reference drew nothing), and AFTER the call known_contact,
date_known_contact and the pending requests per start day as "latest requested end day per agent".
Vaccination ('variants4k': first doses day 5, second doses day 26, a one-dose booster day 32): the agents handed to
BaseVaccination.vaccinate, the arrays it reads and writes before / after, the intervention's own dose counts, and the initial NAb samples
(covasim.utils.sample inside update_peak_nab, immunity.py:178) per agent.
tests/test_gpu_ops.py feeds the tapes to the CUDA kernels (cvb_contact_tracing_taped, cvb_vaccinate_taped) and compares with the
reference's arrays bit for bit (the peak NAb level, a float64 2**x, at 1e-6).
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
import covasim.base as cvb  # noqa: E402
import covasim.immunity as cvimm  # noqa: E402
import scenarios  # noqa: E402

TRACE_PRE = ('diagnosed', 'date_diagnosed', 'dead', 'known_contact', 'date_known_contact', 'quarantined', 'date_quarantined', 'date_end_quarantine',
             'date_pos_test', 'date_tested')
TRACE_DAYS = {12, 25}
VACC_PRE = ('dead', 'vaccinated', 'doses', 'vaccine_source', 'date_vaccinated', 'nab', 'peak_nab', 't_nab_event')
VACC_DAYS = {5, 26, 32}


def pending(people, n, days):
    ''' people._pending_quarantine[day] -> float32[n]: the latest requested end day of every agent (-1 none) '''
    out = {}
    for d in days:
        arr = np.full(n, -1.0, dtype=np.float32)
        for ind, end in people._pending_quarantine.get(d, []):
            arr[ind] = max(arr[ind], end)
        out[d] = arr
    return out


def check_run(sim, name, keys):
    golden = np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz'))
    for k in keys:
        assert np.array_equal(sim.results[k].values, golden[f'results/{k}']), f'{name}: wrapping changed the run ({k})'


def record_tracing():
    name = 'hybrid3k'
    calls, tape, state = [], [], dict(active=False, layer=None)
    orig_apply, orig_filter, orig_find = cvi.contact_tracing.apply, cvu.binomial_filter, cvb.Layer.find_contacts

    def binomial_filter(prob, arr):
        u = np.random.random(len(arr))                          # utils.py:383-396
        if state['active']:
            tape.append((state['layer'], np.array(arr), u.copy()))
        return arr[(u < prob).nonzero()[0]]

    def find_contacts(self, inds, as_array=True):
        if state['active']:
            state['layer'] = [lk for lk, layer in state['sim'].people.contacts.items() if layer is self][0]
        return orig_find(self, inds, as_array=as_array)

    def apply(self, sim):
        if sim.t not in TRACE_DAYS:
            return orig_apply(self, sim)
        P, n, t = sim.people, sim['pop_size'], sim.t
        pre = {k: np.array(P[k]) for k in TRACE_PRE}
        pend_pre = pending(P, n, range(t, t + 3))
        del tape[:]
        state.update(active=True, sim=sim)
        try:
            out = orig_apply(self, sim)
        finally:
            state['active'] = False
        calls.append(dict(t=int(t), pre=pre, pend_pre=pend_pre, pend_post=pending(P, n, range(t, t + 3)),
                          post={k: np.array(P[k]) for k in ('known_contact', 'date_known_contact')}, tape=list(tape)))
        return out

    cvi.contact_tracing.apply, cvu.binomial_filter, cvb.Layer.find_contacts = apply, binomial_filter, find_contacts
    try:
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
        sim.run()
    finally:
        cvi.contact_tracing.apply, cvu.binomial_filter, cvb.Layer.find_contacts = orig_apply, orig_filter, orig_find
    check_run(sim, name, ('cum_infections', 'cum_diagnoses', 'cum_quarantined'))
    lkeys = list(sim.people.contacts.keys())
    n = sim['pop_size']
    out = {}
    for c in calls:
        T = np.full((len(lkeys), n), -1.0)
        for lk, arr, u in c['tape']:
            T[lkeys.index(lk), arr] = u
        pre = f'{name}/t{c["t"]}/'
        out[pre + 'tape'] = T
        for k, v in c['pre'].items():
            out[pre + 'pre/' + k] = v
        for k, v in c['post'].items():
            out[pre + 'post/' + k] = v
        for d, v in c['pend_pre'].items():
            out[pre + f'pend_pre/{d}'] = v
        for d, v in c['pend_post'].items():
            out[pre + f'pend_post/{d}'] = v
        n_new = int((c['post']['known_contact'] & ~c['pre']['known_contact']).sum())
        print(f'tracing day {c["t"]}: draws per layer {[len(u) for _, _, u in c["tape"]]}, {n_new} newly known contacts, '
              f'requests {[int((v >= 0).sum()) for v in c["pend_post"].values()]} (before: {[int((v >= 0).sum()) for v in c["pend_pre"].values()]})')
    path = os.path.join(ROOT, 'tests', 'golden', 'trace_tape.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


def record_vaccination():
    name = 'variants4k'
    calls, tape, state = [], [], dict(active=False)
    orig_vacc, orig_sample = cvi.BaseVaccination.vaccinate, cvu.sample

    def sample(*args, **kwargs):
        out = orig_sample(*args, **kwargs)
        if state['active']:
            tape.append(np.array(out, dtype=np.float64, copy=True))
        return out

    def vaccinate(self, sim, vacc_inds, t=None):
        if sim.t not in VACC_DAYS or not len(vacc_inds):
            return orig_vacc(self, sim, vacc_inds, t=t)
        P = sim.people
        pre = {k: np.array(P[k]) for k in VACC_PRE}
        pre['iv_doses'] = np.array(self.doses)
        flows = {k: P.flows[k] for k in ('new_doses', 'new_vaccinated')}
        del tape[:]
        state['active'] = True
        try:
            out = orig_vacc(self, sim, vacc_inds, t=t)
        finally:
            state['active'] = False
        post = {k: np.array(P[k]) for k in VACC_PRE}
        post['iv_doses'] = np.array(self.doses)
        calls.append(dict(t=int(sim.t), label=self.label, inds=np.array(vacc_inds), given=np.array(out), pre=pre, post=post, tape=[a.copy() for a in tape],
                          flows={k: P.flows[k] - flows[k] for k in flows}))
        return out

    cvi.BaseVaccination.vaccinate, cvu.sample = vaccinate, sample
    cvimm.cvu.sample = sample
    try:
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
        sim.run()
    finally:
        cvi.BaseVaccination.vaccinate, cvu.sample = orig_vacc, orig_sample
    check_run(sim, name, ('cum_infections', 'cum_doses', 'cum_vaccinated'))
    n = sim['pop_size']
    out = {}
    for c in calls:
        given = c['given']
        fresh = given[~(c['pre']['nab'][given] > 0)]            # immunity.py:170-178: the sample covers the agents without prior antibodies, in order
        T = np.full(n, np.nan)
        if len(fresh):
            assert len(c['tape']) == 1 and len(c['tape'][0]) == len(fresh), (len(c['tape']), len(fresh))
            T[fresh] = c['tape'][0]
        else:
            assert not c['tape']
        pre = f'{name}/t{c["t"]}/{c["label"]}/'
        out[pre + 'tape'] = T
        out[pre + 'inds'] = c['inds'].astype(np.int32)
        out[pre + 'flows'] = np.array([c['flows']['new_doses'], c['flows']['new_vaccinated']], dtype=np.float64)
        for k, v in c['pre'].items():
            out[pre + 'pre/' + k] = v
        for k, v in c['post'].items():
            out[pre + 'post/' + k] = v
        print(f'vaccination day {c["t"]} ({c["label"]}): {len(c["inds"])} selected, {len(given)} doses given, {len(fresh)} without prior antibodies, flows {c["flows"]}')
    path = os.path.join(ROOT, 'tests', 'golden', 'vacc_tape.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    record_tracing()
    record_vaccination()
