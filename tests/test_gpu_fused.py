'''
The fused day pipeline (covasim_b200/csrc/day_fused.cu, cvb_run_days): whole blocks of days in one C-ABI call -- five launches per
day, a packed per-agent state word instead of sixteen flag arrays -- against the oracle (Philox mode) and against the per-step
entry points.  Bit-exact for flags / dates / counters / infection log; 1e-6 relative for NAb and immunity floats.
'''
import numpy as np
import pytest

import parity
import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


# scenario -> at least this share of the days must have gone through cvb_run_days
BLOCK_SCENARIOS = {'random2k_nowaning': 1.0, 'default20k': 1.0, 'dynamic2k': 1.0, 'hybrid3k': 0.85, 'baseline20k': 0.9, 'variants4k': 0.0,
                   'sequence3k': 0.0, 'clip3k': 0.9, 'fracsus2k': 1.0}


@pytest.mark.parametrize('name', sorted(BLOCK_SCENARIOS))
def test_run_in_blocks_matches_oracle(cv, name):
    ''' sim.run(): every stretch of days without a host decision is ONE cvb_run_days call; People arrays, results and log vs the oracle '''
    sim, orc = parity.build_pair(cv, name)
    sim.run()
    orc.run()
    assert sim.fused_days >= BLOCK_SCENARIOS[name] * sim.npts - 1e-9, f'only {sim.fused_days} of {sim.npts} days were fused'
    parity.compare_people(sim, orc, 'at the end')
    parity.compare_results(sim, orc)
    parity.compare_log(sim, orc)


C2_SMALL = dict(
    # the BASELINE C2 recipe (bench.py) at 40k agents: test_prob + contact_tracing with their defaults, all four hybrid layers
    pars=dict(pop_size=40000, pop_type='hybrid', n_days=70, pop_infected=200, rand_seed=1, verbose=0),
    interventions=[('test_prob', dict(symp_prob=0.1, asymp_prob=0.01, start_day=20)), ('contact_tracing', dict(trace_probs=0.3, start_day=30))])


def test_blocks_of_seven_days_with_state_checks(cv):
    ''' run(until=...) in weekly blocks: People arrays against the oracle and the packed state words against the arrays at every boundary '''
    sim, orc = parity.build_pair(cv, spec=C2_SMALL)
    sim.set_seed()
    orc.rng.set_seed(orc.pars['rand_seed'])
    first = True
    while not sim.complete:
        until = min(sim.t + 7, sim.npts)
        sim.run(until=until, reset_seed=first)
        first = False
        while orc.t < until:
            orc.step()
        if not sim.complete:
            parity.compare_people(sim, orc, f'after day {until - 1}')
            parity.check_packed_state(sim, f'after day {until - 1}')
    orc.finalize()
    assert sim.fused_days == sim.npts
    parity.compare_people(sim, orc, 'at the end')
    parity.compare_results(sim, orc)
    parity.compare_log(sim, orc)


def test_fused_equals_per_step(cv):
    ''' The same sim through cvb_run_days and through the per-step entry points: every People array and result series identical '''
    a = cv.Sim(**scenarios.build(cv, C2_SMALL), fused=True).run()
    b = cv.Sim(**scenarios.build(cv, C2_SMALL), fused=False).run()
    assert a.fused_days == a.npts and b.fused_days == 0
    for k in a.people.keys():
        x, y = a.people.to_numpy(k), b.people.to_numpy(k)
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == 'f')), k
    for k in a.result_keys():
        assert np.array_equal(a.results[k].values, b.results[k].values, equal_nan=True), k
    la, lb = a.infection_log, b.infection_log
    for k in la:
        assert np.array_equal(la[k], lb[k]), k


def test_host_writes_between_runs_are_seen(cv):
    ''' People arrays written from Python between two run() calls: the packed state is rebuilt (same as the per-step path) '''
    def go(fused):
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS['random2k_nowaning']), fused=fused)
        sim.run(until=10)
        sim.people.quarantined[100:140] = True                      # a user puts forty people into quarantine by hand
        sim.people.date_end_quarantine[100:140] = 15.0
        sim.run()
        return sim
    a, b = go(True), go(False)
    assert a.fused_days == a.npts
    for k in ('n_quarantined', 'new_infections', 'n_exposed', 'cum_deaths'):
        assert np.array_equal(a.results[k].values, b.results[k].values), k
    assert a.results['n_quarantined'].values[10] >= 40


@pytest.mark.parametrize('name', ['rescale3k', 'clip3k', 'hybrid3k', 'variants4k'])
def test_restore_then_run_again(cv, name):
    '''
    snapshot() at day 0, run, restore(), run again: every result series identical -- including the state a run leaves outside the
    People arrays (rescale vector, the edges clip_edges holds back, pending second doses), which restore must rewind too.
    '''
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
    sim.initialize()
    snap = sim.snapshot()
    sim.run()
    first = {k: sim.results[k].values.copy() for k in sim.result_keys()}
    layers = {lk: len(l) for lk, l in sim.people.contacts.items()}
    sim.restore(snap)
    sim.run()
    for k, v in first.items():
        assert np.array_equal(v, sim.results[k].values, equal_nan=True), k
    assert layers == {lk: len(l) for lk, l in sim.people.contacts.items()}


def test_restore_mid_run(cv):
    ''' A snapshot taken in the middle of a run (second doses pending, population partly rescaled) resumes to the same end '''
    spec = scenarios.SCENARIOS['rescale3k']
    sim = cv.Sim(**scenarios.build(cv, spec))
    sim.run(until=16)
    snap = sim.snapshot()
    sim.run(reset_seed=False)
    first = {k: sim.results[k].values.copy() for k in sim.result_keys()}
    sim.restore(snap)
    sim.run(reset_seed=False)
    for k, v in first.items():
        assert np.array_equal(v[16:], sim.results[k].values[16:], equal_nan=True), k


@pytest.mark.parametrize('fused', [True, False])
def test_restore_mid_run_with_pending_quarantine_requests(cv, fused):
    '''
    hybrid3k traces contacts with delays of up to two days, so in the middle of the run quarantine requests are waiting for their start day
    (reference people.py:620-640 _pending_quarantine; here the library's request ring).  A snapshot keeps them and restore puts them back:
    the resumed run ends exactly like the uninterrupted one, People arrays included.
    '''
    spec = scenarios.SCENARIOS['hybrid3k']
    sim = cv.Sim(**scenarios.build(cv, spec), fused=fused)
    sim.run(until=21)
    snap = sim.snapshot()
    assert snap['quar_horizon'] == 3 and len(snap['quar_ring']) >= 1 and sum(int((h >= 0).sum()) for h in snap['quar_ring'].values()) > 10
    sim.run(reset_seed=False)
    first = {k: sim.results[k].values.copy() for k in sim.result_keys()}
    people = {k: sim.people.to_numpy(k).copy() for k in ('quarantined', 'date_quarantined', 'date_end_quarantine', 'known_contact', 'date_exposed')}
    sim.restore(snap)
    sim.run(reset_seed=False)
    for k, v in first.items():
        assert np.array_equal(v[21:], sim.results[k].values[21:], equal_nan=True), k
    for k, v in people.items():
        assert np.array_equal(v, sim.people.to_numpy(k), equal_nan=(v.dtype.kind == 'f')), k


def test_multisim_lockstep_members_equal_solo_runs(cv):
    ''' MultiSim advances its members through cvb_run_days_multi (one stream per member): each member identical to its solo run '''
    base = cv.Sim(**scenarios.build(cv, C2_SMALL, pop_size=12000, n_days=40, pop_infected=120))
    msim = cv.MultiSim(base, n_runs=5)
    msim.run(keep_people=True)
    assert all(s.fused_days == s.npts for s in msim.sims)
    for i, member in enumerate(msim.sims):
        solo = cv.Sim(**scenarios.build(cv, C2_SMALL, pop_size=12000, n_days=40, pop_infected=120, rand_seed=base['rand_seed'] + i), fused=False).run()
        for k in solo.result_keys():
            assert np.array_equal(solo.results[k].values, member.results[k].values, equal_nan=True), (i, k)
            assert np.array_equal(solo.results[k].values, msim.member_results[i][k], equal_nan=True), (i, k)
        for k in ('exposed', 'date_exposed', 'date_recovered', 'nab', 'quarantined'):
            x, y = solo.people.to_numpy(k), member.people.to_numpy(k)
            assert np.array_equal(x, y, equal_nan=(x.dtype.kind == 'f')), (i, k)


class late_quarantine:
    ''' Test-only intervention: on `day`, ask for a quarantine that starts `ahead` days later (People.schedule_quarantine, people.py:620-640) '''
    def __init__(self, day, ahead):
        self.day, self.ahead, self.initialized = day, ahead, False

    def initialize(self, sim):
        self.initialized = True

    def __call__(self, sim):
        if sim.t == self.day:
            sim.people.schedule_quarantine(np.arange(50, 400, 7), start_date=sim.t + self.ahead, period=6)


def test_quarantine_ring_grows_without_losing_requests(cv):
    ''' A request beyond the ring's horizon re-sizes it in the middle of a run; the requests already pending (delayed tracing) survive '''
    def go(pre_size):
        spec = scenarios.SCENARIOS['hybrid3k']                      # contact tracing with trace_time up to 2 days
        kw = scenarios.build(cv, spec)
        kw['interventions'] = kw['interventions'] + [late_quarantine(14, 5)]
        sim = cv.Sim(**kw)
        sim.initialize()
        if pre_size:
            sim._set_quar_horizon(8)
        return sim.run()
    a, b = go(True), go(False)
    assert b._quar_horizon >= 6
    for k in ('new_quarantined', 'n_quarantined', 'new_infections', 'cum_diagnoses'):
        assert np.array_equal(a.results[k].values, b.results[k].values), k
    assert a.results['new_quarantined'].values[19] >= 40


def test_layer_update_with_a_fraction(cv):
    ''' Layer.update(people, frac < 1) (reference base.py:1849-1876): exactly round(E * frac) edges get new endpoints and weight 1 '''
    sim = cv.Sim(pop_size=5000, n_days=10, pop_infected=50, rand_seed=9, verbose=0, dynam_layer=dict(a=1))
    sim.initialize()
    layer = sim.people.contacts['a']
    layer['beta'][:] = 0.5
    before = layer.to_numpy()
    E = len(layer)
    sim.t = 3
    layer.update(sim.people, frac=0.25)
    after = layer.to_numpy()
    touched = after['beta'] == 1.0
    assert int(touched.sum()) == int(np.round(E * 0.25))
    same = (before['p1'] == after['p1']) & (before['p2'] == after['p2'])
    assert same[~touched].all() and (~same[touched]).mean() > 0.99
    assert after['p1'].min() >= 0 and after['p1'].max() < 5000 and after['p2'].max() < 5000
    full = cv.Sim(pop_size=5000, n_days=10, pop_infected=50, rand_seed=9, verbose=0, dynam_layer=dict(a=1))
    full.initialize()
    full.t = 3
    full.people.contacts['a'].update(full.people)                      # an edge gets the same endpoints in a partial and a full regeneration
    f = full.people.contacts['a'].to_numpy()
    assert np.array_equal(f['p1'][touched], after['p1'][touched]) and np.array_equal(f['p2'][touched], after['p2'][touched])


@pytest.mark.parametrize('day', [0, 22])
def test_compact_snapshot_restores_every_byte(cv, day):
    '''
    Sim.snapshot keeps the People arena in compact form too (per array: one value + exceptions, or dense) and Sim.restore sends
    that (cvb_restore_compact + the dense arrays): every array must come back exactly as a plain copy of the arena would leave it,
    at day 0 (almost everything is one value) and in the middle of an epidemic with vaccination (many dense arrays).
    '''
    import torch
    spec = scenarios.SCENARIOS['baseline20k']
    sim = cv.Sim(**scenarios.build(cv, spec))
    sim.initialize()
    if day:
        sim.run(until=day)
    snap = sim.snapshot(pinned=True)
    c = snap['arena_compact']
    assert c['n_seg'] > 20 and c['h2d_bytes'] < 0.6 * snap['arena'].numel()
    want = {k: sim.people.to_numpy(k).copy() for k in sim.people.keys()}
    sim.run(until=day + 15, reset_seed=False)
    sim.people._arena.fill_(0x5A)                                   # nothing may survive from the state being replaced
    sim.restore(snap)
    torch.cuda.synchronize()
    for k, w in want.items():
        got = sim.people.to_numpy(k)
        assert np.array_equal(got, w, equal_nan=(w.dtype.kind == 'f')), f'day {day}: {k} differs after a compact restore'
    plain = dict(snap)
    del plain['arena_compact']
    a = sim.run().summary
    sim.restore(plain)
    b = sim.run().summary
    assert all(a[k] == b[k] or (a[k] != a[k] and b[k] != b[k]) for k in a), 'compact and plain restore lead to different runs'


def test_bulk_copy_staged_dense_pass_gives_the_same_run(cv):
    '''
    CVB_DENSE_VARIANT=3 selects the dense edge pass whose tiles are staged by cp.async.bulk + mbarrier (edge_pass_tma_kernel; slower than
    the default, kept for profiling): same epidemic, agent for agent.  The variant is read once per process, hence the subprocess.
    '''
    import json
    import os
    import subprocess
    import sys
    code = ("import sys, json, hashlib, numpy as np; sys.path.insert(0, 'tests'); import scenarios, covasim_b200 as cv\n"
            "out = {}\n"
            "for name, kw in (('dynamic2k', {}), ('hybrid3k', dict(use_adjacency=False)), ('variants4k', dict(use_adjacency=False))):\n"
            "    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]), **kw); sim.run()\n"
            "    out[name] = [float(sim.summary['cum_infections']), hashlib.sha256(sim.people.to_numpy('date_exposed').tobytes()).hexdigest()]\n"
            "print(json.dumps(out))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runs = []
    for variant in ('0', '3'):
        env = dict(os.environ, CVB_DENSE_VARIANT=variant)
        r = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert runs[0] == runs[1]
    assert all(v[0] > 100 for v in runs[0].values())
