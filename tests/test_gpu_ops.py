'''
GPU parity of the stateless operators (the device forms of the reference's four Numba kernels) against
(a) the inputs/outputs recorded from the unmodified reference (tests/golden) and (b) the oracle on
seeded synthetic inputs, including empty, ragged and unaligned sizes.  All through the C ABI.
'''
import hashlib

import numpy as np
import pytest

from oracle import cvoracle as cvo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


_layers = {}


def layers_of(name):
    ''' The scenario's static contact layers, regenerated from the seed by the oracle's population builder '''
    if name not in _layers:
        import scenarios
        sim = cvo.OracleSim(**scenarios.build(cvo, scenarios.SCENARIOS[name]), rng='mt')
        sim.initialize()
        _layers[name] = sim.contacts
    return _layers[name]


@pytest.mark.parametrize('name,day', [('hybrid3k', 12), ('hybrid3k', 25), ('variants4k', 20)])
def test_reference_kernel_vectors(cv, golden, name, day):
    g = golden(name)
    pre = f'k{day}/'
    vl = cv.ops.compute_viral_load(int(g[pre + 'vl/t']), g[pre + 'vl/date_inf'], g[pre + 'vl/date_rec'], g[pre + 'vl/date_dead'],
                                   float(g[pre + 'vl/frac_time']), float(g[pre + 'vl/load_ratio']), float(g[pre + 'vl/high_cap']))
    assert np.array_equal(vl.cpu().numpy(), g[pre + 'vl/out'])
    contacts = layers_of(name)
    by_len = {len(l['p1']): l for l in contacts.values()}
    n_calls = int(g[pre + 'n_calls'])
    assert n_calls > 0
    for j in range(n_calls):
        a = {k.split('/')[-1]: g[k] for k in g.files if k.startswith(f'{pre}ts{j}/')}
        rt, rs = cv.ops.compute_trans_sus(a['rel_trans'], a['rel_sus'], a['inf'], a['sus'], float(a['beta_layer']), a['viral_load'],
                                          a['symp'], a['iso'], a['quar'], float(a['asymp_factor']), float(a['iso_factor']),
                                          float(a['quar_factor']), a['immunity_factors'])
        assert np.array_equal(rt.cpu().numpy(), a['out_trans'])
        assert np.array_equal(rs.cpu().numpy(), a['out_sus'])
        c = {k.split('/')[-1]: g[k] for k in g.files if k.startswith(f'{pre}ci{j}/')}
        layer = by_len[int(c['n_edges'])]
        assert digest(layer['p1']) == str(c['p1_digest'])
        u = np.concatenate([c['u_dir1'], c['u_dir2']])
        src, tgt, n_draws = cv.ops.compute_infections(float(c['beta']), layer['p1'], layer['p2'], layer['beta'], c['rel_trans'], c['rel_sus'],
                                                      lambda n: u[:n])
        assert n_draws == (len(c['u_dir1']), len(c['u_dir2']))          # consumes exactly the reference's draws
        assert np.array_equal(src.cpu().numpy(), c['out_src'])
        assert np.array_equal(tgt.cpu().numpy(), c['out_tgt'])
    for lk, layer in contacts.items():
        out = cv.ops.find_contacts(layer['p1'], layer['p2'], g[f'{pre}fc/{lk}/inds'], n_agents=len(g[pre + 'vl/out']))
        assert np.array_equal(out.cpu().numpy(), g[f'{pre}fc/{lk}/out'])


@pytest.mark.parametrize('n_agents,n_edges', [(50, 0), (50, 1), (1000, 1023), (1000, 1024), (1000, 1025), (30000, 200003), (200000, 3000001)])
def test_compute_infections_vs_oracle(cv, n_agents, n_edges):
    rng = np.random.RandomState(n_edges + 1)
    p1 = np.sort(rng.randint(0, n_agents, n_edges)).astype(np.int32)
    p2 = rng.randint(0, n_agents, n_edges).astype(np.int32)
    lb = np.where(rng.random_sample(n_edges) < 0.1, 0.0, 1.0).astype(np.float32) * rng.random_sample(n_edges).astype(np.float32)
    infectious = rng.random_sample(n_agents) < 0.15
    rel_trans = np.where(infectious, rng.gamma(0.5, 2.0, n_agents), 0).astype(np.float32)
    rel_sus = np.where(~infectious & (rng.random_sample(n_agents) < 0.7), rng.random_sample(n_agents), 0).astype(np.float32)
    drawn = []
    stream = np.random.RandomState(7)

    def draw_oracle(direction, edges):
        u = stream.random_sample(len(edges))
        drawn.append(u)
        return u
    beta = 0.3
    src_o, tgt_o = cvo.compute_infections(beta, p1, p2, lb, rel_trans, rel_sus, draw_oracle)
    u = np.concatenate(drawn)
    src, tgt, n_draws = cv.ops.compute_infections(beta, p1, p2, lb, rel_trans, rel_sus, lambda n: u[:n])
    assert n_draws == (len(drawn[0]), len(drawn[1]))
    assert np.array_equal(src.cpu().numpy(), src_o)
    assert np.array_equal(tgt.cpu().numpy(), tgt_o)
    if n_edges:
        who = np.nonzero(infectious)[0][:: 7]
        out = cv.ops.find_contacts(p1, p2, who, n_agents=n_agents)
        assert np.array_equal(out.cpu().numpy(), cvo.find_contacts(p1, p2, who))


def test_unaligned_edge_arrays(cv):
    ''' Views that are not 16-byte aligned take the scalar-load path and must give the same answer '''
    import torch
    rng = np.random.RandomState(3)
    n, e = 5000, 40001
    p1 = np.sort(rng.randint(0, n, e + 1)).astype(np.int32)
    p2 = rng.randint(0, n, e + 1).astype(np.int32)
    lb = np.ones(e + 1, dtype=np.float32)
    rel_trans = np.where(rng.random_sample(n) < 0.2, 1.5, 0).astype(np.float32)
    rel_sus = np.where(rel_trans == 0, 0.8, 0).astype(np.float32)
    d_p1, d_p2, d_lb = (torch.as_tensor(x).cuda()[1:] for x in (p1, p2, lb))      # offset by one element = 4 bytes
    stream = np.random.RandomState(11)
    drawn = []
    src_o, tgt_o = cvo.compute_infections(0.4, p1[1:], p2[1:], lb[1:], rel_trans, rel_sus, lambda d, ed: drawn.append(stream.random_sample(len(ed))) or drawn[-1])
    u = np.concatenate(drawn)
    src, tgt, _ = cv.ops.compute_infections(0.4, d_p1, d_p2, d_lb, rel_trans, rel_sus, lambda k: u[:k])
    assert np.array_equal(src.cpu().numpy(), src_o) and np.array_equal(tgt.cpu().numpy(), tgt_o)


@pytest.mark.parametrize('n', [1, 31, 1023, 1024, 1025, 100003, 5000000])
def test_true_indices(cv, n):
    import ctypes as C
    import torch
    rng = np.random.RandomState(n)
    flags = rng.random_sample(n) < (0.5 if n < 2000 else 0.03)
    ws = cv.ops.Workspace(n)
    d = torch.as_tensor(flags).cuda()
    out = torch.empty(n, dtype=torch.int32, device='cuda')
    n_out = C.c_int64(0)
    cv._capi.call('cvb_true_indices', ws.handle, d.data_ptr(), n, out.data_ptr(), C.byref(n_out), None)
    assert np.array_equal(out[:n_out.value].cpu().numpy(), np.nonzero(flags)[0])
    assert np.array_equal(cv.true(d).cpu().numpy(), np.nonzero(flags)[0])


def test_viral_load_edge_cases(cv):
    ''' NaN dates, zero-length infectious periods, negative days (historical interventions run with t < 0) '''
    nan = np.nan
    d_inf = np.array([nan, 3, 3, 3, 10, 0, -5], dtype=np.float32)
    d_rec = np.array([nan, 3, 20, nan, 12, 40, 9], dtype=np.float32)
    d_dead = np.array([nan, nan, nan, 15, nan, nan, nan], dtype=np.float32)
    for t in (-3, 0, 4, 11, 30):
        got = cv.ops.compute_viral_load(t, d_inf, d_rec, d_dead, 0.3, 2.0, 4.0).cpu().numpy()
        want = cvo.compute_viral_load(t, d_inf, d_rec, d_dead, 0.3, 2.0, 4.0)
        assert np.array_equal(got, want)


# ---- People.infect: the CUDA prognosis tree against arrays recorded from the reference, fed the reference's own draws ----------
INFECT_CALLS = [('variants4k', 9), ('variants4k', 20), ('variants4k', 33), ('baseline20k', 35), ('baseline20k', 52)]
_infect_sims = {}


@pytest.mark.parametrize('name,day', INFECT_CALLS)
def test_infect_kernel_against_reference_tape(cv, name, day):
    '''
    tests/golden/infect_tape.npz (oracle/gen_infect_golden.py): a People.infect call of the unmodified reference -- the targets, the
    touched agents' arrays before and after, and every draw the call consumed, re-indexed per agent and prognosis step.  The CUDA
    infect kernel is given the same agents, the same state and the same draws (cvb_infect_list_taped) and must produce the reference's
    arrays: flags, dates, durations, counters, rel_trans bit for bit; the peak NAb level (a float64 2**x) at 1e-6.
    '''
    import ctypes as C
    import os
    import torch
    import scenarios
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'infect_tape.npz'))
    pre = f'{name}/t{day}/'
    t, variant, hosp_max, icu_max = (int(x) for x in g[pre + 'args'])
    if name not in _infect_sims:
        sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS[name]))
        sim.initialize()
        _infect_sims[name] = (sim, sim.snapshot(pinned=False))
    sim, snap = _infect_sims[name]
    sim.restore(snap)
    P, dev = sim.people, sim.people.device
    uniq = torch.as_tensor(g[pre + 'uniq'].astype(np.int64), device=dev)
    for key in g.files:                                               # the touched agents' state before the call
        if key.startswith(pre + 'pre/'):
            k = key.split('/')[-1]
            vals = torch.as_tensor(g[key], device=dev)
            if k in ('symp_imm', 'sev_imm'):
                P[k][variant, uniq] = vals.to(P[k].dtype)
            else:
                P[k][uniq] = vals.to(P[k].dtype)
    inds, infected, D = g[pre + 'inds'], g[pre + 'infected'], g[pre + 'draws']
    tape = np.full((len(inds), 16), 0.5)
    for row, a in enumerate(infected):                                # the kernel looks a draw up by the agent's FIRST position in the list
        tape[int(np.nonzero(inds == a)[0][0])] = D[row]
    d_inds = torch.as_tensor(inds.astype(np.int32), device=dev).contiguous()
    d_tape = torch.as_tensor(tape, dtype=torch.float64, device=dev).contiguous()
    sim.t = t
    sim._push_pars()
    cv._capi.call('cvb_infect_list_taped', sim._handle, d_inds.data_ptr(), len(inds), variant, cv._capi.LAYER_IMPORT, t, 1, hosp_max, icu_max,
                  d_tape.data_ptr(), sim._stream_ptr)
    torch.cuda.synchronize()
    bad = []
    u = g[pre + 'uniq']
    for key in g.files:
        if not key.startswith(pre + 'post/'):
            continue
        k = key.split('/')[-1]
        want = g[key]
        got = (P[k][variant, uniq] if k == 'exposed_by_variant' else P[k][uniq]).cpu().numpy()
        if k == 'peak_nab':
            ok = np.allclose(got, want, rtol=1e-6, atol=0, equal_nan=True)
        else:
            ok = np.array_equal(got.astype(want.dtype), want, equal_nan=(want.dtype.kind == 'f'))
        if not ok:
            j = int(np.nonzero(~np.isclose(got.astype(np.float64), want.astype(np.float64), rtol=1e-6, atol=0, equal_nan=True))[0][0])
            bad.append(f'{k}: agent {u[j]} got {got[j]} want {want[j]}')
    assert not bad, f'{name} day {day} (variant {variant}, hosp_max {hosp_max}, icu_max {icu_max}): ' + '; '.join(bad)
    assert int(sim._counters[t, cv.defaults.COUNTER_IDS['new_infections']].item()) == len(infected)


@pytest.mark.parametrize('day', [12, 25])
def test_test_prob_kernel_against_reference_tape(cv, day):
    '''
    tests/golden/test_tape.npz (oracle/gen_test_golden.py): a test_prob.apply + People.test call of the unmodified reference -- the People
    arrays it reads, the three uniform arrays it consumed (per agent) and the arrays it wrote.  The CUDA test_prob kernel, given the same
    state and the same uniforms (cvb_test_prob_taped), must write the same arrays, and count the same number of tests.
    '''
    import ctypes as C
    import os
    import torch
    import scenarios
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'test_tape.npz'))
    pre = f'hybrid3k/t{day}/'
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS['hybrid3k']))
    sim.initialize()
    P, dev = sim.people, sim.people.device
    for key in g.files:
        if key.startswith(pre + 'pre/'):
            k = key.split('/')[-1]
            P[k] = g[key]
    tp = [iv for iv in sim['interventions'] if isinstance(iv, cv.test_prob)][0]
    tape = torch.as_tensor(g[pre + 'tape'], dtype=torch.float64, device=dev).contiguous()
    sim.t = day
    sim._push_pars()
    cv._capi.call('cvb_test_prob_taped', sim._handle, day, C.byref(tp._c), None, tape.data_ptr(), sim._stream_ptr)
    torch.cuda.synchronize()
    for key in g.files:
        if key.startswith(pre + 'post/'):
            k = key.split('/')[-1]
            got, want = P.to_numpy(k), g[key]
            assert np.array_equal(got, want, equal_nan=(want.dtype.kind == 'f')), f'day {day}: {k} differs at {np.nonzero(~((got == want) | ((got != got) & (want != want))))[0][:5]}'
    assert int(sim._counters[day, cv.defaults.COUNTER_IDS['new_tests']].item()) == int(g[pre + 'n_tests'])


@pytest.mark.parametrize('day', [12, 25])
@pytest.mark.parametrize('adjacency', [True, False])
def test_contact_tracing_kernel_against_reference_tape(cv, day, adjacency):
    '''
    tests/golden/trace_tape.npz (oracle/gen_trace_vacc_golden.py): a contact_tracing.apply call of the unmodified reference -- the arrays it
    reads, the quarantine requests pending before it, the uniforms binomial_filter consumed per (layer, contact), and afterwards
    known_contact / date_known_contact and the pending requests of the next three days.  The CUDA tracing kernels (adjacency rows, or the
    streamed edge lists), given the same state and the same uniforms (cvb_contact_tracing_taped), must leave the same arrays and requests.
    The contact network is the reference's own: replay mode builds the population from the reference's random streams.
    '''
    import ctypes as C
    import hashlib
    import os
    import torch
    import scenarios
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, 'golden', 'trace_tape.npz'))
    ref = np.load(os.path.join(here, 'golden', 'hybrid3k.npz'))
    pre = f'hybrid3k/t{day}/'
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS['hybrid3k']), rng='mt', use_adjacency=adjacency)
    sim.initialize()
    P, dev, n = sim.people, sim.people.device, sim.n
    for lk, layer in P.contacts.items():                              # the same edges as the reference's run
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        assert sha(layer['p1'].cpu().numpy()) + sha(layer['p2'].cpu().numpy()) == str(ref[f'final_contacts_digest/{lk}'])
    for key in g.files:
        if key.startswith(pre + 'pre/'):
            P[key.split('/')[-1]] = g[key]
    ct = [iv for iv in sim['interventions'] if isinstance(iv, cv.contact_tracing)][0]
    sim.t = day
    sim._push_pars()
    if adjacency:                                                     # (replay mode itself walks the edge lists: build the rows the native mode uses)
        sim.rng_mode, sim._adj_dirty = 'philox', True
        sim._build_adjacency()
        assert sim._adj is not None
    sim._set_quar_horizon(3)
    for d in range(day, day + 3):                                     # requests already pending, as (agent, end day)
        pend = g[pre + f'pend_pre/{d}']
        for end in np.unique(pend[pend >= 0]):
            inds = torch.as_tensor(np.nonzero(pend == end)[0].astype(np.int32), device=dev)
            cv._capi.call('cvb_schedule_quarantine', sim._handle, inds.data_ptr(), len(inds), d, float(end), sim._stream_ptr)
    tape = torch.as_tensor(g[pre + 'tape'], dtype=torch.float64, device=dev).contiguous()
    cv._capi.call('cvb_contact_tracing_taped', sim._handle, day, C.byref(ct._c), tape.data_ptr(), sim._stream_ptr)
    torch.cuda.synchronize()
    for k in ('known_contact', 'date_known_contact'):
        got, want = P.to_numpy(k), g[pre + 'post/' + k]
        assert np.array_equal(got, want, equal_nan=(want.dtype.kind == 'f')), f'day {day}: {k} differs at {np.nonzero(~((got == want) | ((got != got) & (want != want))))[0][:5]}'
    out = torch.empty(n, dtype=torch.float32, device=dev)
    n_req = 0
    for d in range(day, day + 3):
        cv._capi.call('cvb_pending_quarantine', sim._handle, d, out.data_ptr(), sim._stream_ptr)
        torch.cuda.synchronize()
        want = g[pre + f'pend_post/{d}']
        got = out.cpu().numpy()
        assert np.array_equal(got, want), f'day {day}: requests starting on day {d} differ at {np.nonzero(got != want)[0][:5]}: got {got[got != want][:5]} want {want[got != want][:5]}'
        n_req += int((want >= 0).sum())
    assert n_req > 300


@pytest.mark.parametrize('day,label', [(5, 'pfizer'), (26, 'pfizer'), (32, 'jj_boost')])
def test_vaccinate_kernel_against_reference_tape(cv, day, label):
    '''
    tests/golden/vacc_tape.npz (oracle/gen_trace_vacc_golden.py): a BaseVaccination.vaccinate call of the unmodified reference (first doses, second
    doses, a booster) -- the agents, the arrays it reads and writes before / after, the intervention's own dose counts and the initial
    NAb samples.  The CUDA dose kernel, given the same agents (as explicit probabilities 1), state and samples (cvb_vaccinate_taped), must
    write the same arrays and count the same flows; the peak NAb level (a float64 2**x) is compared at 1e-6.
    '''
    import ctypes as C
    import os
    import torch
    import scenarios
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'vacc_tape.npz'))
    pre = f'variants4k/t{day}/{label}/'
    sim = cv.Sim(**scenarios.build(cv, scenarios.SCENARIOS['variants4k']))
    sim.initialize()
    P, dev, n = sim.people, sim.people.device, sim.n
    iv = [v for v in sim['interventions'] if isinstance(v, cv.vaccinate_prob) and v.label == label][0]
    for key in g.files:
        if key.startswith(pre + 'pre/'):
            k = key.split('/')[-1]
            if k == 'iv_doses':
                iv.doses.copy_(torch.as_tensor(g[key].astype(np.int32), device=dev))
            else:
                P[k] = g[key]
    prob = torch.zeros(n, dtype=torch.float64, device=dev)
    prob[torch.as_tensor(g[pre + 'inds'].astype(np.int64), device=dev)] = 1.0
    tape = torch.as_tensor(np.nan_to_num(g[pre + 'tape'], nan=0.0), dtype=torch.float64, device=dev).contiguous()
    iv.due_day.fill_(-1)
    iv._c.first_dose_today, iv._c.second_dose_today = 1, 0
    sim.t = day
    sim._push_pars()
    cv._capi.call('cvb_vaccinate_taped', sim._handle, day, C.byref(iv._c), iv.doses.data_ptr(), iv.due_day.data_ptr(), prob.data_ptr(), tape.data_ptr(), sim._stream_ptr)
    torch.cuda.synchronize()
    for key in g.files:
        if not key.startswith(pre + 'post/'):
            continue
        k = key.split('/')[-1]
        want = g[key]
        got = iv.doses.cpu().numpy() if k == 'iv_doses' else P.to_numpy(k)
        if k == 'peak_nab':
            assert np.allclose(got, want, rtol=1e-6, atol=0, equal_nan=True), f'day {day}: peak_nab differs'
        else:
            assert np.array_equal(got.astype(want.dtype), want, equal_nan=(want.dtype.kind == 'f')), f'day {day}: {k} differs at {np.nonzero(got.astype(want.dtype) != want)[0][:5]}'
    flows = g[pre + 'flows']
    assert int(sim._counters[day, cv.defaults.COUNTER_IDS['new_doses']].item()) == int(flows[0])
    assert int(sim._counters[day, cv.defaults.COUNTER_IDS['new_vaccinated']].item()) == int(flows[1])
