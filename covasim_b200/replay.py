'''
Replay mode (``cv.Sim(rng='mt')``): the device path driven by the reference's two MT19937 streams.

In this mode the simulated day follows the reference's call order exactly (reference sim.py:558-685 and Appendix A of
SURVEY.md) and every random draw is taken on the host from ``sim.rng`` (``np_`` = NumPy's global stream, ``nb`` =
Numba's; both verified in lockstep against the reference by oracle/gen_golden.py) with the sizes and in the order the
reference would use, then uploaded.  The numerics stay on the device:

* state transitions, immunity, NAb kinetics, counters: the same kernels as native mode
  (cvb_update_states_pre / _post, cvb_update_nab_count);
* transmission: cvb_compute_viral_load, cvb_compute_trans_sus per (variant, layer), and cvb_infections_count /
  cvb_infections_draw, which consume the uploaded uniforms in the reference's edge order and return the ordered
  (source, target) lists;
* People.infect, People.test, contact tracing and vaccination are written against the People *device tensors* with
  torch indexing -- line for line the reference's logic (people.py:435-617, interventions.py:921-1145, 1428-1662).  This is
  also how a user's own Python intervention works on the device arrays.

The result is a simulation that reproduces the REFERENCE bit for bit (tests/test_gpu_replay.py checks it against the
golden files recorded from the unmodified reference, including the 58 values of its tests/baseline.json).  It synchronises
with the device several times per day and is meant for verification, not speed.
'''
import ctypes as C

import numpy as np
import torch

from . import defaults as cvd
from . import _capi

f32 = np.float32


# ---------------------------------------------------------------------------------------------------
# host samplers on the two streams (reference utils.py:156-237)
# ---------------------------------------------------------------------------------------------------
def sample(rng, dist=None, par1=None, par2=None, size=None, **kw):
    size = int(size)
    np_ = rng.np_
    if dist in ('unif', 'uniform'):
        return np_.uniform(low=par1, high=par2, size=size)
    if dist in ('norm', 'normal'):
        return np_.normal(loc=par1, scale=par2, size=size)
    if dist == 'normal_pos':
        return np.abs(np_.normal(loc=par1, scale=par2, size=size))
    if dist == 'normal_int':
        return np.round(np.abs(np_.normal(loc=par1, scale=par2, size=size)))
    if dist == 'poisson':
        return rng.nb.poisson(f32(par1), size)
    if dist == 'neg_binomial':
        step = kw.get('step', 1)
        return np_.negative_binomial(n=par2, p=par2 / (par1 / step + par2), size=size) * step
    if dist in ('lognorm', 'lognormal', 'lognorm_int', 'lognormal_int'):
        if par1 > 0:
            mean = np.log(par1 ** 2 / np.sqrt(par2 ** 2 + par1 ** 2))
            sigma = np.sqrt(np.log(par2 ** 2 / par1 ** 2 + 1))
            out = np_.lognormal(mean=mean, sigma=sigma, size=size)
        else:
            out = np.zeros(size)
        return np.round(out) if '_int' in dist else out
    raise NotImplementedError(f'The selected distribution "{dist}" is not implemented')


def _dev(sim, arr, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(arr)).to(device=sim.device, dtype=dtype)


def _uniforms(sim, n):
    ''' n float64 uniforms from the NumPy stream, on the device '''
    return _dev(sim, sim.rng.np_.random_sample(int(n)))


def _add_flow(sim, key, value):
    sim._counters[sim.t, cvd.COUNTER_IDS[key]] += int(value)


def _add_vflow(sim, key, variant, value):
    sim._vcounters[sim.t, variant, cvd.VCOUNTER_IDS[key]] += int(value)


# ---------------------------------------------------------------------------------------------------
# People.infect + update_peak_nab (reference people.py:435-586, immunity.py:138-202)
# ---------------------------------------------------------------------------------------------------
def update_peak_nab(sim, inds, nab_pars, symp=None):
    P = sim.people
    has = P.nab[inds] > 0
    prior, fresh = inds[has], inds[~has]
    if len(prior):
        from .interventions import boost_is_f64
        if boost_is_f64(nab_pars['nab_boost']):                  # (a NumPy float64 boost -- target_eff -- multiplies in float64)
            P.peak_nab[prior] = (P.peak_nab[prior].double() * float(nab_pars['nab_boost'])).float()
        else:
            P.peak_nab[prior] = P.peak_nab[prior] * float(nab_pars['nab_boost'])
    if len(fresh):
        if nab_pars['nab_init'] is None:
            raise ValueError(f'Attempt to administer a vaccine without an initial NAb distribution to {len(fresh)} unvaccinated people failed.')
        level = _dev(sim, 2.0 ** sample(sim.rng, size=len(fresh), **nab_pars['nab_init']))      # the power is taken on the host, like the reference
        if symp is not None:
            scale = torch.full((sim.n,), float('nan'), dtype=torch.float64, device=sim.device)
            ris = sim.pars['rel_imm_symp']
            scale[symp['asymp']] = ris['asymp']
            scale[symp['mild']] = ris['mild']
            scale[symp['sev']] = ris['severe']
            level = level * scale[fresh] * (1 + nab_pars['nab_eff']['alpha_inf_diff'])
        P.peak_nab[fresh] = level.to(torch.float32)
    P.t_nab_event[inds] = int(sim.people.t)


def infect(sim, inds, hosp_max=False, icu_max=False, source=None, layer=None, variant=0, count_flows=True):
    P, pars, t, dev = sim.people, sim.pars, sim.t, sim.device
    inds = torch.as_tensor(np.asarray(inds) if not isinstance(inds, torch.Tensor) else inds).to(device=dev, dtype=torch.int64)
    if len(inds) == 0:
        return inds
    # np.unique(inds, return_index=True): sorted unique targets, source of the FIRST occurrence (people.py:465-467)
    uniq, inverse = torch.unique(inds, sorted=True, return_inverse=True)
    first = torch.full((len(uniq),), len(inds), dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inverse, torch.arange(len(inds), device=dev), reduce='amin')
    if source is not None:
        source = torch.as_tensor(source).to(device=dev, dtype=torch.int64)[first]
    keep = P.susceptible[uniq]
    inds = uniq[keep]
    if source is not None:
        source = source[keep]
    n = len(inds)
    rel = {k: pars[k] for k in ('rel_symp_prob', 'rel_severe_prob', 'rel_crit_prob', 'rel_death_prob')}
    if variant:
        vp = pars['variant_pars'][pars['variant_map'][variant]]
        for k in rel:
            rel[k] *= vp[k]
    dur = pars['dur']

    bt = inds[P.peak_nab[inds] != 0]                                   # breakthrough infections (people.py:486-491)
    if len(bt):
        first_bt = bt[P.n_breakthroughs[bt] == 0]
        P.rel_trans[first_bt] = P.rel_trans[first_bt] * float(f32(pars['trans_redux']))
    for k in ('susceptible', 'naive', 'recovered', 'diagnosed'):
        P[k][inds] = False
    P.exposed[inds] = True
    P.n_infections[inds] += 1
    P.n_breakthroughs[bt] += 1
    P.exposed_variant[inds] = float(variant)
    P.exposed_by_variant[variant, inds] = True
    if count_flows:
        _add_flow(sim, 'new_infections', n)
        _add_flow(sim, 'new_reinfections', int(torch.count_nonzero(~torch.isnan(P.date_recovered[inds]))))
        _add_vflow(sim, 'new_infections_by_variant', variant, n)
    sim._log_append(source, inds, layer, variant)

    def d(key, who):
        return _dev(sim, sample(sim.rng, size=len(who), **dur[key]))

    P.dur_exp2inf[inds] = d('exp2inf', inds).to(torch.float32)
    P.date_exposed[inds] = float(t)
    P.date_infectious[inds] = P.dur_exp2inf[inds] + float(t)
    for k in ('date_symptomatic', 'date_severe', 'date_critical', 'date_diagnosed', 'date_recovered'):
        P[k][inds] = float('nan')

    p_symp = float(f32(rel['rel_symp_prob'])) * P.symp_prob[inds] * (1 - P.symp_imm[variant, inds])
    is_symp = _uniforms(sim, n) < p_symp
    symp, asymp = inds[is_symp], inds[~is_symp]
    if count_flows:
        _add_vflow(sim, 'new_symptomatic_by_variant', variant, len(symp))

    dd = d('asym2rec', asymp)
    P.date_recovered[asymp] = (P.date_infectious[asymp] + dd).to(torch.float32)
    P.dur_disease[asymp] = (P.dur_exp2inf[asymp] + dd).to(torch.float32)

    P.dur_inf2sym[symp] = d('inf2sym', symp).to(torch.float32)
    P.date_symptomatic[symp] = P.date_infectious[symp] + P.dur_inf2sym[symp]
    p_sev = float(f32(rel['rel_severe_prob'])) * P.severe_prob[symp] * (1 - P.sev_imm[variant, symp])
    is_sev = _uniforms(sim, len(symp)) < p_sev
    sev, mild = symp[is_sev], symp[~is_sev]
    if count_flows:
        _add_vflow(sim, 'new_severe_by_variant', variant, len(sev))

    dd = d('mild2rec', mild)
    P.date_recovered[mild] = (P.date_symptomatic[mild] + dd).to(torch.float32)
    P.dur_disease[mild] = (P.dur_exp2inf[mild] + P.dur_inf2sym[mild] + dd).to(torch.float32)

    P.dur_sym2sev[sev] = d('sym2sev', sev).to(torch.float32)
    P.date_severe[sev] = P.date_symptomatic[sev] + P.dur_sym2sev[sev]
    p_crit = float(f32(rel['rel_crit_prob'])) * P.crit_prob[sev] * float(f32(pars['no_hosp_factor'] if hosp_max else 1.0))
    is_crit = _uniforms(sim, len(sev)) < p_crit
    crit, noncrit = sev[is_crit], sev[~is_crit]

    dd = d('sev2rec', noncrit)
    P.date_recovered[noncrit] = (P.date_severe[noncrit] + dd).to(torch.float32)
    P.dur_disease[noncrit] = (P.dur_exp2inf[noncrit] + P.dur_inf2sym[noncrit] + P.dur_sym2sev[noncrit] + dd).to(torch.float32)

    P.dur_sev2crit[crit] = d('sev2crit', crit).to(torch.float32)
    P.date_critical[crit] = P.date_severe[crit] + P.dur_sev2crit[crit]
    p_death = float(f32(rel['rel_death_prob'])) * P.death_prob[crit] * float(f32(pars['no_icu_factor'] if icu_max else 1.0))
    is_dead = _uniforms(sim, len(crit)) < p_death
    dead, alive = crit[is_dead], crit[~is_dead]

    dd = d('crit2rec', alive)
    P.date_recovered[alive] = (P.date_critical[alive] + dd).to(torch.float32)
    P.dur_disease[alive] = (P.dur_exp2inf[alive] + P.dur_inf2sym[alive] + P.dur_sym2sev[alive] + P.dur_sev2crit[alive] + dd).to(torch.float32)

    dd = d('crit2die', dead)
    P.date_dead[dead] = (P.date_critical[dead] + dd).to(torch.float32)
    P.dur_disease[dead] = (P.dur_exp2inf[dead] + P.dur_inf2sym[dead] + P.dur_sym2sev[dead] + P.dur_sev2crit[dead] + dd).to(torch.float32)
    P.date_recovered[dead] = float('nan')

    if pars['use_waning']:
        update_peak_nab(sim, inds, pars, symp=dict(asymp=asymp, mild=mild, sev=sev))
    return inds


# ---------------------------------------------------------------------------------------------------
# built-in interventions on the device tensors with host draws
# ---------------------------------------------------------------------------------------------------
def test_people(sim, inds, sensitivity, loss_prob, test_delay):
    ''' reference people.py:589-617 '''
    P, t = sim.people, sim.t
    inds = torch.unique(inds)
    P.tested[inds] = True
    P.date_tested[inds] = float(t)
    is_inf = inds[P.infectious[inds]]
    pos = _uniforms(sim, len(is_inf)) < sensitivity
    is_inf_pos = is_inf[pos]
    not_diag = is_inf_pos[torch.isnan(P.date_diagnosed[is_inf_pos])]
    kept = _uniforms(sim, len(not_diag)) < (1.0 - loss_prob)
    final = not_diag[kept]
    P.date_diagnosed[final] = float(t + test_delay)
    P.date_pos_test[final] = float(t)
    return final


def test_prob_apply(iv, sim):
    ''' reference interventions.py:921-981 '''
    P, t = sim.people, sim.t
    symp = P.symptomatic
    from .interventions import quar_test_mask
    qt = quar_test_mask(sim, iv.quar_policy)                                                # interventions.py:682-712 get_quar_inds
    probs = torch.where(symp, iv.symp_prob, iv.asymp_prob).to(torch.float64)
    probs[qt & symp] = iv.symp_quar_prob
    probs[qt & ~symp] = iv.asymp_quar_prob
    ov = iv.override(sim)                          # ILI symptoms (consumes the Numba stream like the reference) and subtargets
    if ov is not None:
        probs = torch.where(torch.isnan(ov), probs, ov)
    probs[P.diagnosed] = 0.0
    tested = torch.nonzero(_uniforms(sim, sim.n) < probs).flatten()
    test_people(sim, tested, iv.sensitivity, iv.loss_prob, iv.test_delay)
    sim._host_add('new_tests', t, len(tested) * sim.pars['pop_scale'] / sim.rescale_vec[t])


def test_num_apply(iv, sim):
    ''' reference interventions.py:786-854: the weights are built on the device, choose_w runs on the mirrored NumPy stream '''
    n_tests = iv.n_tests_today(sim)
    if not n_tests:
        return
    P, t = sim.people, sim.t
    probs = torch.ones(sim.n, dtype=torch.float64, device=sim.device)
    if iv.pdf is not None and bool(P.symptomatic.any()):                                     # interventions.py:812-819 swab_delay
        from .interventions import swab_terms
        symp_inds, symp_time, dens, count = swab_terms(iv.pdf, sim)
        probs[symp_inds] *= torch.as_tensor(float(iv.symp_test) * (dens * count), dtype=torch.float64, device=sim.device)
    else:
        probs[P.symptomatic] *= iv.symp_test
    rel_t = t - iv.start_day
    if iv.ili_prev is not None and rel_t < len(iv.ili_prev):                                # interventions.py:823-828 (cvu.choose: Numba stream)
        ili = sim.rng.nb.choice(sim['pop_size'], int(iv.ili_prev[rel_t] * sim['pop_size']), replace=False)
        ili = torch.as_tensor(np.asarray(ili), dtype=torch.int64, device=sim.device)
        probs[ili[~P.symptomatic[ili]]] *= iv.symp_test
    from .interventions import quar_test_mask
    qt = quar_test_mask(sim, iv.quar_policy)                                                # interventions.py:682-712 get_quar_inds
    probs[qt] *= iv.quar_test
    if iv.subtarget is not None:                                                             # interventions.py:834-837
        from .interventions import get_subtargets
        s_inds, s_vals = get_subtargets(iv.subtarget, sim)
        s_inds = torch.as_tensor(np.asarray(s_inds), dtype=torch.int64, device=sim.device)
        s_vals = torch.as_tensor(np.asarray(s_vals), dtype=torch.float64, device=sim.device)
        probs[s_inds] = probs[s_inds] * s_vals
    probs[P.diagnosed] = 0.0
    probs = probs.cpu().numpy()
    n_tests = iv.rescaled(sim, n_tests, probs.sum())
    n_tests = min(n_tests, int((probs != 0).sum()))
    total = probs.sum()
    p = probs / total if total else np.ones(len(probs)) / len(probs)                       # reference utils.py:446-483 choose_w
    inds = sim.rng.np_.choice(len(probs), int(n_tests), p=p, replace=False)
    test_people(sim, torch.as_tensor(inds, dtype=torch.int64, device=sim.device), iv.sensitivity, iv.loss_prob, iv.test_delay)


def contact_tracing_apply(iv, sim):
    ''' reference interventions.py:1044-1145 '''
    P, t = sim.people, sim.t
    if not iv.presumptive:
        cases = torch.nonzero(P.date_diagnosed == t).flatten()
    else:
        just = torch.nonzero(P.date_tested == t).flatten()
        cases = just[P.exposed[just]]
    if iv.capacity is not None:                        # interventions.py:1079-1083: np.random.choice on the NumPy stream
        cap = int(iv.capacity / sim.rescale_vec[t])
        if len(cases) > cap:
            cases = torch.as_tensor(sim.rng.np_.choice(cases.cpu().numpy(), cap, replace=False), device=cases.device)
    if not len(cases):
        return
    by_time = {}
    for lkey, prob in iv.trace_probs.items():
        if prob == 0:
            continue
        found = sim.people.contacts[lkey].find_contacts(cases).to(torch.int64)          # device kernel: sorted unique partners
        if len(found):
            u = _uniforms(sim, len(found))
            by_time.setdefault(iv.trace_time[lkey], []).append(found[u < prob])
    for trace_time, parts in by_time.items():
        who = torch.unique(torch.cat(parts))
        who = who[~P.dead[who]]
        P.known_contact[who] = True
        P.date_known_contact[who] = torch.fmin(P.date_known_contact[who], torch.tensor(float(t + trace_time), device=sim.device))
        P.schedule_quarantine(who, start_date=t + trace_time, period=iv.quar_period - trace_time)


def vaccinate_apply(iv, sim):
    ''' reference interventions.py:1428-1482, 1631-1662 '''
    P, t = sim.people, sim.t
    picked = torch.zeros(0, dtype=torch.int64, device=sim.device)
    if t < np.min(iv.days):
        return
    if np.any(iv.days == t):
        probs = torch.zeros(sim.n, dtype=torch.float64, device=sim.device)
        eligible = P.vaccinated if iv.booster else ~P.vaccinated
        probs[eligible] = iv.prob
        if iv.subtarget is not None:
            from .interventions import subtarget_override
            ov = subtarget_override(iv.subtarget, sim)
            probs = torch.where(torch.isnan(ov), probs, ov)
        picked = torch.nonzero(_uniforms(sim, sim.n) < probs).flatten()
        if len(picked) and iv.p['interval'] is not None:
            nxt = t + iv.p['interval']
            if nxt < sim['n_days']:
                iv._second[nxt] = picked
    second = iv._second.get(t)
    if second is not None:
        picked = torch.cat((picked, second))
    vaccinate_inds(iv, sim, picked)


def vaccinate_num_apply(iv, sim):
    '''
    reference interventions.py:1715-1791 vaccinate_num.select_people, literally: the scheduled second doses live in day-keyed
    Python sets whose iteration order decides the order of the shuffle and of the NAb draws, so the selection runs on host
    copies of the three arrays it reads.
    '''
    P, t = sim.people, sim.t
    sched = lambda day: iv._scheduled.setdefault(day, set())
    num_people = iv.n_today(sim)
    if num_people == 0:
        sched(t + 1).update(sched(t))
        return
    num_agents = int(np.floor(num_people / sim['pop_scale'] + sim.rng.np_.random_sample()))         # sc.randround
    dead, vaccinated, doses = np.asarray(P.dead), np.asarray(P.vaccinated), iv.doses.cpu().numpy()
    picked = None
    if sched(t):
        scheduled = np.fromiter(sched(t), dtype=np.int32)
        scheduled = scheduled[(doses[scheduled] < iv.p['doses']) & ~dead[scheduled]]
        if len(scheduled) > num_agents:
            sim.rng.np_.shuffle(scheduled)
            sched(t + 1).update(scheduled[num_agents:])
            picked = scheduled[:num_agents]
    else:
        scheduled = np.array([], dtype=np.int32)
    if picked is None:
        probs = np.ones(sim.n)
        probs[dead] = 0.0
        if iv.subtarget is not None:
            from .interventions import get_subtargets
            s_inds, s_vals = get_subtargets(iv.subtarget, sim)
            s_inds = np.asarray(s_inds)
            probs[s_inds] = probs[s_inds] * np.asarray(s_vals)
        if iv.booster:
            probs[~vaccinated] = 0.0
        else:
            probs[vaccinated] = 0.0
        seq = iv._sequence_host
        eligible = seq[sim.rng.np_.random_sample(sim.n) < probs[seq]]
        if len(eligible) == 0:
            picked = scheduled
        else:
            eligible = eligible[:num_agents]
            eligible = eligible[~np.isin(eligible, scheduled)]
            first = eligible[:num_agents - len(scheduled)] if len(eligible) + len(scheduled) > num_agents else eligible
            if iv.p['doses'] > 1:
                sched(t + iv.p['interval']).update(first)
            picked = np.concatenate([scheduled, first])
    vaccinate_inds(iv, sim, torch.as_tensor(np.asarray(picked, dtype=np.int64), device=sim.device))


def vaccinate_inds(iv, sim, picked):
    ''' reference interventions.py:1428-1482 BaseVaccination.vaccinate '''
    P, t = sim.people, sim.t
    if not len(picked):
        return
    inds = picked[~P.dead[picked]]
    inds = inds[iv.doses[inds] < iv.p['doses']]
    new_vacc = inds[~P.vaccinated[inds]]
    n_new = len(torch.unique(new_vacc))
    if len(inds):
        iv.doses[inds] += 1
        P.vaccinated[inds] = True
        P.vaccine_source[inds] = iv.index
        P.doses[inds] += 1
        P.date_vaccinated[inds] = float(t)
        update_peak_nab(sim, inds, iv.p, symp=None)
        factor = sim.pars['pop_scale'] / sim.rescale_vec[t]
        sim._host_add('new_doses', t, len(inds) * factor)
        sim._host_add('new_vaccinated', t, n_new * factor)


# ---------------------------------------------------------------------------------------------------
# one day, in the reference's order
# ---------------------------------------------------------------------------------------------------
def update_dynamic_layer(sim, layer):
    ''' reference base.py:1849-1876 with the Numba stream '''
    n = len(layer)
    inds = sim.rng.nb.choice(n, n, replace=False)
    p1, p2 = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    p1[inds] = np.array(sim.rng.nb.choice(sim.n, n, replace=True), dtype=np.int32)
    p2[inds] = np.array(sim.rng.nb.choice(sim.n, n, replace=True), dtype=np.int32)
    layer['p1'] = p1
    layer['p2'] = p2
    layer['beta'] = np.ones(n, dtype=np.float32)


def step(sim):
    from .interventions import test_prob, contact_tracing, vaccinate_prob, vaccinate_num
    t, pars, P, h, st = sim.t, sim.pars, sim.people, sim._handle, sim._stream_ptr
    call = _capi.call
    P.t = t
    sim.rescale()
    sim._push_pars()
    call('cvb_update_states_pre', h, t, st)
    for lkey, dyn in pars['dynam_layer'].items():
        if dyn:
            update_dynamic_layer(sim, P.contacts[lkey])
    hosp_max = bool(P.count('severe') > pars['n_beds_hosp']) if pars['n_beds_hosp'] is not None else False
    icu_max = bool(P.count('critical') > pars['n_beds_icu']) if pars['n_beds_icu'] is not None else False
    if pars['n_imports']:
        n_imports = int(sim.rng.nb.poisson(f32(pars['n_imports'] / sim.rescale_vec[t]), 1)[0])
        if n_imports > 0:
            who = sim.rng.nb.choice(pars['pop_size'], n_imports, replace=False)
            infect(sim, who, hosp_max, icu_max, layer='importation')
            sim._host_add('n_imports', t, n_imports)
    for v in pars['variants']:
        v.apply(sim)
    from .interventions import sequence, test_num

    def apply_intervention(iv):
        if isinstance(iv, sequence):                      # the intervention in force today, through the same dispatch
            inner = iv.active(sim)
            if inner is not None:
                apply_intervention(inner)
        elif isinstance(iv, test_prob):
            if not (t < iv.start_day or (iv.end_day is not None and t > iv.end_day)):
                test_prob_apply(iv, sim)
        elif isinstance(iv, test_num):
            test_num_apply(iv, sim)
        elif isinstance(iv, contact_tracing):
            if not (t < iv.start_day or (iv.end_day is not None and t > iv.end_day)):
                contact_tracing_apply(iv, sim)
        elif isinstance(iv, vaccinate_num):
            vaccinate_num_apply(iv, sim)
        elif isinstance(iv, vaccinate_prob):
            vaccinate_apply(iv, sim)
        else:
            iv(sim)
    for iv in pars['interventions']:
        apply_intervention(iv)
    sim._push_pars()
    call('cvb_update_states_post', h, t, st)

    # transmission, per variant and layer, with the Numba stream (reference sim.py:602-649)
    vd = pars['viral_dist']
    n = sim.n
    vl = torch.empty(n, dtype=torch.float32, device=sim.device)
    call('cvb_compute_viral_load', t, P.date_infectious.data_ptr(), P.date_recovered.data_ptr(), P.date_dead.data_ptr(),
         float(vd['frac_time']), float(vd['load_ratio']), float(vd['high_cap']), vl.data_ptr(), n, st)
    rt = torch.empty(n, dtype=torch.float32, device=sim.device)
    rs = torch.empty(n, dtype=torch.float32, device=sim.device)
    for v in range(pars['n_variants']):
        vlabel = pars['variant_map'][v]
        beta = float(f32(pars['beta'] * pars['rel_beta'] * pars['variant_pars'][vlabel]['rel_beta']))
        inf_v = (P.infectious & (P.infectious_variant == v)).contiguous()
        if not bool(inf_v.any()):
            continue
        for lkey, layer in P.contacts.items():
            imm = P.sus_imm[v]
            call('cvb_compute_trans_sus', P.rel_trans.data_ptr(), P.rel_sus.data_ptr(), inf_v.data_ptr(), P.susceptible.data_ptr(),
                 float(pars['beta_layer'][lkey]), vl.data_ptr(), P.symptomatic.data_ptr(), P.isolated.data_ptr(), P.quarantined.data_ptr(),
                 float(pars['asymp_factor']), float(pars['iso_factor'][lkey]), float(pars['quar_factor'][lkey]), imm.data_ptr(),
                 rt.data_ptr(), rs.data_ptr(), n, st)
            E = len(layer)
            n_draws = (C.c_int64 * 2)()
            call('cvb_infections_count', h, beta, layer['p1'].data_ptr(), layer['p2'].data_ptr(), layer['beta'].data_ptr(), E,
                 rt.data_ptr(), rs.data_ptr(), n_draws, st)
            total = int(n_draws[0] + n_draws[1])
            u = _dev(sim, sim.rng.nb.random_sample(total))                  # direction p1->p2 first, then p2->p1 (utils.py:112-127)
            src = torch.empty(max(total, 1), dtype=torch.int32, device=sim.device)
            tgt = torch.empty(max(total, 1), dtype=torch.int32, device=sim.device)
            n_out = C.c_int64(0)
            call('cvb_infections_draw', h, beta, layer['p1'].data_ptr(), layer['p2'].data_ptr(), layer['beta'].data_ptr(), E,
                 rt.data_ptr(), rs.data_ptr(), u.data_ptr(), src.data_ptr(), tgt.data_ptr(), C.byref(n_out), st)
            k = n_out.value
            infect(sim, tgt[:k], hosp_max, icu_max, source=src[:k], layer=lkey, variant=v)

    call('cvb_update_nab_count', h, t, st)
    for an in pars['analyzers']:
        an(sim)
    sim.t += 1
    if sim.t == sim.npts:
        sim.complete = True
