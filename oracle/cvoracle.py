'''
TEST INFRASTRUCTURE -- CPU (NumPy) restatement of Covasim's per-timestep hot path.

This module is the *oracle* for covasim_b200: a plain-NumPy restatement of the algorithm of the
reference (Covasim 3.1.7, /root/reference) for the path SURVEY.md section 8 names.  Every function
cites the reference file:line it follows.  It is deliberately structured differently from the
reference (one flat state dict, free functions) -- it restates the arithmetic, it is not a copy.

PARITY PINNED: in 'mt' RNG mode (two MT19937 streams seeded like the reference) this oracle
reproduces the reference bit-for-bit; oracle/gen_golden.py ran the unmodified reference in the
build container and committed its outputs under tests/golden/, and tests/test_oracle_golden.py
checks the oracle against them (including the 58 values of the reference's own
tests/baseline.json).  In 'philox' RNG mode the same code draws from counter-based keyed streams
(oracle/philox.py); that is the mode the CUDA fast path is checked against bit-for-bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under covasim_b200/ does.
'''
import copy
import datetime as dt
import numpy as np

from . import ref_defaults as cvd       # the oracle's OWN copy of the field/result inventories (never the product package:
from . import ref_parameters as cvpar   # importing covasim_b200 would dlopen the CUDA library); tests/test_oracle_golden.py diffs the two
from . import philox as ph

f32 = np.float32
f64 = np.float64
i32 = np.int32


# =================================================================================================
# RNG providers
# =================================================================================================

class MTStreams:
    '''
    The reference's two MT19937 streams (reference utils.py:271-298): ``np_`` mirrors NumPy's global
    stream, ``nb`` mirrors Numba's.  Same algorithm, same seed, independent state.
    '''
    kind = 'mt'

    def __init__(self, seed=None):
        self.np_ = np.random.RandomState()
        self.nb = np.random.RandomState()
        self.seed = None
        if seed is not None:
            self.set_seed(seed)

    def set_seed(self, seed):
        self.seed = int(seed)
        self.np_.seed(self.seed)
        self.nb.seed(self.seed)

    # -- rare host-side set choices (seed infections sim.py:528, importations sim.py:586, variant imports immunity.py:127):
    #    the reference's choice(n, k, replace=False), i.e. the first k entries of a permutation of n -- O(n) per call
    def choose(self, which, n, k):
        return getattr(self, which).choice(int(n), int(k), replace=False)

    # -- per-edge transmission draws (Numba stream; utils.py:123)
    def edge_uniforms(self, t, layer_idx, direction, edge_inds):
        return self.nb.random_sample(len(edge_inds))

    # -- per-agent Bernoulli uniforms (NumPy stream; utils.py:321-357)
    def agent_uniforms(self, t, purpose, sub, inds, slot=0):
        return self.np_.random_sample(len(inds))

    # -- per-agent standard normals -> handled through sample() below
    def sample(self, t, purpose, sub, inds, slot, dist=None, par1=None, par2=None, **kw):
        return sample(self.np_, self.nb, dist=dist, par1=par1, par2=par2, size=len(inds), **kw)


class PhiloxStreams(MTStreams):
    '''
    Native-RNG mode: every hot-path draw is a pure function of (seed, purpose, sub, day, index, slot).
    Rare host-side set choices (seed infections, imports) still use the MT streams.
    '''
    kind = 'philox'

    def choose(self, which, n, k):
        return choose_distinct(getattr(self, which), n, k)

    def edge_uniforms(self, t, layer_idx, direction, edge_inds):
        u1, u2 = ph.keyed_uniform2(self.seed, ph.P_EDGE, layer_idx, t, edge_inds)
        return u1 if direction == 0 else u2

    def agent_uniforms(self, t, purpose, sub, inds, slot=0):
        return ph.keyed_uniform(self.seed, purpose, sub, t, inds, slot)

    def sample(self, t, purpose, sub, inds, slot, dist=None, par1=None, par2=None, **kw):
        z = ph.keyed_normal(self.seed, purpose, sub, t, inds, slot)
        if dist in ('normal',):
            return par1 + par2 * z
        if dist == 'normal_pos':
            return np.abs(par1 + par2 * z)
        if dist == 'normal_int':
            return np.round(np.abs(par1 + par2 * z))
        if dist in ('lognormal', 'lognormal_int'):
            if par1 > 0:
                mean, sigma = lognormal_pars(par1, par2)
                out = np.exp(mean + sigma * z)
            else:
                out = np.zeros(len(inds))
            return np.round(out) if dist.endswith('_int') else out
        raise NotImplementedError(f'distribution {dist} has no keyed form')


def choose_distinct(stream, n, k):
    '''
    k distinct integers in [0, n) in O(k) for k << n (native-RNG mode; covasim_b200/utils.py:choose_distinct is the same
    function): uniform draws from ``stream``, first occurrences kept in draw order, repeated until k are found.  The
    reference's choice(replace=False) permutes all n -- 30 ms per call at 2M agents, on every importation day.
    '''
    n, k = int(n), int(k)
    if k > n // 8:
        return stream.choice(n, k, replace=False)
    out = np.zeros(0, dtype=np.int64)
    while len(out) < k:
        out = np.concatenate([out, stream.randint(0, n, size=int(1.2 * (k - len(out))) + 8)])
        _, first = np.unique(out, return_index=True)
        out = out[np.sort(first)]
    return out[:k]


def poisson_cdf(lam):
    ''' Cumulative distribution of Poisson(lam) in float64 (same table as covasim_b200/population.py:poisson_cdf) '''
    lam = float(lam)
    kmax = int(lam + 12 * np.sqrt(lam + 1) + 40)
    k = np.arange(1, kmax + 1, dtype=np.float64)
    logp = np.concatenate([[-lam], -lam + np.cumsum(np.log(lam) - np.log(k))]) if lam > 0 else np.concatenate([[0.0], np.full(kmax, -np.inf)])
    return np.cumsum(np.exp(logp))


def make_keyed_pop(pars, seed, school_ages=(6, 22), work_ages=(22, 65)):
    '''
    NumPy restatement of the keyed (device-side) population generator, covasim_b200/population.py:KeyedPop: the
    distributions of the reference's make_randpop / make_random_contacts / make_microstructured_contacts /
    make_hybrid_contacts (population.py:143-364), every draw a keyed Philox uniform (purpose 10).
    '''
    P_POP = 10
    subs = dict(h=1, s=2, w=3, c=4, a=5)
    n = int(pars['pop_size'])
    u = lambda sub, index0, m, slot: ph.keyed_uniform(seed, P_POP, sub, 0, np.arange(index0, index0 + m, dtype=np.int64), slot)
    age_data = cvd.default_age_data
    lo, width = age_data[:, 0].astype(np.float64), (age_data[:, 1] + 1 - age_data[:, 0]).astype(np.float64)
    probs = age_data[:, 2] / age_data[:, 2].sum()
    sexes = (u(0, 0, n, 0) < 0.5).astype(np.int32)
    bins = np.minimum(np.searchsorted(np.cumsum(probs), u(0, 0, n, 1)), len(probs) - 1)
    ages = lo[bins] + width[bins] * u(0, 0, n, 2)

    def poisson(lam, sub, index0, m):
        return np.searchsorted(poisson_cdf(lam), u(sub, index0, m, 0), side='right')

    def random_layer(lk, mapping, nc):
        m = n if mapping is None else len(mapping)
        counts = np.round(poisson(nc, subs.get(lk, 5), 0, m).astype(np.float64) / 2.0).astype(np.int64)
        total = int(counts.sum())
        pos = np.repeat(np.arange(m, dtype=np.int64), counts)
        tpos = np.minimum(np.floor(u(subs.get(lk, 5), 0, total, 1) * m).astype(np.int64), m - 1)
        if mapping is not None:
            pos, tpos = mapping[pos], mapping[tpos]
        return dict(p1=pos.astype(i32), p2=tpos.astype(i32), beta=np.ones(total, dtype=f32))

    def households(cluster_size):
        sizes, covered, c0 = [], 0, 0
        while covered < n:
            block = max(1024, int((n - covered) / max(float(cluster_size), 0.5) * 1.2) + 16)
            draw = poisson(cluster_size, subs['h'], c0, block)
            sizes.append(draw)
            covered += int(draw.sum())
            c0 += block
        ends = np.cumsum(np.concatenate(sizes))
        k = int(np.searchsorted(ends, n)) + 1
        ends = ends[:k].copy()
        ends[-1] = n
        p1, p2 = [], []
        start = 0
        for end in ends.tolist():                      # plain loops: the restatement the vectorised device code is checked against
            for a in range(start, end):
                for b in range(a + 1, end):
                    p1.append(a)
                    p2.append(b)
            start = end
        return dict(p1=np.array(p1, dtype=i32), p2=np.array(p2, dtype=i32), beta=np.ones(len(p1), dtype=f32))

    contacts = {}
    if pars['pop_type'] == 'random':
        for lk, nc in pars['contacts'].items():
            contacts[lk] = random_layer(lk, None, nc)
    else:
        nc = dict(h=4, s=20, w=20, c=20)
        nc.update(pars['contacts'])
        ages32 = ages.astype(np.float32)
        for lk in pars['contacts'].keys():
            if lk == 'h':
                contacts[lk] = households(nc['h'])
            elif lk == 'c':
                contacts[lk] = random_layer(lk, None, nc['c'])
            else:
                a0, a1 = school_ages if lk == 's' else work_ages
                contacts[lk] = random_layer(lk, np.nonzero((ages32 >= a0) & (ages32 < a1))[0], nc[lk])
    return dict(age=ages, sex=sexes, contacts=contacts, layer_keys=list(contacts.keys()))


def lognormal_pars(par1, par2):
    ''' Mean/sigma of the underlying normal (reference utils.py:223-225) '''
    mean = np.log(par1 ** 2 / np.sqrt(par2 ** 2 + par1 ** 2))
    sigma = np.sqrt(np.log(par2 ** 2 / par1 ** 2 + 1))
    return mean, sigma


def sample(np_, nb, dist=None, par1=None, par2=None, size=None, **kw):
    ''' Distribution sampler on the two MT streams (reference utils.py:156-237) '''
    size = int(size)
    if dist in ('unif', 'uniform'):
        return np_.uniform(low=par1, high=par2, size=size)
    if dist in ('norm', 'normal'):
        return np_.normal(loc=par1, scale=par2, size=size)
    if dist == 'normal_pos':
        return np.abs(np_.normal(loc=par1, scale=par2, size=size))
    if dist == 'normal_int':
        return np.round(np.abs(np_.normal(loc=par1, scale=par2, size=size)))
    if dist == 'poisson':
        return nb.poisson(f32(par1), size)                       # utils.py:393-406 (Numba stream)
    if dist == 'neg_binomial':                                    # utils.py:409-426
        step = kw.get('step', 1)
        p = par2 / (par1 / step + par2)
        return np_.negative_binomial(n=par2, p=p, size=size) * step
    if dist in ('lognorm', 'lognormal', 'lognorm_int', 'lognormal_int'):
        if par1 > 0:
            mean, sigma = lognormal_pars(par1, par2)
            out = np_.lognormal(mean=mean, sigma=sigma, size=size)
        else:
            out = np.zeros(size)
        return np.round(out) if '_int' in dist else out
    raise NotImplementedError(f'The selected distribution "{dist}" is not implemented')


# =================================================================================================
# Population (init-time; reference population.py)
# =================================================================================================

def random_contacts(rng, pop_size, n, overshoot=1.2, mapping=None):
    ''' One random layer as an edge list (reference population.py:239-283) '''
    pop_size = int(pop_size)
    n_all = int(pop_size * n * overshoot)
    pool = rng.nb.choice(pop_size, n_all, replace=True) if pop_size > 0 else np.zeros(0, dtype=np.int64)
    counts = rng.nb.poisson(f32(n), pop_size)
    counts = np.array((counts / 2.0).round(), dtype=i32)
    total = int(counts.sum())
    p1 = np.repeat(np.arange(pop_size, dtype=i32), counts)
    p2 = np.array(pool[:total], dtype=i32)
    if mapping is not None:
        mapping = np.array(mapping, dtype=i32)
        p1, p2 = mapping[p1], mapping[p2]
    return p1, p2


def household_contacts(rng, pop_size, cluster_size):
    ''' Cliques of consecutive agents with Poisson sizes (reference population.py:286-329) '''
    pop_size = int(pop_size)
    p1, p2 = [], []
    start = 0
    while start < pop_size:
        size = int(rng.nb.poisson(f32(cluster_size), 1)[0])
        size = min(size, pop_size - start)
        if size > 1:
            members = range(start, start + size)
            for a in members:
                # The reference collects the partners of ``a`` in a Python set and appends list(set):
                # the edge order within a household is therefore CPython's set iteration order
                # (not always ascending), which we reproduce by doing the same thing.
                partners = set()
                for b in members:
                    if b > a:
                        partners.add(b)
                p1.append(np.full(len(partners), a, dtype=i32))
                p2.append(np.array(list(partners), dtype=i32))
        start += size
    if len(p1):
        return np.concatenate(p1).astype(i32), np.concatenate(p2).astype(i32)
    return np.zeros(0, dtype=i32), np.zeros(0, dtype=i32)


def make_population(pars, rng):
    ''' Ages, sexes and contact layers (reference population.py:143-236, 332-364) '''
    n = int(pars['pop_size'])
    sexes = rng.np_.binomial(1, 0.5, n)
    age_data = cvd.default_age_data
    lo = age_data[:, 0]
    width = age_data[:, 1] + 1 - lo
    probs = age_data[:, 2] / age_data[:, 2].sum()
    bins = np.searchsorted(np.cumsum(probs), rng.np_.random_sample(n))        # utils.py:360-375
    ages = lo[bins] + width[bins] * rng.np_.random_sample(n)
    layers = {}
    if pars['pop_type'] == 'random':
        for lkey, nc in pars['contacts'].items():
            layers[lkey] = random_contacts(rng, n, nc)
    elif pars['pop_type'] == 'hybrid':
        nc = dict(h=4, s=20, w=20, c=20)
        nc.update(pars['contacts'])
        gen = {}
        gen['h'] = household_contacts(rng, n, nc['h'])           # generation order h, c, s, w
        gen['c'] = random_contacts(rng, n, nc['c'])
        s_inds = np.nonzero((ages >= 6) * (ages < 22))[0]
        w_inds = np.nonzero((ages >= 22) * (ages < 65))[0]
        gen['s'] = random_contacts(rng, len(s_inds), nc['s'], mapping=s_inds)
        gen['w'] = random_contacts(rng, len(w_inds), nc['w'], mapping=w_inds)
        for lkey in pars['contacts'].keys():                      # stored in parameter order h, s, w, c
            layers[lkey] = gen[lkey]
    else:
        raise NotImplementedError(pars['pop_type'])
    contacts = {lk: dict(p1=p1, p2=p2, beta=np.ones(len(p1), dtype=f32)) for lk, (p1, p2) in layers.items()}
    return dict(age=ages, sex=sexes, contacts=contacts)


# =================================================================================================
# People state: a flat dict of arrays
# =================================================================================================

def new_people(pars, age, sex):
    ''' Allocate the structure-of-arrays state (reference people.py:47-118) '''
    n = int(pars['pop_size'])
    nv = int(pars['n_variants'])
    P = {}
    for k in cvd.person_fields:
        if k == 'uid':
            P[k] = np.arange(n, dtype=i32)
        elif k in cvd.person_int_fields:
            P[k] = np.zeros(n, dtype=i32)
        else:
            P[k] = np.full(n, np.nan, dtype=f32)
    for k in cvd.states:
        P[k] = np.full(n, k in ('susceptible', 'naive'), dtype=bool)
    for k in cvd.variant_states:
        P[k] = np.full(n, np.nan, dtype=f32)
    for k in cvd.by_variant_states:
        P[k] = np.zeros((nv, n), dtype=bool)
    for k in cvd.imm_states:
        P[k] = np.zeros((nv, n), dtype=f32)
    for k in cvd.nab_states:
        P[k] = np.zeros(n, dtype=i32 if k == 't_nab_event' else f32)
    for k in cvd.vacc_states:
        P[k] = np.zeros(n, dtype=i32)
    for k in cvd.dates + cvd.durs:
        P[k] = np.full(n, np.nan, dtype=f32)
    P['age'][:] = age
    P['sex'][:] = sex
    return P


def set_prognoses(P, pars, rng):
    ''' Age-dependent prognosis probabilities and transmissibility draw (reference people.py:139-161) '''
    rng.set_seed(pars['rand_seed'])
    progs = pars['prognoses']
    inds = np.digitize(P['age'], progs['age_cutoffs']) - 1
    P['symp_prob'][:] = progs['symp_probs'][inds]
    P['severe_prob'][:] = progs['severe_probs'][inds] * progs['comorbidities'][inds]
    P['crit_prob'][:] = progs['crit_probs'][inds]
    P['death_prob'][:] = progs['death_probs'][inds]
    P['rel_sus'][:] = progs['sus_ORs'][inds]
    P['rel_trans'][:] = progs['trans_ORs'][inds] * sample(rng.np_, rng.nb, size=len(inds), **pars['beta_dist'])


# =================================================================================================
# The numeric kernels (reference utils.py:39-147)
# =================================================================================================

def compute_viral_load(t, time_start, time_recovered, time_dead, frac_time, load_ratio, high_cap):
    '''
    Two-level viral load (reference utils.py:39-79).  float32 arithmetic except the early/late
    comparison, which the reference evaluates in float64 (int32 - float32 promotes): SURVEY App. C.
    '''
    frac_time, load_ratio, high_cap = f32(frac_time), f32(load_ratio), f32(high_cap)
    with np.errstate(all='ignore'):
        stop = np.where(np.isnan(time_dead), time_recovered, time_dead).astype(f32)
        total = (stop - time_start).astype(f32)
        trans_day = (frac_time * total).astype(f32)
        trans_point = np.where(trans_day > high_cap, (high_cap / total).astype(f32), frac_time).astype(f32)
        early = (f64(t) - time_start.astype(f64)) / total.astype(f64) < trans_point.astype(f64)
        one = f32(1)
        denom = f32(one + f32(frac_time * f32(load_ratio - one)))
        load = np.where(early, f32(load_ratio / denom), f32(one / denom)).astype(f32)
    return load


def compute_trans_sus(rel_trans, rel_sus, inf, sus, beta_layer, viral_load, symp, iso, quar,
                      asymp_factor, iso_factor, quar_factor, immunity_factors):
    '''
    Per-agent transmissibility / susceptibility for one (variant, layer) (reference utils.py:82-90).
    rel_trans is a float32 chain evaluated left to right; rel_sus is a float32 product multiplied in
    float64 by (1 - immunity) and rounded to float32.
    '''
    one = f32(1)
    asymp_factor, iso_factor, quar_factor, beta_layer = f32(asymp_factor), f32(iso_factor), f32(quar_factor), f32(beta_layer)
    f_asymp = np.where(symp, one, asymp_factor).astype(f32)
    f_iso = np.where(iso, iso_factor, one).astype(f32)
    f_quar = np.where(quar, quar_factor, one).astype(f32)
    rt = (rel_trans * inf.astype(f32)).astype(f32)
    rt = (rt * f_quar).astype(f32)
    rt = (rt * f_asymp).astype(f32)
    rt = (rt * f_iso).astype(f32)
    rt = (rt * beta_layer).astype(f32)
    rt = (rt * viral_load).astype(f32)
    rs = (rel_sus * sus.astype(f32)).astype(f32)
    rs = (rs * f_quar).astype(f32)
    rs = (rs.astype(f64) * (1.0 - immunity_factors.astype(f64))).astype(f32)
    return rt, rs


def compute_infections(beta, p1, p2, layer_betas, rel_trans, rel_sus, draw):
    '''
    Both directions of every edge: probability, Bernoulli draw, ordered compaction (reference
    utils.py:93-128).  ``draw(direction, edge_inds)`` returns one float64 uniform per surviving edge
    (edges whose float32 probability is exactly zero consume no draw).
    '''
    beta = f32(beta)
    src_out, tgt_out = [], []
    for direction, (sources, targets) in enumerate(((p1, p2), (p2, p1))):
        strans = rel_trans[sources]
        live = np.nonzero(strans)[0]
        prob = (beta * layer_betas[live]).astype(f32)
        prob = (prob * strans[live]).astype(f32)
        prob = (prob * rel_sus[targets[live]]).astype(f32)
        nz = np.nonzero(prob)[0]
        edges = live[nz]
        u = draw(direction, edges)
        hit = np.nonzero(u < prob[nz])[0]
        src_out.append(sources[edges[hit]])
        tgt_out.append(targets[edges[hit]])
    return np.concatenate(src_out).astype(i32), np.concatenate(tgt_out).astype(i32)


def find_contacts(p1, p2, inds):
    ''' Sorted unique partners of ``inds`` over both edge columns (reference utils.py:131-147, base.py:1808-1846) '''
    n = int(max(p1.max(initial=-1), p2.max(initial=-1), np.max(inds, initial=-1))) + 1
    member = np.zeros(n, dtype=bool)
    member[inds] = True
    out = np.zeros(n, dtype=bool)
    out[p2[member[p1]]] = True
    out[p1[member[p2]]] = True
    return np.nonzero(out)[0].astype(i32)


# =================================================================================================
# Immunity (reference immunity.py:138-350)
# =================================================================================================

def update_peak_nab(P, pars, rng, t, inds, nab_pars, symp=None, purpose=ph.P_INFECT, sub=0, slot=9):
    ''' Boost or initialise peak NAbs at an infection / vaccination event (reference immunity.py:138-202) '''
    has = P['nab'][inds] > 0
    prior = inds[has]
    fresh = inds[~has]
    if len(prior):
        boost = nab_pars['nab_boost']
        if isinstance(boost, np.floating) and not isinstance(boost, (np.float32, np.float16)):      # NumPy scalar rules: a float64 scalar (target_eff) multiplies in float64
            P['peak_nab'][prior] = (P['peak_nab'][prior].astype(f64) * f64(boost)).astype(f32)
        else:
            P['peak_nab'][prior] = (P['peak_nab'][prior] * f32(boost)).astype(f32)
    if len(fresh):
        if nab_pars['nab_init'] is None:
            raise ValueError(f'Attempt to administer a vaccine without an initial NAb distribution to {len(fresh)} unvaccinated people failed.')
        level = 2.0 ** rng.sample(t, purpose, sub, fresh, slot, **nab_pars['nab_init'])
        if symp is not None:
            scale = np.full(len(P['nab']), np.nan)
            scale[symp['asymp']] = pars['rel_imm_symp']['asymp']
            scale[symp['mild']] = pars['rel_imm_symp']['mild']
            scale[symp['sev']] = pars['rel_imm_symp']['severe']
            level = level * scale[fresh] * (1 + nab_pars['nab_eff']['alpha_inf_diff'])
        P['peak_nab'][fresh] = level
    P['t_nab_event'][inds] = t


def update_nab(P, pars, t, inds):
    ''' One day of NAb kinetics for agents with a peak NAb (reference immunity.py:205-213) '''
    dt_ = t - P['t_nab_event'][inds]
    peak = P['peak_nab'][inds]
    nab = (P['nab'][inds].astype(f64) + pars['nab_kin'][dt_] * peak.astype(f64)).astype(f32)
    nab = np.where(nab < 0, f32(0), nab)
    nab = np.where(nab > peak, peak, nab)
    P['nab'][inds] = nab


def calc_VE(nab, alpha, beta):
    ''' NAb -> protection, logistic in log-NAb (reference immunity.py:216-247) '''
    with np.errstate(all='ignore'):
        lo = np.exp(alpha) * nab ** beta
        return lo / (1 + lo)


def check_immunity(P, pars, t):
    ''' Per-variant protection factors from NAbs (reference immunity.py:303-350); float64 -> float32 '''
    eff = pars['nab_eff']
    n = len(P['nab'])
    with np.errstate(invalid='ignore'):
        was_inf = np.nonzero(t >= P['date_recovered'])[0]
    is_vacc = np.nonzero(P['vaccinated'])[0]
    for v in range(pars['n_variants']):
        natural = np.zeros(n)
        vaccine = np.zeros(n)
        natural[was_inf] = pars['immunity'][v, :][P['recovered_variant'][was_inf].astype(int)]
        if len(is_vacc) and len(pars['vaccine_pars']):
            vmap = pars['vaccine_map']
            table = np.zeros(max(vmap.keys()) + 1)
            for num, key in vmap.items():
                table[num] = pars['vaccine_pars'][key][pars['variant_map'][v]]
            vaccine[is_vacc] = table[P['vaccine_source'][is_vacc]]
        enab = P['nab'] * np.maximum(natural, vaccine)
        P['sus_imm'][v, :] = calc_VE(enab, eff['alpha_inf'], eff['beta_inf'])
        P['symp_imm'][v, :] = calc_VE(enab, eff['alpha_symp_inf'], eff['beta_symp_inf'])
        P['sev_imm'][v, :] = calc_VE(enab, eff['alpha_sev_symp'], eff['beta_sev_symp'])


def nab_growth_decay(length, growth_time, decay_rate1, decay_time1, decay_rate2, decay_time2):
    ''' Per-day NAb increments: linear growth then slowing exponential decay (reference immunity.py:404-448) '''
    length = length + 1
    t2 = np.arange(length - growth_time, dtype=i32)
    y1 = np.arange(growth_time, dtype=i32) / growth_time
    rate = np.full(len(t2), decay_rate1, dtype=float)
    rate[t2 > decay_time2] = decay_rate2
    mid = np.nonzero((t2 > decay_time1) * (t2 <= decay_time2))[0]
    slowing = (1 / (decay_time2 - decay_time1)) * (decay_rate1 - decay_rate2)
    rate[mid] = decay_rate1 - slowing * np.arange(len(mid), dtype=i32)
    titre = np.zeros(len(t2))
    for i in range(1, len(t2)):
        titre[i] = titre[i - 1] + rate[i]
    y = np.concatenate([y1, np.exp(-titre)])
    return np.diff(y)[0:length]


def init_immunity(pars, npts):
    ''' Cross-immunity matrix and NAb kinetics table (reference immunity.py:269-300) '''
    nv = pars['n_variants']
    if pars['immunity'] is None:
        imm = np.ones((nv, nv), dtype=f32)
        cross = cvpar.get_cross_immunity()
        for i in range(nv):
            li = pars['variant_map'][i]
            for j in range(nv):
                lj = pars['variant_map'][j]
                if li in cross and lj in cross:
                    imm[j][i] = cross[lj][li]
        pars['immunity'] = imm
    decay = dict(pars['nab_decay'])
    form = decay.pop('form')
    if form not in (None, 'nab_growth_decay'):
        raise NotImplementedError(form)
    pars['nab_kin'] = nab_growth_decay(npts, **decay)


# =================================================================================================
# State transitions (reference people.py:164-374)
# =================================================================================================

def _due(P, t, current, date, base=None):
    ''' Agents (optionally within ``base``) whose state is False and whose date is defined and <= t (people.py:211-219) '''
    with np.errstate(invalid='ignore'):
        cond = (~current) & (t >= date)          # NaN compares False
    if base is not None:
        return base[cond[base]]
    return np.nonzero(cond)[0]


def update_states_pre(P, pars, t, flows, vflows):
    ''' Start-of-day transitions in the reference's order (people.py:164-186) '''
    is_exp = np.nonzero(P['exposed'])[0]
    nv = pars['n_variants']
    # infectious (people.py:222-232)
    inds = _due(P, t, P['infectious'], P['date_infectious'], is_exp)
    P['infectious'][inds] = True
    P['infectious_variant'][inds] = P['exposed_variant'][inds]
    for v in range(nv):
        vi = inds[P['infectious_variant'][inds] == v]
        vflows['new_infectious_by_variant'][v] += len(vi)
        P['infectious_by_variant'][v, vi] = True
    flows['new_infectious'] += len(inds)
    # symptomatic / severe / critical (people.py:235-253)
    for state, flow in (('symptomatic', 'new_symptomatic'), ('severe', 'new_severe'), ('critical', 'new_critical')):
        inds = _due(P, t, P[state], P['date_' + state], is_exp)
        P[state][inds] = True
        flows[flow] += len(inds)
    # recovery (people.py:256-291)
    inds = _due(P, t, P['recovered'], P['date_recovered'], is_exp)
    for k in ('exposed', 'infectious', 'symptomatic', 'severe', 'critical'):
        P[k][inds] = False
    P['recovered'][inds] = True
    P['recovered_variant'][inds] = P['exposed_variant'][inds]
    P['infectious_variant'][inds] = np.nan
    P['exposed_variant'][inds] = np.nan
    P['exposed_by_variant'][:, inds] = False
    P['infectious_by_variant'][:, inds] = False
    if pars['use_waning']:
        P['susceptible'][inds] = True
        P['diagnosed'][inds] = False
    flows['new_recoveries'] += len(inds)
    # leave isolation (people.py:368-374)
    inds = _due(P, t, ~P['isolated'], P['date_end_isolation'])
    P['isolated'][inds] = False
    # death (people.py:294-312)
    inds = _due(P, t, P['dead'], P['date_dead'], is_exp)
    P['dead'][inds] = True
    known = inds[P['diagnosed'][inds]]
    P['known_dead'][known] = True
    for k in ('susceptible', 'exposed', 'infectious', 'symptomatic', 'severe', 'critical', 'known_contact',
              'quarantined', 'recovered'):
        P[k][inds] = False
    for k in ('infectious_variant', 'exposed_variant', 'recovered_variant'):
        P[k][inds] = np.nan
    flows['new_deaths'] += len(inds)
    flows['new_known_deaths'] += len(known)
    if pars['use_waning']:
        check_immunity(P, pars, t)


def update_states_post(P, pars, t, flows, pending_quar):
    ''' Transitions after interventions (people.py:189-196, 315-366) '''
    # diagnoses (people.py:315-332)
    pos = _due(P, t, P['diagnosed'], P['date_pos_test'])
    P['date_pos_test'][pos] = np.nan
    diag = _due(P, t, P['diagnosed'], P['date_diagnosed'])
    P['diagnosed'][diag] = True
    flows['new_diagnoses'] += len(pos)
    # quarantine (people.py:335-358); requests are order-independent (max end day, count once)
    n_quar = 0
    for ind, end_day in pending_quar.pop(t, []):
        if P['quarantined'][ind]:
            P['date_end_quarantine'][ind] = max(P['date_end_quarantine'][ind], end_day)
        elif not (P['dead'][ind] or P['recovered'][ind] or P['diagnosed'][ind] or P['isolated'][ind]):
            P['quarantined'][ind] = True
            P['date_quarantined'][ind] = t
            P['date_end_quarantine'][ind] = end_day
            n_quar += 1
    flows['new_quarantined'] += n_quar
    d = np.nonzero(P['quarantined'] & (P['date_diagnosed'] == t))[0]
    P['date_end_quarantine'][d] = t
    rel = _due(P, t, ~P['quarantined'], P['date_end_quarantine'])
    P['quarantined'][rel] = False
    # isolation (people.py:361-366)
    iso = np.nonzero(P['date_diagnosed'] == t)[0]
    P['isolated'][iso] = True
    P['date_end_isolation'][iso] = P['date_recovered'][iso]
    flows['new_isolated'] += len(iso)


def schedule_quarantine(pending_quar, inds, start_date, period):
    ''' Queue quarantine requests (people.py:620-640) '''
    start_date = int(start_date)
    period = int(period)
    lst = pending_quar.setdefault(start_date, [])
    for ind in inds:
        lst.append((ind, start_date + period))


def test_people(P, rng, t, inds, sensitivity, loss_prob, test_delay, sub=0):
    ''' Administer tests (people.py:589-617) '''
    inds = np.unique(inds)
    P['tested'][inds] = True
    P['date_tested'][inds] = t
    is_inf = inds[P['infectious'][inds]]
    pos = rng.agent_uniforms(t, ph.P_TEST_SENS, sub, is_inf) < sensitivity
    is_inf_pos = is_inf[pos]
    not_diag = is_inf_pos[np.isnan(P['date_diagnosed'][is_inf_pos])]
    kept = rng.agent_uniforms(t, ph.P_TEST_LOSS, sub, not_diag) < (1.0 - loss_prob)
    final = not_diag[kept]
    P['date_diagnosed'][final] = t + test_delay
    P['date_pos_test'][final] = t
    return final


def infect(P, pars, rng, t, flows, vflows, log, inds, hosp_max=False, icu_max=False, source=None, layer=None, variant=0):
    '''
    Infect agents and sample their whole disease course (reference people.py:435-586).
    ``inds`` may contain duplicates and non-susceptibles; the first occurrence of a target wins.
    '''
    inds = np.asarray(inds)
    if len(inds) == 0:
        return inds
    inds, first = np.unique(inds, return_index=True)
    if source is not None:
        source = np.asarray(source)[first]
    keep = P['susceptible'][inds]
    inds = inds[keep]
    if source is not None:
        source = source[keep]
    n = len(inds)

    rel = {k: pars[k] for k in ('rel_symp_prob', 'rel_severe_prob', 'rel_crit_prob', 'rel_death_prob')}
    vlabel = pars['variant_map'][variant]
    if variant:
        for k in rel:
            rel[k] *= pars['variant_pars'][vlabel][k]
    dur = pars['dur']

    # breakthrough infections (people.py:486-491)
    bt = inds[P['peak_nab'][inds] != 0]
    if len(bt):
        first_bt = bt[P['n_breakthroughs'][bt] == 0]
        P['rel_trans'][first_bt] = (P['rel_trans'][first_bt] * f32(pars['trans_redux'])).astype(f32)

    # flags and flows (people.py:494-506)
    for k in ('susceptible', 'naive', 'recovered', 'diagnosed'):
        P[k][inds] = False
    P['exposed'][inds] = True
    P['n_infections'][inds] += 1
    P['n_breakthroughs'][bt] += 1
    P['exposed_variant'][inds] = variant
    P['exposed_by_variant'][variant, inds] = True
    flows['new_infections'] += n
    flows['new_reinfections'] += int(np.count_nonzero(~np.isnan(P['date_recovered'][inds])))
    vflows['new_infections_by_variant'][variant] += n
    if log is not None:
        log.append(dict(source=None if source is None else source.astype(i32), target=inds.astype(i32),
                        date=t, layer=layer, variant=variant))

    def draw_dur(key, who, slot):
        return rng.sample(t, ph.P_INFECT, 0, who, slot, **dur[key])

    def draw_u(who, slot):
        return rng.agent_uniforms(t, ph.P_INFECT, 0, who, slot)

    # exposure -> infectious (people.py:513-520)
    P['dur_exp2inf'][inds] = draw_dur('exp2inf', inds, 0)
    P['date_exposed'][inds] = t
    P['date_infectious'][inds] = P['dur_exp2inf'][inds] + f32(t)
    for k in ('date_symptomatic', 'date_severe', 'date_critical', 'date_diagnosed', 'date_recovered'):
        P[k][inds] = np.nan

    # symptomatic? (people.py:522-527)
    p_symp = (f32(rel['rel_symp_prob']) * P['symp_prob'][inds]).astype(f32) * (f32(1) - P['symp_imm'][variant, inds]).astype(f32)
    is_symp = draw_u(inds, 1) < p_symp.astype(f32)
    symp, asymp = inds[is_symp], inds[~is_symp]
    vflows['new_symptomatic_by_variant'][variant] += len(symp)

    # asymptomatic course (people.py:529-532)
    d = draw_dur('asym2rec', asymp, 2)
    P['date_recovered'][asymp] = P['date_infectious'][asymp] + d
    P['dur_disease'][asymp] = P['dur_exp2inf'][asymp] + d

    # symptomatic course (people.py:534-543)
    P['dur_inf2sym'][symp] = draw_dur('inf2sym', symp, 2)
    P['date_symptomatic'][symp] = P['date_infectious'][symp] + P['dur_inf2sym'][symp]
    p_sev = (f32(rel['rel_severe_prob']) * P['severe_prob'][symp]).astype(f32) * (f32(1) - P['sev_imm'][variant, symp]).astype(f32)
    is_sev = draw_u(symp, 3) < p_sev.astype(f32)
    sev, mild = symp[is_sev], symp[~is_sev]
    vflows['new_severe_by_variant'][variant] += len(sev)

    # mild (people.py:545-548)
    d = draw_dur('mild2rec', mild, 4)
    P['date_recovered'][mild] = P['date_symptomatic'][mild] + d
    P['dur_disease'][mild] = P['dur_exp2inf'][mild] + P['dur_inf2sym'][mild] + d

    # severe (people.py:550-556)
    P['dur_sym2sev'][sev] = draw_dur('sym2sev', sev, 4)
    P['date_severe'][sev] = P['date_symptomatic'][sev] + P['dur_sym2sev'][sev]
    p_crit = (f32(rel['rel_crit_prob']) * P['crit_prob'][sev]).astype(f32) * f32(pars['no_hosp_factor'] if hosp_max else 1.0)
    is_crit = draw_u(sev, 5) < p_crit.astype(f32)
    crit, noncrit = sev[is_crit], sev[~is_crit]

    # severe, not critical (people.py:558-561)
    d = draw_dur('sev2rec', noncrit, 6)
    P['date_recovered'][noncrit] = P['date_severe'][noncrit] + d
    P['dur_disease'][noncrit] = P['dur_exp2inf'][noncrit] + P['dur_inf2sym'][noncrit] + P['dur_sym2sev'][noncrit] + d

    # critical (people.py:563-569)
    P['dur_sev2crit'][crit] = draw_dur('sev2crit', crit, 6)
    P['date_critical'][crit] = P['date_severe'][crit] + P['dur_sev2crit'][crit]
    p_death = (f32(rel['rel_death_prob']) * P['death_prob'][crit]).astype(f32) * f32(pars['no_icu_factor'] if icu_max else 1.0)
    is_dead = draw_u(crit, 7) < p_death.astype(f32)
    dead, alive = crit[is_dead], crit[~is_dead]

    # critical, survives (people.py:571-574)
    d = draw_dur('crit2rec', alive, 8)
    P['date_recovered'][alive] = P['date_critical'][alive] + d
    P['dur_disease'][alive] = (P['dur_exp2inf'][alive] + P['dur_inf2sym'][alive] + P['dur_sym2sev'][alive]
                               + P['dur_sev2crit'][alive] + d)

    # critical, dies (people.py:576-580)
    d = draw_dur('crit2die', dead, 8)
    P['date_dead'][dead] = P['date_critical'][dead] + d
    P['dur_disease'][dead] = (P['dur_exp2inf'][dead] + P['dur_inf2sym'][dead] + P['dur_sym2sev'][dead]
                              + P['dur_sev2crit'][dead] + d)
    P['date_recovered'][dead] = np.nan

    if pars['use_waning']:
        update_peak_nab(P, pars, rng, t, inds, pars, symp=dict(asymp=asymp, mild=mild, sev=sev))
    return inds


# =================================================================================================
# Interventions (the callers of the path; reference interventions.py)
# =================================================================================================

class Intervention:
    def initialize(self, sim):
        self.initialized = True

    def apply(self, sim):
        raise NotImplementedError

    def __call__(self, sim):
        return self.apply(sim)


class dynamic_pars(Intervention):
    ''' Set parameters on given days (reference interventions.py:411-479) '''
    def __init__(self, pars=None, **kwargs):
        self.pars = dict(pars or {})
        self.pars.update(kwargs)
        for spec in self.pars.values():
            for sub in ('days', 'vals'):
                if np.isscalar(spec[sub]):
                    spec[sub] = np.atleast_1d(spec[sub])

    def initialize(self, sim):
        pass

    def apply(self, sim):
        for parkey, spec in self.pars.items():
            hit = np.nonzero(np.asarray(spec['days']) == sim.t)[0]
            for ind in hit[:1]:
                val = spec['vals'][ind]
                if isinstance(val, dict):
                    sim.pars[parkey].update(val)
                else:
                    sim.pars[parkey] = val


class sequence(Intervention):
    ''' interventions[i] is in force from days[i] until days[i+1] (reference interventions.py:482-523) '''
    def __init__(self, days, interventions):
        self.days, self.interventions = list(np.atleast_1d(days)), list(interventions)

    def initialize(self, sim):
        self.days = [sim.day(d) for d in self.days]
        self.days_arr = np.array(self.days + [sim.npts])
        for iv in self.interventions:
            iv.initialize(sim)

    def apply(self, sim):
        hit = np.nonzero(self.days_arr <= sim.t)[0]
        if len(hit):
            self.interventions[hit[-1]].apply(sim)


class change_beta(Intervention):
    ''' Scale beta (overall or per layer) on given days (reference interventions.py:533-586) '''
    def __init__(self, days, changes, layers=None):
        self.days, self.changes, self.layers = days, changes, layers

    def initialize(self, sim):
        self.days = np.sort(np.atleast_1d(np.array([sim.day(d) for d in np.atleast_1d(self.days)])))
        self.changes = np.atleast_1d(np.array(self.changes, dtype=float))
        layers = self.layers if isinstance(self.layers, (list, tuple)) else [self.layers]
        self.orig = {}
        for lk in layers:
            self.orig['overall' if lk is None else lk] = sim.pars['beta'] if lk is None else sim.pars['beta_layer'][lk]

    def apply(self, sim):
        hit = np.nonzero(self.days == sim.t)[0]
        for ind in hit[:1]:
            for lk, b in self.orig.items():
                if lk == 'overall':
                    sim.pars['beta'] = b * self.changes[ind]
                else:
                    sim.pars['beta_layer'][lk] = b * self.changes[ind]


class clip_edges(Intervention):
    ''' Move a fraction of a layer's edges out of the simulation and back (reference interventions.py:589-667) '''
    def __init__(self, days, changes, layers=None):
        self.days, self.changes, self.layers = days, changes, layers

    def initialize(self, sim):
        self.days = np.sort(np.atleast_1d(np.array([sim.day(d) for d in np.atleast_1d(self.days)])))
        self.changes = np.atleast_1d(np.array(self.changes, dtype=float))
        self.layers = list(sim.contacts.keys()) if self.layers is None else ([self.layers] if isinstance(self.layers, str) else list(self.layers))
        self.contacts = {lk: dict(p1=np.zeros(0, dtype=i32), p2=np.zeros(0, dtype=i32), beta=np.zeros(0, dtype=f32)) for lk in self.layers}

    @staticmethod
    def _move(src, dst, inds):                      # base.py:1742-1771 pop_inds + append
        for k in ('p1', 'p2', 'beta'):
            dst[k] = np.concatenate([dst[k], src[k][inds]])
            src[k] = np.delete(src[k], inds)

    def apply(self, sim):
        hit = np.nonzero(self.days == sim.t)[0]
        for ind in hit[:1]:
            for lkey in self.layers:
                s_layer, i_layer = sim.contacts[lkey], self.contacts[lkey]
                n_sim, n_int = len(s_layer['p1']), len(i_layer['p1'])
                n_contacts = n_sim + n_int
                if not n_contacts:
                    continue
                n_to_move = int((n_sim / n_contacts - self.changes[ind]) * n_contacts)
                if n_to_move > 0:
                    self._move(s_layer, i_layer, sim.rng.choose('nb', n_sim, n_to_move))
                else:
                    self._move(i_layer, s_layer, sim.rng.choose('nb', n_int, abs(n_to_move)))


class test_num(Intervention):
    ''' Number-based testing (reference interventions.py:718-854) '''
    def __init__(self, daily_tests, symp_test=100.0, quar_test=1.0, quar_policy=None, sensitivity=1.0, loss_prob=0, test_delay=0,
                 start_day=0, end_day=None, subtarget=None, ili_prev=None, swab_delay=None):
        self.subtarget, self.ili_prev = subtarget, ili_prev
        self.pdf = get_pdf(**swab_delay) if swab_delay else None
        self.daily_tests, self.symp_test, self.quar_test = daily_tests, symp_test, quar_test
        self.quar_policy = quar_policy if quar_policy else 'start'
        self.sensitivity, self.loss_prob, self.test_delay = sensitivity, loss_prob, test_delay
        self.start_day, self.end_day = start_day, end_day

    def initialize(self, sim):
        self.start_day, self.end_day = sim.day(self.start_day), sim.day(self.end_day)
        self.daily_tests = np.array([self.daily_tests] * sim.npts) if np.isscalar(self.daily_tests) else np.asarray(self.daily_tests)
        if self.ili_prev is not None:
            self.ili_prev = np.array([self.ili_prev] * sim.npts) if np.isscalar(self.ili_prev) else np.asarray(self.ili_prev)
        self.index = sim.intervention_index(self)

    def apply(self, sim):
        t, P, pars = sim.t, sim.P, sim.pars
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return
        rel_t = t - self.start_day
        if rel_t >= len(self.daily_tests):
            return
        n_tests = int(np.floor(self.daily_tests[rel_t] / sim.rescale_vec[t] + sim.rng.np_.random_sample()))           # sc.randround
        if not (n_tests and np.isfinite(n_tests)):
            return
        sim.results['new_tests'][t] += n_tests
        n = pars['pop_size']
        probs = np.ones(n)
        if self.pdf is not None and P['symptomatic'].any():                     # interventions.py:812-819: weight by the time since symptom onset
            symp_inds = np.nonzero(P['symptomatic'])[0]
            symp_time, dens, count = swab_terms(self.pdf, t, P['date_symptomatic'], symp_inds)
            probs[symp_inds] *= self.symp_test * (dens * count)               # (`symp_test *= pdf * count`: the product on the right is taken first)
        else:
            probs[P['symptomatic']] *= self.symp_test
        if self.ili_prev is not None and rel_t < len(self.ili_prev):           # interventions.py:823-828: people with other illnesses test like symptomatic ones
            chosen = sim.rng.choose('nb', n, int(self.ili_prev[rel_t] * n))
            probs[np.setdiff1d(chosen, np.nonzero(P['symptomatic'])[0])] *= self.symp_test
        probs[get_quar_mask(P, t, self.quar_policy)] *= self.quar_test
        if self.subtarget is not None:                                          # interventions.py:834-837: weights multiply
            inds = np.asarray(self.subtarget['inds'])
            probs[inds] = probs[inds] * self.subtarget['vals']
        probs[P['diagnosed']] = 0.0
        if sim.rescale_vec[t] / pars['pop_scale'] < 1:                          # interventions.py:838-842
            in_tot = probs.sum() * sim.rescale_vec[t]
            out_tot = pars['pop_size'] * pars['pop_scale'] - sim.rescale_vec[t] * pars['pop_size']
            n_tests = int(np.floor(n_tests * in_tot / (in_tot + out_tot) + sim.rng.np_.random_sample()))
        n_tests = min(n_tests, int((probs != 0).sum()))
        if sim.rng.kind == 'mt':                                                # utils.py:446-483 choose_w on the NumPy stream
            total = probs.sum()
            p = probs / total if total else np.ones(n) / n
            inds = sim.rng.np_.choice(n, int(n_tests), p=p, replace=False)
        else:                                                                   # exponential clocks: the n smallest -log(1-u)/w
            if n_tests <= 0:
                return
            u = sim.rng.agent_uniforms(t, ph.P_TEST, self.index, np.arange(n))
            with np.errstate(divide='ignore'):
                key = np.where(probs > 0, -np.log(1.0 - u) / probs, np.inf)
            inds = np.argpartition(key, n_tests - 1)[:n_tests]
        test_people(P, sim.rng, t, inds, self.sensitivity, self.loss_prob, self.test_delay, sub=self.index)


def get_pdf(dist=None, par1=None, par2=None):
    ''' Symptom-onset-to-swab delay density (reference utils.py:240-268) '''
    import scipy.stats as sps
    if dist in ('None', 'none', None):
        return None
    if dist == 'uniform':
        return sps.uniform(loc=par1, scale=par2)
    if dist == 'lognormal':
        mean = np.log(par1 ** 2 / np.sqrt(par2 + par1 ** 2))
        sigma = np.sqrt(np.log(par2 / par1 ** 2 + 1))
        return sps.lognorm(sigma, loc=-0.5, scale=np.exp(mean))
    raise NotImplementedError(f'The selected distribution "{dist}" is not implemented')


def swab_terms(pdf, t, date_symptomatic, symp_inds):
    ''' Days since symptom onset of the symptomatic, the density there, and the inverse share of each delay (interventions.py:812-817, 935-939) '''
    symp_time = (f32(t) - date_symptomatic[symp_inds]).astype(i32)
    inv_count = np.bincount(symp_time) / len(symp_time)
    count = np.nan * np.ones(inv_count.shape)
    count[inv_count != 0] = 1 / inv_count[inv_count != 0]
    return symp_time, pdf.pdf(symp_time), count[symp_time]


def get_quar_mask(P, t, policy):
    ''' interventions.py:691-715 get_quar_inds as a boolean mask '''
    if policy == 'start':
        return P['date_quarantined'] == t - 1
    if policy == 'end':
        return P['date_end_quarantine'] == t + 1
    if policy == 'both':
        return (P['date_quarantined'] == t - 1) | (P['date_end_quarantine'] == t + 1)
    if policy == 'daily':
        return P['quarantined'].copy()
    if isinstance(policy, str):
        raise ValueError(f'Quarantine policy "{policy}" not recognized')
    mask = np.zeros(len(P['quarantined']), dtype=bool)
    if callable(policy):                                        # a function returning the people to test
        mask[np.asarray(policy(P), dtype=int)] = True
        return mask
    for q in np.atleast_1d(policy):                             # days after the start of quarantine on which a test is done
        mask |= P['date_quarantined'] == t - 1 - q
    return mask


class test_prob(Intervention):
    ''' Probability-based testing (reference interventions.py:857-981) '''
    def __init__(self, symp_prob, asymp_prob=0.0, symp_quar_prob=None, asymp_quar_prob=None, quar_policy=None,
                 sensitivity=1.0, loss_prob=0.0, test_delay=0, start_day=0, end_day=None, subtarget=None, ili_prev=None, swab_delay=None):
        self.subtarget, self.ili_prev = subtarget, ili_prev
        self.pdf = get_pdf(**swab_delay) if swab_delay else None
        self.symp_prob, self.asymp_prob = symp_prob, asymp_prob
        self.symp_quar_prob = symp_prob if symp_quar_prob is None else symp_quar_prob
        self.asymp_quar_prob = asymp_prob if asymp_quar_prob is None else asymp_quar_prob
        self.quar_policy = quar_policy if quar_policy else 'start'
        self.sensitivity, self.loss_prob, self.test_delay = sensitivity, loss_prob, test_delay
        self.start_day, self.end_day = start_day, end_day

    def initialize(self, sim):
        self.start_day = sim.day(self.start_day)
        self.end_day = sim.day(self.end_day)
        self.index = sim.intervention_index(self)
        if self.ili_prev is not None:
            self.ili_prev = np.array([self.ili_prev] * sim.npts) if np.isscalar(self.ili_prev) else np.asarray(self.ili_prev)

    def apply(self, sim):
        t, P = sim.t, sim.P
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return
        n = len(P['uid'])
        symp = P['symptomatic']
        ili = np.zeros(n, dtype=bool)                           # interventions.py:946-953: ILI symptoms, independent of COVID
        if self.ili_prev is not None and t - self.start_day < len(self.ili_prev):
            chosen = sim.rng.choose('nb', n, int(self.ili_prev[t - self.start_day] * n))
            ili[chosen] = True
            ili &= ~symp
        qt = get_quar_mask(P, t, self.quar_policy)
        probs = np.where(symp, self.symp_prob, self.asymp_prob).astype(float)
        if self.pdf is not None and symp.any():                 # interventions.py:934-943: symptomatic people test by the time since onset
            symp_inds = np.nonzero(symp)[0]
            symp_time, dens, count = swab_terms(self.pdf, t, P['date_symptomatic'], symp_inds)
            sp = np.ones(len(symp_time))
            early = 1 > (symp_time * self.symp_prob)
            sp[early] = self.symp_prob / (1 - symp_time[early] * self.symp_prob)
            probs[symp_inds] = dens * sp * count
        probs[qt & symp] = self.symp_quar_prob
        probs[qt & ~symp] = self.asymp_quar_prob
        probs[ili] = self.symp_prob                             # ILI people test like symptomatic ones, in quarantine or not (:962-967)
        if self.subtarget is not None:                          # interventions.py:971-973: explicit probabilities win
            probs[np.asarray(self.subtarget['inds'])] = self.subtarget['vals']
        probs[P['diagnosed']] = 0.0
        everyone = np.arange(n)
        tested = np.nonzero(sim.rng.agent_uniforms(t, ph.P_TEST, self.index, everyone) < probs)[0]
        test_people(P, sim.rng, t, tested, self.sensitivity, self.loss_prob, self.test_delay, sub=self.index)
        sim.results['new_tests'][t] += len(tested) * sim.pars['pop_scale'] / sim.rescale_vec[t]


class contact_tracing(Intervention):
    ''' Trace contacts of newly diagnosed agents and schedule quarantine (reference interventions.py:984-1145) '''
    def __init__(self, trace_probs=None, trace_time=None, start_day=0, end_day=None, presumptive=False, quar_period=None, capacity=None):
        self.capacity = capacity
        self.trace_probs, self.trace_time = trace_probs, trace_time
        self.start_day, self.end_day, self.presumptive, self.quar_period = start_day, end_day, presumptive, quar_period

    def initialize(self, sim):
        self.start_day = sim.day(self.start_day)
        self.end_day = sim.day(self.end_day)
        lkeys = list(sim.contacts.keys())
        tp = 1.0 if self.trace_probs is None else self.trace_probs
        tt = 0.0 if self.trace_time is None else self.trace_time
        self.trace_probs = dict(tp) if isinstance(tp, dict) else {k: tp for k in lkeys}
        self.trace_time = dict(tt) if isinstance(tt, dict) else {k: tt for k in lkeys}
        if self.quar_period is None:
            self.quar_period = sim.pars['quar_period']
        self.index = sim.intervention_index(self)

    def apply(self, sim):
        t, P = sim.t, sim.P
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return
        if not self.presumptive:
            cases = np.nonzero(P['date_diagnosed'] == t)[0]
        else:
            just = np.nonzero(P['date_tested'] == t)[0]
            cases = just[P['exposed'][just]]
        if self.capacity is not None:                           # interventions.py:1079-1083: at most `capacity` cases, picked at random
            cap = int(self.capacity / sim.rescale_vec[t])
            if len(cases) > cap:
                cases = sim.rng.np_.choice(cases, cap, replace=False) if sim.rng.kind == 'mt' else cases[sim.rng.choose('np_', len(cases), cap)]
        if not len(cases):
            return
        by_time = {}
        for lkey, prob in self.trace_probs.items():
            if prob == 0:
                continue
            layer = sim.contacts[lkey]
            found = find_contacts(layer['p1'], layer['p2'], cases)
            if len(found):
                lidx = sim.layer_index(lkey)
                u = sim.rng.agent_uniforms(t, ph.P_TRACE, (self.index << 8) | lidx, found)
                by_time.setdefault(self.trace_time[lkey], []).extend(found[np.nonzero(u < prob)[0]])
        dead = np.nonzero(P['dead'])[0]
        for trace_time, who in by_time.items():
            who = np.setdiff1d(np.array(who, dtype=i32), dead)
            P['known_contact'][who] = True
            P['date_known_contact'][who] = np.fmin(P['date_known_contact'][who], f32(t + trace_time))
            schedule_quarantine(sim.pending_quar, who, t + trace_time, self.quar_period - trace_time)


class vaccinate_prob(Intervention):
    ''' Probability-based vaccination with optional second dose (reference interventions.py:1257-1662) '''
    def __init__(self, vaccine, days, label=None, prob=1.0, booster=False, subtarget=None):
        self.vaccine, self.days, self.label, self.prob, self.booster = vaccine, days, label, prob, booster
        self.subtarget = subtarget

    def initialize(self, sim):
        if isinstance(self.vaccine, str):
            _, mapping = cvpar.get_vaccine_choices()
            key = self.vaccine.lower()
            for txt in ['.', ' ', '&', '-', 'vaccine']:
                key = key.replace(txt, '')
            key = mapping[key]
            self.p = dict(cvpar.get_vaccine_variant_pars(vaccine=key))
            self.p.update(cvpar.get_vaccine_dose_pars(vaccine=key))
            if self.label is None:
                self.label = key
        else:
            self.p = copy.deepcopy(self.vaccine)
            if self.label is None:
                self.label = 'custom'
        for k, v in cvpar.get_vaccine_dose_pars(default=True).items():
            self.p.setdefault(k, v)
        dflt = cvpar.get_vaccine_variant_pars(default=True)
        for k in sim.pars['variant_pars'].keys():
            self.p.setdefault(k, dflt.get(k, 1.0))
        if 'target_eff' in self.p:                              # interventions.py:1383-1397: the NAb level that gives the wanted efficacy
            assert self.p['doses'] == len(self.p['target_eff']), 'Provided mismatching efficacies and doses.'
            nabs = np.arange(-8, 4, 0.1)
            ne = sim.pars['nab_eff']
            lo_inf = np.exp(ne['alpha_inf']) * (2 ** nabs) ** ne['beta_inf']                   # immunity.py:250-262 calc_VE_symp
            lo_symp = np.exp(ne['alpha_symp_inf']) * (2 ** nabs) ** ne['beta_symp_inf']
            ve_symp = 1 - ((1 - lo_inf / (1 + lo_inf)) * (1 - lo_symp / (1 + lo_symp)))
            peak = nabs[np.argmax(ve_symp > self.p['target_eff'][0])]
            self.p['nab_init'] = dict(dist='normal', par1=peak, par2=2)
            if self.p['doses'] == 2:
                boosted = nabs[np.argmax(ve_symp > self.p['target_eff'][1])]
                self.p['nab_boost'] = (2 ** boosted) / (2 ** peak)
        n = len(sim.P['uid'])
        self.doses = np.zeros(n, dtype=i32)
        sim.pars['vaccine_pars'][self.label] = self.p
        self.index = list(sim.pars['vaccine_pars'].keys()).index(self.label)
        sim.pars['vaccine_map'][self.index] = self.label
        self.days = np.sort(np.atleast_1d(np.array([sim.day(d) for d in np.atleast_1d(self.days)])))
        self.second_dose_days = [None] * sim.npts
        self.iindex = sim.intervention_index(self)

    def select_people(self, sim):
        t, P = sim.t, sim.P
        picked = np.array([], dtype=int)
        if t >= np.min(self.days):
            if np.any(self.days == t):
                n = len(P['uid'])
                probs = np.zeros(n)
                eligible = P['vaccinated'] if self.booster else ~P['vaccinated']
                probs[eligible] = self.prob
                if self.subtarget is not None:                  # interventions.py:1644-1647: explicit probabilities win
                    probs[np.asarray(self.subtarget['inds'])] = self.subtarget['vals']
                picked = np.nonzero(sim.rng.agent_uniforms(t, ph.P_VACC, self.iindex, np.arange(n)) < probs)[0]
                if len(picked) and self.p['interval'] is not None:
                    nxt = t + self.p['interval']
                    if nxt < sim.pars['n_days']:
                        self.second_dose_days[nxt] = picked
            second = self.second_dose_days[t]
            if second is not None:
                picked = np.concatenate((picked, second), axis=None)
        return picked

    def apply(self, sim):
        t, P = sim.t, sim.P
        inds = self.select_people(sim)
        if not len(inds):
            return
        inds = inds[~P['dead'][inds]]
        inds = inds[self.doses[inds] < self.p['doses']]
        new_vacc = np.setdiff1d(inds, np.nonzero(P['vaccinated'])[0])
        if len(inds):
            self.doses[inds] += 1
            P['vaccinated'][inds] = True
            P['vaccine_source'][inds] = self.index
            P['doses'][inds] += 1
            P['date_vaccinated'][inds] = t
            update_peak_nab(P, sim.pars, sim.rng, t, inds, self.p, symp=None, purpose=ph.P_NAB_VACC, sub=self.iindex, slot=0)
            factor = sim.pars['pop_scale'] / sim.rescale_vec[t]
            sim.flows['new_doses'] += len(inds) * factor
            sim.flows['new_vaccinated'] += len(new_vacc) * factor


class vaccinate_num(vaccinate_prob):
    '''
    A number of doses per day handed out along a priority sequence, second doses first (reference interventions.py:1665-1791).
    mt mode restates the reference literally, including its day-keyed Python sets of scheduled second doses (whose iteration
    order decides the order of the shuffle and of the NAb draws).  Native-RNG mode: the same decisions with keyed draws --
    deferral picks the ``num_agents`` scheduled people with the smallest uniform (P_VACC, slot 1), first doses go to the first
    eligible people of the sequence whose uniform (P_VACC, slot 0) is below their weight.
    '''
    def __init__(self, vaccine, num_doses, booster=False, subtarget=None, sequence=None, label=None):
        vaccinate_prob.__init__(self, vaccine, days=0, label=label, prob=1.0, booster=booster, subtarget=subtarget)
        self.num_doses, self.sequence = num_doses, sequence

    def initialize(self, sim):
        vaccinate_prob.initialize(self, sim)
        n = len(sim.P['uid'])
        if isinstance(self.num_doses, dict):
            self.num_doses = {sim.day(k): v for k, v in self.num_doses.items()}
        if callable(self.sequence):                             # interventions.py:1539-1552 process_sequence
            self.sequence = np.asarray(self.sequence(sim.P))
        elif isinstance(self.sequence, str) and self.sequence == 'age':
            self.sequence = np.argsort(-sim.P['age']) if sim.rng.kind == 'mt' else np.argsort(-sim.P['age'], kind='stable')
        elif self.sequence is None:
            self.sequence = sim.rng.np_.permutation(n)
        else:
            self.sequence = np.asarray(self.sequence)
        self.days = np.array([0])                               # (acts from day 0 on)
        self._scheduled = {}                                    # mt mode: day -> set, like the reference's ddict(set)
        self.due_day = np.full(n, -1, dtype=np.int64)           # native mode: the day an agent's second dose is due

    def _sched(self, day):
        return self._scheduled.setdefault(day, set())

    def n_today(self, sim):                                     # interventions.py:1526-1536 process_doses
        if np.isscalar(self.num_doses):
            return self.num_doses
        if callable(self.num_doses):
            return self.num_doses(sim)
        return self.num_doses.get(sim.t, 0)

    def select_people(self, sim):
        t, P = sim.t, sim.P
        n = len(P['uid'])
        native = sim.rng.kind != 'mt'
        num_people = self.n_today(sim)
        if num_people == 0:                                     # defer everyone due today
            if native:
                self.due_day[self.due_day == t] = t + 1
            else:
                self._sched(t + 1).update(self._sched(t))
            return np.array([], dtype=int)
        num_agents = int(np.floor(num_people / sim.pars['pop_scale'] + sim.rng.np_.random_sample()))     # sc.randround
        if native:
            scheduled = np.nonzero((self.due_day == t) & (self.doses < self.p['doses']) & ~P['dead'])[0]
            if len(scheduled) > num_agents:
                u = sim.rng.agent_uniforms(t, ph.P_VACC, self.iindex, scheduled, slot=1)
                order = np.argsort(u, kind='stable')
                self.due_day[scheduled[order[num_agents:]]] = t + 1
                return scheduled[order[:num_agents]]
        elif self._sched(t):
            scheduled = np.fromiter(self._sched(t), dtype=i32)
            scheduled = scheduled[(self.doses[scheduled] < self.p['doses']) & ~P['dead'][scheduled]]
            if len(scheduled) > num_agents:
                sim.rng.np_.shuffle(scheduled)
                self._sched(t + 1).update(scheduled[num_agents:])
                return scheduled[:num_agents]
        else:
            scheduled = np.array([], dtype=i32)
        probs = np.ones(n)
        probs[P['dead']] = 0.0
        if self.subtarget is not None:                          # interventions.py:1745-1747: weights multiply
            inds = np.asarray(self.subtarget['inds'])
            probs[inds] = probs[inds] * self.subtarget['vals']
        if self.booster:
            probs[~P['vaccinated']] = 0.0
        else:
            probs[P['vaccinated']] = 0.0
        if native:
            mask = sim.rng.agent_uniforms(t, ph.P_VACC, self.iindex, np.arange(n)) < probs
            eligible = self.sequence[mask[self.sequence]]
        else:
            eligible = self.sequence[sim.rng.np_.random_sample(n) < probs[self.sequence]]
        if len(eligible) == 0:
            return scheduled
        eligible = eligible[:num_agents]
        eligible = eligible[~np.isin(eligible, scheduled)]
        first = eligible[:num_agents - len(scheduled)] if len(eligible) + len(scheduled) > num_agents else eligible
        if self.p['doses'] > 1:
            if native:
                self.due_day[first] = t + self.p['interval']
            else:
                self._sched(t + self.p['interval']).update(first)
        return np.concatenate([scheduled, first])


def vaccinate(*args, **kwargs):
    return vaccinate_num(*args, **kwargs) if 'num_doses' in kwargs else vaccinate_prob(*args, **kwargs)


class variant:
    ''' A variant introduced by importation on given days (reference immunity.py:18-130) '''
    def __init__(self, variant, days, label=None, n_imports=1, rescale=True):
        self.days, self.n_imports, self.rescale = days, int(n_imports), rescale
        if isinstance(variant, str):
            _, mapping = cvpar.get_variant_choices()
            key = variant.lower()
            for txt in ['.', ' ', 'variant', 'voc']:
                key = key.replace(txt, '')
            self.label = mapping[key]
            self.p = dict(cvpar.get_variant_pars(variant=self.label))
        else:
            self.p = dict(cvpar.get_variant_pars(default=True))
            self.p.update(variant)
            self.label = self.p.pop('label', label) or 'custom'
        self.index = None

    def initialize(self, sim):
        self.days = np.sort(np.atleast_1d(np.array([sim.day(d) for d in np.atleast_1d(self.days)])))
        sim.pars['variant_pars'][self.label] = self.p
        self.index = list(sim.pars['variant_pars'].keys()).index(self.label)
        sim.pars['variant_map'][self.index] = self.label

    def apply(self, sim):
        if np.any(self.days == sim.t):
            sus = np.nonzero(sim.P['susceptible'])[0]
            scale = sim.rescale_vec[sim.t] if self.rescale else 1.0
            n_imports = int(np.floor(self.n_imports / scale + sim.rng.np_.random_sample()))   # sc.randround
            if sim.rng.kind == 'mt':
                who = sim.rng.np_.choice(sus, n_imports, replace=False)
            else:
                who = sus[sim.rng.choose('np_', len(sus), min(n_imports, len(sus)))]
            sim.infect(who, layer='importation', variant=self.index)
            sim.results['n_imports'][sim.t] += n_imports


# =================================================================================================
# The simulation loop (reference sim.py)
# =================================================================================================

class OracleSim:
    ''' Host-side orchestration of one simulation (reference sim.py:94-125, 558-685, 764-1072) '''

    def __init__(self, pars=None, rng='mt', popdict=None, **kwargs):
        kw = dict(pars or {})
        kw.update(kwargs)
        self.interventions = kw.pop('interventions', [])
        if not isinstance(self.interventions, list):
            self.interventions = [self.interventions]
        self.variants = kw.pop('variants', [])
        if not isinstance(self.variants, list):
            self.variants = [self.variants]
        self.analyzers = kw.pop('analyzers', [])
        self.pars = cvpar.make_pars(**kw)
        self.pars['pop_size'] = int(self.pars['pop_size'])
        self.rng = MTStreams() if rng == 'mt' else PhiloxStreams()
        self.popdict = popdict
        self.initialized = False
        self.t = None
        self.keep_log = True

    # -- date helpers
    def day(self, d):
        if d is None:
            return None
        if isinstance(d, (int, np.integer, float)):
            return int(d)
        if isinstance(d, str):
            d = dt.datetime.strptime(d, '%Y-%m-%d').date()
        start = dt.datetime.strptime(self.pars['start_day'], '%Y-%m-%d').date() if isinstance(self.pars['start_day'], str) else self.pars['start_day']
        return (d - start).days

    def intervention_index(self, obj):
        flat = []                                      # top-level interventions first, then the ones nested in a sequence

        def walk(ivs):
            flat.extend(ivs)
            for iv in ivs:
                walk(list(getattr(iv, 'interventions', None) or []))
        walk(list(self.interventions))
        return [id(i) for i in flat].index(id(obj))

    def layer_index(self, lkey):
        return list(self.contacts.keys()).index(lkey)

    def initialize(self):
        pars = self.pars
        self.npts = int(pars['n_days']) + 1
        self.t = 0
        self.rng.set_seed(pars['rand_seed'])                                    # sim.py:113
        for v in self.variants:                                                # sim.py:354-371
            v.initialize(self)
        pars['n_variants'] = len(pars['variant_pars'])
        if pars['use_waning']:
            init_immunity(pars, self.npts)
        self._init_results()
        if pars['prognoses'] is None:
            pars['prognoses'] = cvpar.get_prognoses(pars['prog_by_age'])
        pop = self.popdict if self.popdict is not None else make_population(pars, self.rng)
        self.P = new_people(pars, pop['age'], pop['sex'])
        self.contacts = {lk: dict(p1=np.array(l['p1'], dtype=i32), p2=np.array(l['p2'], dtype=i32),
                                  beta=np.array(l['beta'], dtype=f32)) for lk, l in pop['contacts'].items()}
        set_prognoses(self.P, pars, self.rng)                                   # people.py:130-161 (re-seeds)
        self.pending_quar = {}
        self.infection_log = []
        self.flows = {k: 0 for k in cvd.new_result_flows}
        self.vflows = {k: np.zeros(pars['n_variants']) for k in cvd.new_result_flows_by_variant}
        self.init_infections()                                                  # sim.py:117 -> :416-417 (inside init_people)
        for iv in self.interventions:
            iv.initialize(self)
        self.rng.set_seed(pars['rand_seed'])                                    # sim.py:121
        self.initialized = True
        self.complete = False
        return self

    def _init_results(self):
        nv, npts = self.pars['n_variants'], self.npts
        R = {}
        for k in cvd.cum_result_flows + cvd.new_result_flows + tuple(f'n_{s}' for s in cvd.result_stocks) + cvd.other_results:
            R[k] = np.zeros(npts)
        V = {}
        for k in ('prevalence_by_variant', 'incidence_by_variant') + cvd.cum_result_flows_by_variant + cvd.new_result_flows_by_variant + tuple(f'n_{s}' for s in cvd.result_stocks_by_variant):
            V[k] = np.zeros((nv, npts))
        R['variant'] = V
        self.results = R
        scale = 1 if self.pars['rescale'] else self.pars['pop_scale']
        self.rescale_vec = scale * np.ones(npts)

    def infect(self, inds, hosp_max=False, icu_max=False, source=None, layer=None, variant=0):
        return infect(self.P, self.pars, self.rng, self.t, self.flows, self.vflows,
                      self.infection_log if self.keep_log else None, inds, hosp_max, icu_max, source, layer, variant)

    def init_infections(self):
        ''' Seed infections (reference sim.py:505-532); draws from the Numba stream '''
        pars = self.pars
        if pars['frac_susceptible'] < 1:                       # sim.py:519-521: a random share of the population is not susceptible
            n = int(np.round((1 - pars['frac_susceptible']) * pars['pop_size']))
            inds = self.rng.choose('nb', pars['pop_size'], n)
            self.make_naive(inds)                              # people.py:412-431 make_nonnaive
            self.P['susceptible'][inds] = False
            self.P['naive'][inds] = False
        if pars['pop_infected']:
            inds = self.rng.choose('nb', pars['pop_size'], int(pars['pop_infected']))
            self.infect(inds, layer='seed_infection')

    def make_naive(self, inds, reset_vx=False):
        ''' Back to the never-infected state (reference people.py:378-409); used by dynamic rescaling '''
        P = self.P
        for key in cvd.states:
            if key in ('susceptible', 'naive'):
                P[key][inds] = True
            elif key != 'vaccinated' or reset_vx:
                P[key][inds] = False
        for key in cvd.variant_states:
            P[key][inds] = np.nan
        for key in cvd.by_variant_states:
            P[key][:, inds] = False
        non_vx = inds if reset_vx else inds[~P['vaccinated'][inds]]
        for key in cvd.imm_states:
            P[key][:, non_vx] = 0
        for key in cvd.nab_states + cvd.vacc_states:
            P[key][non_vx] = 0
        for key in cvd.dates + cvd.durs:
            if key != 'date_vaccinated' or reset_vx:
                P[key][inds] = np.nan

    def rescale(self):
        ''' Dynamic rescaling (reference sim.py:535-555): when too many agents are no longer naive, each agent starts to stand for more people '''
        pars = self.pars
        if not pars['rescale']:
            return
        pop_scale, current = pars['pop_scale'], self.rescale_vec[self.t]
        if current < pop_scale:
            not_naive = np.nonzero(~self.P['naive'])[0]
            n_not_naive, n_people = len(not_naive), pars['pop_size']
            ratio, threshold = n_not_naive / n_people, pars['rescale_threshold']
            if ratio > threshold:
                scaling = min(max(ratio / threshold, pars['rescale_factor']), pop_scale / current)
                self.rescale_vec[self.t:] *= scaling
                n = int(round(n_not_naive * (1.0 - 1.0 / scaling)))
                self.make_naive(not_naive[self.rng.choose('nb', n_not_naive, n)])

    def step(self):
        ''' One simulated day (reference sim.py:558-685) '''
        t, P, pars = self.t, self.P, self.pars
        self.rescale()
        self.flows = {k: 0 for k in cvd.new_result_flows}
        self.vflows = {k: np.zeros(pars['n_variants']) for k in cvd.new_result_flows_by_variant}
        update_states_pre(P, pars, t, self.flows, self.vflows)

        # dynamic layers (people.py:199-206, base.py:1849-1876)
        for lidx, (lkey, dyn) in enumerate(pars['dynam_layer'].items()):
            if dyn:
                self.update_layer(lkey, lidx)

        hosp_max = bool(np.count_nonzero(P['severe']) > pars['n_beds_hosp']) if pars['n_beds_hosp'] is not None else False
        icu_max = bool(np.count_nonzero(P['critical']) > pars['n_beds_icu']) if pars['n_beds_icu'] is not None else False

        if pars['n_imports']:                                                   # sim.py:583-588
            n_imports = int(self.rng.nb.poisson(f32(pars['n_imports'] / self.rescale_vec[t]), 1)[0])
            if n_imports > 0:
                who = self.rng.choose('nb', pars['pop_size'], n_imports)
                self.infect(who, hosp_max, icu_max, layer='importation')
                self.results['n_imports'][t] += n_imports
        for v in self.variants:
            v.apply(self)
        for iv in self.interventions:
            iv(self)
        update_states_post(P, pars, t, self.flows, self.pending_quar)

        vd = pars['viral_dist']
        viral_load = compute_viral_load(t, P['date_infectious'], P['date_recovered'], P['date_dead'],
                                        vd['frac_time'], vd['load_ratio'], vd['high_cap'])
        nv = pars['n_variants']
        for v in range(nv):
            vlabel = pars['variant_map'][v]
            beta = f32(pars['beta'] * pars['rel_beta'] * pars['variant_pars'][vlabel]['rel_beta'])
            with np.errstate(invalid='ignore'):
                inf_v = P['infectious'] & (P['infectious_variant'] == v)
            if not inf_v.any():
                continue
            for lidx, (lkey, layer) in enumerate(self.contacts.items()):
                rt, rs = compute_trans_sus(P['rel_trans'], P['rel_sus'], inf_v, P['susceptible'], pars['beta_layer'][lkey],
                                           viral_load, P['symptomatic'], P['isolated'], P['quarantined'],
                                           pars['asymp_factor'], pars['iso_factor'][lkey], pars['quar_factor'][lkey],
                                           P['sus_imm'][v, :])
                draw = lambda direction, edges, lidx=lidx: self.rng.edge_uniforms(t, lidx, direction, edges)
                src, tgt = compute_infections(beta, layer['p1'], layer['p2'], layer['beta'], rt, rs, draw)
                self.infect(tgt, hosp_max, icu_max, source=src, layer=lkey, variant=v)

        R = self.results
        for k in cvd.result_stocks:
            R[f'n_{k}'][t] = np.count_nonzero(P[k])
        for k in cvd.result_stocks_by_variant:
            for v in range(nv):
                R['variant'][f'n_{k}'][v, t] = np.count_nonzero(P[k][v, :])
        for k, c in self.flows.items():
            R[k][t] += c
        for k, c in self.vflows.items():
            for v in range(nv):
                R['variant'][k][v][t] += c[v]

        if pars['use_waning']:
            has = np.nonzero(P['peak_nab'])[0]
            if len(has):
                update_nab(P, pars, t, has)
        alive = np.nonzero(~P['dead'])[0]
        nab_alive = P['nab'][alive]
        R['pop_nabs'][t] = np.sum(nab_alive[np.nonzero(nab_alive)[0]]) / len(alive)
        R['pop_protection'][t] = np.nanmean(P['sus_imm'])
        R['pop_symp_protection'][t] = np.nanmean(P['symp_imm'])
        for an in self.analyzers:
            an(self)
        self.t += 1
        if self.t == self.npts:
            self.complete = True

    def update_layer(self, lkey, lidx):
        ''' Regenerate a dynamic layer in place (reference base.py:1849-1876); Numba stream '''
        layer = self.contacts[lkey]
        n = len(layer['p1'])
        if self.rng.kind == 'mt':
            inds = self.rng.nb.choice(n, n, replace=False)
            layer['p1'][inds] = np.array(self.rng.nb.choice(self.pars['pop_size'], n, replace=True), dtype=i32)
            layer['p2'][inds] = np.array(self.rng.nb.choice(self.pars['pop_size'], n, replace=True), dtype=i32)
            layer['beta'][inds] = f32(1)
        else:
            u1, u2 = ph.keyed_uniform2(self.rng.seed, ph.P_DYNLAYER, lidx, self.t, np.arange(n))
            layer['p1'][:] = np.minimum((u1 * self.pars['pop_size']).astype(np.int64), self.pars['pop_size'] - 1)
            layer['p2'][:] = np.minimum((u2 * self.pars['pop_size']).astype(np.int64), self.pars['pop_size'] - 1)
            layer['beta'][:] = f32(1)

    def run(self, until=None):
        if not self.initialized:
            self.initialize()
        self.rng.set_seed(self.pars['rand_seed'])                               # sim.py:713-716
        until = self.npts if until is None else until
        while self.t < until:
            self.step()
        if self.complete:
            self.finalize()
        return self

    def finalize(self):
        ''' Cumulative and derived results (reference sim.py:764-1072) '''
        R, pars, P = self.results, self.pars, self.P
        rv = self.rescale_vec
        for k in list(R.keys()):
            if k != 'variant' and k not in cvd.unscaled_results:
                R[k] = R[k] * rv
        for k in list(R['variant'].keys()):
            if k not in ('prevalence_by_variant', 'incidence_by_variant'):
                R['variant'][k] = R['variant'][k] * rv[None, :]
        for k in cvd.result_flows:
            R[f'cum_{k}'] = np.cumsum(R[f'new_{k}'])
        for k in cvd.result_flows_by_variant:
            R['variant'][f'cum_{k}'] = np.cumsum(R['variant'][f'new_{k}'], axis=1)
        R['cum_infections'] = R['cum_infections'] + pars['pop_infected'] * rv[0]
        R['variant']['cum_infections_by_variant'] = R['variant']['cum_infections_by_variant'] + pars['pop_infected'] * rv[0]
        self.t -= 1
        scaled_pop = pars['pop_size'] * pars['pop_scale']
        count_recov = 1 - pars['use_waning']
        with np.errstate(all='ignore'):
            R['n_alive'] = scaled_pop - R['cum_deaths']
            R['n_naive'] = scaled_pop - R['cum_deaths'] - R['n_recovered'] - R['n_exposed']
            R['n_susceptible'] = R['n_alive'] - R['n_exposed'] - count_recov * R['cum_recoveries']
            R['n_preinfectious'] = R['n_exposed'] - R['n_infectious']
            R['n_removed'] = count_recov * R['cum_recoveries'] + R['cum_deaths']
            R['prevalence'] = R['n_exposed'] / R['n_alive']
            R['incidence'] = R['new_infections'] / R['n_susceptible']
            R['frac_vaccinated'] = R['n_vaccinated'] / R['n_alive']
            R['variant']['incidence_by_variant'] = R['variant']['new_infections_by_variant'] / R['n_susceptible'][None, :]
            R['variant']['prevalence_by_variant'] = R['variant']['new_infections_by_variant'] / R['n_alive'][None, :]
            # yield (sim.py:840-855)
            ty = np.zeros(self.npts)
            nz = np.nonzero(R['new_tests'])[0]
            ty[nz] = R['new_diagnoses'][nz] / R['new_tests'][nz]
            R['test_yield'] = ty
            rty = np.zeros(self.npts)
            nz = np.nonzero(R['n_infectious'])[0]
            denom = R['n_infectious'][nz] / (R['n_alive'][nz] - R['cum_diagnoses'][nz])
            rty[nz] = ty[nz] / denom
            R['rel_test_yield'] = rty
            # doubling time (sim.py:858-885)
            window, cap = 3, 30
            ci = R['cum_infections']
            now, prev = ci[window:], ci[:-window]
            use = (prev > 0) & (now > prev)
            dbl = np.full(self.npts, np.nan)
            tail = dbl[window:]
            tail[use] = np.minimum(window * np.log(2) / np.log(now[use] / prev[use]), cap)
            R['doubling_time'] = dbl
            # r_eff, 'daily' method (sim.py:888-946)
            rec = np.nonzero(~np.isnan(P['date_recovered']))[0]
            dead = np.nonzero(~np.isnan(P['date_dead']))[0]
            outcome = np.concatenate((P['date_recovered'][rec], P['date_dead'][dead]))
            both = np.concatenate((rec, dead))
            mean_inf = outcome.mean() - P['date_infectious'][both].mean() if len(outcome) else 0
            new_inf = R['new_infections'] - R['n_imports']
            n_inf = R['n_infectious']
            raw = mean_inf * np.divide(new_inf, n_inf, out=np.zeros(self.npts), where=n_inf > 0)
            if len(raw) >= 3:
                initial = int(min(len(raw), pars['dur']['exp2inf']['par1'] + pars['dur']['asym2rec']['par1']))
                for i in range(initial):
                    raw[i] = raw[i:initial].mean()
                sm = raw.copy()
                for _ in range(2):
                    padded = np.concatenate([[sm[0]], sm, [sm[-1]]])
                    sm = np.convolve(padded, [0.25, 0.5, 0.25], mode='valid')
                sm[:2] = raw[:2]
                sm[-2:] = raw[-2:]
                raw = sm
            R['r_eff'] = raw
        self.summary = {k: float(v[self.t]) for k, v in R.items() if k != 'variant'}
        self.results_ready = True
        return self
