'''
Population and contact-network construction (init-time, host side; reference covasim/population.py).

Out of the timed path, but needed to feed it.  Two generators:

* ``exact=True`` consumes the two MT19937 streams in exactly the reference's order
  (population.py:143-236 ages/sexes, :239-283 random layers, :286-329 household cliques,
  :332-364 hybrid), so that a replay-mode sim reproduces the reference's population bit for bit.
  Random layers are built with ``np.repeat`` instead of the reference's per-agent Python loop (same
  result); household cliques still need CPython's set iteration order, hence a Python loop.
* ``exact=False`` draws the same distributions but builds households fully vectorised (edge order
  within a household ascending) -- ~20x faster at 1M agents; this is what benchmarks use.
'''
import numpy as np

from . import defaults as cvd

__all__ = ['make_randpop', 'make_random_contacts', 'make_microstructured_contacts', 'make_hybrid_contacts']

i32 = np.int32
f32 = np.float32


def make_random_contacts(rng, pop_size, n, overshoot=1.2, mapping=None):
    ''' Random layer: each agent p gets round(Poisson(n)/2) edges (p, random other) (population.py:239-283) '''
    pop_size = int(pop_size)
    n_all = int(pop_size * n * overshoot)
    pool = rng.nb.choice(pop_size, n_all, replace=True) if pop_size > 0 else np.zeros(0, dtype=np.int64)
    counts = np.array((rng.nb.poisson(f32(n), pop_size) / 2.0).round(), dtype=i32)
    total = int(counts.sum())
    p1 = np.repeat(np.arange(pop_size, dtype=i32), counts)
    p2 = np.array(pool[:total], dtype=i32)
    if mapping is not None:
        mapping = np.array(mapping, dtype=i32)
        p1, p2 = mapping[p1], mapping[p2]
    return dict(p1=p1, p2=p2)


def _cluster_sizes(rng, pop_size, cluster_size, exact):
    ''' Poisson cluster sizes covering pop_size agents, consuming exactly as many draws as a one-at-a-time loop '''
    sizes = []
    covered = 0
    lam = f32(cluster_size)
    if exact:
        # Draw in blocks, but rewind so that the stream ends exactly after the last draw the loop would make
        while covered < pop_size:
            state = rng.nb.get_state()
            block = max(1024, int((pop_size - covered) / max(float(cluster_size), 0.5) * 1.1) + 16)
            draw = rng.nb.poisson(lam, block)
            csum = covered + np.cumsum(draw)
            hit = np.nonzero(csum >= pop_size)[0]
            if len(hit):
                k = int(hit[0]) + 1
                rng.nb.set_state(state)
                draw = rng.nb.poisson(lam, k)
                sizes.append(draw)
                covered = pop_size
            else:
                sizes.append(draw)
                covered = int(csum[-1])
    else:
        while covered < pop_size:
            block = max(1024, int((pop_size - covered) / max(float(cluster_size), 0.5) * 1.2) + 16)
            draw = rng.nb.poisson(lam, block)
            sizes.append(draw)
            covered += int(draw.sum())
    sizes = np.concatenate(sizes).astype(np.int64)
    ends = np.cumsum(sizes)
    k = int(np.nonzero(ends >= pop_size)[0][0]) + 1
    sizes = sizes[:k].copy()
    sizes[-1] -= ends[k - 1] - pop_size          # the last cluster is truncated (population.py:309-310)
    return sizes


def make_microstructured_contacts(rng, pop_size, cluster_size, exact=True):
    ''' Households: consecutive agents form cliques of Poisson(cluster_size) members (population.py:286-329) '''
    pop_size = int(pop_size)
    sizes = _cluster_sizes(rng, pop_size, cluster_size, exact)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    p1_parts, p2_parts = [], []
    if exact:
        for start, size in zip(starts.tolist(), sizes.tolist()):
            if size < 2:
                continue
            members = range(start, start + size)
            for a in members:
                partners = set()               # the reference appends list(set(...)): CPython set order, not always ascending
                for b in members:
                    if b > a:
                        partners.add(b)
                p1_parts.append(np.full(len(partners), a, dtype=i32))
                p2_parts.append(np.array(list(partners), dtype=i32))
    else:
        for size in np.unique(sizes):
            if size < 2:
                continue
            a, b = np.triu_indices(int(size), k=1)
            base = starts[sizes == size]
            p1_parts.append((base[:, None] + a[None, :]).ravel().astype(i32))
            p2_parts.append((base[:, None] + b[None, :]).ravel().astype(i32))
    if p1_parts:
        p1, p2 = np.concatenate(p1_parts), np.concatenate(p2_parts)
        if not exact:
            order = np.lexsort((p2, p1))
            p1, p2 = p1[order], p2[order]
    else:
        p1, p2 = np.zeros(0, dtype=i32), np.zeros(0, dtype=i32)
    return dict(p1=p1, p2=p2)


def make_hybrid_contacts(rng, pop_size, ages, contacts, school_ages=(6, 22), work_ages=(22, 65), exact=True):
    ''' Households + age-banded random school/work layers + random community layer (population.py:332-364) '''
    nc = dict(h=4, s=20, w=20, c=20)
    nc.update(contacts)
    out = {}
    out['h'] = make_microstructured_contacts(rng, pop_size, nc['h'], exact=exact)       # generation order: h, c, s, w
    out['c'] = make_random_contacts(rng, pop_size, nc['c'])
    ages = np.asarray(ages)
    s_inds = np.nonzero((ages >= school_ages[0]) * (ages < school_ages[1]))[0]
    w_inds = np.nonzero((ages >= work_ages[0]) * (ages < work_ages[1]))[0]
    out['s'] = make_random_contacts(rng, len(s_inds), nc['s'], mapping=s_inds)
    out['w'] = make_random_contacts(rng, len(w_inds), nc['w'], mapping=w_inds)
    return out


def make_randpop(pars, rng, exact=True, sex_ratio=0.5):
    ''' Ages (default age pyramid), sexes and contact layers (population.py:143-236) '''
    n = int(pars['pop_size'])
    sexes = rng.np_.binomial(1, sex_ratio, n)
    age_data = cvd.default_age_data
    lo = age_data[:, 0]
    width = age_data[:, 1] + 1 - lo
    probs = age_data[:, 2] / age_data[:, 2].sum()
    bins = np.searchsorted(np.cumsum(probs), rng.np_.random_sample(n))
    ages = lo[bins] + width[bins] * rng.np_.random_sample(n)
    if pars['pop_type'] == 'random':
        layers = {lk: make_random_contacts(rng, n, nc) for lk, nc in pars['contacts'].items()}
    elif pars['pop_type'] == 'hybrid':
        gen = make_hybrid_contacts(rng, n, ages, pars['contacts'], exact=exact)
        layers = {lk: gen[lk] for lk in pars['contacts'].keys()}                           # stored in parameter order h, s, w, c
    else:
        raise NotImplementedError(f'Population type "{pars["pop_type"]}" is not built (choices: random, hybrid)')
    for layer in layers.values():
        layer['beta'] = np.ones(len(layer['p1']), dtype=f32)
    return dict(uid=np.arange(n, dtype=i32), age=ages, sex=sexes, contacts=layers, layer_keys=list(pars['contacts'].keys()))


# ===================================================================================================
# Device-side generation: the same distributions, drawn from keyed Philox uniforms
# ===================================================================================================
#
# The reference builds a population with Python loops over agents (population.py:143-364: 11.7 s at 1M agents, tens of GB
# of Python lists at 100M).  Here every draw is a pure function of (seed, P_POP, sub, index, slot), so the population can
# be generated ON THE DEVICE, in agent chunks, identically on every rank of an agent-partitioned run, and without ever
# holding all edges at once.  Everything except the uniforms themselves is torch index arithmetic (init-time plumbing);
# the uniforms come from the library's Philox kernel (cvb_keyed_uniform).  oracle/cvoracle.py:make_keyed_pop is the NumPy
# restatement this is tested against, bit for bit (tests/test_gpu_popgen.py, tests/test_popgen_cpu.py).
#
#   agent i          : sex = u(sub 0, i, slot 0) < 0.5;  age bin from u(sub 0, i, slot 1) through the cumulative age
#                      pyramid, age = lo + width * u(sub 0, i, slot 2)            (population.py:186-204)
#   random layer L   : position i of the layer's eligible list gets round(Poisson(n) / 2) edges, Poisson by inverse CDF of
#                      u(sub L, i, slot 0); edge e (running index within the layer) goes to eligible position
#                      floor(u(sub L, e, slot 1) * m)                               (population.py:239-283)
#   household layer  : cluster c has Poisson(n) members, inverse CDF of u(sub H, c, slot 0); consecutive agents; the last
#                      cluster is truncated; all pairs a < b, ordered by (a, b)     (population.py:286-329)

P_POP = 10
SUB_AGENT = 0
_LAYER_SUB = dict(h=1, s=2, w=3, c=4, a=5)


def poisson_cdf(lam):
    ''' Cumulative distribution of Poisson(lam) in float64, long enough that the tail is below 2^-60 '''
    lam = float(lam)
    kmax = int(lam + 12 * np.sqrt(lam + 1) + 40)
    k = np.arange(1, kmax + 1, dtype=np.float64)
    logp = np.concatenate([[-lam], -lam + np.cumsum(np.log(lam) - np.log(k))]) if lam > 0 else np.concatenate([[0.0], np.full(kmax, -np.inf)])
    return np.cumsum(np.exp(logp))


def device_uniforms(seed, device):
    ''' uniforms(sub, index0, n, slot) -> float64[n] on ``device`` from the library's Philox kernel '''
    import torch
    from . import _capi

    def fn(sub, index0, n, slot):
        out = torch.empty(int(n), dtype=torch.float64, device=device)
        if n:
            with torch.cuda.device(device):
                _capi.call('cvb_keyed_uniform', int(seed), P_POP, int(sub), 0, int(index0), int(n), int(slot), out.data_ptr(), None)
        return out
    return fn


class KeyedPop:
    '''
    A population defined by keyed draws.  ``ages`` / ``sexes`` are whole-population device tensors; a layer's edges are
    produced on demand for any range of its eligible positions (``layer_edges``), each with its running index inside the
    layer -- which is also the index the transmission kernels key their per-edge draws on.
    '''

    def __init__(self, pars, seed, device, uniforms, school_ages=(6, 22), work_ages=(22, 65)):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.seed = int(seed)
        self.u = uniforms
        n = self.n = int(pars['pop_size'])
        dev = self.device
        age_data = cvd.default_age_data
        lo = torch.as_tensor(age_data[:, 0].astype(np.float64), device=dev)
        width = torch.as_tensor((age_data[:, 1] + 1 - age_data[:, 0]).astype(np.float64), device=dev)
        probs = age_data[:, 2] / age_data[:, 2].sum()
        cum = torch.as_tensor(np.cumsum(probs), device=dev)
        self.sexes = (self.u(SUB_AGENT, 0, n, 0) < 0.5).to(torch.int32)
        bins = torch.searchsorted(cum, self.u(SUB_AGENT, 0, n, 1)).clamp_(max=len(probs) - 1)
        self.ages = lo[bins] + width[bins] * self.u(SUB_AGENT, 0, n, 2)
        self.plans = {}
        contacts = pars['contacts']
        if pars['pop_type'] == 'random':
            for lk, nc in contacts.items():
                self.plans[lk] = self._plan_random(lk, None, nc)
        elif pars['pop_type'] == 'hybrid':
            nc = dict(h=4, s=20, w=20, c=20)
            nc.update(contacts)
            ages32 = self.ages.to(torch.float32)              # the People array is float32; band membership uses what the sim sees
            for lk in contacts.keys():                         # parameter order h, s, w, c = layer index order
                if lk == 'h':
                    self.plans[lk] = self._plan_households(nc['h'])
                elif lk == 'c':
                    self.plans[lk] = self._plan_random(lk, None, nc['c'])
                elif lk in ('s', 'w'):
                    a0, a1 = school_ages if lk == 's' else work_ages
                    elig = torch.nonzero((ages32 >= a0) & (ages32 < a1)).flatten()
                    self.plans[lk] = self._plan_random(lk, elig, nc[lk])
                else:
                    raise NotImplementedError(f'hybrid populations have layers h, s, w, c; got "{lk}"')
        else:
            raise NotImplementedError(f'Population type "{pars["pop_type"]}" is not built (choices: random, hybrid)')

    # ---- plans: per eligible position, how many edges it starts and where they sit in the layer ----------------
    def _poisson(self, lam, sub, index0, n):
        torch = self.torch
        cdf = torch.as_tensor(poisson_cdf(lam), device=self.device)
        return torch.searchsorted(cdf, self.u(sub, index0, n, 0), right=True)

    def _finish_plan(self, plan, counts):
        torch = self.torch
        offsets = torch.zeros(len(counts) + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        plan.update(counts=counts, offsets=offsets, n_edges=int(offsets[-1].item()))
        if plan['n_edges'] >= 2 ** 31:
            raise ValueError(f'layer with {plan["n_edges"]} edges: at most 2^31 - 1 edges per layer')
        return plan

    def _plan_random(self, lk, mapping, n_contacts):
        torch = self.torch
        m = self.n if mapping is None else int(mapping.numel())
        sub = _LAYER_SUB.get(lk, 5)
        counts = torch.round(self._poisson(n_contacts, sub, 0, m).to(torch.float64) / 2.0).to(torch.int64)      # half to even, like np.round
        return self._finish_plan(dict(kind='random', sub=sub, mapping=mapping, m=m), counts)

    def _plan_households(self, cluster_size):
        torch = self.torch
        n, sub = self.n, _LAYER_SUB['h']
        sizes, covered, c0 = [], 0, 0
        while covered < n:
            block = max(1024, int((n - covered) / max(float(cluster_size), 0.5) * 1.2) + 16)
            draw = self._poisson(cluster_size, sub, c0, block)
            sizes.append(draw)
            covered += int(draw.sum().item())
            c0 += block
        ends = torch.cumsum(torch.cat(sizes), 0)
        k = int(torch.searchsorted(ends, torch.tensor([n], device=self.device, dtype=ends.dtype))[0].item()) + 1      # first end >= n
        ends = ends[:k].clone()
        ends[-1] = n                                            # the last cluster is truncated (population.py:309-310)
        agents = torch.arange(n, dtype=torch.int64, device=self.device)
        cluster_end = ends[torch.searchsorted(ends, agents, right=True)]
        counts = cluster_end - agents - 1                       # agent a pairs with a+1 .. end-1
        return self._finish_plan(dict(kind='households', sub=sub, mapping=None, m=n, ends=ends), counts)

    # ---- edges -------------------------------------------------------------------------------------------
    def layer_edges(self, lk, i0=0, i1=None):
        '''
        Edges started by eligible positions [i0, i1) of layer ``lk``: (p1:int32, p2:int32, e0) where e0 is the running
        index of the first of them inside the layer (edges are ordered by position, then by draw).
        '''
        torch = self.torch
        plan = self.plans[lk]
        i1 = plan['m'] if i1 is None else min(int(i1), plan['m'])
        counts = plan['counts'][i0:i1]
        e0, e1 = int(plan['offsets'][i0].item()), int(plan['offsets'][i1].item())
        pos = torch.repeat_interleave(torch.arange(i0, i1, dtype=torch.int64, device=self.device), counts, output_size=e1 - e0)
        if plan['kind'] == 'households':
            run = torch.arange(e0, e1, dtype=torch.int64, device=self.device) - plan['offsets'][pos]
            p1, p2 = pos, pos + 1 + run
        else:
            m = plan['m']
            tpos = torch.floor(self.u(plan['sub'], e0, e1 - e0, 1) * m).to(torch.int64).clamp_(max=m - 1)
            if plan['mapping'] is None:
                p1, p2 = pos, tpos
            else:
                p1, p2 = plan['mapping'][pos], plan['mapping'][tpos]
        return p1.to(torch.int32), p2.to(torch.int32), e0

    def layer_keys(self):
        return list(self.plans.keys())

    def materialize(self):
        ''' The whole population in make_randpop's format (device tensors) '''
        torch = self.torch
        contacts = {}
        for lk in self.plans:
            p1, p2, _ = self.layer_edges(lk)
            contacts[lk] = dict(p1=p1, p2=p2, beta=torch.ones(p1.numel(), dtype=torch.float32, device=self.device))
        return dict(age=self.ages, sex=self.sexes, contacts=contacts, layer_keys=list(self.plans.keys()))
