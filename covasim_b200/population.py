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
