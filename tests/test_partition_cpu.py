'''
CPU tests of the host logic of the agent-partitioned run (covasim_b200/partition.py): the partition plan, the
partitioned adjacency against a brute-force construction, and the two communicators -- torch.distributed with a
world_size-2 gloo group, and the in-process LocalComm.  The kernels themselves are covered on the GPU
(tests/test_gpu_partition.py).
'''
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan():
    from covasim_b200 import partition as cvpart
    chunk, ranges = cvpart.plan(1_000_000, 8)
    assert chunk == 125_024 and chunk % 32 == 0
    assert ranges[0] == (0, chunk) and ranges[-1] == (7 * chunk, 1_000_000)
    assert all(hi - lo == chunk for lo, hi in ranges[:-1])
    chunk, ranges = cvpart.plan(5003, 3)
    assert chunk == 1696 and ranges == [(0, 1696), (1696, 3392), (3392, 5003)]
    with pytest.raises(ValueError):
        cvpart.plan(64, 3)


def random_layers(rng, n, sizes):
    out = []
    for e in sizes:
        out.append(dict(p1=np.sort(rng.randint(0, n, e)).astype(np.int32), p2=rng.randint(0, n, e).astype(np.int32), beta=rng.random_sample(e).astype(np.float32)))
    return out


def brute_rows(layers, ids, lo, hi):
    ''' {source: sorted list of (local target, edge, meta, beta bits)} by plain loops '''
    rows = {}
    for l, layer in zip(ids, layers):
        for e, (a, b, w) in enumerate(zip(layer['p1'], layer['p2'], layer['beta'])):
            wb = int(np.float32(w).view(np.int32))
            if lo <= b < hi:
                rows.setdefault(int(a), []).append((int(b) - lo, e, (l << 1) | 0, wb))
            if lo <= a < hi:
                rows.setdefault(int(b), []).append((int(a) - lo, e, (l << 1) | 1, wb))
    return {k: sorted(v) for k, v in rows.items()}


@pytest.mark.parametrize('world', [1, 2, 3])
def test_partition_adjacency_matches_brute_force(world):
    from covasim_b200 import partition as cvpart
    rng = np.random.RandomState(world)
    n = 700
    layers = random_layers(rng, n, [900, 0, 1500])
    layers[0]['p2'][:5] = layers[0]['p1'][:5]                 # self-loops and duplicate edges are legal
    layers[2]['p1'][10:14] = layers[2]['p1'][10]
    layers[2]['p2'][10:14] = layers[2]['p2'][10]
    ids = [0, 2, 3]
    chunk, ranges = cvpart.plan(n, world)
    total = 0
    for lo, hi in ranges:
        ptr, adj, M = cvpart.build_partition_adjacency(layers, ids, lo, hi, world * chunk, 'cpu')
        ptr, adj = ptr.numpy(), adj.numpy()
        assert ptr.shape == (world * chunk + 1,) and ptr[-1] == M
        want = brute_rows(layers, ids, lo, hi)
        for src in range(world * chunk):
            got = sorted(tuple(int(x) for x in row) for row in adj[ptr[src]:ptr[src + 1]])
            assert got == want.get(src, []), f'row {src} of range [{lo},{hi})'
        total += M
    assert total == 2 * sum(len(l['p1']) for l in layers)     # every directed edge lives on exactly one rank


def _dist_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from covasim_b200 import partition as cvpart
    comm = cvpart.DistComm()
    chunk = 64
    codes_local = torch.full((chunk,), rank + 1, dtype=torch.uint8)
    codes_global = torch.zeros(world * chunk, dtype=torch.uint8)
    comm.all_gather(codes_global, codes_local)
    bits_local = torch.full((chunk // 32,), 100 + rank, dtype=torch.int32)
    bits_global = torch.zeros(world * chunk // 32, dtype=torch.int32)
    comm.all_gather(bits_global, bits_local)
    table = torch.arange(6, dtype=torch.int64).reshape(2, 3) * (rank + 1)
    comm.all_reduce_sum(table)
    objs = comm.gather_objects(dict(rank=rank, ids=np.arange(rank + 2)))
    peer = comm.peer_exchange(torch.device('cpu'), dict(codes=chunk, cases=chunk // 8))      # gloo: no peer memory -> every rank agrees on the fallback
    assert peer is None
    q.put((rank, codes_global.numpy(), bits_global.numpy(), table.numpy(), [o['rank'] for o in objs], [len(o['ids']) for o in objs]))
    dist.barrier()
    dist.destroy_process_group()


def test_distcomm_gloo_world2():
    world = 2
    port = 29600 + (os.getpid() % 2000)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, codes, bits, table, ranks, lens in got:
        assert np.array_equal(codes, np.repeat([1, 2], 64))
        assert np.array_equal(bits, np.repeat([100, 101], 2))
        assert np.array_equal(table, np.arange(6).reshape(2, 3) * 3)
        assert ranks == [0, 1] and lens == [2, 3]


def test_localcomm_threads():
    import threading
    from covasim_b200 import partition as cvpart
    world = 3
    comms = cvpart.LocalComm.make(world)
    out = [None] * world

    def work(r):
        c = comms[r]
        res = []
        for day in range(20):                                  # many rounds: the barriers must keep the ranks in step
            loc = torch.full((32,), 10 * day + r, dtype=torch.uint8)
            glob = torch.zeros(32 * world, dtype=torch.uint8)
            c.all_gather(glob, loc)
            res.append(glob.numpy().copy())
        t = torch.tensor([r + 1.0, 2.0 * r], dtype=torch.float64)
        c.all_reduce_sum(t)
        objs = c.gather_objects(('rank', r))
        out[r] = (res, t.numpy(), objs)
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    for r in range(world):
        res, t, objs = out[r]
        for day, g in enumerate(res):
            assert np.array_equal(g, np.repeat([10 * day, 10 * day + 1, 10 * day + 2], 32))
        assert np.array_equal(t, [6.0, 6.0])
        assert objs == [('rank', 0), ('rank', 1), ('rank', 2)]


# ---- properties over many shapes (hypothesis): ragged sizes, empty layers, self-loops, any world size ---------------------------------
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(n=st.integers(33, 3000), world=st.integers(1, 8), sizes=st.lists(st.integers(0, 400), min_size=1, max_size=4), seed=st.integers(0, 10**6))
def test_partition_properties(n, world, sizes, seed):
    from covasim_b200 import partition as cvpart
    try:
        chunk, ranges = cvpart.plan(n, world)
    except ValueError:
        assert (world - 1) * (-(-(-(-n // world)) // 32) * 32) >= n          # only when the last rank would own nobody
        return
    assert chunk % 32 == 0 and ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:])) and all(hi - lo == chunk for lo, hi in ranges[:-1])
    rng = np.random.RandomState(seed)
    layers = random_layers(rng, n, sizes)
    ids = list(range(len(sizes)))
    seen = 0
    for lo, hi in ranges:
        ptr, adj, M = cvpart.build_partition_adjacency(layers, ids, lo, hi, world * chunk, 'cpu')
        ptr, adj = ptr.numpy(), adj.numpy()[:M]
        assert ptr[0] == 0 and ptr[-1] == M and np.all(np.diff(ptr) >= 0)
        assert np.all((adj[:, 0] >= 0) & (adj[:, 0] < hi - lo))               # targets are local
        src = np.repeat(np.arange(world * chunk), np.diff(ptr))
        for l, layer in zip(ids, layers):                                      # every entry is a real edge, seen from the right side
            mine = (adj[:, 2] >> 1) == l
            e, d = adj[mine, 1], adj[mine, 2] & 1
            p_src = np.where(d == 0, layer['p1'][e], layer['p2'][e])
            p_tgt = np.where(d == 0, layer['p2'][e], layer['p1'][e])
            assert np.array_equal(p_src, src[mine]) and np.array_equal(p_tgt - lo, adj[mine, 0])
            assert np.array_equal(adj[mine, 3].view(np.float32), layer['beta'][e])
        seen += M
    assert seen == 2 * sum(sizes)                                              # each directed edge on exactly one rank


@pytest.mark.parametrize('world,k', [(1, 5), (3, 0), (3, 7), (3, 40), (4, 1000)])
def test_k_smallest_over_ranks(world, k):
    '''
    Sim._k_smallest_mask / _pick_positions / _global_counts: the population-wide selections of partitioned runs (test_num's n smallest keys,
    vaccinate_num's first k of a sequence, rescaling's positions in the global list of non-naive agents) from every rank's own offers.
    Ranks of different sizes, some with no candidate at all; float keys and int64 positions; k above the number of candidates.
    '''
    import threading
    import types
    from covasim_b200 import partition as cvpart
    from covasim_b200.sim import Sim
    rng = np.random.default_rng(world * 100 + k)
    sizes = [int(x) for x in rng.integers(0, 60, world)]
    sizes[0] = 50
    vals = [rng.permutation(10_000)[:n].astype(np.int64) for n in sizes]            # distinct within a rank ...
    for r in range(world):
        vals[r] = vals[r] * world + r                                                   # ... and across ranks
    fvals = [v.astype(np.float64) / 7.0 for v in vals]
    masks = [rng.random(n) < 0.7 for n in sizes]
    if world > 1:
        masks[1][:] = False                                                             # a rank without candidates
    comms = cvpart.LocalComm.make(world) if world > 1 else [None]
    got_i, got_f, picked = [None] * world, [None] * world, [None] * world
    n_flagged = [int(m.sum()) for m in masks]
    pos = rng.permutation(sum(n_flagged))[:min(9, sum(n_flagged))]

    def work(r):
        stub = types.SimpleNamespace(_comm=comms[r], device=torch.device('cpu'))
        stub._global_counts = lambda n: Sim._global_counts(stub, n)
        m = torch.as_tensor(masks[r])
        got_i[r] = Sim._k_smallest_mask(stub, torch.as_tensor(vals[r]), m, k).numpy()
        got_f[r] = Sim._k_smallest_mask(stub, torch.as_tensor(fvals[r]), m, k).numpy()
        counts = Sim._global_counts(stub, int(m.sum()))
        assert counts == n_flagged
        picked[r] = Sim._pick_positions(stub, torch.nonzero(m).flatten(), counts, pos).numpy()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    allv = np.concatenate([v[m] for v, m in zip(vals, masks)])
    want = set(np.sort(allv)[:k].tolist())
    for got in (got_i, got_f):
        chosen = set(np.concatenate([v[g] for v, g in zip(vals, got)]).tolist())
        assert chosen == want
        assert all(not (g & ~m).any() for g, m in zip(got, masks))
    # positions in the global ascending list of flagged agents -> (rank, local index)
    flat = [(r, i) for r in range(world) for i in np.nonzero(masks[r])[0]]
    assert sorted((r, int(i)) for r in range(world) for i in picked[r]) == sorted(flat[p] for p in pos)


def test_peer_exchange_layout_and_double_buffering(monkeypatch):
    '''
    partition.PeerExchange on the CPU: symmetric memory is replaced by plain host tensors whose addresses every "rank" (thread) knows, and
    cvb_peer_push by a memmove into each of them.  Checks what the GPU runs rely on: chunk r of an exchange lands at rank r's slot of the
    SAME half on every rank, the two halves alternate per exchange name, names do not overlap, the returned address is the half just
    filled, and the buffer filled two exchanges ago is the one overwritten (never the one the readers of the previous exchange use).
    '''
    import ctypes
    import sys
    import threading
    import types
    from covasim_b200 import partition as cvpart
    from covasim_b200 import _capi
    world, chunk = 3, 96
    sections = dict(codes=chunk, cases=chunk // 8)
    bufs, barrier = {}, threading.Barrier(world)

    class Handle:
        def __init__(self, rank):
            self.rank = rank

        @property
        def buffer_ptrs(self):
            return [bufs[r].data_ptr() for r in range(world)]

        def barrier(self, channel=0, timeout_ms=0):
            barrier.wait()

    fake = types.ModuleType('torch.distributed._symmetric_memory')
    tls = threading.local()

    def empty(n, dtype=None, device=None):
        t = torch.full((n,), 0xEE, dtype=torch.uint8)
        bufs[tls.rank] = t
        barrier.wait()                                       # every rank has allocated before anybody asks for the peers' addresses
        return t
    fake.empty = empty
    fake.rendezvous = lambda t, group: Handle(tls.rank)
    monkeypatch.setitem(sys.modules, 'torch.distributed._symmetric_memory', fake)
    import torch.distributed as tdist
    monkeypatch.setattr(tdist, '_symmetric_memory', fake, raising=False)

    def fake_call(name, *args):
        assert name == 'cvb_peer_push'
        src, nbytes, ptrs, w, off, _stream = args
        for p in range(w):
            ctypes.memmove(int(ptrs[p]) + int(off), int(src), int(nbytes))
    monkeypatch.setattr(_capi, 'call', fake_call)

    seen = [[] for _ in range(world)]

    def work(r):
        tls.rank = r
        comm = types.SimpleNamespace(world=world, rank=r, group=None, dist=types.SimpleNamespace(group=types.SimpleNamespace(WORLD='world')))
        ex = cvpart.PeerExchange(comm, torch.device('cpu'), sections)
        for day in range(5):
            for name, nbytes in sections.items():
                if name == 'cases' and day % 2:
                    continue                                 # tracing is not active every day: the two exchanges alternate independently
                inp = torch.full((nbytes,), 16 * day + r + (100 if name == 'cases' else 0), dtype=torch.uint8)
                ptr = ex.all_gather(name, inp, None)
                got = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(world * nbytes,)).copy()
                seen[r].append((name, day, ptr - bufs[r].data_ptr(), got))
                barrier.wait()                               # (stands for the kernels that read the buffer before the next exchange)
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert all(len(s) == len(seen[0]) > 0 for s in seen)
    offsets = {}
    for r in range(world):
        for name, day, off, got in seen[r]:
            nbytes = sections[name]
            want = np.repeat([16 * day + q + (100 if name == 'cases' else 0) for q in range(world)], nbytes).astype(np.uint8)
            assert np.array_equal(got, want), (r, name, day)
            offsets.setdefault(name, []).append(off) if r == 0 else None
            assert off % 256 == 0
    for name, offs in offsets.items():                      # two halves per name, strictly alternating
        assert len(set(offs)) == 2 and all(a != b for a, b in zip(offs, offs[1:])) and all(a == b for a, b in zip(offs, offs[2:]))
    spans = sorted((o, o + world * sections[n]) for n, offs in offsets.items() for o in set(offs))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))          # no two halves overlap


@pytest.mark.parametrize('seed', range(12))
def test_vaccinate_num_partitioned_selection_equals_single(seed):
    '''
    vaccinate_num: the partitioned form of the day's selection (population-wide "first k of the sequence" / "k smallest uniforms" through
    Sim._k_smallest_mask) picks exactly the people the single-process form picks and defers exactly the same second doses -- on random states:
    more second doses due than doses, nobody eligible, all candidates on one rank, doses for everybody, subtarget weights, boosters.
    CPU tensors; the keyed uniforms are replaced by one global random array that every rank slices.
    '''
    import threading
    import types
    from covasim_b200 import partition as cvpart
    from covasim_b200.interventions import vaccinate_num
    from covasim_b200.sim import Sim
    rng = np.random.default_rng(1000 + seed)
    world = int(rng.integers(2, 5))
    n = int(rng.integers(40, 400))
    bounds = np.sort(rng.choice(np.arange(1, n), world - 1, replace=False)) if world > 1 else np.zeros(0, dtype=int)
    starts = np.concatenate([[0], bounds]).astype(int)
    ends = np.concatenate([bounds, [n]]).astype(int)
    t = 7
    booster = bool(seed % 4 == 3)
    dead = rng.random(n) < 0.05
    vaccinated = rng.random(n) < (0.6 if booster else 0.3)
    doses = np.where(vaccinated & (not booster), rng.integers(1, 3, n), 0).astype(np.int32)
    due = np.where(vaccinated & (rng.random(n) < 0.5), t, -1).astype(np.int32)
    if seed % 3 == 1:
        due[:] = -1                                            # nobody scheduled
    seq_len = n if seed % 5 else max(3, n // 10)               # sometimes a short priority list (all candidates near the front of the id range)
    sequence = rng.permutation(n)[:seq_len] if seed % 5 else np.arange(seq_len)
    num_people = [0, 3, n // 4, 10 * n][seed % 4] if seed % 7 else 1
    subtarget = dict(inds=np.arange(0, n, 3), vals=rng.random(len(np.arange(0, n, 3)))) if seed % 2 else None
    u = {0: rng.random(n), 1: rng.random(n)}
    round_u = float(rng.random())

    def make_iv(lo, hi, comm):
        iv = vaccinate_num.__new__(vaccinate_num)
        iv.num_doses, iv.booster, iv.subtarget, iv.iindex = num_people, booster, subtarget, 0
        iv.p = dict(doses=2, interval=21)
        iv.doses = torch.as_tensor(doses[lo:hi].copy())
        iv.due_day = torch.as_tensor(due[lo:hi].copy())
        iv._prob = torch.empty(hi - lo, dtype=torch.float64)
        iv._uniforms = lambda sim, slot: torch.as_tensor(u[slot][lo:hi].copy())
        if comm is None:
            iv.sequence = torch.as_tensor(sequence.astype(np.int64))
        else:
            pos = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
            pos[sequence[::-1]] = np.arange(len(sequence) - 1, -1, -1)
            iv._pos = torch.as_tensor(pos[lo:hi])
        return iv

    def make_sim(lo, hi, comm):
        class StubSim(types.SimpleNamespace):
            def __getitem__(self, k):
                return {'pop_scale': 1.0, 'pop_size': n}[k]
        sim = StubSim(t=t, id0=lo, n_local=hi - lo, n=n, _comm=comm, device=torch.device('cpu'),
                      people=types.SimpleNamespace(dead=torch.as_tensor(dead[lo:hi].copy()), vaccinated=torch.as_tensor(vaccinated[lo:hi].copy()), device=torch.device('cpu')),
                      rng=types.SimpleNamespace(np_=types.SimpleNamespace(random_sample=lambda: round_u)))
        for name in ('_global_counts', '_k_smallest_mask', '_pick_positions'):
            setattr(sim, name, (lambda f: (lambda *a: f(sim, *a)))(getattr(Sim, name)))
        return sim

    iv0 = make_iv(0, n, None)
    sched0, first0 = vaccinate_num.select_people(iv0, make_sim(0, n, None))
    comms = cvpart.LocalComm.make(world)
    got = [None] * world
    ivs = [make_iv(int(starts[r]), int(ends[r]), comms[r]) for r in range(world)]

    def work(r):
        s, f = vaccinate_num.select_people(ivs[r], make_sim(int(starts[r]), int(ends[r]), comms[r]))
        got[r] = (s.numpy() + int(starts[r]), f.numpy() + int(starts[r]))
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert all(g is not None for g in got)
    assert sorted(np.concatenate([g[0] for g in got]).tolist()) == sorted(sched0.numpy().tolist())
    assert sorted(np.concatenate([g[1] for g in got]).tolist()) == sorted(first0.numpy().tolist())
    assert np.array_equal(np.concatenate([iv.due_day.numpy() for iv in ivs]), iv0.due_day.numpy())          # the same second doses were deferred


@pytest.mark.parametrize('seed', range(8))
def test_test_num_partitioned_selection_equals_single(seed, monkeypatch):
    '''
    test_num: the people tested by the ranks of a partitioned run (everybody at or below the n-th smallest key of the whole population, found
    from every rank's own n smallest) are the n people a single process picks with topk -- including ranks without any candidate, more tests
    than candidates, and weights of zero (keys +inf).  CPU tensors, kernels stubbed out.
    '''
    import threading
    import types
    from covasim_b200 import partition as cvpart
    from covasim_b200.interventions import test_num
    import ctypes
    monkeypatch.setattr(ctypes, 'byref', lambda x: x)        # (the stub parameter block is not a ctypes structure)
    rng = np.random.default_rng(2000 + seed)
    world = int(rng.integers(2, 5))
    n = int(rng.integers(30, 300))
    bounds = np.sort(rng.choice(np.arange(1, n), world - 1, replace=False))
    starts = np.concatenate([[0], bounds]).astype(int)
    ends = np.concatenate([bounds, [n]]).astype(int)
    weight = np.where(rng.random(n) < 0.2, 0.0, rng.choice([1.0, 3.0, 50.0], n))
    if seed % 3 == 0:
        weight[starts[1]:ends[1]] = 0.0                       # a rank whose agents cannot be tested
    key = np.where(weight > 0, -np.log(1 - rng.random(n)) / np.where(weight > 0, weight, 1.0), np.inf)
    n_tests = [1, 5, n // 3, 10 * n][seed % 4]

    def run(lo, hi, comm):
        tested = []
        iv = test_num.__new__(test_num)
        iv.subtarget = iv.ili_prev = iv.pdf = None
        iv._quar_code, iv.sensitivity, iv.loss_prob, iv.test_delay, iv.index = 0, 1.0, 0.0, 0, 0
        iv._c = types.SimpleNamespace()
        iv._weight, iv._key = torch.as_tensor(weight[lo:hi].copy()), torch.as_tensor(key[lo:hi].copy())
        iv.n_tests_today = lambda sim: n_tests

        class StubSim(types.SimpleNamespace):
            def __getitem__(self, k):
                return {'pop_scale': 1.0}[k]
        sim = StubSim(t=4, _comm=comm, _handle=None, _stream_ptr=None, rescale_vec=np.ones(10))

        def call(name, *args):
            if name == 'cvb_test_list':
                tested.append(int(args[3]))
        sim._call = call
        inds = test_num.apply(iv, sim)
        inds = np.zeros(0, dtype=np.int64) if inds is None else inds.numpy().astype(np.int64)
        assert (not tested and len(inds) == 0) or tested == [len(inds)]
        return inds + lo
    want = run(0, n, None)
    assert len(want) == min(n_tests, int((weight > 0).sum()))
    comms = cvpart.LocalComm.make(world)
    got = [None] * world

    def work(r):
        got[r] = run(int(starts[r]), int(ends[r]), comms[r])
    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert all(g is not None for g in got)
    assert sorted(np.concatenate(got).tolist()) == sorted(want.tolist())


@pytest.mark.parametrize('seed', range(12))
def test_vaccinate_num_selection_product_equals_oracle(seed):
    '''
    The day's selection of vaccinate_num in the product (device-array set algebra, here on CPU tensors) against the oracle's native-RNG
    restatement (oracle/cvoracle.py) on random states, with the same uniforms: same recipients, same deferred second doses.
    '''
    import types
    from covasim_b200.interventions import vaccinate_num
    from covasim_b200.sim import Sim
    from oracle import cvoracle as cvo
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(40, 400))
    t = 9
    booster = bool(seed % 4 == 3)
    dead = rng.random(n) < 0.05
    vaccinated = rng.random(n) < (0.6 if booster else 0.3)
    doses = np.where(vaccinated & (not booster), rng.integers(1, 3, n), 0).astype(np.int32)
    due = np.where(vaccinated & (rng.random(n) < 0.5), t, -1).astype(np.int64)
    if seed % 3 == 1:
        due[:] = -1
    sequence = rng.permutation(n)[:n if seed % 5 else max(3, n // 10)]
    num_people = [0, 3, n // 4, 10 * n][seed % 4] if seed % 7 else 1
    subtarget = dict(inds=np.arange(0, n, 3), vals=rng.random(len(np.arange(0, n, 3)))) if seed % 2 else None
    u = {0: rng.random(n), 1: rng.random(n)}
    round_u = float(rng.random())
    # product
    iv = vaccinate_num.__new__(vaccinate_num)
    iv.num_doses, iv.booster, iv.subtarget, iv.iindex = num_people, booster, subtarget, 0
    iv.p = dict(doses=2, interval=21)
    iv.doses, iv.due_day = torch.as_tensor(doses.copy()), torch.as_tensor(due.astype(np.int32))
    iv._prob = torch.empty(n, dtype=torch.float64)
    iv._uniforms = lambda sim, slot: torch.as_tensor(u[slot].copy())
    iv.sequence = torch.as_tensor(sequence.astype(np.int64))

    class StubSim(types.SimpleNamespace):
        def __getitem__(self, k):
            return {'pop_scale': 1.0, 'pop_size': n}[k]
    sim = StubSim(t=t, id0=0, n_local=n, n=n, _comm=None, device=torch.device('cpu'),
                  people=types.SimpleNamespace(dead=torch.as_tensor(dead.copy()), vaccinated=torch.as_tensor(vaccinated.copy()), device=torch.device('cpu')),
                  rng=types.SimpleNamespace(np_=types.SimpleNamespace(random_sample=lambda: round_u)))
    sched, first = vaccinate_num.select_people(iv, sim)
    # oracle (native-RNG mode)
    ov = cvo.vaccinate_num.__new__(cvo.vaccinate_num)
    ov.num_doses, ov.booster, ov.subtarget, ov.iindex = num_people, booster, subtarget, 0
    ov.p = dict(doses=2, interval=21)
    ov.doses, ov.due_day, ov.sequence = doses.copy(), due.copy(), sequence.copy()
    ov._scheduled = {}
    osim = types.SimpleNamespace(t=t, P=dict(uid=np.arange(n), dead=dead.copy(), vaccinated=vaccinated.copy()), pars={'pop_scale': 1.0},
                                 rng=types.SimpleNamespace(kind='philox', np_=types.SimpleNamespace(random_sample=lambda: round_u),
                                                           agent_uniforms=lambda t_, purpose, sub, inds, slot=0: u[slot][np.asarray(inds)]))
    want = cvo.vaccinate_num.select_people(ov, osim)
    got = np.concatenate([sched.numpy(), first.numpy()])
    assert sorted(got.tolist()) == sorted(np.asarray(want, dtype=np.int64).tolist())
    others = np.ones(n, dtype=bool)                             # (the oracle books today's first doses for their second dose here, the product in the dose kernel)
    others[first.numpy()] = False
    assert np.array_equal(iv.due_day.numpy().astype(np.int64)[others], ov.due_day[others])
    assert np.all(ov.due_day[first.numpy()] == t + 21)
