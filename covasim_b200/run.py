'''
Ensembles: MultiSim / multi_run (reference covasim/run.py:36-374, 1326-1519).

The reference runs ensemble members in a process pool and pickles whole sims back (run.py:1477).  Members are
independent (seeds differ by ``rand_seed + i``, run.py:1363-1365), so here they are SHARDED:

* one process, several GPUs: member i lives on ``cuda:(i % n_gpus)``; the members are stepped round-robin, and
  because Sim.step() never synchronises with the device, all GPUs work concurrently;
* one process per GPU (torchrun): rank r runs members r, r + world, ...; there is no data-path collective, only one
  gather of the finished result tables (``gather_results``; NCCL when the process group is NCCL, gloo in the CPU
  tests).

``reduce / mean / median / combine`` keep the reference's semantics (run.py:220-374) and operate on host arrays.
'''
import copy

import numpy as np

from . import defaults as cvd
from .base import Result

__all__ = ['MultiSim', 'multi_run', 'shard_indices', 'gather_results', 'pack_results', 'unpack_results']


# ---------------------------------------------------------------------------------------------------
# distributed plumbing (device independent; tested with gloo on CPU)
# ---------------------------------------------------------------------------------------------------
def shard_indices(n_items, rank, world):
    ''' Members owned by ``rank``: rank, rank + world, ... (round-robin keeps the shards balanced for any n_items) '''
    return list(range(rank, n_items, world))


def pack_results(results):
    ''' Flatten one member's results into a float64 vector: every main series, then every by-variant series (row-major) '''
    keys = [k for k, v in results.items() if isinstance(v, Result)]
    vkeys = list(results['variant'].keys())
    parts = [np.asarray(results[k].values, dtype=np.float64).ravel() for k in keys]
    parts += [np.asarray(results['variant'][k].values, dtype=np.float64).ravel() for k in vkeys]
    return np.concatenate(parts), keys, vkeys


def unpack_results(vec, keys, vkeys, npts, nv):
    ''' Inverse of pack_results -> dict of arrays '''
    out, off = {}, 0
    for k in keys:
        out[k] = vec[off:off + npts].copy()
        off += npts
    out['variant'] = {}
    for k in vkeys:
        out['variant'][k] = vec[off:off + nv * npts].reshape(nv, npts).copy()
        off += nv * npts
    return out


def gather_results(local, n_items, device=None):
    '''
    All-gather per-member float64 vectors across the process group.  ``local`` maps member index -> 1-D array (all
    the same length); returns the full list in member order on every rank.  One collective for the whole ensemble.
    '''
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local[i] for i in range(n_items)]
    world, rank = dist.get_world_size(), dist.get_rank()
    length = len(next(iter(local.values()))) if local else 0
    lt = torch.tensor([length], dtype=torch.int64)
    backend = dist.get_backend()
    dev = torch.device('cpu') if backend == 'gloo' else (device or torch.device('cuda', torch.cuda.current_device()))
    lt = lt.to(dev)
    dist.all_reduce(lt, op=dist.ReduceOp.MAX)
    length = int(lt.item())
    per_rank = (n_items + world - 1) // world
    buf = torch.zeros((per_rank, length), dtype=torch.float64, device=dev)
    for slot, i in enumerate(shard_indices(n_items, rank, world)):
        buf[slot] = torch.as_tensor(local[i], dtype=torch.float64)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = [None] * n_items
    for r in range(world):
        host = out[r].cpu().numpy()
        for slot, i in enumerate(shard_indices(n_items, r, world)):
            full[i] = host[slot]
    return full


# ---------------------------------------------------------------------------------------------------
# MultiSim
# ---------------------------------------------------------------------------------------------------
class MultiSim:
    ''' An ensemble of sims (reference run.py:36-374) '''

    def __init__(self, sims=None, base_sim=None, label=None, n_runs=4, **kwargs):
        if sims is None and base_sim is None:
            raise ValueError('Must supply at least one sim or a base sim')
        if isinstance(sims, list):
            self.sims = sims
            self.base_sim = base_sim if base_sim is not None else sims[0]
        else:
            self.base_sim = sims if sims is not None else base_sim
            self.sims = None
        self.label = label if label is not None else getattr(self.base_sim, 'label', None)
        self.n_runs = n_runs
        self.run_args = dict(kwargs)
        self.results = None
        self.which = None
        self.member_results = None

    def __len__(self):
        return len(self.sims) if self.sims is not None else self.n_runs

    def init_sims(self, noise=0.0, noisepar='beta'):
        ''' Create the members from the base sim: same parameters, seed + i (reference run.py:1340-1379) '''
        if self.sims is not None:
            return
        self.sims = []
        base_seed = self.base_sim['rand_seed']
        for i in range(self.n_runs):
            sim = self.base_sim.copy()
            sim['rand_seed'] = base_seed + i
            sim.label = f'Sim {i}'
            if noise:
                scale = 1 + noise * np.random.RandomState(base_seed + i).normal()
                sim[noisepar] = sim[noisepar] * scale
            self.sims.append(sim)

    def run(self, n_gpus=None, keep_people=False, **kwargs):
        '''
        Run every member (reference run.py:142-180, multi_run 1406-1519).  Members are placed on the visible GPUs
        round-robin and stepped in an interleaved loop; under torch.distributed each rank runs its own shard and the
        finished result tables are all-gathered once.
        '''
        import torch
        import torch.distributed as dist
        self.init_sims(**{k: kwargs[k] for k in ('noise', 'noisepar') if k in kwargs})
        n = len(self.sims)
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank, world = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
        mine = shard_indices(n, rank, world)
        if distributed:
            devices = [torch.device('cuda', torch.cuda.current_device())]
        else:
            n_dev = torch.cuda.device_count() if n_gpus is None else min(n_gpus, torch.cuda.device_count())
            devices = [torch.device('cuda', d) for d in range(max(n_dev, 1))]
        # initialise (population build + upload), then advance all local members in lockstep
        active = []
        for slot, i in enumerate(mine):
            sim = self.sims[i]
            if not sim.initialized:                        # (an initialised member keeps the device its arrays live on)
                sim.device = devices[slot % len(devices)]
            with torch.cuda.device(sim.device):
                if not sim.initialized:
                    sim.initialize()
                sim.set_seed()
            active.append(sim)
        self._advance_members(active)
        for sim in active:
            with torch.cuda.device(sim.device):
                sim.finalize()
        # gather the result tables (the only communication of the whole ensemble)
        local = {}
        keys = vkeys = None
        for i in mine:
            vec, keys, vkeys = pack_results(self.sims[i].results)
            local[i] = vec
        if keys is None:                                   # a rank without members still takes part in the collective
            _, keys, vkeys = pack_results(_empty_results(self.sims[0]))
        vectors = gather_results(local, n)
        # sizes from a member this rank has initialised (n_variants / n_days are only final after initialize()); ranks without
        # members derive them from the base parameters the same way initialize() does
        ref = active[0] if active else self.sims[0]
        if not ref.initialized:
            ref._validate_pars()
        npts = ref.npts
        nv = ref['n_variants'] if ref.initialized else 1 + len(ref['variants'])
        self.member_results = [unpack_results(v, keys, vkeys, npts, nv) for v in vectors]
        if not keep_people:
            for sim in active:                             # reference sim.shrink(): drop the big arrays (run.py:1400-1401)
                sim.people = None
                sim._destroy()
        return self

    def _advance_members(self, sims):
        '''
        Run the members to their end in lockstep.  Stretches of days that need no host decision in ANY member go through ONE
        cvb_run_days_multi call: a single host thread issues every member's five launches per day to the member's own stream, so
        small members (whose kernels are far too short to fill a B200) overlap on the GPU instead of queueing behind each other.
        Days on which some member needs its host (an intervention that acts through Python, a variant import, ...) are stepped
        member by member.
        '''
        import ctypes as C
        import torch
        from . import _capi
        if not sims:
            return
        lockstep = len({(s.npts, s.t) for s in sims}) == 1 and all(s.rng_mode == 'philox' for s in sims)
        if not lockstep:
            remaining = list(sims)
            while remaining:
                for sim in list(remaining):
                    with torch.cuda.device(sim.device):
                        sim.step()
                    if sim.complete:
                        remaining.remove(sim)
            return
        streams = []
        for sim in sims:
            with torch.cuda.device(sim.device):
                streams.append(torch.cuda.Stream(device=sim.device))
        n = len(sims)
        handles = (C.c_void_p * n)(*[s._handle for s in sims])
        stream_ptrs = (C.c_void_p * n)(*[st.cuda_stream for st in streams])
        devices = sorted({s.device for s in sims}, key=str)
        npts = sims[0].npts
        plain = all(not s.pars['analyzers'] and not s.pars['stopping_func'] for s in sims)
        while not sims[0].complete:
            t = sims[0].t
            if plain and all(s._fusable_day(t) for s in sims):
                t1 = t + 1
                while t1 < npts and all(s._fusable_day(t1) for s in sims):
                    t1 += 1
                for sim in sims:                           # parameters / adjacency up to date, pending edge copies done
                    with torch.cuda.device(sim.device):
                        sim._push_pars()
                        if sim._adj_dirty:
                            sim._build_adjacency()
                            sim._build_plan_keep_days()
                        if sim._plan['needs_edges']:
                            sim._sync_edges()
                for d in devices:
                    torch.cuda.synchronize(d)              # the members' streams do not synchronise with the default stream
                _capi.call('cvb_run_days_multi', handles, n, int(t), int(t1), stream_ptrs)
                for d in devices:
                    torch.cuda.synchronize(d)
                for sim in sims:
                    sim.fused_days += t1 - t
                    sim.people.t = t1 - 1
                    sim.t = t1
                    sim.complete = t1 == npts
            else:
                for sim in sims:
                    with torch.cuda.device(sim.device):
                        sim.step()

    # ---- reductions (reference run.py:220-374) ----------------------------------------------------
    def _check(self):
        if self.member_results is None:
            raise RuntimeError('MultiSim has not been run')

    def reduce(self, quantiles=None, use_mean=False, bounds=None):
        ''' Median (or mean) and low/high bands over the members for every result '''
        self._check()
        if use_mean:
            bounds = 2 if bounds is None else bounds
        else:
            quantiles = dict(low=0.1, high=0.9) if quantiles is None else quantiles
        keys = [k for k in self.member_results[0].keys() if k != 'variant']
        vkeys = list(self.member_results[0]['variant'].keys())
        out = dict(variant={})
        for group, names, dest in ((None, keys, out), ('variant', vkeys, out['variant'])):
            for k in names:
                stack = np.stack([(m[k] if group is None else m['variant'][k]) for m in self.member_results], axis=-1)
                r = Result(k, npts=stack.shape[-2], n_variants=stack.shape[0] if stack.ndim == 3 else 0)
                if use_mean:
                    mean, std = stack.mean(axis=-1), stack.std(axis=-1)
                    r.values[:], r.low, r.high = mean, mean - bounds * std, mean + bounds * std
                else:
                    r.values[:] = np.quantile(stack, q=0.5, axis=-1)
                    r.low = np.quantile(stack, q=quantiles['low'], axis=-1)
                    r.high = np.quantile(stack, q=quantiles['high'], axis=-1)
                dest[k] = r
        self.results = out
        self.which = 'reduced'
        return self

    def mean(self, bounds=None):
        return self.reduce(use_mean=True, bounds=bounds)

    def median(self, quantiles=None):
        return self.reduce(use_mean=False, quantiles=quantiles)

    def combine(self):
        ''' Sum the members as if they were one large population (reference run.py:300-374): counts add, fractions average '''
        self._check()
        n = len(self.member_results)
        keys = [k for k in self.member_results[0].keys() if k != 'variant']
        out = dict(variant={})
        for k in keys:
            stack = np.stack([m[k] for m in self.member_results], axis=-1)
            r = Result(k, npts=stack.shape[0])
            r.values[:] = stack.mean(axis=-1) if k in cvd.unscaled_results else stack.sum(axis=-1)
            out[k] = r
        # the reference combines the main result series only (run.py:353-358 loops over result_keys()): the by-variant series of
        # the combined sim stay those of the first member
        for k, vals in self.member_results[0]['variant'].items():
            r = Result(k, npts=vals.shape[1], n_variants=vals.shape[0])
            r.values[:] = vals
            out['variant'][k] = r
        self.results = out
        self.which = 'combined'
        return self

    def summarize(self, key='cum_infections'):
        self._check()
        return np.array([m[key][-1] for m in self.member_results])


def _empty_results(sim):
    s = copy.copy(sim)
    s._init_results()
    return s.results


def multi_run(sim, n_runs=4, **kwargs):
    ''' Convenience wrapper (reference run.py:1406-1519): run n_runs copies of ``sim`` with consecutive seeds; returns the sims '''
    msim = MultiSim(sim, n_runs=n_runs)
    msim.run(keep_people=kwargs.pop('keep_people', False), **kwargs)
    return msim.sims
