'''
Sim: host-side orchestration of one simulation (reference covasim/sim.py).

The host stays Python and keeps the reference's public contract -- ``Sim(pars, **kwargs)``,
``initialize() / step() / run(until) / finalize()``, ``sim[par]``, ``sim.results[key].values``,
``sim.summary``, ``sim.people``, interventions and analyzers called once per day from inside ``step``
(reference sim.py:558-685) -- while every per-agent and per-edge computation of the day runs as CUDA
kernels on device-resident People / Layer arrays through the C ABI (include/covasim_b200.h).

Per-day results are accumulated on the device (one row of int64 counters per day) and copied to the host
once, in ``finalize`` (or on demand through ``sync_results``); ``step`` never synchronises with the GPU.
'''
import copy
import ctypes as C
import datetime as dt

import numpy as np
import torch

from . import defaults as cvd
from . import parameters as cvpar
from . import population as cvpop
from . import immunity as cvimm
from . import utils as cvu
from . import _capi
from .base import Result, AlreadyRunError
from .people import People

__all__ = ['Sim', 'r_eff_windowed', 'gen_time']

f32 = np.float32


def r_eff_windowed(method, date_infectious, date_recovered, date_dead, log_source, npts, window=7):
    '''
    The 'infectious' / 'outcome' r_eff of the reference (sim.py:949-981) as a function of plain arrays: every source is dated
    by the day it became infectious (or recovered / died); r_eff[t] = infections caused by the sources of the last ``window``
    days / number of those sources.
    '''
    window = int(window)
    n = len(date_infectious)
    source_date = np.full(n, -1, dtype=np.int64)
    dates = [date_infectious] if method == 'infectious' else [date_recovered, date_dead]
    for d in dates:                                      # t == date for an integer day t inside the run
        d = np.asarray(d, dtype=np.float64)
        ok = np.isfinite(d) & (d >= 0) & (d < npts) & (d == np.floor(d))
        source_date[ok] = d[ok].astype(np.int64)
    sources = np.bincount(source_date[source_date >= 0], minlength=npts).astype(np.float64)
    src = np.asarray(log_source)
    src = src[src >= 0]                                  # seed infections and importations have no source
    sd = source_date[src]
    targets = np.bincount(sd[sd >= 0], minlength=npts).astype(np.float64)
    r_eff = np.divide(targets, sources, out=np.full(npts, np.nan), where=sources > 0)
    num = np.nancumsum(r_eff * sources)
    num[window:] = num[window:] - num[:-window]
    den = np.cumsum(sources)
    den[window:] = den[window:] - den[:-window]
    return np.divide(num, den, out=np.full(npts, np.nan), where=den > 0)


def gen_time(date_exposed, date_symptomatic, log_source, log_target):
    ''' Generation-time statistics from the infection log (reference sim.py:990-1025) as a function of plain arrays '''
    src, tgt = np.asarray(log_source), np.asarray(log_target)
    has = src >= 0
    src, tgt = src[has], tgt[has]
    de, ds = np.asarray(date_exposed, dtype=np.float64), np.asarray(date_symptomatic, dtype=np.float64)
    true = de[tgt] - de[src]
    both = np.isfinite(ds[src]) & np.isfinite(ds[tgt])
    clin = ds[tgt][both] - ds[src][both]
    with np.errstate(all='ignore'):
        return {'true': np.mean(true), 'true_std': np.std(true), 'clinical': np.mean(clin), 'clinical_std': np.std(clin)}


class Sim:

    def __init__(self, pars=None, popdict=None, label=None, device=None, rng='philox', pop_exact=None, use_adjacency=True, partition=None,
                 hit_capacity=0, pop_gen='host', fused=True, log_capacity=None, **kwargs):
        kw = dict(pars or {})
        kw.update(kwargs)
        for alias, key in (('n_agents', 'pop_size'), ('init_infected', 'pop_infected')):       # reference base.py:266-273
            if alias in kw:
                kw[key] = kw.pop(alias)
        self.pars = cvpar.make_pars(**kw)
        self.pars['pop_size'] = int(self.pars['pop_size'])
        for key in ('interventions', 'analyzers', 'variants'):
            if not isinstance(self.pars[key], list):
                self.pars[key] = [self.pars[key]]
        if rng not in ('philox', 'mt'):
            raise NotImplementedError(f'rng mode "{rng}" is not available (choices: philox (native), mt (replay of the reference streams))')
        self.rng_mode = rng
        self.label = label
        self.popdict = popdict
        self.pop_exact = pop_exact
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.rng = cvu.HostStreams()
        self.people = None
        self.results = {}
        self.summary = None
        self.t = None
        self.initialized = False
        self.complete = False
        self.results_ready = False
        self._handle = None
        self._stream_ptr = None
        self._orig_pars = None
        self._pars_key = None
        self._pars_dirty = True
        self._quar_horizon = 1
        self._host_adds = {}
        self.kernel_timers = None          # set to {} to time every C-ABI call of step() with CUDA events
        # fused=True: days on which no host decision is needed run through cvb_run_days (five launches per day, whole blocks of
        # days per call from run()); False: every day through the per-step entry points.  Same results either way.
        self.fused = bool(fused)
        self.log_capacity = log_capacity   # entries of the device infection log (default 4 per agent); overflow is reported, not silent
        self._plan = None
        self.fused_days = 0                # days that went through cvb_run_days (diagnostics / tests)
        self.use_adjacency = use_adjacency   # False: stream every layer densely each day (the reference's access pattern)
        self._adj = None
        self._adj_mask = 0
        self._edges_event = None
        self._copy_stream = None
        self._adj_dirty = False
        # agent partition over several GPUs (partition.py): None = the whole population on this GPU; True = one rank of
        # the torch.distributed world; or a communicator object (partition.DistComm / partition.LocalComm)
        # 'host': the reference's generators on the host (population.py; bit-exact with the reference when pop_exact);
        # 'device': the same distributions from keyed Philox draws, generated on the GPU (population.KeyedPop)
        if pop_gen not in ('host', 'device'):
            raise ValueError(f'pop_gen must be "host" or "device", not "{pop_gen}"')
        self.pop_gen = pop_gen
        self._keyed_pop = None
        self._partition = partition
        self._comm = None
        self._peer = None
        self._hit_capacity = int(hit_capacity)
        self._part_bufs = None
        if partition is not None and partition is not False and rng != 'philox':
            raise NotImplementedError('agent partitioning needs the native RNG (rng="philox"): replay mode follows the reference\'s sequential streams')

    # ---- dict-like parameter access (reference base.py:63-114) -----------------------------------
    def __getitem__(self, key):
        try:
            return self.pars[key]
        except KeyError:
            raise KeyError(f'Key "{key}" not found; available keys: {", ".join(self.pars.keys())}')

    def __setitem__(self, key, value):
        if key not in self.pars:
            raise KeyError(f'Key "{key}" not found; available keys: {", ".join(self.pars.keys())}')
        self.pars[key] = value

    def __del__(self):
        self._destroy()

    def _destroy(self):
        h, self._handle = getattr(self, '_handle', None), None
        if h is not None:
            try:
                _capi.lib.cvb_destroy(h)
            except Exception:
                pass

    # ---- time helpers (reference base.py:325-441) --------------------------------------------------
    @property
    def n(self):
        return int(self.pars['pop_size'])

    @property
    def n_local(self):
        ''' Agents held on this GPU: all of them unless the simulation is agent-partitioned '''
        return self.people.n if self.people is not None else self.n

    @property
    def id0(self):
        return self.people.id0 if self.people is not None else 0

    @property
    def npts(self):
        return int(self.pars['n_days']) + 1

    @property
    def tvec(self):
        return np.arange(self.npts)

    @property
    def scaled_pop_size(self):
        return self.pars['pop_size'] * self.pars['pop_scale']

    def _start_date(self):
        sd = self.pars['start_day']
        if isinstance(sd, str):
            return dt.datetime.strptime(sd, '%Y-%m-%d').date()
        if isinstance(sd, dt.datetime):
            return sd.date()
        if isinstance(sd, dt.date):
            return sd
        return dt.date(2020, 3, 1) + dt.timedelta(days=int(sd))

    def day(self, day, *args):
        ''' Convert a date / string / int to a day index (reference base.py:346-389) '''
        if day is None:
            return None
        if isinstance(day, (int, np.integer, float)):
            return int(day)
        if isinstance(day, str):
            day = dt.datetime.strptime(day, '%Y-%m-%d').date()
        if isinstance(day, dt.datetime):
            day = day.date()
        return (day - self._start_date()).days

    def date(self, ind, as_date=False):
        ''' Convert a day index to a date string (reference base.py:392-430) '''
        d = self._start_date() + dt.timedelta(days=int(ind))
        return d if as_date else d.strftime('%Y-%m-%d')

    @property
    def datevec(self):
        return [self.date(i) for i in range(self.npts)]

    def result_keys(self, which='main'):
        if which == 'variant':
            return list(self.results['variant'].keys())
        return [k for k in self.results.keys() if isinstance(self.results[k], Result)]

    def intervention_index(self, obj):
        ''' Position in the intervention list; interventions nested in others (cv.sequence) are numbered after the top-level ones '''
        flat = []

        def walk(ivs):
            flat.extend(ivs)
            for iv in ivs:
                walk(list(getattr(iv, 'interventions', None) or []))
        walk(list(self.pars['interventions']))
        return [id(i) for i in flat].index(id(obj))

    def get_interventions(self, which=None):
        ivs = self.pars['interventions']
        if which is None:
            return list(ivs)
        if isinstance(which, int):
            return ivs[which]
        return [i for i in ivs if (isinstance(which, type) and isinstance(i, which)) or getattr(i, 'label', None) == which]

    def get_analyzers(self, which=None):
        ''' The analyzers, or those of a given class / label / position (reference base.py:783-880 get_analyzers) '''
        ans = self.pars['analyzers']
        if which is None:
            return list(ans)
        if isinstance(which, int):
            return ans[which]
        return [a for a in ans if (isinstance(which, type) and isinstance(a, which)) or getattr(a, 'label', None) == which]

    def get_analyzer(self, which=None):
        found = self.get_analyzers(which)
        return found if not isinstance(found, list) else (found[-1] if found else None)

    def copy(self):
        ''' A deep copy (reference base.py:444-446): of a running simulation too -- the copy gets its own device arrays and handle '''
        return copy.deepcopy(self)

    _NO_COPY = ('_handle', '_adj', '_beds', '_part_bufs', '_copy_stream', '_edges_event', '_plan', '_cpars', '_counters', '_vcounters', '_sums', '_log',
                '_stream_ptr', '_comm', '_keyed_pop', '_peer', '_compact_table')

    def __deepcopy__(self, memo):
        if self._comm is not None:
            raise NotImplementedError('an agent-partitioned simulation cannot be copied (its arrays are spread over several processes)')
        if self._handle is not None:
            self._sync_edges()
            torch.cuda.synchronize(self.device)
        new = type(self).__new__(type(self))
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, None if k in self._NO_COPY else copy.deepcopy(v, memo))
        new._adj_mask = 0
        if self._handle is not None:
            new._clone_device_state(self)
        return new

    def _clone_device_state(self, src):
        ''' Second half of a deep copy: a handle of its own, bound to the copied arrays, with the source's run state '''
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        _capi.call('cvb_create', C.byref(h), self.n_local, self.pars['n_variants'], self.npts, int(self.pars['rand_seed']))
        self._handle = h
        self._stream_ptr = None
        self._counters, self._vcounters, self._sums = src._counters.clone(), src._vcounters.clone(), src._sums.clone()
        _capi.call('cvb_bind_results', h, self._counters.data_ptr(), self._vcounters.data_ptr(), self._sums.data_ptr())
        self._log = {k: v.clone() for k, v in src._log.items()}
        L = self._log
        _capi.call('cvb_bind_log', h, L['source'].data_ptr(), L['target'].data_ptr(), L['date'].data_ptr(), L['layer'].data_ptr(),
                   L['variant'].data_ptr(), len(L['source']), L['count'].data_ptr())
        _capi.call('cvb_set_seed', h, int(self.rng.seed))
        self.people._bind(self)
        if self.pars['use_waning']:
            kin = np.ascontiguousarray(self.pars['nab_kin'], dtype=np.float64)
            _capi.call('cvb_set_nab_kin', h, kin.ctypes.data, len(kin))
        self._pars_dirty = True
        self._pars_key = None
        self._push_pars()
        _capi.call('cvb_clone_scratch', h, src._handle, self._stream_ptr)
        self._build_adjacency()
        self._build_plan_keep_days()

    # ---- initialisation (reference sim.py:94-125) --------------------------------------------------
    def set_seed(self, seed=-1):
        if seed != -1:
            self.pars['rand_seed'] = seed
        self.rng.set_seed(self.pars['rand_seed'])
        if self._handle is not None:
            _capi.call('cvb_set_seed', self._handle, int(self.rng.seed))

    def initialize(self, reset=False, init_infections=True, **kwargs):
        if not torch.cuda.is_available():
            raise _capi.CvbError('covasim_b200 needs a CUDA device: there is no CPU fallback')
        pars = self.pars
        self.t = 0
        self._validate_pars()
        self.set_seed()
        for v in pars['variants']:
            if not isinstance(v, cvimm.variant):
                raise TypeError(f'Variant {v} is not a cv.variant object; please create using cv.variant()')
            if not v.initialized:
                v.initialize(self)
        pars['n_variants'] = len(pars['variant_pars'])
        if pars['n_variants'] > _capi.MAX_VARIANTS:
            raise ValueError(f'at most {_capi.MAX_VARIANTS} variants are supported')
        cvimm.init_immunity(self)
        self._init_results()
        self._init_people(**kwargs)
        self._create_handle()
        if init_infections:
            self.init_infections()
        for iv in pars['interventions']:
            if hasattr(iv, 'initialize'):
                iv.initialize(self)
        for an in pars['analyzers']:
            if hasattr(an, 'initialize'):
                an.initialize(self)
        self.set_seed()
        self.initialized = True
        self.complete = False
        self.results_ready = False
        self._orig_pars = None
        self._build_plan()
        return self

    # ---- the fused day pipeline (cvb_run_days) ---------------------------------------------------------------------------
    def _build_plan(self):
        '''
        Decide which days can run without the host (reference sim.py:558-685 needs the host only where Python decides something):
        register the built-in interventions that run inside the fused kernels with the handle and collect the days on which
        some intervention / variant needs its apply().  ``self._plan`` is None when no day can be fused.
        '''
        self._plan = None
        self.fused_days = 0
        self._make_plan()
        if self._comm is not None:                       # every rank of a partitioned run must take the same path on the same day
            if not all(self._comm.gather_objects(self._plan is not None)):
                self._plan = None

    def _make_plan(self):
        pars = self.pars
        if not self.fused or self.rng_mode != 'philox' or self.n_local % 4 != 0 or self._handle is None:
            return
        if self._comm is not None and (pars['n_beds_hosp'] is not None or pars['n_beds_icu'] is not None):
            return                                       # bed limits compare global counts: the per-step path sums them over the ranks
        host_days, test, trace, vaccs = set(), None, None, []
        for iv in pars['interventions']:
            plan = iv._device_plan(self) if hasattr(iv, '_device_plan') else None
            if plan is None:
                return                                   # a plug-in that must be called every day
            if plan[0] == 'host':
                host_days.update(int(d) for d in plan[1])
            elif plan[0] == 'test':
                if test is not None or trace is not None:
                    return                               # one test_prob, applied before tracing
                test = plan
            elif plan[0] == 'trace':
                if trace is not None:
                    return
                trace = plan
            elif plan[0] == 'vacc':
                if len(vaccs) == 4:
                    return
                vaccs.append(plan)
        for v in pars['variants']:
            host_days.update(int(d) for d in np.atleast_1d(v.days))
        lkeys = self.people.layer_keys()
        regen = 0
        from .base import Layer
        for i, lk in enumerate(lkeys):
            if pars['dynam_layer'].get(lk):
                if type(self.people.contacts[lk]).update is not Layer.update:
                    return                               # a user-defined Layer.update runs on the host
                regen |= 1 << i
        if trace is not None and self._comm is None:
            traced = [i for i, lk in enumerate(lkeys) if trace[1].trace_prob[i] > 0 and len(self.people.contacts[lk]) > 0]
            if any(pars['dynam_layer'].get(lkeys[i]) for i in traced) or (traced and not self.use_adjacency):
                return                                   # tracing over a streamed layer keeps the per-step path
        h = self._handle
        _capi.call('cvb_plan_clear', h)
        if test is not None:
            _capi.call('cvb_plan_test_prob', h, C.byref(test[1]), test[2], test[3])
        if trace is not None:
            _capi.call('cvb_plan_contact_tracing', h, C.byref(trace[1]), trace[2], trace[3])
        for plan in vaccs:
            _capi.call('cvb_plan_vaccinate', h, C.byref(plan[1]), plan[2].ctypes.data, plan[3].data_ptr(), plan[4].data_ptr())
        _capi.call('cvb_plan_dynamic_layers', h, regen)
        # the fused kernels of static layers read only the adjacency: the raw edge arrays matter when a layer is streamed densely
        dense = self._comm is None and (regen != 0 or self._adj is None or any(len(self.people.contacts[lk]) > 0 and not ((self._adj_mask >> i) & 1) for i, lk in enumerate(lkeys)))
        trace_days = None if trace is None else (trace[2], trace[3])
        self._plan = dict(host_days=host_days, needs_edges=bool(dense), trace_days=trace_days)

    def _build_plan_keep_days(self):
        ''' The adjacency changed (a layer was edited): recompute what depends on it '''
        days = self.fused_days
        self._build_plan()
        self.fused_days = days

    def _fusable_day(self, t):
        ''' True if day t needs no host decision: no intervention / variant acts through Python, no importations, no rescaling '''
        plan, pars = self._plan, self.pars
        if plan is None or t in plan['host_days'] or pars['n_imports'] or self.kernel_timers is not None:
            return False
        if pars['rescale'] and self.rescale_vec[t] < pars['pop_scale']:
            return False
        return True

    def _run_block(self, t0, t1):
        ''' Days [t0, t1) in one C-ABI call '''
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        self._push_pars()
        if self._adj_dirty:
            self._build_adjacency()
            self._build_plan_keep_days()
        if self._plan['needs_edges']:
            self._sync_edges()
        if self._comm is not None:
            self._run_block_partitioned(int(t0), int(t1))
        else:
            _capi.call('cvb_run_days', self._handle, int(t0), int(t1), self._stream_ptr)
        self.fused_days += t1 - t0
        self.people.t = t1 - 1
        self.t = t1
        if self.t == self.npts:
            self.complete = True

    def _run_block_partitioned(self, t0, t1):
        '''
        Days [t0, t1) of an agent-partitioned simulation through the fused kernels: the same five launches per day, with the two
        exchanges of the partition (partition.py) between them -- the case bitmap on tracing days, the 1-byte transmit codes every day.
        '''
        h, st = self._handle, self._stream_ptr
        td = self._plan['trace_days']
        for t in range(t0, t1):
            first = int(t == t0)
            _capi.call('cvb_fused_phase', h, t, 0, first, st)
            if td is not None and t >= td[0] and (td[1] < 0 or t <= td[1]):
                self._exchange_cases()
                _capi.call('cvb_fused_phase', h, t, 1, first, st)
            _capi.call('cvb_fused_phase', h, t, 2, first, st)
            self._exchange_codes()
            _capi.call('cvb_fused_phase', h, t, 3, first, st)
        _capi.call('cvb_fused_phase', h, t1, 4, 0, st)

    def check_packed_state(self):
        ''' Verification hook: the library's packed per-agent state word against the People arrays; returns (inexpressible, mismatches, examples) '''
        out = np.zeros(26, dtype=np.int64)
        _capi.call('cvb_state_check', self._handle, int(self.t) - 1, out.ctypes.data, self._stream_ptr)
        return int(out[0]), int(out[1]), out[2:].reshape(8, 3)

    def _validate_pars(self):
        pars = self.pars
        if pars['end_day'] is not None:
            pars['n_days'] = self.day(pars['end_day'])
        pars['n_days'] = int(pars['n_days'])
        pars['end_day'] = self.date(pars['n_days'])       # reference sim.py:240-256: end_day and n_days always agree

    def _init_results(self):
        ''' Result containers (reference sim.py:284-351) '''
        npts, nv = self.npts, self.pars['n_variants']
        R = {}
        for k in cvd.cum_result_flows + cvd.new_result_flows + tuple(f'n_{s}' for s in cvd.result_stocks):
            R[k] = Result(k, npts=npts)
        for k in cvd.other_results:
            R[k] = Result(k, npts=npts, scale=k not in cvd.unscaled_results)
        V = {}
        for k in ('prevalence_by_variant', 'incidence_by_variant'):
            V[k] = Result(k, npts=npts, scale=False, n_variants=nv)
        for k in cvd.cum_result_flows_by_variant + cvd.new_result_flows_by_variant + tuple(f'n_{s}' for s in cvd.result_stocks_by_variant):
            V[k] = Result(k, npts=npts, n_variants=nv)
        R['variant'] = V
        R['date'] = self.datevec
        R['t'] = self.tvec
        self.results = R
        scale = 1 if self.pars['rescale'] else self.pars['pop_scale']
        self.rescale_vec = scale * np.ones(npts)
        self.results_ready = False
        self._host_adds = {}

    def _init_people(self, **kwargs):
        pars = self.pars
        if pars['prognoses'] is None:
            pars['prognoses'] = cvpar.get_prognoses(pars['prog_by_age'])
        pop = self.popdict
        partitioned = self._partition is not None and self._partition is not False
        if pop is None and self.pop_gen == 'device':
            if self.rng_mode == 'mt':
                raise NotImplementedError('pop_gen="device" draws the population from keyed Philox uniforms; replay mode needs the reference\'s generators (pop_gen="host")')
            torch.cuda.set_device(self.device)
            gen = cvpop.KeyedPop(pars, pars['rand_seed'], self.device, cvpop.device_uniforms(pars['rand_seed'], self.device))
            if partitioned:                            # edges are streamed in chunks into the partitioned adjacency (_bind_partition)
                self._keyed_pop = gen
                pop = dict(age=gen.ages, sex=gen.sexes, contacts={lk: None for lk in gen.layer_keys()})
            else:
                pop = gen.materialize()
        if pop is None:
            exact = self.pop_exact if self.pop_exact is not None else (pars['pop_size'] <= 200_000)
            pop = cvpop.make_randpop(pars, self.rng, exact=exact)
        if partitioned:
            from . import partition as cvpart
            self._comm = cvpart.DistComm() if self._partition is True else self._partition
            if any(pars['dynam_layer'].get(lk) for lk in pop['contacts'].keys()):
                raise NotImplementedError('dynamic layers cannot be agent-partitioned (their edges are regenerated over the whole population every day)')
            self._chunk, ranges = cvpart.plan(self.n, self._comm.world)
            lo, hi = ranges[self._comm.rank]
            # every rank builds the same population from the same seed and keeps its own agents; the edge lists stay on the
            # host until the partitioned adjacency is built from them (_create_handle)
            self._global_layers = pop['contacts']
            empty = {lk: dict(p1=np.zeros(0, dtype=np.int32), p2=np.zeros(0, dtype=np.int32), beta=np.zeros(0, dtype=np.float32)) for lk in pop['contacts'].keys()}
            self.people = People(pars, self.device, age=pop['age'], sex=pop['sex'], contacts=empty, local_range=(lo, hi))
        else:
            self.people = People(pars, self.device, age=pop['age'], sex=pop['sex'], contacts=pop['contacts'])
        self.popdict = None
        lkeys = self.people.layer_keys()
        if len(lkeys) > _capi.MAX_LAYERS:
            raise ValueError(f'at most {_capi.MAX_LAYERS} contact layers are supported')
        cvpar.reset_layer_pars(pars, layer_keys=lkeys, force=False)
        self.people.set_prognoses(self.rng)

    def _create_handle(self):
        self._destroy()
        torch.cuda.set_device(self.device)
        pars = self.pars
        npts, nv = self.npts, pars['n_variants']
        h = C.c_void_p()
        _capi.call('cvb_create', C.byref(h), self.n_local, nv, npts, int(pars['rand_seed']))
        self._handle = h
        self._stream_ptr = None            # legacy default stream; torch's current stream is the same unless changed
        dev = self.device
        self._counters = torch.zeros((npts, cvd.N_COUNTERS), dtype=torch.int64, device=dev)
        self._vcounters = torch.zeros((npts, nv, cvd.N_VCOUNTERS), dtype=torch.int64, device=dev)
        self._sums = torch.zeros((npts, 4), dtype=torch.float64, device=dev)
        _capi.call('cvb_bind_results', h, self._counters.data_ptr(), self._vcounters.data_ptr(), self._sums.data_ptr())
        self._beds = None
        if self._comm is not None:         # bed limits compare GLOBAL counts (sim.py:579-580): the day's row is summed over the ranks
            self._beds = torch.zeros((npts, 2), dtype=torch.int64, device=dev)
            _capi.call('cvb_bind_beds', h, self._beds.data_ptr())
        cap = int(self.log_capacity) if self.log_capacity else int(max(4 * self.n_local, 1024))
        self._log = dict(source=torch.empty(cap, dtype=torch.int32, device=dev), target=torch.empty(cap, dtype=torch.int32, device=dev),
                         date=torch.empty(cap, dtype=torch.int32, device=dev), layer=torch.empty(cap, dtype=torch.int8, device=dev),
                         variant=torch.empty(cap, dtype=torch.int8, device=dev), count=torch.zeros(1, dtype=torch.int64, device=dev))
        L = self._log
        _capi.call('cvb_bind_log', h, L['source'].data_ptr(), L['target'].data_ptr(), L['date'].data_ptr(), L['layer'].data_ptr(),
                   L['variant'].data_ptr(), cap, L['count'].data_ptr())
        self.people._bind(self)
        if pars['use_waning']:
            kin = np.ascontiguousarray(pars['nab_kin'], dtype=np.float64)
            _capi.call('cvb_set_nab_kin', h, kin.ctypes.data, len(kin))
        self._quar_horizon = 1
        self._pars_dirty = True
        self._push_pars()
        if self._comm is not None:
            self._bind_partition()
        else:
            self._build_adjacency()

    def _bind_partition(self):
        '''
        Agent-partitioned run: exchange buffers + the adjacency of the edges that end in a local target, one row per
        GLOBAL source (partition.build_partition_adjacency; reference Contacts base.py:1509-1876).
        '''
        from . import partition as cvpart
        comm, dev, chunk = self._comm, self.device, self._chunk
        n_slots = comm.world * chunk
        B = dict(codes_local=torch.zeros(chunk, dtype=torch.uint8, device=dev), codes_global=torch.zeros(n_slots, dtype=torch.uint8, device=dev),
                 case_local=torch.zeros(chunk // 32, dtype=torch.int32, device=dev), case_global=torch.zeros(n_slots // 32, dtype=torch.int32, device=dev))
        self._part_bufs = B
        self._peer = comm.peer_exchange(dev, dict(codes=chunk, cases=chunk // 8)) if hasattr(comm, 'peer_exchange') else None
        _capi.call('cvb_set_partition', self._handle, self.id0, self.n, chunk, comm.world, self.people.rel_trans_global.data_ptr(),
                   B['codes_local'].data_ptr(), B['codes_global'].data_ptr(), B['case_local'].data_ptr(), B['case_global'].data_ptr(),
                   self._hit_capacity)
        lkeys = self.people.layer_keys()
        if self._keyed_pop is not None:
            gen, step = self._keyed_pop, 4_000_000

            def chunks():                              # agent chunks of every layer, never the whole edge list
                for i, lk in enumerate(lkeys):
                    for a in range(0, gen.plans[lk]['m'], step):
                        p1, p2, e0 = gen.layer_edges(lk, a, a + step)
                        yield i, p1, p2, None, e0
            ptr, adj, M = cvpart.build_partition_adjacency_chunks(chunks(), self.id0, self.id0 + self.n_local, n_slots, dev)
            self._keyed_pop = None
        else:
            ids = [i for i, lk in enumerate(lkeys) if len(self._global_layers[lk]['p1']) > 0]
            ptr, adj, M = cvpart.build_partition_adjacency([self._global_layers[lkeys[i]] for i in ids], ids, self.id0, self.id0 + self.n_local, n_slots, dev)
        self._global_layers = None
        mask = 0
        for i in range(len(lkeys)):
            mask |= 1 << i
        self._adj = (ptr, adj)
        self._adj_dirty = False
        _capi.call('cvb_bind_partition_adjacency', self._handle, ptr.data_ptr(), adj.data_ptr(), M, mask)

    def _exchange_codes(self):
        ''' One fixed-size all-gather per day: every agent's 1-byte transmit code (stream-ordered, no host synchronisation) '''
        B = self._part_bufs
        self._timed_collective('allgather_codes', B['codes_global'], B['codes_local'])

    def _exchange_cases(self):
        ''' All-gather of today's case bitmap (1 bit per agent), on days a contact_tracing intervention is active '''
        B = self._part_bufs
        self._timed_collective('allgather_cases', B['case_global'], B['case_local'])

    def _timed_collective(self, name, out, inp):
        timers = self.kernel_timers if self.kernel_timers is not None else getattr(self, 'collective_timers', None)
        if timers is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        if self._peer is not None:         # stores into every rank's buffer over NVLink + a signal barrier (partition.PeerExchange)
            which = 'codes' if name == 'allgather_codes' else 'cases'
            ptr = self._peer.all_gather(which, inp, self._stream_ptr)
            _capi.call('cvb_set_exchange_buffers', self._handle, ptr if which == 'codes' else None, ptr if which == 'cases' else None)
        else:
            self._comm.all_gather(out, inp)
        if timers is not None:
            b.record()
            timers.setdefault(name, []).append((a, b))

    def _choose_true(self, key, stream, k):
        '''
        Up to ``k`` distinct agents among those whose flag ``key`` is set, over all ranks: positions in the ascending list
        of flagged agents are drawn on the host (cvu.choose_distinct), the flagged list itself never leaves the device.
        Returns (global ids owned by this rank as a device tensor, number chosen overall).
        '''
        nz = torch.nonzero(self.people[key]).flatten()
        counts = self._global_counts(int(nz.numel()))
        k = min(int(k), sum(counts))
        who = self._pick_positions(nz, counts, cvu.choose_distinct(stream, sum(counts), k) if k > 0 else np.zeros(0, dtype=np.int64)) + self.id0
        return who.to(torch.int32), k

    def _global_counts(self, n_local_flagged):
        ''' How many flagged agents every rank holds (one object gather under a partition; [n] otherwise) '''
        return [int(n_local_flagged)] if self._comm is None else [int(c) for c in self._comm.gather_objects(int(n_local_flagged))]

    def _k_smallest_mask(self, vals, mask, k):
        '''
        Among this rank's agents with ``mask`` set, those whose value is one of the ``k`` smallest over ALL ranks (values are distinct:
        sequence positions, continuous keys).  Every rank offers its own k smallest; the k-th smallest offer is the threshold.
        '''
        k = int(k)
        if k <= 0:
            return torch.zeros_like(mask)
        cand = vals[mask]
        k_local = min(k, int(cand.numel()))
        mine = torch.topk(cand, k_local, largest=False, sorted=False).values.cpu().numpy() if k_local else np.zeros(0, dtype=cand.cpu().numpy().dtype)
        offers = np.concatenate(self._comm.gather_objects(mine)) if self._comm is not None else mine
        if len(offers) < k:                                  # fewer candidates than k over all ranks (a rank with k or more offers k by itself)
            return mask.clone()
        threshold = np.partition(offers, k - 1)[k - 1].item()
        return mask & (vals <= threshold)

    def _pick_positions(self, flagged, counts, pos):
        ''' ``pos``: positions in the ascending list of flagged agents over ALL ranks (identical on every rank) -> this rank's local indices '''
        rank = 0 if self._comm is None else self._comm.rank
        off = sum(counts[:rank])
        pos = np.asarray(pos, dtype=np.int64)
        mine = pos[(pos >= off) & (pos < off + counts[rank])] - off
        return flagged[torch.as_tensor(mine, dtype=torch.int64, device=self.device)]

    def _build_adjacency(self):
        '''
        Device-resident bidirectional adjacency (CSR over agents) of the static layers -- the CSR form of the
        reference's Contacts (base.py:1509-1876).  For agent i, entries adj[ptr[i]:ptr[i+1]] are 16 bytes each:
        (neighbour, edge index within its layer, (layer << 1) | direction, beta).  Built once (and again if a layer's
        edge list is changed through the Layer API); transmission and tracing then visit only the edges of
        today's transmitters / cases.  Dynamic layers (regenerated every day) keep the dense streaming passes.
        '''
        self._adj_dirty = False
        self._adj_mask = 0
        self._sync_edges()
        if self._comm is not None:
            raise NotImplementedError('contact layers of an agent-partitioned simulation cannot be edited after initialisation')
        people, pars = self.people, self.pars
        if self.rng_mode == 'mt':                      # replay mode walks the edge lists in the reference's order
            self._adj = None
            return
        static = [i for i, lk in enumerate(people.layer_keys()) if not pars['dynam_layer'].get(lk) and len(people.contacts[lk]) > 0]
        if not self.use_adjacency or not static:
            self._adj = None
            _capi.call('cvb_bind_adjacency', self._handle, None, None, 0, 0)
            return
        dev = self.device
        src, nbr, eid, meta, wts = [], [], [], [], []
        for i in static:
            layer = list(people.contacts.values())[i]
            E = len(layer)
            if E >= 2 ** 31:
                raise ValueError('a layer with 2^31 or more edges cannot be indexed by the adjacency')
            e = torch.arange(E, dtype=torch.int32, device=dev)
            for d, (a, b) in enumerate(((layer['p1'], layer['p2']), (layer['p2'], layer['p1']))):
                src.append(a)
                nbr.append(b)
                eid.append(e)
                meta.append(torch.full((E,), (i << 1) | d, dtype=torch.int32, device=dev))
                wts.append(layer['beta'])
        src = torch.cat(src)
        order = torch.sort(src, stable=True).indices
        M = int(src.numel())
        adj = torch.empty((M, 4), dtype=torch.int32, device=dev)
        adj[:, 0] = torch.cat(nbr)[order]
        adj[:, 1] = torch.cat(eid)[order]
        adj[:, 2] = torch.cat(meta)[order]
        adj[:, 3] = torch.cat(wts)[order].view(torch.int32)
        ptr = torch.zeros(self.n + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(torch.bincount(src.to(torch.int64), minlength=self.n), 0)
        mask = 0
        for i in static:
            mask |= 1 << i
        self._adj = (ptr, adj)                      # keep the tensors alive while they are bound
        self._adj_mask = mask
        _capi.call('cvb_bind_adjacency', self._handle, ptr.data_ptr(), adj.data_ptr(), M, mask)

    def _set_quar_horizon(self, horizon):
        if horizon > self._quar_horizon:
            _capi.call('cvb_set_quar_horizon', self._handle, int(horizon))
            self._quar_horizon = int(horizon)

    # ---- checkpoint / restore (reference base.py:682-741 Sim.save/load, sim.py:688-761 resume) ----------
    _IV_HOST_TYPES = (set, dict, list, tuple, int, float, bool, str, type(None), np.ndarray, np.integer, np.floating)

    def snapshot(self, pinned=True, compact=None):
        '''
        Host copy of everything a run mutates: the People arena (every per-agent array, one buffer), every layer's edge list, the
        RNG streams, the clock, the result tables, the rescale vector, and the state interventions keep (device arrays and host
        bookkeeping such as pending second doses or the edges clip_edges holds back).  With ``pinned=True`` the buffers are
        page-locked so ``restore`` is one asynchronous H2D copy for the People and one per edge array.  ``compact`` (default: with
        pinned buffers) also keeps the arena in compact form -- per array either "one value + exceptions" or dense -- which is what
        ``restore`` then sends: at day 0 of the 1M-agent benchmark sim 33 of 202 bytes per agent are dense.
        '''
        self._sync_edges()
        torch.cuda.synchronize(self.device)

        def host(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=pinned)
            h.copy_(t)
            return h

        def layer_cols(layer):
            return {c: host(layer[c]) for c in layer.columns}

        iv_dev, iv_host, iv_layers = [], [], []
        for iv in self.pars['interventions']:
            attrs = vars(iv) if hasattr(iv, '__dict__') else {}
            iv_dev.append({k: host(v) for k, v in attrs.items() if isinstance(v, torch.Tensor)})
            iv_host.append({k: copy.deepcopy(v) for k, v in attrs.items()
                            if isinstance(v, self._IV_HOST_TYPES) and not k.startswith('__') and k not in ('label',)})
            held = attrs.get('contacts')                    # clip_edges: the edges it has taken out of the simulation
            iv_layers.append({lk: layer_cols(l) for lk, l in held.items()} if isinstance(held, dict) else None)
        snap = dict(t=self.t, complete=self.complete, arena=host(self.people._arena),
                    layers={lk: layer_cols(l) for lk, l in self.people.contacts.items()},
                    counters=host(self._counters), vcounters=host(self._vcounters), sums=host(self._sums),
                    log_count=host(self._log['count']), host_adds={k: v.copy() for k, v in self._host_adds.items()},
                    rescale_vec=self.rescale_vec.copy(),
                    rng=(self.rng.seed, self.rng.np_.get_state(), self.rng.nb.get_state()),
                    pars={k: copy.deepcopy(v) for k, v in self.pars.items() if k not in ('interventions', 'analyzers', 'variants', 'prognoses', 'nab_kin')},
                    iv=iv_dev, iv_host=iv_host, iv_layers=iv_layers, quar_horizon=self._quar_horizon)
        if self._quar_horizon > 1:                          # quarantine requests that start on one of the coming days (contact tracing with a delay)
            buf = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            ring = {}
            for d in range(self.t, self.t + self._quar_horizon):
                _capi.call('cvb_pending_quarantine', self._handle, int(d), buf.data_ptr(), self._stream_ptr)
                if bool((buf >= 0).any()):
                    ring[d] = host(buf)
            snap['quar_ring'] = ring
        torch.cuda.synchronize(self.device)
        if pinned if compact is None else compact:
            snap['arena_compact'] = self._compact_arena(snap['arena'], pinned)
        return snap

    def _compact_arena(self, arena_host, pinned=True):
        '''
        The saved arena as {fill segments + exceptions, dense runs}: per array, the most common 32-bit word and the words that
        differ from it if those are few (8 + 8 bytes each against 4 bytes per word of a dense copy), else the array as it is.
        '''
        words = arena_host.numpy().view(np.uint32)
        segs, exc_idx, exc_val, dense = [], [], [], []
        for name, dt, shape, off, nbytes in self.people._layout:
            w0, w1 = off // 4, (off + nbytes + 3) // 4
            fw = words[w0:w1]
            if len(fw) == 0:
                continue
            vals, counts = np.unique(fw[::max(1, len(fw) // 4096)], return_counts=True)
            v = vals[np.argmax(counts)]
            idx = np.flatnonzero(fw != v)
            if len(idx) * 16 <= len(fw):                    # at most a quarter of the dense bytes
                segs.append((w0, w1 - w0, int(v)))
                exc_idx.append(idx.astype(np.int64) + w0)
                exc_val.append(fw[idx].astype(np.int64))
            elif dense and dense[-1][1] + 256 >= off:        # next to the previous dense array (only alignment padding between)
                dense[-1] = (dense[-1][0], (w1 * 4 + 255) // 256 * 256)
            else:
                dense.append((off, (w1 * 4 + 255) // 256 * 256))
        n_exc = int(sum(len(i) for i in exc_idx))
        table = np.concatenate([np.asarray(segs, dtype=np.int64).reshape(-1), *exc_idx, *exc_val]) if (segs or n_exc) else np.zeros(0, dtype=np.int64)
        t = torch.empty(max(len(table), 1), dtype=torch.int64, pin_memory=pinned)
        t[:len(table)] = torch.from_numpy(table)
        total = arena_host.numel()
        dense = [(a, min(b, total)) for a, b in dense]
        return dict(table=t, n_seg=len(segs), n_exc=n_exc, dense=dense, arena_bytes=total,
                    h2d_bytes=int(len(table) * 8 + sum(b - a for a, b in dense)))

    def _restore_arena(self, snap):
        ''' The People arena of a snapshot back on the device: compact form if the snapshot has one (one small table + the dense arrays), else one copy '''
        arena, c = self.people._arena, snap.get('arena_compact')
        if c is None or c['arena_bytes'] != arena.numel():
            arena.copy_(snap['arena'], non_blocking=True)
            return
        dev = getattr(self, '_compact_table', None)
        if dev is None or dev.numel() < c['table'].numel():
            dev = self._compact_table = torch.empty(c['table'].numel(), dtype=torch.int64, device=self.device)
        dev[:c['table'].numel()].copy_(c['table'], non_blocking=True)
        _capi.call('cvb_restore_compact', arena.data_ptr(), arena.numel(), dev.data_ptr(), c['n_seg'], c['n_exc'], self._stream_ptr)
        for a, b in c['dense']:
            arena[a:b].copy_(snap['arena'][a:b], non_blocking=True)

    def restore_light(self, snap):
        ''' restore() without the People / Layer copies: for callers that rewound those arrays on the device themselves '''
        return self.restore(snap, arrays=False)

    def _sync_edges(self):
        ''' Make the current stream wait for the edge lists restore() is still copying on the side stream '''
        ev, self._edges_event = getattr(self, '_edges_event', None), None
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def restore(self, snap, arrays=True):
        '''
        Put a snapshot back on the device and rewind the clock.  The People arena is ONE asynchronous copy from pinned memory on
        the current stream; the edge lists are copied on a side stream, concurrently with the days that follow: the transmission
        and tracing kernels of static layers read the adjacency, which does not change, so only code that touches the raw edge
        arrays (dynamic layers, the per-step path, Python access through Layer, finalize) waits for them (_sync_edges).
        '''
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        if arrays:
            self._sync_edges()
            self._restore_arena(snap)
            cur = torch.cuda.current_stream(self.device)
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            cs = self._copy_stream
            cs.wait_stream(cur)                            # earlier kernels may still be reading the arrays that are overwritten
            resized = False
            with torch.cuda.stream(cs):
                for lk, cols in snap['layers'].items():
                    layer = self.people.contacts[lk]
                    for c, h in cols.items():
                        if layer._cols[c].shape == h.shape:
                            layer._cols[c].copy_(h, non_blocking=True)
                        else:
                            layer._cols[c] = h.to(self.device, non_blocking=True)
                            resized = True
                self._edges_event = cs.record_event()
            if resized:                                     # a layer changed size since the snapshot (clip_edges): bind the new arrays
                self._sync_edges()
                for layer in self.people.contacts.values():
                    layer._rebind()
        self._counters.copy_(snap['counters'], non_blocking=True)
        self._vcounters.copy_(snap['vcounters'], non_blocking=True)
        self._sums.copy_(snap['sums'], non_blocking=True)
        self._log['count'].copy_(snap['log_count'], non_blocking=True)
        from .base import Layer
        for iv, dev, hst, held in zip(self.pars['interventions'], snap['iv'], snap['iv_host'], snap['iv_layers']):
            for k, h in dev.items():
                getattr(iv, k).copy_(h, non_blocking=True)
            for k, v in hst.items():
                setattr(iv, k, copy.deepcopy(v))
            if held is not None:
                iv.contacts = {lk: Layer(cols['p1'], cols['p2'], cols['beta'], label=lk, device=self.device) for lk, cols in held.items()}
        _capi.call('cvb_reset', self._handle, self._stream_ptr)
        if snap.get('quar_ring'):                            # put back the requests that were pending when the snapshot was taken
            self._set_quar_horizon(snap['quar_horizon'])
            buf = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            for d, h in snap['quar_ring'].items():
                buf.copy_(h, non_blocking=True)
                _capi.call('cvb_set_pending_quarantine', self._handle, int(d), buf.data_ptr(), self._stream_ptr)
        _capi.call('cvb_state_invalidate', self._handle)
        self.fused_days = 0
        self._host_adds = {k: v.copy() for k, v in snap['host_adds'].items()}
        self.rescale_vec = snap['rescale_vec'].copy()
        seed, np_state, nb_state = snap['rng']
        self.rng.seed = seed
        self.rng.np_.set_state(np_state)
        self.rng.nb.set_state(nb_state)
        for k, v in snap['pars'].items():
            self.pars[k] = copy.deepcopy(v)
        self._pars_dirty = True
        self.t = snap['t']
        self.complete = snap['complete']
        self.results_ready = False
        self._orig_pars = None
        for k in self.result_keys():
            self.results[k].values[:] = 0
        for k in self.result_keys('variant'):
            self.results['variant'][k].values[:] = 0
        return self

    def edge_work(self):
        ''' Per-day (adjacency entries visited, transmitters) of the sparse edge pass, int64[npts, 2] '''
        out = np.zeros((self.npts, 2), dtype=np.int64)
        _capi.call('cvb_get_edge_work', self._handle, out.ctypes.data)
        return out

    def h2d_bytes(self, snap):
        ''' Bytes restore() copies host -> device '''
        n = snap['arena_compact']['h2d_bytes'] if snap.get('arena_compact') is not None else snap['arena'].numel() * snap['arena'].element_size()
        n += sum(h.numel() * h.element_size() for cols in snap['layers'].values() for h in cols.values())
        n += sum(snap[k].numel() * snap[k].element_size() for k in ('counters', 'vcounters', 'sums', 'log_count'))
        n += sum(h.numel() * h.element_size() for saved in snap['iv'] for h in saved.values())
        return int(n)

    def d2h_bytes(self):
        ''' Bytes finalize() reads device -> host (the result tables + the three sums compute_r_eff needs) '''
        n = sum(t.numel() * t.element_size() for t in (self._counters, self._vcounters, self._sums))
        return int(n + 24)

    # ---- parameters -> device struct ---------------------------------------------------------------
    def _pars_fingerprint(self):
        p = self.pars
        return (p['beta'], p['rel_beta'], p['asymp_factor'], tuple(p['beta_layer'].values()), tuple(p['iso_factor'].values()),
                tuple(p['quar_factor'].values()), p['n_beds_hosp'], p['n_beds_icu'], p['no_hosp_factor'], p['no_icu_factor'],
                p['rel_symp_prob'], p['rel_severe_prob'], p['rel_crit_prob'], p['rel_death_prob'], p['trans_redux'], p['nab_boost'],
                tuple(tuple(v.values()) for v in p['variant_pars'].values()), len(p['vaccine_pars']), p['quar_period'],
                tuple(p['viral_dist'].values()), tuple(tuple(d.values()) for d in p['dur'].values()),
                tuple(p['nab_init'].values()) if p['nab_init'] else None, tuple(p['nab_eff'].values()) if p['nab_eff'] else None,
                None if p['immunity'] is None else np.asarray(p['immunity']).tobytes(),
                tuple(tuple(sorted((k, float(x)) for k, x in v.items() if isinstance(x, (int, float)))) for v in p['vaccine_pars'].values()))

    def _push_pars(self):
        ''' Re-send the scalar block when an intervention changed a parameter (reference sim.py:602-642 re-reads them daily) '''
        key = self._pars_fingerprint()
        if not self._pars_dirty and key == self._pars_key:
            return
        p = self.pars
        nv = p['n_variants']
        lkeys = self.people.layer_keys()
        c = _capi.cvb_pars()
        c.n_variants, c.n_layers, c.use_waning, c.n_vaccines = nv, len(lkeys), int(bool(p['use_waning'])), len(p['vaccine_pars'])
        c.quar_period = int(p['quar_period'])
        c.has_vaccine_pars = int(len(p['vaccine_pars']) > 0)
        c.n_beds_hosp = -1 if p['n_beds_hosp'] is None else int(p['n_beds_hosp'])
        c.n_beds_icu = -1 if p['n_beds_icu'] is None else int(p['n_beds_icu'])
        c.asymp_factor = p['asymp_factor']
        vd = p['viral_dist']
        c.frac_time, c.load_ratio, c.high_cap = vd['frac_time'], vd['load_ratio'], vd['high_cap']
        c.trans_redux, c.no_hosp_factor, c.no_icu_factor, c.nab_boost = p['trans_redux'], p['no_hosp_factor'], p['no_icu_factor'], p['nab_boost']
        for v in range(nv):
            vp = p['variant_pars'][p['variant_map'][v]]
            c.beta[v] = float(f32(p['beta'] * p['rel_beta'] * vp['rel_beta']))                # reference sim.py:627
            for name, key_ in (('rel_symp', 'rel_symp_prob'), ('rel_severe', 'rel_severe_prob'), ('rel_crit', 'rel_crit_prob'), ('rel_death', 'rel_death_prob')):
                val = p[key_] * (vp[key_] if v else 1.0)                                       # reference people.py:476-481
                getattr(c, name)[v] = float(f32(val))
        for i, lk in enumerate(lkeys):
            c.beta_layer[i], c.iso_factor[i], c.quar_factor[i] = p['beta_layer'][lk], p['iso_factor'][lk], p['quar_factor'][lk]
        if p['use_waning']:
            imm = np.asarray(p['immunity'], dtype=np.float32)
            for a in range(nv):
                for b in range(nv):
                    c.immunity[a][b] = float(imm[a, b])
            for num, vkey in p['vaccine_map'].items():
                if num >= _capi.MAX_VACCINES:
                    raise ValueError(f'at most {_capi.MAX_VACCINES} vaccines are supported')
                for v in range(nv):
                    c.vaccine_imm[num][v] = float(p['vaccine_pars'][vkey][p['variant_map'][v]])
            e = p['nab_eff']
            c.exp_alpha_inf, c.beta_inf = float(np.exp(e['alpha_inf'])), e['beta_inf']
            c.exp_alpha_symp_inf, c.beta_symp_inf = float(np.exp(e['alpha_symp_inf'])), e['beta_symp_inf']
            c.exp_alpha_sev_symp, c.beta_sev_symp = float(np.exp(e['alpha_sev_symp'])), e['beta_sev_symp']
            c.nab_norm = 1 + e['alpha_inf_diff']
            ris = p['rel_imm_symp']
            c.rel_imm_asymp, c.rel_imm_mild, c.rel_imm_severe = ris['asymp'], ris['mild'], ris['severe']
            c.nab_init = _capi.dist_struct(p['nab_init'])
        for i, name in enumerate(_capi.DUR_ORDER):
            c.dur[i] = _capi.dist_struct(p['dur'][name])
        _capi.call('cvb_set_pars', self._handle, C.byref(c))
        self._cpars = c
        self._pars_key = key
        self._pars_dirty = False

    # ---- seeding (reference sim.py:505-532) ----------------------------------------------------------
    def init_infections(self, force=False):
        pars = self.pars
        if pars['frac_susceptible'] < 1:                       # reference sim.py:519-521: a random share is not susceptible
            n = int(np.round((1 - pars['frac_susceptible']) * pars['pop_size']))
            if self.rng_mode == 'mt':
                inds = self.rng.nb.choice(pars['pop_size'], n, replace=False)                           # cvu.choose: Numba stream
            else:
                inds = cvu.choose_distinct(self.rng.nb, pars['pop_size'], n)
            self.people.make_nonnaive(inds)
        if pars['pop_infected']:
            if self.rng_mode == 'mt':
                inds = self.rng.nb.choice(pars['pop_size'], int(pars['pop_infected']), replace=False)   # cvu.choose: Numba stream
            else:
                inds = cvu.choose_distinct(self.rng.nb, pars['pop_size'], int(pars['pop_infected']))
            self.people.infect(inds, layer='seed_infection', count_flows=False)

    def _log_append(self, source, target, layer, variant):
        ''' Append transmissions to the device infection log from Python (replay mode); the kernels append directly '''
        L = self._log
        n = len(target)
        if n == 0:
            return
        pos = int(L['count'].item())
        end = min(pos + n, len(L['source']))
        k = end - pos
        lkeys = self.people.layer_keys()
        code = lkeys.index(layer) if layer in lkeys else {'seed_infection': _capi.LAYER_SEED}.get(layer, _capi.LAYER_IMPORT)
        L['target'][pos:end] = target[:k].to(torch.int32)
        L['source'][pos:end] = -1 if source is None else source[:k].to(torch.int32)
        L['date'][pos:end] = int(self.t)
        L['layer'][pos:end] = code
        L['variant'][pos:end] = int(variant)
        L['count'] += n

    def _host_add(self, key, t, value):
        ''' Host-side contribution to a result (e.g. n_imports), merged with the device counters at sync time '''
        arr = self._host_adds.setdefault(key, np.zeros(self.npts))
        arr[t] += value

    # ---- one day (reference sim.py:558-685) ----------------------------------------------------------
    def step(self):
        if self.complete:
            raise AlreadyRunError('Simulation already complete (call sim.initialize() to re-run)')
        if self.rng_mode == 'mt':
            from . import replay
            if torch.cuda.current_device() != self.device.index:
                torch.cuda.set_device(self.device)
            return replay.step(self)
        t, pars, people, h, st = self.t, self.pars, self.people, self._handle, self._stream_ptr
        if torch.cuda.current_device() != self.device.index:      # ensembles keep members on several GPUs in one process
            torch.cuda.set_device(self.device)
        if self._fusable_day(t):                                   # nothing for the host to decide today: the fused day kernels
            people.t = t
            self._run_block(t, t + 1)
            if pars['analyzers']:
                self.t, self.complete = t, False                   # analyzers see the day that has just been simulated (sim.py:677-678)
                for an in pars['analyzers']:
                    an(self)
                _capi.call('cvb_state_invalidate', h)              # an analyzer may have written People arrays
                self.t = t + 1
                self.complete = self.t == self.npts
            return
        _capi.call('cvb_state_invalidate', h)                      # Python (interventions, rescaling) may write People arrays today
        self._sync_edges()
        people.t = t
        call = _capi.call if self.kernel_timers is None else self._timed_call
        self.rescale()
        self._push_pars()
        if self._adj_dirty:
            self._build_adjacency()
        call('cvb_update_states_pre', h, t, st)
        if self._beds is not None and (pars['n_beds_hosp'] is not None or pars['n_beds_icu'] is not None):
            self._comm.all_reduce_sum(self._beds[t])             # 16 bytes: today's severe / critical counts over all ranks
        for lkey, dyn in pars['dynam_layer'].items():                                 # reference people.py:199-206
            if dyn:
                people.contacts[lkey].update(people)
        if pars['n_imports']:                                                          # reference sim.py:583-588
            n_imports = int(self.rng.nb.poisson(f32(pars['n_imports'] / self.rescale_vec[t]), 1)[0])
            if n_imports > 0:
                who = cvu.choose_distinct(self.rng.nb, pars['pop_size'], n_imports)          # (replay mode has its own step)
                people.infect(who, hosp_max='auto', icu_max='auto', layer='importation')
                self._host_add('n_imports', t, n_imports)
        for v in pars['variants']:
            v.apply(self)
        for iv in pars['interventions']:
            iv(self)
        self._push_pars()
        if self._adj_dirty:                    # an intervention edited a layer's edge list
            self._build_adjacency()
        call('cvb_post_and_prepare', h, t, st)
        if self._comm is not None:
            self._exchange_codes()
        call('cvb_edge_pass', h, t, st)
        call('cvb_infect_winners', h, t, st)
        call('cvb_update_nab_count', h, t, st)
        for an in pars['analyzers']:
            an(self)
        self.t += 1
        if self.t == self.npts:
            self.complete = True

    def _call(self, name, *args):
        ''' C-ABI call, timed with CUDA events when ``kernel_timers`` is a dict (interventions use this too) '''
        if self.kernel_timers is None:
            return _capi.call(name, *args)
        return self._timed_call(name, *args)

    def rescale(self):
        '''
        Dynamic rescaling (reference sim.py:535-555): once more than rescale_threshold of the agents are no longer naive, a
        random share of them is made naive again and every agent stands for more people from today on.  Needs one count per
        day (a device synchronisation; plus one object gather over the ranks of a partitioned run) while there is still room to rescale;
        the chosen agents are reset on the device.
        '''
        pars = self.pars
        if not pars['rescale']:
            return
        pop_scale, current = pars['pop_scale'], self.rescale_vec[self.t]
        if current < pop_scale:
            not_naive = torch.nonzero(~self.people.naive).flatten()                         # (this rank's agents under a partition)
            counts = self._global_counts(int(not_naive.numel()))
            n_not_naive, n_people = sum(counts), pars['pop_size']
            ratio, threshold = n_not_naive / n_people, pars['rescale_threshold']
            if ratio > threshold:
                scaling = min(max(ratio / threshold, pars['rescale_factor']), pop_scale / current)
                self.rescale_vec[self.t:] *= scaling
                n = int(round(n_not_naive * (1.0 - 1.0 / scaling)))
                if self.rng_mode == 'mt':
                    choices = self.rng.nb.choice(n_not_naive, n, replace=False)                # cvu.choose: Numba stream
                else:
                    choices = cvu.choose_distinct(self.rng.nb, n_not_naive, n)                 # the same positions on every rank (host streams in step)
                self.people.make_naive(self._pick_positions(not_naive, counts, choices))

    def _timed_call(self, name, *args):
        ''' _capi.call bracketed by CUDA events on the launching stream (bench.py's per-kernel timing) '''
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _capi.call(name, *args)
        b.record()
        self.kernel_timers.setdefault(name, []).append((a, b))

    def run(self, until=None, reset_seed=True, restore_pars=True, verbose=None, **kwargs):
        ''' Run to the end (or to ``until``) and finalize (reference sim.py:688-761) '''
        if not self.initialized:
            self.initialize()
        if self._orig_pars is None:
            self._orig_pars = {k: copy.deepcopy(v) for k, v in self.pars.items() if k not in ('interventions', 'analyzers', 'variants', 'prognoses', 'nab_kin')}
        if reset_seed:
            self.set_seed()
        until = self.npts if until is None else self.day(until)
        if until > self.npts:
            raise AlreadyRunError(f'Requested to run until t={until} but the simulation end is t={self.npts}')
        if self.t >= until:
            raise AlreadyRunError(f'Simulation is currently at t={self.t}, requested to run until t={until} which has already been reached')
        if self.complete:
            raise AlreadyRunError('Simulation is already complete (call sim.initialize() to re-run)')
        if self._advance(until):
            return self
        if self.complete:
            self.finalize(restore_pars=restore_pars)
        return self

    def _advance(self, until):
        ''' The day loop of run(): stretches of days that need no host decision are ONE C-ABI call each; True if stopping_func fired '''
        _capi.call('cvb_state_invalidate', self._handle)          # the caller may have written People arrays since the last run
        while self.t < until:
            if self.pars['stopping_func'] and self.pars['stopping_func'](self):
                return True
            if self.rng_mode == 'philox' and not self.pars['analyzers'] and not self.pars['stopping_func'] and self._fusable_day(self.t):
                t1 = self.t + 1
                while t1 < until and self._fusable_day(t1):
                    t1 += 1
                self._run_block(self.t, t1)
            else:
                self.step()
        return False

    def fused_timing(self, enable=None):
        '''
        Per-kernel device time of the fused day loop: ``fused_timing(True)`` / ``(False)`` switches the CUDA-event timing of
        cvb_run_days on / off; ``fused_timing()`` returns {kernel: (milliseconds, launches)} accumulated since the last read.
        '''
        if enable is not None:
            _capi.call('cvb_timing_enable', self._handle, int(bool(enable)))
            return None
        ms = np.zeros(len(_capi.TIMED_KINDS), dtype=np.float64)
        cnt = np.zeros(len(_capi.TIMED_KINDS), dtype=np.int64)
        _capi.call('cvb_timing_read', self._handle, ms.ctypes.data, cnt.ctypes.data)
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_capi.TIMED_KINDS)}

    # ---- results ---------------------------------------------------------------------------------------
    def sync_results(self):
        ''' Copy the device counter tables into the host Result arrays (raw per-day values, unscaled) '''
        torch.cuda.synchronize(self.device)
        counters, vcounters, sums_t = self._counters, self._vcounters, self._sums
        if self._comm is not None:             # partitioned: every table is a sum over ranks (one collective per table, once per run)
            status = np.zeros(2, dtype=np.int64)
            _capi.call('cvb_partition_status', self._handle, status.ctypes.data)
            counters, vcounters, sums_t = counters.clone(), vcounters.clone(), sums_t.clone()
            flags = torch.as_tensor(status, device=self.device)
            for tsr in (counters, vcounters, sums_t, flags):
                self._comm.all_reduce_sum(tsr)
            dropped, bad_trans = (int(x) for x in flags.cpu().numpy())
            if dropped:
                raise RuntimeError(f'{dropped} transmissions were dropped: raise hit_capacity (Sim(..., hit_capacity=...))')
            if bad_trans:
                raise RuntimeError(f'rel_trans of {bad_trans} infectious agent-days was neither its initial value nor that value times trans_redux: '
                                   'agent-partitioned runs rebuild remote transmissibility from the initial value')
        cnt = counters.cpu().numpy().astype(np.float64)
        vcnt = vcounters.cpu().numpy().astype(np.float64)
        sums = sums_t.cpu().numpy()
        R = self.results
        nv = self.pars['n_variants']
        for k, cid in cvd.COUNTER_IDS.items():
            if k in R:
                R[k].values[:] = cnt[:, cid]
        factor = self.pars['pop_scale'] / self.rescale_vec             # reference interventions.py:979, 1477
        for k in ('new_tests', 'new_doses', 'new_vaccinated'):
            R[k].values[:] = R[k].values * factor
        for k, add in self._host_adds.items():
            R[k].values[:] = R[k].values + add
        for k, vid in cvd.VCOUNTER_IDS.items():
            R['variant'][k].values[:] = vcnt[:, :, vid].T
        n_alive = cnt[:, cvd.COUNTER_IDS['n_alive_agents']]
        with np.errstate(all='ignore'):
            R['pop_nabs'].values[:] = np.where(n_alive > 0, sums[:, 0] / n_alive, 0.0)
            R['pop_protection'].values[:] = sums[:, 1] / (nv * self.n)
            R['pop_symp_protection'].values[:] = sums[:, 2] / (nv * self.n)
        return self.results

    @property
    def infection_log(self):
        ''' The infection log as host arrays sorted by (date, variant, layer, target) (reference people.py:508-511) '''
        torch.cuda.synchronize(self.device)
        L = self._log
        count, cap = int(L['count'].item()), len(L['source'])
        if count > cap:
            import warnings
            warnings.warn(f'the infection log holds {cap} entries but {count} infections happened: the last {count - cap} are missing '
                          f'(raise it with Sim(..., log_capacity=...)); compute_r_eff / compute_gen_time over the log are affected', RuntimeWarning)
        n = min(count, cap)
        out = {k: L[k][:n].cpu().numpy() for k in ('source', 'target', 'date', 'layer', 'variant')}
        kept = out['target'] >= 0                 # (a candidate that was not infected after all leaves a placeholder in its slot)
        out = {k: v[kept] for k, v in out.items()}
        if self._comm is not None:             # every rank logs the infections of its own agents (global ids)
            parts = self._comm.gather_objects(out)
            out = {k: np.concatenate([p[k] for p in parts]) for k in out}
        order = np.lexsort((out['target'], out['layer'], out['variant'], out['date']))
        return {k: v[order] for k, v in out.items()}

    def finalize(self, restore_pars=True, **kwargs):
        ''' Cumulative and derived results (reference sim.py:764-1072) '''
        if self.results_ready:
            raise AlreadyRunError('Simulation has already been finalized')
        self._sync_edges()
        self.sync_results()
        R, pars = self.results, self.pars
        rv = self.rescale_vec
        for k in self.result_keys():
            if R[k].scale:
                R[k].values *= rv
        for k in self.result_keys('variant'):
            if R['variant'][k].scale:
                R['variant'][k].values = R['variant'][k].values * rv[None, :]
        for k in cvd.result_flows:
            R[f'cum_{k}'].values[:] = np.cumsum(R[f'new_{k}'].values)
        for k in cvd.result_flows_by_variant:
            R['variant'][f'cum_{k}'].values[:] = np.cumsum(R['variant'][f'new_{k}'].values, axis=1)
        R['cum_infections'].values += pars['pop_infected'] * rv[0]
        R['variant']['cum_infections_by_variant'].values += pars['pop_infected'] * rv[0]
        for iv in pars['interventions']:
            if hasattr(iv, 'finalize'):
                iv.finalize(self)
        for an in pars['analyzers']:
            if hasattr(an, 'finalize'):
                an.finalize(self)
        self.results_ready = True
        self.t -= 1
        self.compute_results()
        if restore_pars and self._orig_pars:
            for k, v in self._orig_pars.items():
                self.pars[k] = v
            self._orig_pars = None
            self._pars_dirty = True
        return self

    def compute_results(self):
        self.compute_states()
        self.compute_yield()
        self.compute_doubling()
        self.compute_r_eff()
        self.compute_summary()

    def compute_states(self):
        ''' reference sim.py:808-837 '''
        R = self.results
        v = lambda k: R[k].values
        count_recov = 1 - self.pars['use_waning']
        with np.errstate(all='ignore'):
            R['n_alive'].values[:] = self.scaled_pop_size - v('cum_deaths')
            R['n_naive'].values[:] = self.scaled_pop_size - v('cum_deaths') - v('n_recovered') - v('n_exposed')
            R['n_susceptible'].values[:] = v('n_alive') - v('n_exposed') - count_recov * v('cum_recoveries')
            R['n_preinfectious'].values[:] = v('n_exposed') - v('n_infectious')
            R['n_removed'].values[:] = count_recov * v('cum_recoveries') + v('cum_deaths')
            R['prevalence'].values[:] = v('n_exposed') / v('n_alive')
            R['incidence'].values[:] = v('new_infections') / v('n_susceptible')
            R['frac_vaccinated'].values[:] = v('n_vaccinated') / v('n_alive')
            V = R['variant']
            V['incidence_by_variant'].values[:] = V['new_infections_by_variant'].values / v('n_susceptible')[None, :]
            V['prevalence_by_variant'].values[:] = V['new_infections_by_variant'].values / v('n_alive')[None, :]

    def compute_yield(self):
        ''' reference sim.py:840-855 '''
        R = self.results
        v = lambda k: R[k].values
        with np.errstate(all='ignore'):
            nz = np.nonzero(v('new_tests'))[0]
            R['test_yield'].values[nz] = v('new_diagnoses')[nz] / v('new_tests')[nz]
            nz = np.nonzero(v('n_infectious'))[0]
            denom = v('n_infectious')[nz] / (v('n_alive')[nz] - v('cum_diagnoses')[nz])
            R['rel_test_yield'].values[nz] = v('test_yield')[nz] / denom

    def compute_doubling(self, window=3, max_doubling_time=30):
        ''' reference sim.py:858-885 '''
        ci = self.results['cum_infections'].values
        now, prev = ci[window:], ci[:-window]
        use = (prev > 0) & (now > prev)
        out = np.full(self.npts, np.nan)
        tail = out[window:]
        with np.errstate(all='ignore'):
            tail[use] = np.minimum(window * np.log(2) / np.log(now[use] / prev[use]), max_doubling_time)
        self.results['doubling_time'].values[:] = out
        return out

    def compute_r_eff(self, method='daily', smoothing=2, window=7):
        '''
        Effective reproduction number (reference sim.py:888-987): 'daily' from daily infections, 'infectious' / 'outcome' by
        counting, through the infection log, how many people each person infected, dated by when the source became infectious
        / recovered or died, over a sliding window.
        '''
        if method in ('infectious', 'outcome'):
            P = self.people
            if self._comm is not None:
                raise NotImplementedError("r_eff methods 'infectious' / 'outcome' need every agent's dates on one rank and are not built for agent-partitioned runs")
            log = self.infection_log
            values = r_eff_windowed(method, P.to_numpy('date_infectious'), P.to_numpy('date_recovered'), P.to_numpy('date_dead'),
                                    log['source'], self.npts, window)
            self.results['r_eff'].values[:] = values
            return self.results['r_eff'].values
        if method != 'daily':
            raise ValueError(f'Method must be "daily", "infectious", or "outcome", not "{method}"')
        # mean duration of infectiousness (sim.py:916-925): mean outcome date - mean date of becoming infectious over everybody with
        # an outcome.  Three float64 sums on the device and one 24-byte read instead of three per-agent arrays on the host (the
        # reference takes float32 means: the two agree to ~1e-7 relative)
        P = self.people
        d_rec, d_dead, d_inf = P['date_recovered'], P['date_dead'], P['date_infectious']
        rec, dead = ~torch.isnan(d_rec), ~torch.isnan(d_dead)
        f64 = torch.float64
        acc = torch.stack([d_rec[rec].to(f64).sum() + d_dead[dead].to(f64).sum(), d_inf[rec].to(f64).sum() + d_inf[dead].to(f64).sum(),
                           (rec.sum() + dead.sum()).to(f64)])
        if self._comm is not None:             # over the whole population
            self._comm.all_reduce_sum(acc)
        so, si, cnt_ = (float(x) for x in acc.cpu().numpy())
        mean_inf = so / cnt_ - si / cnt_ if cnt_ else 0
        if self.rng_mode == 'mt' and self._comm is None:      # replay mode reproduces the reference to the last bit: float32 NumPy means
            h_rec, h_dead, h_inf = P.to_numpy_many(('date_recovered', 'date_dead', 'date_infectious'))
            ri, di_ = np.nonzero(~np.isnan(h_rec))[0], np.nonzero(~np.isnan(h_dead))[0]
            outcome, both = np.concatenate((h_rec[ri], h_dead[di_])), np.concatenate((ri, di_))
            mean_inf = outcome.mean() - h_inf[both].mean() if len(outcome) else 0
        R = self.results
        new_inf = R['new_infections'].values - R['n_imports'].values
        n_inf = R['n_infectious'].values
        raw = mean_inf * np.divide(new_inf, n_inf, out=np.zeros(self.npts), where=n_inf > 0)
        if len(raw) >= 3:
            dur = self.pars['dur']
            initial = int(min(len(raw), dur['exp2inf']['par1'] + dur['asym2rec']['par1']))
            for i in range(initial):
                raw[i] = raw[i:initial].mean()
            sm = raw.copy()
            for _ in range(smoothing):
                sm = np.convolve(np.concatenate([[sm[0]], sm, [sm[-1]]]), [0.25, 0.5, 0.25], mode='valid')
            sm[:smoothing] = raw[:smoothing]
            sm[-smoothing:] = raw[-smoothing:]
            raw = sm
        R['r_eff'].values[:] = raw
        return raw

    def compute_gen_time(self):
        ''' Generation time: exposure to exposure ('true') and symptom onset to symptom onset ('clinical') (reference sim.py:990-1025) '''
        if self._comm is not None:
            raise NotImplementedError('compute_gen_time needs every agent\'s dates on one rank and is not built for agent-partitioned runs')
        log, P = self.infection_log, self.people
        self.results['gen_time'] = gen_time(P.to_numpy('date_exposed'), P.to_numpy('date_symptomatic'), log['source'], log['target'])
        return self.results['gen_time']

    def make_transtree(self, **kwargs):
        ''' The transmission tree of the finished run as arrays (reference sim.py:1075-1089 / analysis.py:1772) '''
        from .analysis import TransTree
        return TransTree(self, **kwargs)

    def compute_summary(self, t=None):
        ''' reference sim.py:1040-1072 '''
        if t is None:
            t = self.t
        self.summary = {k: float(self.results[k].values[t]) for k in self.result_keys()}
        return self.summary
