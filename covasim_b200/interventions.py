'''
Interventions: the plug-in boundary of the hot path (reference covasim/interventions.py).

``Intervention`` keeps the reference's contract -- ``initialize(sim)``, ``apply(sim)`` called once per
day from inside ``Sim.step`` in list order, ``finalize(sim)`` -- and plain callables are accepted too
(reference sim.py:596-597).  Custom interventions read and write ``sim.people.<field>`` device tensors.

The built-ins that appear in the benchmark configurations run as device passes in native-RNG mode:
``test_prob`` (one per-agent kernel), ``contact_tracing`` (case bitmap + one edge pass per traced layer)
and ``vaccinate_prob`` (one per-agent kernel); ``change_beta`` only edits parameters.
'''
import ctypes as C

import numpy as np
import torch

from . import parameters as cvpar
from . import _capi

__all__ = ['Intervention', 'dynamic_pars', 'sequence', 'change_beta', 'clip_edges', 'test_num', 'test_prob', 'contact_tracing', 'vaccinate_prob', 'find_day', 'process_days']


def find_day(arr, t=None, interv=None, sim=None, which='first'):
    ''' Indices of ``arr`` equal to day t (reference interventions.py:22-53) '''
    if callable(arr):
        arr = np.atleast_1d(arr(interv, sim))
    all_inds = np.nonzero(np.asarray(arr))[0] if t is None else np.nonzero(np.asarray(arr) == t)[0]     # t None: indices of True entries
    if len(all_inds) == 0 or which == 'all':
        return all_inds
    return [all_inds[0]] if which == 'first' else [all_inds[-1]]


def process_days(sim, days):
    ''' Sorted integer day array (reference interventions.py:78-99) '''
    if callable(days):
        return days
    days = list(np.atleast_1d(days)) if not isinstance(days, str) else [days]
    out = []
    for d in days:
        if d in ['end', -1]:
            d = sim['end_day']
        out.append(sim.day(d))
    return np.sort(np.array(out))


def get_subtargets(subtarget, sim):
    ''' (indices, values) of a ``subtarget`` option: a dict with 'inds' and 'vals', either of which may be a function of the sim (reference interventions.py:153-187) '''
    if callable(subtarget):
        subtarget = subtarget(sim)
    if 'inds' not in subtarget:
        raise ValueError(f'The subtarget dict must have keys "inds" and "vals", but you supplied {subtarget}')
    inds = subtarget['inds'](sim) if callable(subtarget['inds']) else subtarget['inds']
    vals = subtarget['vals'](sim) if callable(subtarget['vals']) else subtarget['vals']
    if np.ndim(vals) and len(vals) != len(inds):
        raise ValueError(f'Length of subtargeting indices ({len(inds)}) does not match length of values ({len(vals)})')
    return inds, vals


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
    raise NotImplementedError(f'The selected distribution "{dist}" is not implemented; choices are: none, uniform, lognormal')


def swab_terms(pdf, sim):
    '''
    swab_delay (interventions.py:812-817, 935-939): today's symptomatic agents (device indices), their whole days since symptom onset, the
    delay density there and the inverse share of each delay among them.  A host computation on a few thousand values (one synchronisation).
    '''
    if sim._comm is not None:
        raise NotImplementedError('swab_delay weighs every symptomatic agent by the share of its onset day in the whole population and is not built for agent-partitioned runs')
    P = sim.people
    symp_inds = torch.nonzero(P.symptomatic.as_subclass(torch.Tensor)).flatten()
    if not len(symp_inds):
        return symp_inds, None, None, None
    symp_time = (np.float32(sim.t) - P.date_symptomatic.as_subclass(torch.Tensor)[symp_inds].cpu().numpy()).astype(np.int32)
    inv_count = np.bincount(symp_time) / len(symp_time)
    count = np.nan * np.ones(inv_count.shape)
    count[inv_count != 0] = 1 / inv_count[inv_count != 0]
    return symp_inds, symp_time, pdf.pdf(symp_time), count[symp_time]


def quar_test_mask(sim, policy):
    ''' interventions.py:691-715 get_quar_inds as a device mask '''
    P, t = sim.people, sim.t
    if policy == 'start':
        return (P.date_quarantined == t - 1).as_subclass(torch.Tensor)
    if policy == 'end':
        return (P.date_end_quarantine == t + 1).as_subclass(torch.Tensor)
    if policy == 'both':
        return ((P.date_quarantined == t - 1) | (P.date_end_quarantine == t + 1)).as_subclass(torch.Tensor)
    if policy == 'daily':
        return P.quarantined.as_subclass(torch.Tensor).clone()
    mask = torch.zeros(sim.n_local, dtype=torch.bool, device=P.device)
    if callable(policy):                                   # a function of the sim returning the people to test (global ids under a partition)
        inds = policy(sim)
        inds = torch.as_tensor(np.asarray(inds) if not isinstance(inds, torch.Tensor) else inds).to(device=P.device, dtype=torch.int64) - int(sim.id0)
        mask[inds[(inds >= 0) & (inds < sim.n_local)]] = True
        return mask
    dq = P.date_quarantined.as_subclass(torch.Tensor)
    for q in np.atleast_1d(policy):                        # days after the start of quarantine on which a test is done
        mask |= dq == float(t - 1 - q)
    return mask


def boost_is_f64(boost):
    ''' NumPy's scalar rules for ``peak_nab[inds] *= boost`` (immunity.py:170): a NumPy float64 scalar forces the product into float64, a Python number does not '''
    return isinstance(boost, np.floating) and not isinstance(boost, (np.float32, np.float16))


def subtarget_override(subtarget, sim):
    ''' The subtarget as a device array float64[n_local]: the explicit probability of every subtargeted agent, NaN elsewhere '''
    inds, vals = get_subtargets(subtarget, sim)
    dev = sim.people.device
    inds = torch.as_tensor(inds, device=dev).to(torch.int64)
    vals = torch.as_tensor(vals, dtype=torch.float64, device=dev)
    if vals.ndim == 0:
        vals = vals.expand(len(inds))
    if sim._comm is not None:                          # agent-partitioned: global ids, this rank keeps its own agents
        keep = (inds >= sim.id0) & (inds < sim.id0 + sim.n_local)
        inds, vals = inds[keep] - sim.id0, vals[keep]
    out = torch.full((sim.n_local,), float('nan'), dtype=torch.float64, device=dev)
    out[inds] = vals
    return out


class Intervention:
    ''' Base class (reference interventions.py:223-410) '''

    def __init__(self, label=None, **kwargs):
        self.label = label if label is not None else self.__class__.__name__
        self.days = []
        self.initialized = False
        self.finalized = False

    def __call__(self, *args, **kwargs):
        if not self.initialized:
            raise RuntimeError(f'Intervention (label={self.label}, {type(self)}) has not been initialized')
        return self.apply(*args, **kwargs)

    def initialize(self, sim=None):
        self.initialized = True
        self.finalized = False

    def finalize(self, sim=None):
        if self.finalized:
            raise RuntimeError('Intervention already finalized')
        self.finalized = True

    def apply(self, sim):
        raise NotImplementedError

    def shrink(self, in_place=False):
        return self

    def _device_plan(self, sim):
        '''
        What Sim's block runner (cvb_run_days) needs to know about this intervention: None = apply() must be called every day
        (the default, and what every user-written intervention gets); ('host', days) = apply() only matters on ``days``;
        ('test', pars, start, end) / ('trace', pars, start, end) = runs inside the fused day kernels.
        '''
        return None


class dynamic_pars(Intervention):
    '''
    Change simulation parameters on given days (reference interventions.py:411-479): ``pars`` maps a parameter name to
    ``dict(days=..., vals=...)``; dict values update a nested parameter (e.g. ``beta_layer``).  Parameters may also be given as
    keyword arguments.  The device copy of the scalar block is refreshed the same day (Sim.step re-sends it when it changed).
    '''

    def __init__(self, pars=None, **kwargs):
        pars = dict(pars or {})
        sim_par_keys = list(cvpar.make_pars().keys())
        for k in [k for k in kwargs if k in sim_par_keys]:
            pars[k] = kwargs.pop(k)
        super().__init__(**kwargs)
        for parkey, spec in pars.items():
            for subkey in ('days', 'vals'):
                if subkey not in spec:
                    raise KeyError(f'Parameter {parkey} is missing subkey {subkey}')
                if isinstance(spec[subkey], (int, float, np.integer, np.floating)):
                    spec[subkey] = np.atleast_1d(spec[subkey])
            if len(spec['days']) != len(spec['vals']):
                raise ValueError(f'Length of days ({len(spec["days"])}) does not match length of values ({len(spec["vals"])}) for parameter {parkey}')
        self.pars = pars

    def _device_plan(self, sim):
        if any(callable(spec['days']) for spec in self.pars.values()):
            return None
        return ('host', sorted({int(d) for spec in self.pars.values() for d in np.atleast_1d(spec['days'])}))

    def apply(self, sim):
        t = sim.t
        for parkey, spec in self.pars.items():
            for ind in find_day(spec['days'], t, interv=self, sim=sim):
                self.days.append(t)
                val = spec['vals'][ind]
                if isinstance(val, dict):
                    sim[parkey].update(val)
                else:
                    sim[parkey] = val
                sim._pars_dirty = True


class sequence(Intervention):
    '''
    Switch between interventions: ``interventions[i]`` is applied from ``days[i]`` until ``days[i+1]`` (reference
    interventions.py:482-523).  The nested interventions are initialised with the sequence and get their own index (after the
    top-level ones, see Sim.intervention_index), which keys their random draws.
    '''

    def __init__(self, days, interventions, **kwargs):
        super().__init__(**kwargs)
        if len(np.atleast_1d(days)) != len(interventions):
            raise ValueError('sequence: one start day per intervention')
        self.days = list(np.atleast_1d(days))
        self.interventions = list(interventions)

    def initialize(self, sim):
        super().initialize()
        self.days = [sim.day(d) for d in self.days]
        self.days_arr = np.array(self.days + [sim.npts])
        for iv in self.interventions:
            iv.initialize(sim)

    def active(self, sim):
        ''' The intervention in force today, or None '''
        inds = find_day(self.days_arr <= sim.t, which='last')
        return self.interventions[inds[0]] if len(inds) else None

    def apply(self, sim):
        iv = self.active(sim)
        if iv is not None:
            return iv.apply(sim)

    def finalize(self, sim=None):
        super().finalize()
        for iv in self.interventions:
            if hasattr(iv, 'finalize'):
                iv.finalize(sim)


class change_beta(Intervention):
    ''' Scale overall or per-layer beta on given days (reference interventions.py:533-586) '''

    def __init__(self, days, changes, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.days, self.changes, self.layers = days, changes, layers
        self.orig_betas = None

    def initialize(self, sim):
        super().initialize()
        self.days = process_days(sim, self.days)
        self.changes = np.atleast_1d(np.array(self.changes, dtype=float))
        if len(self.days) != len(self.changes):
            raise ValueError(f'Number of days supplied ({len(self.days)}) does not match number of changes ({len(self.changes)})')
        layers = self.layers if isinstance(self.layers, (list, tuple)) else [self.layers]
        self.orig_betas = {}
        for lk in layers:
            self.orig_betas['overall' if lk is None else lk] = sim['beta'] if lk is None else sim['beta_layer'][lk]

    def _device_plan(self, sim):
        return None if callable(self.days) else ('host', [int(d) for d in np.atleast_1d(self.days)])

    def apply(self, sim):
        for ind in find_day(self.days, sim.t, interv=self, sim=sim):
            for lk, b in self.orig_betas.items():
                if lk == 'overall':
                    sim['beta'] = b * self.changes[ind]
                else:
                    sim['beta_layer'][lk] = b * self.changes[ind]


class clip_edges(Intervention):
    '''
    Remove a fraction of a layer's contacts on given days and put them back later (reference interventions.py:589-667).
    The edges move between the simulation's device-resident layer and a layer owned by the intervention
    (``Layer.pop_inds`` / ``Layer.append``); the adjacency is rebuilt before the next transmission pass.  Which edges move is a
    host-side set choice from the Numba stream (the reference's cvu.choose; the O(k) sampler in native-RNG mode).
    '''

    def __init__(self, days, changes, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.days, self.changes, self.layers = days, changes, layers
        self.contacts = None

    def initialize(self, sim):
        super().initialize()
        if sim._comm is not None:
            raise NotImplementedError('clip_edges edits contact layers during the run, which agent-partitioned simulations do not support')
        self.days = process_days(sim, self.days)
        self.changes = np.atleast_1d(np.array(self.changes, dtype=float))
        if len(self.days) != len(self.changes):
            raise ValueError(f'Number of days supplied ({len(self.days)}) does not match number of changes ({len(self.changes)})')
        lkeys = sim.people.layer_keys()
        self.layers = lkeys if self.layers is None else ([self.layers] if isinstance(self.layers, str) else list(self.layers))
        from .base import Layer
        self.contacts = {lk: Layer(label=lk, device=sim.people.device) for lk in self.layers}

    def _device_plan(self, sim):
        return None if callable(self.days) else ('host', [int(d) for d in np.atleast_1d(self.days)])

    def apply(self, sim):
        for ind in find_day(self.days, sim.t, interv=self, sim=sim):
            for lkey in self.layers:
                s_layer, i_layer = sim.people.contacts[lkey], self.contacts[lkey]
                n_sim, n_int = len(s_layer), len(i_layer)
                n_contacts = n_sim + n_int
                if not n_contacts:
                    continue
                prop_to_move = n_sim / n_contacts - self.changes[ind]
                n_to_move = int(prop_to_move * n_contacts)
                src, dst, n_src = (s_layer, i_layer, n_sim) if n_to_move > 0 else (i_layer, s_layer, n_int)
                k = abs(n_to_move)
                if sim.rng_mode == 'mt':
                    inds = sim.rng.nb.choice(n_src, k, replace=False)                  # cvu.choose: Numba stream
                else:
                    from . import utils as cvu
                    inds = cvu.choose_distinct(sim.rng.nb, n_src, k)
                dst.append(src.pop_inds(inds))

    def finalize(self, sim=None):
        super().finalize()
        self.contacts = None


_QUAR_POLICY = dict(start=0, end=1, both=2, daily=3)
_QUAR_POLICY_HOST = 4     # a number / list of days since the start of quarantine, or a function of the sim: the set is built here, not in the kernel


def quar_policy_code(policy, who):
    ''' Kernel code of a quarantine-testing policy (reference interventions.py:682-712 get_quar_inds) '''
    if isinstance(policy, str):
        if policy not in _QUAR_POLICY:
            raise ValueError(f'{who}: quarantine policy "{policy}" not recognized: must be a string (start, end, both, daily), int, list, array, set, tuple, or function')
        return _QUAR_POLICY[policy]
    if callable(policy) or np.ndim(policy) <= 1:
        return _QUAR_POLICY_HOST
    raise ValueError(f'{who}: quarantine policy {policy!r} not recognized')



class test_num(Intervention):
    '''
    A given number of tests per day, handed out by weight: symptomatic people ``symp_test`` times as likely, people in
    quarantine ``quar_test`` times, diagnosed people never (reference interventions.py:718-854 + people.py:589-617).
    Native-RNG mode: one device pass writes the weights and the exponential-clock keys ``-log(1 - u) / w`` of every agent, the
    ``n_tests`` smallest keys are a weighted sample without replacement (the reference's choose_w), and a second pass administers
    the tests.  One device synchronisation per day (the number of agents with non-zero weight caps ``n_tests``).
    ``subtarget`` and ``ili_prev`` multiply the weights of the agents they name after the kernel pass (the keys of those agents are
    recomputed from the same uniforms); so do ``swab_delay`` for the symptomatic and a ``quar_policy`` given as days since the start of quarantine
    or as a function.  Not built: daily_tests from a data file.
    '''

    def __init__(self, daily_tests, symp_test=100.0, quar_test=1.0, quar_policy=None, subtarget=None, ili_prev=None, sensitivity=1.0,
                 loss_prob=0, test_delay=0, start_day=0, end_day=None, swab_delay=None, **kwargs):
        super().__init__(**kwargs)
        self.subtarget, self.ili_prev = subtarget, ili_prev
        self.pdf = get_pdf(**swab_delay) if swab_delay else None
        if isinstance(daily_tests, str):
            raise NotImplementedError('test_num: daily_tests from a data file is outside the built path (pass numbers)')
        self.daily_tests = daily_tests
        self.symp_test, self.quar_test = symp_test, quar_test
        self.quar_policy = quar_policy if quar_policy else 'start'
        self._quar_code = quar_policy_code(self.quar_policy, 'test_num')
        self.sensitivity, self.loss_prob, self.test_delay = sensitivity, loss_prob, test_delay
        self.start_day, self.end_day = start_day, end_day

    def initialize(self, sim):
        super().initialize()
        self.start_day = sim.day(self.start_day)
        self.end_day = sim.day(self.end_day)
        self.days = [self.start_day, self.end_day]
        dt_ = self.daily_tests
        self.daily_tests = np.array([dt_] * sim.npts) if isinstance(dt_, (int, float, np.integer, np.floating)) else np.asarray(dt_)
        if self.ili_prev is not None:                      # process_daily_data: a number applies every day
            ip = self.ili_prev
            self.ili_prev = np.array([ip] * sim.npts) if isinstance(ip, (int, float, np.integer, np.floating)) else np.asarray(ip)
        self.index = sim.intervention_index(self)
        self._c = _capi.cvb_test_num_pars(symp_test=float(self.symp_test), quar_test=float(self.quar_test),
                                          quar_policy=self._quar_code, index=self.index)
        dev = sim.people.device
        self._weight = torch.empty(sim.n_local, dtype=torch.float64, device=dev)
        self._key = torch.empty(sim.n_local, dtype=torch.float64, device=dev)

    def n_tests_today(self, sim):
        ''' The day's number of tests (randround of the scaled daily number; one NumPy-stream draw), or 0 '''
        t = sim.t
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return 0
        rel_t = t - self.start_day
        if rel_t >= len(self.daily_tests):
            return 0
        n_tests = int(np.floor(self.daily_tests[rel_t] / sim.rescale_vec[t] + sim.rng.np_.random_sample()))       # sc.randround
        if not (n_tests and np.isfinite(n_tests)):
            return 0
        sim._host_add('new_tests', t, n_tests)
        return n_tests

    def _extra_weights(self, sim):
        '''
        ili_prev (interventions.py:823-828: today's people with other illnesses, not symptomatic, test like symptomatic ones) and subtarget
        (:834-837) multiply the weights the kernel wrote; the exponential-clock keys of those agents are recomputed from the SAME uniforms
        (-log(1 - u) / w: the kernel's formula), so the draw an agent gets does not depend on the options.
        '''
        host_policy = self._quar_code == _QUAR_POLICY_HOST
        if self.ili_prev is None and self.subtarget is None and self.pdf is None and not host_policy:
            return
        t, dev, id0, n_local = sim.t, sim.people.device, int(sim.id0), sim.n_local
        touched = []
        rel_t = t - self.start_day
        if self.pdf is not None:                               # interventions.py:812-819: the symptomatic weigh symp_test x density x inverse share of their delay
            symp_inds, symp_time, dens, count = swab_terms(self.pdf, sim)
            if len(symp_inds):
                w = torch.as_tensor(1.0 * (float(self.symp_test) * (dens * count)), dtype=torch.float64, device=dev)
                qt = quar_test_mask(sim, self.quar_policy)[symp_inds]
                w = torch.where(qt, w * float(self.quar_test), w)
                w[sim.people.diagnosed.as_subclass(torch.Tensor)[symp_inds]] = 0.0
                self._weight[symp_inds] = w
                touched.append(symp_inds)
        if self.ili_prev is not None and rel_t < len(self.ili_prev):
            n_ili = int(self.ili_prev[rel_t] * sim['pop_size'])
            if sim.rng_mode == 'mt':
                ili = sim.rng.nb.choice(sim['pop_size'], n_ili, replace=False)               # cvu.choose
            else:
                from . import utils as cvu
                ili = cvu.choose_distinct(sim.rng.nb, sim['pop_size'], n_ili)
            ili = torch.as_tensor(np.asarray(ili), dtype=torch.int64, device=dev)
            ili = ili[(ili >= id0) & (ili < id0 + n_local)] - id0                               # (global ids: this rank's agents)
            ili = ili[~sim.people.symptomatic.as_subclass(torch.Tensor)[ili]]
            self._weight[ili] = self._weight[ili] * float(self.symp_test)
            touched.append(ili)
        if host_policy:                                        # (the kernel applied no quarantine factor: policy code 4)
            qt = quar_test_mask(sim, self.quar_policy)
            if self.pdf is not None and len(symp_inds):        # (the swab-delay branch above has already dealt with the symptomatic)
                qt[symp_inds] = False
            qi = torch.nonzero(qt).flatten()
            self._weight[qi] = self._weight[qi] * float(self.quar_test)
            touched.append(qi)
        if self.subtarget is not None:
            inds, vals = get_subtargets(self.subtarget, sim)
            inds = torch.as_tensor(np.asarray(inds) if not isinstance(inds, torch.Tensor) else inds).to(device=dev, dtype=torch.int64)
            vals = torch.as_tensor(np.asarray(vals) if not isinstance(vals, torch.Tensor) else vals).to(device=dev, dtype=torch.float64)
            if vals.ndim == 0:
                vals = vals.expand(len(inds))
            keep = (inds >= id0) & (inds < id0 + n_local)
            inds, vals = inds[keep] - id0, vals[keep]
            self._weight[inds] = self._weight[inds] * vals
            touched.append(inds)
        sel = torch.cat(touched)
        if len(sel):
            u = torch.empty(n_local, dtype=torch.float64, device=dev)
            sim._call('cvb_keyed_uniform', int(sim.rng.seed), _P_TEST, self.index, t, id0, n_local, 0, u.data_ptr(), sim._stream_ptr)
            w = self._weight[sel]
            self._key[sel] = torch.where(w > 0, -torch.log(1.0 - u[sel]) / w, torch.full_like(w, float('inf')))

    def rescaled(self, sim, n_tests, weight_sum):
        ''' Share of the tests that falls inside the simulated sample while the population is still being rescaled (interventions.py:838-842) '''
        t = sim.t
        if sim.rescale_vec[t] / sim['pop_scale'] < 1:
            in_tot = weight_sum * sim.rescale_vec[t]
            out_tot = sim.scaled_pop_size - sim.rescale_vec[t] * sim['pop_size']
            n_tests = int(np.floor(n_tests * in_tot / (in_tot + out_tot) + sim.rng.np_.random_sample()))              # sc.randround
        return n_tests

    def apply(self, sim):
        n_tests = self.n_tests_today(sim)
        if not n_tests:
            return
        t = sim.t
        sim._call('cvb_test_num_keys', sim._handle, t, C.byref(self._c), self._weight.data_ptr(), self._key.data_ptr(), sim._stream_ptr)
        self._extra_weights(sim)
        comm = sim._comm
        n_nonzero = int(torch.count_nonzero(self._weight).item())
        if sim.rescale_vec[t] / sim['pop_scale'] < 1:
            wsum = float(self._weight.sum().item())
            if comm is not None:
                wsum = float(sum(comm.gather_objects(wsum)))
            n_tests = self.rescaled(sim, n_tests, wsum)
        total_nonzero = n_nonzero if comm is None else int(sum(comm.gather_objects(n_nonzero)))
        n_tests = min(n_tests, total_nonzero)
        if n_tests <= 0:
            return
        if comm is None:
            inds = torch.topk(self._key, n_tests, largest=False, sorted=False).indices.to(torch.int32).contiguous()
        else:
            # agent-partitioned: the n_tests smallest keys of the WHOLE population.  Every rank offers its own n_tests smallest, the
            # n_tests-th smallest of all offers is the threshold, and each rank tests its agents at or below it (keys are continuous:
            # no ties; agents with weight 0 have key +inf and n_tests never exceeds the number of finite keys)
            k_local = min(n_tests, n_nonzero)
            mine = torch.topk(self._key, k_local, largest=False, sorted=False).values.cpu().numpy() if k_local else np.zeros(0)
            offers = np.concatenate(comm.gather_objects(mine))
            threshold = np.partition(offers, n_tests - 1)[n_tests - 1]
            inds = torch.nonzero(self._key <= float(threshold)).flatten().to(torch.int32).contiguous()
        if len(inds):
            sim._call('cvb_test_list', sim._handle, t, inds.data_ptr(), len(inds), float(self.sensitivity), float(self.loss_prob), int(self.test_delay),
                      self.index, sim._stream_ptr)
        return inds


class test_prob(Intervention):
    '''
    Probability-based testing (reference interventions.py:857-981 + people.py:589-617).  One per-agent
    device pass: test probability from symptom / quarantine / diagnosis state, keyed Bernoulli draws
    for "tests today", "test is positive" (sensitivity) and "not lost to follow-up".
    ``subtarget`` (explicit probabilities for given agents) is passed to the kernel as a per-agent override array.
    ``ili_prev``, ``subtarget`` and ``swab_delay`` and a ``quar_policy`` given as days since the start of quarantine or as a function become a per-agent
    override array built every day.
    '''

    def __init__(self, symp_prob, asymp_prob=0.0, symp_quar_prob=None, asymp_quar_prob=None, quar_policy=None, subtarget=None,
                 ili_prev=None, sensitivity=1.0, loss_prob=0.0, test_delay=0, start_day=0, end_day=None, swab_delay=None, **kwargs):
        super().__init__(**kwargs)
        self.pdf = get_pdf(**swab_delay) if swab_delay else None
        self.subtarget = subtarget
        self.ili_prev = ili_prev
        self.symp_prob, self.asymp_prob = symp_prob, asymp_prob
        self.symp_quar_prob = symp_prob if symp_quar_prob is None else symp_quar_prob
        self.asymp_quar_prob = asymp_prob if asymp_quar_prob is None else asymp_quar_prob
        self.quar_policy = quar_policy if quar_policy else 'start'
        self._quar_code = quar_policy_code(self.quar_policy, 'test_prob')
        self.sensitivity, self.loss_prob, self.test_delay = sensitivity, loss_prob, test_delay
        self.start_day, self.end_day = start_day, end_day

    def initialize(self, sim):
        super().initialize()
        self.start_day = sim.day(self.start_day)
        self.end_day = sim.day(self.end_day)
        self.days = [self.start_day, self.end_day]
        if self.ili_prev is not None:                      # process_daily_data: a number applies every day
            ip = self.ili_prev
            self.ili_prev = np.array([ip] * sim.npts) if isinstance(ip, (int, float, np.integer, np.floating)) else np.asarray(ip)
            if sim._comm is not None:
                raise NotImplementedError('test_prob(ili_prev=...) draws a population-wide set on the host and is not built for agent-partitioned runs')
        self.index = sim.intervention_index(self)
        self._c = _capi.cvb_test_prob_pars(symp_prob=self.symp_prob, asymp_prob=self.asymp_prob, symp_quar_prob=self.symp_quar_prob,
                                           asymp_quar_prob=self.asymp_quar_prob, sensitivity=self.sensitivity, loss_prob=self.loss_prob,
                                           quar_policy=self._quar_code, test_delay=int(self.test_delay), index=self.index)

    def _device_plan(self, sim):
        if self.subtarget is not None or self.ili_prev is not None or self.pdf is not None or self._quar_code == _QUAR_POLICY_HOST:
            return None                                    # per-agent overrides are built on the host every day
        return ('test', self._c, int(self.start_day), -1 if self.end_day is None else int(self.end_day))

    def apply(self, sim):
        t = sim.t
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return
        override = self.override(sim)                      # kept alive until the call returns
        sim._call('cvb_test_prob', sim._handle, t, C.byref(self._c), None if override is None else override.data_ptr(), sim._stream_ptr)


    def ili_inds(self, sim):
        ''' Today's people with influenza-like illness: a host-side set choice from the Numba stream (interventions.py:946-953), or None '''
        if self.ili_prev is None:
            return None
        rel_t = sim.t - self.start_day
        if rel_t >= len(self.ili_prev):
            return None
        n_ili = int(self.ili_prev[rel_t] * sim['pop_size'])
        if sim.rng_mode == 'mt':
            return sim.rng.nb.choice(sim['pop_size'], n_ili, replace=False)                  # cvu.choose
        from . import utils as cvu
        return cvu.choose_distinct(sim.rng.nb, sim['pop_size'], n_ili)

    def override(self, sim):
        '''
        Per-agent explicit probabilities (NaN = none) for the kernel: people with ILI symptoms who are not symptomatic test like
        symptomatic people whatever their quarantine state (interventions.py:962-967), then the subtarget on top (:971-973).
        '''
        ili = self.ili_inds(sim)
        host_policy = self._quar_code == _QUAR_POLICY_HOST
        if ili is None and self.subtarget is None and self.pdf is None and not host_policy:
            return None
        dev = sim.people.device
        out = torch.full((sim.n_local,), float('nan'), dtype=torch.float64, device=dev)
        if self.pdf is not None:                           # interventions.py:934-943: the symptomatic test by the time since onset ...
            symp_inds, symp_time, dens, count = swab_terms(self.pdf, sim)
            if len(symp_inds):
                sp = np.ones(len(symp_time))
                early = 1 > (symp_time * self.symp_prob)
                sp[early] = self.symp_prob / (1 - symp_time[early] * self.symp_prob)
                sp = torch.as_tensor(dens * sp * count, dtype=torch.float64, device=dev)
                free = ~quar_test_mask(sim, self.quar_policy)[symp_inds]      # ... unless the quarantine probability applies to them (:959-966)
                out[symp_inds[free]] = sp[free]
        if host_policy:                                    # quarantine-testing probabilities for a set the kernel does not know (policy code 4)
            qt = quar_test_mask(sim, self.quar_policy)
            symp = sim.people.symptomatic.as_subclass(torch.Tensor)
            out[qt & symp] = float(self.symp_quar_prob)
            out[qt & ~symp] = float(self.asymp_quar_prob)
        if ili is not None and len(ili):
            ili = torch.as_tensor(ili, dtype=torch.int64, device=dev)
            ili = ili[~sim.people.symptomatic[ili]]
            out[ili] = float(self.symp_prob)
        if self.subtarget is not None:
            sub = subtarget_override(self.subtarget, sim)
            out = torch.where(torch.isnan(sub), out, sub)
        return out


class contact_tracing(Intervention):
    '''
    Trace the contacts of today's newly diagnosed agents and queue their quarantine (reference
    interventions.py:984-1145).  Device form: bitmap of today's cases, then one streaming pass over each
    traced layer's (p1, p2) that notifies the partner of every case with a Bernoulli(trace_prob) keyed
    per (layer, contact) -- the reference's binomial_filter over the unique contact set.
    With a ``capacity`` the cases are counted on the host every day (one synchronisation) and, when there are too many, the
    traced ones are picked there.
    '''

    def __init__(self, trace_probs=None, trace_time=None, start_day=0, end_day=None, presumptive=False, quar_period=None, capacity=None, **kwargs):
        super().__init__(**kwargs)
        self.capacity = capacity
        self.trace_probs, self.trace_time = trace_probs, trace_time
        self.start_day, self.end_day, self.presumptive, self.quar_period = start_day, end_day, presumptive, quar_period

    def initialize(self, sim):
        super().initialize()
        self.start_day = sim.day(self.start_day)
        self.end_day = sim.day(self.end_day)
        self.days = [self.start_day, self.end_day]
        lkeys = sim.people.layer_keys()
        tp = 1.0 if self.trace_probs is None else self.trace_probs
        tt = 0.0 if self.trace_time is None else self.trace_time
        self.trace_probs = dict(tp) if isinstance(tp, dict) else {k: tp for k in lkeys}
        self.trace_time = dict(tt) if isinstance(tt, dict) else {k: tt for k in lkeys}
        if self.quar_period is None:
            self.quar_period = sim['quar_period']
        self.index = sim.intervention_index(self)
        c = _capi.cvb_trace_pars(presumptive=int(bool(self.presumptive)), quar_period=int(self.quar_period), index=self.index)
        for i, lk in enumerate(lkeys):
            c.trace_prob[i] = float(self.trace_probs.get(lk, 0.0))
            c.trace_time[i] = int(self.trace_time.get(lk, 0))
        self._c = c
        sim._set_quar_horizon(int(max(self.trace_time.values(), default=0)) + 1)

    def _device_plan(self, sim):
        if self.capacity is not None or self.presumptive:
            return None
        return ('trace', self._c, int(self.start_day), -1 if self.end_day is None else int(self.end_day))

    def apply(self, sim):
        t = sim.t
        if t < self.start_day or (self.end_day is not None and t > self.end_day):
            return
        if self.capacity is not None:      # at most `capacity` of today's cases are traced, picked at random (interventions.py:1079-1083)
            if sim._comm is not None:
                raise NotImplementedError('contact_tracing(capacity=...) picks among all of today\'s cases and is not built for agent-partitioned runs')
            P = sim.people
            if not self.presumptive:
                cases = torch.nonzero(P.date_diagnosed == t).flatten()
            else:
                just = torch.nonzero(P.date_tested == t).flatten()
                cases = just[P.exposed[just]]
            cap = int(self.capacity / sim.rescale_vec[t])
            if len(cases) > cap:
                if sim.rng_mode == 'mt':
                    chosen = sim.rng.np_.choice(cases.cpu().numpy(), cap, replace=False)
                    cases = torch.as_tensor(chosen, device=cases.device)
                else:
                    from . import utils as cvu
                    pick = cvu.choose_distinct(sim.rng.np_, len(cases), cap)
                    cases = cases[torch.as_tensor(pick, dtype=torch.int64, device=cases.device)]
                if sim._adj_dirty:
                    sim._build_adjacency()
                cases = cases.to(torch.int32).contiguous()
                sim._call('cvb_contact_tracing_list', sim._handle, t, C.byref(self._c), cases.data_ptr(), len(cases), sim._stream_ptr)
                return
        if sim._comm is not None:          # agent-partitioned: local cases -> all-gathered bitmap -> local contacts of every case
            sim._call('cvb_trace_select_cases', sim._handle, t, C.byref(self._c), sim._stream_ptr)
            sim._exchange_cases()
            sim._call('cvb_trace_notify_contacts', sim._handle, t, C.byref(self._c), sim._stream_ptr)
            return
        if sim._adj_dirty:
            sim._build_adjacency()
        sim._call('cvb_contact_tracing', sim._handle, t, C.byref(self._c), sim._stream_ptr)


class vaccinate_prob(Intervention):
    '''
    Probability-based vaccination with scheduled second doses (reference interventions.py:1257-1662).
    One per-agent device pass on first-dose days and on days when second doses are due.
    ``subtarget`` is passed to the kernel as a per-agent override array; ``target_eff`` in a vaccine dict is turned into the
    initial NAb level and the boost at initialisation, like the reference.
    '''

    def __init__(self, vaccine, days, label=None, prob=None, subtarget=None, booster=False, **kwargs):
        super().__init__(**kwargs)
        self.subtarget = subtarget
        self.vaccine, self.days, self.label = vaccine, days, label
        self.prob = 1.0 if prob is None else prob
        self.booster = booster
        self.index = None
        self.p = None

    def initialize(self, sim):
        super().initialize()
        if isinstance(self.vaccine, str):
            choices, mapping = cvpar.get_vaccine_choices()
            key = self.vaccine.lower()
            for txt in ['.', ' ', '&', '-', 'vaccine']:
                key = key.replace(txt, '')
            if key not in mapping:
                raise NotImplementedError(f'The selected vaccine "{self.vaccine}" is not implemented; choices are: {choices}')
            key = mapping[key]
            self.p = dict(cvpar.get_vaccine_variant_pars(vaccine=key))
            self.p.update(cvpar.get_vaccine_dose_pars(vaccine=key))
            if self.label is None:
                self.label = key
        elif isinstance(self.vaccine, dict):
            self.p = dict(self.vaccine)
            if self.label is None:
                self.label = self.p.pop('label', 'custom')
        else:
            raise ValueError(f'Could not understand vaccine of type {type(self.vaccine)}')
        for k, v in cvpar.get_vaccine_dose_pars(default=True).items():
            self.p.setdefault(k, v)
        dflt = cvpar.get_vaccine_variant_pars(default=True)
        for k in sim['variant_pars'].keys():
            self.p.setdefault(k, dflt.get(k, 1.0))
        if 'target_eff' in self.p:                              # interventions.py:1383-1397: the NAb level that gives the wanted efficacy against symptomatic disease
            if self.p['doses'] != len(self.p['target_eff']):
                raise ValueError('Provided mismatching efficacies and doses.')
            nabs = np.arange(-8, 4, 0.1)
            ne = sim['nab_eff']
            lo_inf = np.exp(ne['alpha_inf']) * (2 ** nabs) ** ne['beta_inf']                   # immunity.py:250-262 calc_VE_symp
            lo_symp = np.exp(ne['alpha_symp_inf']) * (2 ** nabs) ** ne['beta_symp_inf']
            ve_symp = 1 - ((1 - lo_inf / (1 + lo_inf)) * (1 - lo_symp / (1 + lo_symp)))
            peak = nabs[np.argmax(ve_symp > self.p['target_eff'][0])]
            self.p['nab_init'] = dict(dist='normal', par1=peak, par2=2)
            if self.p['doses'] == 2:
                boosted = nabs[np.argmax(ve_symp > self.p['target_eff'][1])]
                self.p['nab_boost'] = (2 ** boosted) / (2 ** peak)
        doses, interval = self.p['doses'], self.p['interval']
        if doses == 1 and interval is not None:
            raise ValueError("Can't use dosing intervals for vaccines with only one dose.")
        if doses == 2 and interval is None:
            raise ValueError('Must specify a dosing interval if using a vaccine with more than one dose.')
        if doses > 2:
            raise NotImplementedError('Scheduling three or more doses not yet supported; use a booster vaccine instead')
        sim['vaccine_pars'][self.label] = self.p
        self.index = list(sim['vaccine_pars'].keys()).index(self.label)
        sim['vaccine_map'][self.index] = self.label
        self.days = self._process_own_days(sim)
        self.iindex = sim.intervention_index(self)
        dev = sim.people.device
        self.doses = torch.zeros(sim.n_local, dtype=torch.int32, device=dev)              # doses given by *this* intervention
        self.due_day = torch.full((sim.n_local,), -1, dtype=torch.int32, device=dev)     # device form of second_dose_days
        self._second = {}                          # replay mode: day -> indices due for their second dose
        self._c = _capi.cvb_vaccinate_pars(prob=float(self.prob), nab_init=_capi.dist_struct(self.p['nab_init']), nab_boost=float(self.p['nab_boost']), nab_boost_f64=float(self.p['nab_boost']), nab_boost_is_f64=int(boost_is_f64(self.p['nab_boost'])),
                                           booster=int(bool(self.booster)), vaccine_index=self.index, max_doses=int(doses), index=self.iindex,
                                           interval=-1 if interval is None else int(interval), n_days=int(sim['n_days']))
        sim._pars_dirty = True

    def _process_own_days(self, sim):
        return process_days(sim, self.days)

    def _device_plan(self, sim):
        if callable(self.days):
            return None
        days = {int(d) for d in np.atleast_1d(self.days)}
        due = {d + int(self.p['interval']) for d in days if d + int(self.p['interval']) < sim['n_days']} if self.p['interval'] is not None else set()
        if self.subtarget is not None:                     # per-agent probabilities are built on the host on first-dose days
            return ('host', sorted(days | due))
        flags = np.zeros(sim.npts, dtype=np.uint8)         # bit 0: first doses offered, bit 1: second doses fall due (interventions.py:1631-1660)
        for d in days:
            if 0 <= d < sim.npts:
                flags[d] |= 1
        for d in due:
            if 0 <= d < sim.npts:
                flags[d] |= 2
        return ('vacc', self._c, flags, self.doses, self.due_day)

    def apply(self, sim):
        t = sim.t
        if t < np.min(self.days):
            return
        first = bool(np.any(self.days == t))
        second = self.p['interval'] is not None and t < sim['n_days'] and bool(np.any(self.days + int(self.p['interval']) == t))   # interventions.py:1655-1660
        if not (first or second):
            return
        self._c.first_dose_today = int(first)
        self._c.second_dose_today = int(second)
        override = subtarget_override(self.subtarget, sim) if (first and self.subtarget is not None) else None
        sim._call('cvb_vaccinate_prob', sim._handle, t, C.byref(self._c), self.doses.data_ptr(), self.due_day.data_ptr(),
                  None if override is None else override.data_ptr(), sim._stream_ptr)


_P_VACC = 7          # enum purpose (csrc/cvb_device.cuh): the vaccination draws
_P_TEST = 3          # ... the testing draws


class vaccinate_num(vaccinate_prob):
    '''
    A number of doses per day, handed out along a priority sequence with second doses first (reference
    interventions.py:1665-1791).  The day's selection is set algebra over device arrays (who is due, who is eligible, the
    first ``num_agents`` of the sequence) and ends in the same dose kernel as vaccinate_prob (cvb_vaccinate_prob with the
    chosen agents as explicit probabilities 1).  Native-RNG mode draws keyed uniforms: slot 0 of P_VACC against the
    subtarget weight, slot 1 to pick who keeps a scheduled second dose when doses run short.  Scheduled second doses are one
    day per agent (``due_day``) instead of the reference's day-keyed sets; people the reference would drop from a set
    (dead, fully dosed) are dropped when their day comes.  Replay mode (rng='mt') is not built: the reference's draw order
    there depends on the iteration order of Python sets.
    '''

    def __init__(self, vaccine, num_doses, booster=False, subtarget=None, sequence=None, **kwargs):
        super().__init__(vaccine, days=0, subtarget=subtarget, booster=booster, prob=1.0, **kwargs)
        self.num_doses, self.sequence = num_doses, sequence

    def _process_own_days(self, sim):
        return np.array([0])

    def initialize(self, sim):
        super().initialize(sim)
        if isinstance(self.num_doses, dict):
            self.num_doses = {sim.day(k): v for k, v in self.num_doses.items()}
        seq = self.sequence                                      # interventions.py:1539-1552 process_sequence
        dev, comm, n_global = sim.people.device, sim._comm, int(sim['pop_size'])
        if callable(seq):
            if comm is not None:
                raise NotImplementedError('vaccinate_num: a callable sequence sees one rank\'s people only; pass "age", None or an array of global ids to an agent-partitioned run')
            seq = seq(sim.people)
        elif isinstance(seq, str) and seq == 'age':              # oldest first; equal ages in index order (native RNG) / as NumPy's default sort leaves them (replay)
            age = np.asarray(sim.people.age)
            if comm is not None:                                 # every rank needs the order of the WHOLE population (once, at initialisation)
                age = np.concatenate(comm.gather_objects(age))
            seq = np.argsort(-age, kind='stable') if sim.rng_mode == 'philox' else np.argsort(-age)
        elif seq is None:
            seq = sim.rng.np_.permutation(n_global)
        elif isinstance(seq, str):
            raise TypeError(f'Unable to interpret sequence {seq!r}: must be None, "age", callable, or an array')
        self._sequence_host = np.asarray(seq.cpu() if isinstance(seq, torch.Tensor) else seq).astype(np.int64)
        self._scheduled = {}                                     # replay mode: day -> set of agents due (the reference's ddict(set))
        self._u = torch.empty(sim.n_local, dtype=torch.float64, device=dev)
        self._prob = torch.empty(sim.n_local, dtype=torch.float64, device=dev)
        if comm is None:
            self.sequence = torch.as_tensor(self._sequence_host).to(device=dev)
        else:                                                    # this rank's agents with their position in the sequence (none: never chosen)
            pos = np.full(n_global, np.iinfo(np.int64).max, dtype=np.int64)
            pos[self._sequence_host[::-1]] = np.arange(len(self._sequence_host) - 1, -1, -1)      # (an id listed twice keeps its first position)
            self._pos = torch.as_tensor(pos[sim.id0:sim.id0 + sim.n_local]).to(device=dev)
            self.sequence = None

    def _device_plan(self, sim):
        if isinstance(self.num_doses, dict) and self.num_doses:
            return ('host', range(min(self.num_doses), sim.npts))           # (second doses may be deferred from day to day)
        return None

    def n_today(self, sim):
        nd = self.num_doses                                      # interventions.py:1526-1536 process_doses
        if callable(nd):
            return nd(sim)
        if isinstance(nd, dict):
            return nd.get(sim.t, 0)
        return nd

    def _uniforms(self, sim, slot):
        sim._call('cvb_keyed_uniform', int(sim.rng.seed), _P_VACC, self.iindex, sim.t, int(sim.id0), sim.n_local, slot, self._u.data_ptr(), sim._stream_ptr)
        return self._u

    def _weights(self, sim):
        ''' First-dose weights of this rank's agents (interventions.py:1738-1752): 0 for the dead and the (un)vaccinated, subtarget values multiply '''
        P = sim.people
        prob = self._prob
        prob.fill_(1.0)
        prob[P.dead.as_subclass(torch.Tensor)] = 0.0
        if self.subtarget is not None:
            inds, vals = get_subtargets(self.subtarget, sim)
            inds = torch.as_tensor(np.asarray(inds) if not isinstance(inds, torch.Tensor) else inds).to(device=P.device, dtype=torch.int64)
            vals = torch.as_tensor(np.asarray(vals) if not isinstance(vals, torch.Tensor) else vals).to(device=P.device, dtype=torch.float64)
            if sim._comm is not None:                            # global ids: this rank keeps its own agents
                if vals.ndim == 0:
                    vals = vals.expand(len(inds))
                keep = (inds >= sim.id0) & (inds < sim.id0 + sim.n_local)
                inds, vals = inds[keep] - sim.id0, vals[keep]
            prob[inds] = prob[inds] * vals
        vaccinated = P.vaccinated.as_subclass(torch.Tensor)
        if self.booster:
            prob[~vaccinated] = 0.0
        else:
            prob[vaccinated] = 0.0
        return prob

    def _select_people_partitioned(self, sim):
        '''
        The same selection when the agents are spread over several ranks: every "the first k of ..." becomes "the k smallest sequence
        positions (or uniforms) over all ranks" (Sim._k_smallest_mask: each rank offers its own k smallest, the k-th smallest of all
        offers is the threshold).  Same sets as select_people, so partitioned runs stay bit-identical to the single-GPU run.
        '''
        t, P = sim.t, sim.people
        none = torch.zeros(0, dtype=torch.int64, device=P.device)
        num_people = self.n_today(sim)
        if num_people == 0:
            self.due_day[self.due_day == t] = t + 1
            return none, none
        num_agents = int(np.floor(num_people / sim['pop_scale'] + sim.rng.np_.random_sample()))        # sc.randround (host stream: same on every rank)
        sched_mask = (self.due_day == t) & (self.doses < int(self.p['doses'])) & ~P.dead.as_subclass(torch.Tensor)
        n_sched = sum(sim._global_counts(int(sched_mask.sum().item())))
        if n_sched > num_agents:
            keep = sim._k_smallest_mask(self._uniforms(sim, 1), sched_mask, num_agents)
            self.due_day[sched_mask & ~keep] = t + 1
            return torch.nonzero(keep).flatten(), none
        scheduled = torch.nonzero(sched_mask).flatten()
        mask = (self._uniforms(sim, 0) < self._weights(sim)) & (self._pos < np.iinfo(np.int64).max)
        if sum(sim._global_counts(int(mask.sum().item()))) == 0:
            return scheduled, none
        eligible = sim._k_smallest_mask(self._pos, mask, num_agents) & ~sched_mask
        n_elig = sum(sim._global_counts(int(eligible.sum().item())))
        first = sim._k_smallest_mask(self._pos, eligible, max(num_agents - n_sched, 0)) if n_elig + n_sched > num_agents else eligible
        return scheduled, torch.nonzero(first).flatten()

    def select_people(self, sim):
        ''' Today's recipients as (scheduled second doses, first doses): device index arrays '''
        if sim._comm is not None:
            return self._select_people_partitioned(sim)
        t, P = sim.t, sim.people
        none = torch.zeros(0, dtype=torch.int64, device=P.device)
        num_people = self.n_today(sim)
        if num_people == 0:
            self.due_day[self.due_day == t] = t + 1              # defer everyone due today
            return none, none
        num_agents = int(np.floor(num_people / sim['pop_scale'] + sim.rng.np_.random_sample()))        # sc.randround
        sched_mask = (self.due_day == t) & (self.doses < int(self.p['doses'])) & ~P.dead.as_subclass(torch.Tensor)
        scheduled = torch.nonzero(sched_mask).flatten()
        if len(scheduled) > num_agents:                          # more second doses due than doses: the rest wait a day
            order = torch.argsort(self._uniforms(sim, 1)[scheduled], stable=True)
            self.due_day[scheduled[order[num_agents:]]] = t + 1
            return scheduled[order[:num_agents]], none
        mask = self._uniforms(sim, 0) < self._weights(sim)
        eligible = self.sequence[mask[self.sequence]]
        if len(eligible) == 0:
            return scheduled, none
        eligible = eligible[:num_agents]
        eligible = eligible[~sched_mask[eligible]]
        first = eligible[:max(num_agents - len(scheduled), 0)] if len(eligible) + len(scheduled) > num_agents else eligible
        return scheduled, first

    def apply(self, sim):
        t = sim.t
        scheduled, first = self.select_people(sim)
        if len(scheduled) + len(first) == 0:
            return
        # the dose kernel takes the recipients as explicit probabilities: 1 for today's first doses (it schedules their second
        # dose, interventions.py:1786-1787), 0 for everyone else; who is due today (due_day == t) is dosed by the same pass
        prob = self._prob
        prob.fill_(0.0)
        prob[first] = 1.0
        self._c.first_dose_today = 1
        self._c.second_dose_today = 1
        sim._call('cvb_vaccinate_prob', sim._handle, t, C.byref(self._c), self.doses.data_ptr(), self.due_day.data_ptr(), prob.data_ptr(), sim._stream_ptr)
        return torch.cat((scheduled, first))


def vaccinate(*args, **kwargs):
    ''' vaccinate_num if ``num_doses`` is given, else vaccinate_prob (reference interventions.py:1555-1568) '''
    return vaccinate_num(*args, **kwargs) if 'num_doses' in kwargs else vaccinate_prob(*args, **kwargs)
