'''
People: the structure-of-arrays agent state, resident on the GPU (reference covasim/people.py,
base.py:877-1465).

Every field of ``defaults.all_states`` is one device tensor (bool -> torch.bool, one byte per agent, the
same layout as the reference's NumPy bool arrays; float32 dates with NaN = "not set"; int32 counters)
bound once to the simulation handle, so kernels and Python interventions see the same memory.
``people.<field>`` / ``people[<field>]`` return the tensor; assignment copies in place so the bound
pointer never changes.
'''
import ctypes as C

import numpy as np
import torch

from . import defaults as cvd
from . import utils as cvu
from . import _capi
from .base import Contacts, Layer
from .devarray import DeviceArray

__all__ = ['People']

_TORCH_DTYPE = {np.dtype(np.bool_): torch.bool, np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32}


class People:

    def __init__(self, pars, device, uid=None, age=None, sex=None, contacts=None, local_range=None):
        object.__setattr__(self, '_arrays', {})
        self.pars = pars
        self.device = torch.device(device)
        self.t = 0
        # agent-partitioned runs (partition.py): this object holds the agents [id0, id0 + n) of pop_size
        lo, hi = (0, int(pars['pop_size'])) if local_range is None else (int(local_range[0]), int(local_range[1]))
        n = hi - lo
        nv = int(pars['n_variants'])
        self.n, self.nv, self.id0, self.n_global = n, nv, lo, int(pars['pop_size'])
        self.rel_trans_global = None
        self._sim = None
        self.infection_log_cap = 0
        A = self._arrays
        # One device arena for every field (each field a 256-byte-aligned view): kernels bind the views' pointers as before, and a
        # checkpoint / restore of the whole People state is ONE copy (Sim.snapshot / Sim.restore)
        layout, total = [], 0
        for name in cvd.all_states:
            dt = _TORCH_DTYPE[np.dtype(cvd.field_dtype(name))]
            shape = (nv, n) if cvd.field_is_2d(name) else (n,)
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            layout.append((name, dt, shape, total, nbytes))
            total += (nbytes + 255) // 256 * 256
        self._arena = torch.empty(max(total, 256), dtype=torch.uint8, device=self.device)
        self._layout = layout
        for name, dt, shape, off, nbytes in layout:
            A[name] = self._arena[off:off + nbytes].view(dt).view(shape).as_subclass(DeviceArray)      # NumPy idioms work on it (devarray.py)
            if name == 'uid':
                A[name].copy_(torch.arange(lo, hi, dtype=torch.int32, device=self.device))
            elif name in cvd.states:
                A[name].fill_(name in ('susceptible', 'naive'))
            elif dt == torch.float32 and name not in cvd.imm_states and name not in ('peak_nab', 'nab'):
                A[name].fill_(float('nan'))
            else:
                A[name].zero_()
        self._age_global = None
        if age is not None:
            if not isinstance(age, torch.Tensor):
                age = np.asarray(age)
            if local_range is not None and len(age) == self.n_global:
                self._age_global = age                  # prognoses are drawn for the whole population, then sliced
                age = age[lo:hi]
            self['age'] = age
        if sex is not None:
            if not isinstance(sex, torch.Tensor):
                sex = np.asarray(sex)
            self['sex'] = sex[lo:hi] if (local_range is not None and len(sex) == self.n_global) else sex
        self.contacts = Contacts()
        if contacts is not None:
            for lk, layer in contacts.items():
                if not isinstance(layer, Layer):
                    layer = Layer(layer['p1'], layer['p2'], layer.get('beta'), label=lk, device=self.device)
                else:
                    layer.to(self.device)
                self.contacts[lk] = layer

    def __deepcopy__(self, memo):
        ''' A copy with its own device arena (the fields stay views of ONE buffer) and its own contact layers '''
        import copy
        new = object.__new__(People)
        memo[id(self)] = new
        object.__setattr__(new, '_arrays', {})
        for k, v in self.__dict__.items():
            if k in ('_arrays', '_arena', '_stage'):
                continue
            object.__setattr__(new, k, copy.deepcopy(v, memo))
        object.__setattr__(new, '_arena', self._arena.clone())
        for name, dt, shape, off, nbytes in self._layout:
            new._arrays[name] = new._arena[off:off + nbytes].view(dt).view(shape).as_subclass(DeviceArray)
        return new

    # ---- array access -------------------------------------------------------------------------
    def __getattr__(self, name):
        arrays = object.__getattribute__(self, '_arrays')
        if name in arrays:
            return arrays[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._arrays:
            self[name] = value
        else:
            object.__setattr__(self, name, value)

    def __getitem__(self, key):
        return self._arrays[key]

    def __setitem__(self, key, value):
        ''' In-place assignment: keeps the device pointer bound to the kernels valid '''
        dst = self._arrays[key]
        src = value if isinstance(value, torch.Tensor) else torch.as_tensor(np.asarray(value))
        dst.copy_(src.to(device=dst.device, dtype=dst.dtype) if src.shape == dst.shape else src.to(device=dst.device, dtype=dst.dtype).expand_as(dst))

    def __len__(self):
        return self.n

    def keys(self):
        return list(self._arrays.keys())

    def layer_keys(self):
        return list(self.contacts.keys())

    def to_numpy(self, key):
        return self._arrays[key].cpu().numpy()

    def to_numpy_many(self, keys):
        ''' Several fields to the host through ONE pinned staging buffer (one synchronisation instead of one pageable copy per field) '''
        arrs = [self._arrays[k] for k in keys]
        if not arrs or not arrs[0].is_cuda:
            return [a.cpu().numpy() for a in arrs]
        sizes = [a.numel() * a.element_size() for a in arrs]
        offs = np.concatenate([[0], np.cumsum([(sz + 255) // 256 * 256 for sz in sizes])]).astype(np.int64)
        need = int(offs[-1])
        stage = getattr(self, '_stage', None)
        if stage is None or stage.numel() < need:
            stage = torch.empty(need, dtype=torch.uint8, pin_memory=True)
            object.__setattr__(self, '_stage', stage)
        views = []
        for a, off, sz in zip(arrs, offs[:-1], sizes):
            v = stage[int(off):int(off) + sz].view(a.dtype).view(a.shape)
            v.copy_(a, non_blocking=True)
            views.append(v)
        torch.cuda.current_stream(self.device).synchronize()
        return [v.numpy().copy() for v in views]

    # ---- counting helpers (reference base.py:1091-1145) -----------------------------------------
    def true(self, key):
        return cvu.true(self[key])

    def false(self, key):
        return cvu.false(self[key])

    def defined(self, key):
        return cvu.defined(self[key])

    def undefined(self, key):
        return cvu.undefined(self[key])

    def count(self, key):
        return int(torch.count_nonzero(self[key]).item())

    def count_by_variant(self, key, variant):
        return int(torch.count_nonzero(self[key][variant, :]).item())

    def count_not(self, key):
        return self.n - self.count(key)

    # ---- binding --------------------------------------------------------------------------------
    def _bind(self, sim):
        self._sim = sim
        for name, fid in cvd.FIELD_IDS.items():
            _capi.call('cvb_bind_field', sim._handle, fid, self._arrays[name].data_ptr())
        for i, layer in enumerate(self.contacts.values()):
            layer._bind(sim, i)

    # ---- prognoses (reference people.py:139-161) ------------------------------------------------
    def set_prognoses(self, rng):
        pars = self.pars
        rng.set_seed(pars['rand_seed'])
        progs = pars['prognoses']
        part = self._age_global is not None
        if part:                                           # the People array is float32
            age = self._age_global.to(torch.float32).cpu().numpy() if isinstance(self._age_global, torch.Tensor) else np.asarray(self._age_global, dtype=np.float32)
        else:
            age = self.to_numpy('age')
        inds = np.digitize(age, progs['age_cutoffs']) - 1
        loc = slice(self.id0, self.id0 + self.n) if part else slice(None)
        self['symp_prob'] = progs['symp_probs'][inds][loc]
        self['severe_prob'] = (progs['severe_probs'][inds] * progs['comorbidities'][inds])[loc]
        self['crit_prob'] = progs['crit_probs'][inds][loc]
        self['death_prob'] = progs['death_probs'][inds][loc]
        self['rel_sus'] = progs['sus_ORs'][inds][loc]
        bd = pars['beta_dist']
        if bd['dist'] != 'neg_binomial':
            raise NotImplementedError('beta_dist must be neg_binomial')
        step = bd.get('step', 1)
        p = bd['par2'] / (bd['par1'] / step + bd['par2'])                                    # reference utils.py:409-426
        draws = rng.np_.negative_binomial(n=bd['par2'], p=p, size=len(inds)) * step
        rel_trans = progs['trans_ORs'][inds] * draws
        self['rel_trans'] = rel_trans[loc]
        if part:            # every rank keeps the whole population's initial transmissibility (4 B per agent)
            self.rel_trans_global = torch.as_tensor(np.asarray(rel_trans, dtype=np.float32), device=self.device)
            self._age_global = None

    # ---- events ---------------------------------------------------------------------------------
    def infect(self, inds, hosp_max=None, icu_max=None, source=None, layer=None, variant=0, count_flows=True):
        '''
        Infect agents and sample their disease course on the device (reference people.py:435-586).
        ``hosp_max`` / ``icu_max``: None or False = beds available (the reference's default, used for seed
        infections and variant importations), True = no beds, 'auto' = decided on the device from today's
        severe / critical counts against n_beds_* (what Sim.step passes; reference sim.py:579-580).
        Duplicates and non-susceptible agents are dropped.  ``source`` is not needed: importations and seed
        infections have none, and transmissions are infected by the fused edge pass.
        '''
        sim = self._sim
        if sim.rng_mode == 'mt':                       # replay mode: the reference's stream order, host draws
            from . import replay
            return replay.infect(sim, inds, hosp_max=bool(hosp_max) if not isinstance(hosp_max, str) else False,
                                 icu_max=bool(icu_max) if not isinstance(icu_max, str) else False, source=source, layer=layer,
                                 variant=variant, count_flows=count_flows)
        inds = torch.as_tensor(inds, dtype=torch.int32, device=self.device).contiguous()
        if sim._comm is not None:                      # partitioned: `inds` are global ids; this rank infects the ones it owns
            inds = inds[(inds >= self.id0) & (inds < self.id0 + self.n)] - self.id0
            inds = inds.contiguous()
        if len(inds) == 0:
            return inds
        code = {'seed_infection': _capi.LAYER_SEED, 'importation': _capi.LAYER_IMPORT}.get(layer, _capi.LAYER_IMPORT)
        beds = lambda x: -1 if isinstance(x, str) and x == 'auto' else int(bool(x))
        _capi.call('cvb_infect_list', sim._handle, inds.data_ptr(), len(inds), int(variant), code, int(sim.t), int(bool(count_flows)),
                   beds(hosp_max), beds(icu_max), sim._stream_ptr)
        return inds

    def schedule_quarantine(self, inds, start_date=None, period=None):
        ''' Queue quarantine requests on the device ring (reference people.py:620-640) '''
        sim = self._sim
        start_date = sim.t if start_date is None else int(start_date)
        period = self.pars['quar_period'] if period is None else int(period)
        if start_date - sim.t + 1 > sim._quar_horizon:
            sim._set_quar_horizon(start_date - sim.t + 1)
        inds = torch.as_tensor(inds, dtype=torch.int32, device=self.device).contiguous()
        if len(inds):
            _capi.call('cvb_schedule_quarantine', sim._handle, inds.data_ptr(), len(inds), start_date, float(start_date + period), sim._stream_ptr)

    def story(self, uid, *args, quiet=False):
        '''
        A short history of the events in the life of the given person / people (reference people.py:666-778): the lines the
        reference prints, returned as a list of strings (and printed unless ``quiet``).
        '''
        labels = dict(a='default contact', h='household', s='school', w='workplace', c='community')
        dates = dict(date_critical='became critically ill and needed ICU care', date_dead='died', date_diagnosed='was diagnosed with COVID',
                     date_end_quarantine='ended quarantine', date_infectious='became infectious',
                     date_known_contact='was notified they may have been exposed to COVID', date_pos_test='recieved their positive test result',
                     date_quarantined='entered quarantine', date_recovered='recovered', date_severe='developed severe symptoms and needed hospitalization',
                     date_symptomatic='became symptomatic', date_tested='was tested for COVID', date_vaccinated='was vaccinated against COVID')
        log = self._sim.infection_log if self._sim is not None else None
        lkeys = self.layer_keys()
        lines = []
        for person in [uid] + list(args):
            i = int(person) - self.id0
            get = lambda k: self._arrays[k][i].item()
            sex = 'female' if get('sex') == 0 else 'male'
            intro = f'This is the story of {person}, a {get("age"):.0f} year old {sex}'
            if not get('susceptible'):
                lines.append(f'{intro}, who had {"asymptomatic" if np.isnan(get("date_symptomatic")) else "symptomatic"} COVID.')
            else:
                lines.append(f'{intro}, who did not contract COVID.')
            total, none = 0, []
            for lk, layer in self.contacts.items():
                k = int(((layer['p1'] == person) | (layer['p2'] == person)).sum().item()) if len(layer) else 0
                total += k
                llabel = labels.get(lk.lower(), f'"{lk}"')
                if k:
                    lines.append(f'{person} is connected to {k} people in the {llabel} layer')
                else:
                    none.append(llabel)
            if none:
                lines.append(f'{person} has no contacts in the {", ".join(none)} layer(s)')
            lines.append(f'{person} has {total} contacts in total')
            events = [(get(k), msg) for k, msg in dates.items() if not np.isnan(get(k))]
            if log is not None:
                layer_name = lambda code: labels.get(lkeys[code].lower(), f'"{lkeys[code]}"') if code >= 0 else None
                for k in np.nonzero(log['target'] == person)[0]:
                    code = int(log['layer'][k])
                    events.append((float(log['date'][k]), f'was infected with COVID by {int(log["source"][k])} via the {layer_name(code)} layer' if code >= 0 else
                                   ('was infected with COVID as a seed infection' if code == -1 else 'was infected with COVID by an importation')))
                for k in np.nonzero(log['source'] == person)[0]:
                    x = int((log['source'] == log['target'][k]).sum())
                    events.append((float(log['date'][k]), f'gave COVID to {int(log["target"][k])} via the {layer_name(int(log["layer"][k]))} layer ({x} secondary infections)'))
            if events:
                lines += [f'On day {day:.0f}, {person} {event}' for day, event in sorted(events, key=lambda e: e[0])]
            else:
                lines.append(f'Nothing happened to {person} during the simulation.')
        if not quiet:
            print('\n'.join(lines))
        return lines

    def make_nonnaive(self, inds):
        ''' Reset agents and mark them neither susceptible nor naive (reference people.py:412-431); ``inds`` are global ids '''
        inds = torch.as_tensor(inds, dtype=torch.int64, device=self.device)
        if self._sim is not None and self._sim._comm is not None:      # partitioned: this rank resets the agents it owns
            inds = inds[(inds >= self.id0) & (inds < self.id0 + self.n)] - self.id0
        self.make_naive(inds)
        self._arrays['susceptible'][inds] = False
        self._arrays['naive'][inds] = False

    def make_naive(self, inds, reset_vx=False):
        ''' Reset agents to the never-infected state (reference people.py:378-409); ``inds`` index this object's arrays '''
        inds = torch.as_tensor(inds, dtype=torch.int64, device=self.device)
        A = self._arrays
        for key in cvd.states:
            if key in ('susceptible', 'naive'):
                A[key][inds] = True
            elif key != 'vaccinated' or reset_vx:
                A[key][inds] = False
        for key in cvd.variant_states:
            A[key][inds] = float('nan')
        for key in cvd.by_variant_states:
            A[key][:, inds] = False
        non_vx = inds if reset_vx else inds[~A['vaccinated'][inds]]
        for key in cvd.imm_states:
            A[key][:, non_vx] = 0
        for key in cvd.nab_states + cvd.vacc_states:
            A[key][non_vx] = 0
        for key in cvd.dates + cvd.durs:
            if key != 'date_vaccinated' or reset_vx:
                A[key][inds] = float('nan')
