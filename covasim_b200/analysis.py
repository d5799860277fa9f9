'''
Goodness of fit between simulated and observed time series (reference covasim/analysis.py:991-1222 Fit, misc.py:707-795
compute_gof): what a calibration sweep (BASELINE config 5) evaluates for every ensemble member.  Pure host arithmetic on the
result series the device produced; ``fit_members`` evaluates all members of an ensemble at once.

Terminology follows the reference: difference (sim - data per matched day), goodness of fit (normalised absolute difference by
default), loss (gof x weight), mismatch (sum of the losses over days and keys -- the number a calibration minimises).
'''
import datetime as dt

import numpy as np

__all__ = ['Analyzer', 'snapshot', 'age_histogram', 'compute_gof', 'Fit', 'fit_members']


class Analyzer:
    '''
    Base class of analyzers (reference analysis.py:23-129): objects called once per day at the END of ``Sim.step`` (after the
    day's transmission and counts, reference sim.py:677-678) with full access to the sim; ``sim.people.<field>`` are device
    tensors.  Plain callables are accepted as analyzers too.
    '''

    def __init__(self, label=None):
        self.label = label if label is not None else self.__class__.__name__
        self.initialized = False
        self.finalized = False

    def __call__(self, *args, **kwargs):
        if not self.initialized:
            raise RuntimeError(f'Analyzer (label={self.label}, {type(self)}) has not been initialized')
        return self.apply(*args, **kwargs)

    def initialize(self, sim=None):
        self.initialized = True
        self.finalized = False

    def finalize(self, sim=None):
        if self.finalized:
            raise RuntimeError('Analyzer already finalized')
        self.finalized = True

    def apply(self, sim):
        raise NotImplementedError

    def shrink(self, in_place=False):
        import copy
        return self if in_place else copy.deepcopy(self)


def _process_days(sim, days):
    days = [days] if isinstance(days, (str, int, np.integer, dt.date)) else list(days)
    days = sorted(sim.day(d) for d in days)
    return days, [sim.date(d) for d in days]


class snapshot(Analyzer):
    '''
    Host copies of every People array on the given days (reference analysis.py:149-223): ``snap.snapshots[date]`` /
    ``snap.get(day)`` is a dict of NumPy arrays (the reference deep-copies the People object; here the device arrays are read
    back once per requested day).
    '''

    def __init__(self, days, *args, die=True, **kwargs):
        super().__init__(**kwargs)
        days = [days] if isinstance(days, (str, int, np.integer, dt.date)) else list(days)
        self.days = days + list(args)
        self.die = die
        self.dates = None
        self.snapshots = {}

    def initialize(self, sim):
        self.days, self.dates = _process_days(sim, self.days)
        if self.days[-1] > sim.npts - 1:
            raise ValueError(f'Cannot create snapshot for {self.dates[-1]} (day {self.days[-1]}) because the simulation ends on day {sim.npts - 1}')
        self._sim_day, self._sim_date = sim.day, sim.date
        self.initialized = True

    def apply(self, sim):
        for ind, day in enumerate(self.days):
            if day == sim.t:
                self.snapshots[self.dates[ind]] = {k: sim.people.to_numpy(k) for k in sim.people.keys()}

    def finalize(self, sim=None):
        super().finalize()
        missing = [d for d in self.dates if d not in self.snapshots]
        if missing and self.die:
            raise RuntimeError(f'The dates {missing} were requested but not recorded')

    def get(self, key=None):
        date = self._sim_date(self._sim_day(self.days[0] if key is None else key))
        if date not in self.snapshots:
            raise KeyError(f'Could not find snapshot date {date}: choices are {", ".join(self.snapshots.keys())}')
        return self.snapshots[date]


class age_histogram(Analyzer):
    '''
    Age distribution of the agents who have ever been in the given states, on the given days (reference analysis.py:226-425,
    the plotting and data-comparison parts left out): ``hist.hists[date][state]`` counts, per age bin, the agents whose
    ``date_<state>`` is set, times the day's rescaling factor.  Counted on the device.
    '''

    def __init__(self, days=None, states=None, edges=None, die=True, **kwargs):
        super().__init__(**kwargs)
        self.days, self.states, self.edges, self.die = days, states, edges, die
        self.hists = {}

    def initialize(self, sim):
        super().initialize()
        self.days, self.dates = _process_days(sim, sim.npts - 1 if self.days is None else self.days)
        if self.days[-1] > sim.npts - 1:
            raise ValueError(f'Cannot create histogram for day {self.days[-1]} because the simulation ends on day {sim.npts - 1}')
        self.edges = np.linspace(0, 100, 11) if self.edges is None else np.asarray(self.edges, dtype=float)
        self.bins = self.edges[:-1]
        states = ['exposed', 'severe', 'dead', 'tested', 'diagnosed'] if self.states is None else ([self.states] if isinstance(self.states, str) else list(self.states))
        self.states = [s.replace('date_', '') for s in states]

    def apply(self, sim):
        import torch
        for ind, day in enumerate(self.days):
            if day != sim.t:
                continue
            age = sim.people.age
            edges = torch.as_tensor(self.edges, dtype=age.dtype, device=age.device)
            nb = len(self.edges) - 1
            hist = dict(bins=self.bins)
            for state in self.states:
                who = ~torch.isnan(sim.people[f'date_{state}'])
                a = age[who]
                # np.histogram: [e_i, e_i+1) except the last bin, which also holds its right edge
                b = torch.bucketize(a, edges, right=True) - 1
                b = torch.where(a == edges[-1], torch.full_like(b, nb - 1), b)
                b = b[(b >= 0) & (b < nb)]
                hist[state] = torch.bincount(b, minlength=nb).cpu().numpy() * sim.rescale_vec[sim.t]
            self.hists[self.dates[ind]] = hist

    def finalize(self, sim=None):
        super().finalize()
        missing = [d for d in self.dates if d not in self.hists]
        if missing and self.die:
            raise RuntimeError(f'The dates {missing} were requested but not recorded')

    def get(self, key=None):
        date = self.dates[0] if key is None else (key if key in self.hists else self.dates[self.days.index(int(key))])
        return self.hists[date]


def compute_gof(actual, predicted, normalize=True, use_frac=False, use_squared=False, as_scalar='none', eps=1e-9, estimator=None, **kwargs):
    ''' Goodness of fit of ``predicted`` against ``actual`` (reference misc.py:707-795; the scikit-learn estimators are not wired in) '''
    actual = np.array(actual, dtype=float)
    predicted = np.array(predicted, dtype=float)
    if estimator is not None:
        return estimator(actual, predicted, **kwargs)
    gofs = abs(actual - predicted)
    if normalize and not use_frac:
        actual_max = abs(actual).max()
        if actual_max > 0:
            gofs /= actual_max
    if use_frac:
        if (actual < 0).any() or (predicted < 0).any():
            print('Warning: Calculating fractional errors for non-positive quantities is ill-advised!')
        else:
            gofs /= np.maximum(actual, predicted) + eps
    if use_squared:
        gofs = gofs ** 2
    if as_scalar == 'sum':
        gofs = np.sum(gofs)
    elif as_scalar == 'mean':
        gofs = np.mean(gofs)
    elif as_scalar == 'median':
        gofs = np.median(gofs)
    return gofs


def _data_days(data, start_date):
    ''' Day index of every data row: a 'day' column of integers, or a 'date' column / index of dates or ISO strings '''
    if 'day' in data:
        return [int(d) for d in data['day']]
    dates = data['date'] if 'date' in data else getattr(data, 'index', None)
    if dates is None:
        raise ValueError('data needs a "day" or a "date" column (or a date index)')
    out = []
    for d in dates:
        if isinstance(d, str):
            d = dt.datetime.strptime(d, '%Y-%m-%d').date()
        if isinstance(d, dt.datetime):
            d = d.date()
        out.append((d - start_date).days)
    return out


class Fit:
    '''
    Fit between a finished simulation and data (reference analysis.py:991-1222).  ``data`` maps column names to arrays (a dict
    or a DataFrame) plus a 'day' / 'date' column or a date index; by default every cumulative result that is also a data column
    is used, deaths weighted 10 and diagnoses 5.
    '''

    def __init__(self, sim=None, data=None, weights=None, keys=None, custom=None, compute=True, results=None, npts=None, start_date=None, **kwargs):
        self.weights = dict({'cum_deaths': 10, 'cum_diagnoses': 5}, **(weights or {}))
        self.custom = dict(custom or {})
        self.keys = keys
        self.gof_kwargs = kwargs
        if sim is not None:
            if not sim.results_ready:
                raise RuntimeError('Model fit cannot be calculated until results are run')
            results = {k: sim.results[k].values for k in sim.result_keys()}
            npts, start_date = sim.npts, sim._start_date()
        if data is None:
            raise RuntimeError('Model fit cannot be calculated until data are loaded')
        self.data, self.sim_results, self.sim_npts, self.start_date = data, results, npts, start_date
        self.inds = dict(sim={}, data={})
        self.pair, self.diffs, self.gofs, self.losses, self.mismatches = {}, {}, {}, {}, {}
        self.mismatch = None
        if compute:
            self.compute()

    def compute(self):
        self.reconcile_inputs()
        self.compute_diffs()
        self.compute_gofs()
        self.compute_losses()
        return self.compute_mismatch()

    def reconcile_inputs(self):
        ''' Matching keys and days between the model and the data (analysis.py:1087-1169) '''
        data_cols = [c for c in (self.data.columns if hasattr(self.data, 'columns') else self.data.keys()) if c not in ('day', 'date')]
        if self.keys is None:
            self.keys = [k for k in self.sim_results.keys() if k.startswith('cum_') and k in data_cols]
            if not self.keys:
                raise KeyError(f'No matches found between simulation result keys and data columns {data_cols}')
        missing = [k for k in self.keys if k not in data_cols]
        if missing:
            raise KeyError(f'The following requested key(s) were not found in the data: {", ".join(missing)}')
        days = _data_days(self.data, self.start_date)
        matches = 0
        for key in self.keys:
            col = np.asarray(self.data[key], dtype=float)
            sel = [(d, j) for j, d in enumerate(days) if np.isfinite(col[j]) and 0 <= d < self.sim_npts]
            self.inds['sim'][key] = np.array([d for d, _ in sel], dtype=int)
            self.inds['data'][key] = np.array([j for _, j in sel], dtype=int)
            self.pair[key] = dict(sim=np.asarray(self.sim_results[key], dtype=float)[self.inds['sim'][key]], data=col[self.inds['data'][key]])
            matches += len(sel)
        for key, custom in self.custom.items():
            if 'sim' not in custom or 'data' not in custom:
                raise KeyError(f'Custom input must have "sim" and "data" keys, not {list(custom.keys())}')
            if len(custom['data']) != len(custom['sim']):
                raise ValueError('Custom data and sim must be arrays of the same length')
            if key in self.pair:
                raise ValueError(f'You cannot use a custom key "{key}" that matches one of the existing keys')
            self.pair[key] = dict(sim=np.asarray(custom['sim'], dtype=float), data=np.asarray(custom['data'], dtype=float))
            self.weights[key] = custom.get('weights', custom.get('weight', 1.0))
            matches += 1
        if matches == 0:
            raise ValueError('No paired data points were found between the supplied data and the simulation; please check the dates for each')

    def compute_diffs(self, absolute=False):
        for key, pair in self.pair.items():
            self.diffs[key] = np.abs(pair['sim'] - pair['data']) if absolute else pair['sim'] - pair['data']

    def compute_gofs(self, **kwargs):
        kw = dict(self.gof_kwargs, **kwargs)
        for key, pair in self.pair.items():
            self.gofs[key] = compute_gof(pair['data'], pair['sim'], **kw)

    def compute_losses(self):
        for key, gof in self.gofs.items():
            weight = self.weights.get(key, 1.0)
            if np.ndim(weight):
                weight = np.asarray(weight, dtype=float)
                if len(weight) == self.sim_npts and len(weight) != len(gof):
                    weight = weight[self.inds['sim'][key]]
                elif len(weight) != len(gof):
                    raise ValueError(f'Could not map weight array of length {len(weight)} onto simulation of length {self.sim_npts} or data-model matches of length {len(gof)}')
            self.losses[key] = gof * weight

    def compute_mismatch(self, use_median=False):
        for key, loss in self.losses.items():
            self.mismatches[key] = np.median(loss) if use_median else np.sum(loss)
        self.mismatch = float(np.sum(list(self.mismatches.values())))
        return self.mismatch


def fit_members(member_results, data, npts, start_date=None, **kwargs):
    '''
    Mismatch of every member of an ensemble (``MultiSim.member_results``: one dict of result arrays per member) against the same
    data -- the objective of a calibration sweep -- as a float64 array in member order.
    '''
    start_date = start_date or dt.date(2020, 3, 1)
    return np.array([Fit(data=data, results=res, npts=npts, start_date=start_date, **kwargs).mismatch for res in member_results])


class TransTree:
    '''
    The transmission tree as arrays: a host view of the device infection log (reference analysis.py:1772-1866; the reference
    builds Python lists of dicts per person, which does not scale to the populations this package runs).

    ``source[k] -> target[k]`` on ``date[k]`` over layer ``layer[k]`` (index into the sim's layer keys; -1 seed infection, -2
    importation) with variant ``variant[k]``; ``source`` is -1 for seed infections and importations.  Per-person views follow the
    reference's attributes: ``sources`` / ``source_dates`` (who infected each person, and when; -1 / NaN if nobody),
    ``n_targets`` (count_targets), ``transmissions`` / ``source_inds`` / ``target_inds`` (count_transmissions), ``r0()``.
    The reference's loop tests ``if source:`` (analysis.py:1822), so transmissions whose source is person 0 do not enter
    ``sources`` / ``targets``; ``count_targets`` reproduces that, ``count_transmissions`` and ``r0`` count every transmission.
    '''

    def __init__(self, sim=None, log=None, pop_size=None, n_days=None, date_exposed=None, date_recovered=None, layer_keys=None):
        if sim is not None:
            log = sim.infection_log
            pop_size, n_days = sim.n, int(sim.people.t)
            date_exposed, date_recovered = sim.people.to_numpy_many(('date_exposed', 'date_recovered'))
            layer_keys = sim.people.layer_keys()
        self.pop_size, self.n_days, self.layer_keys = int(pop_size), int(n_days), layer_keys
        self.source = np.asarray(log['source'], dtype=np.int64)
        self.target = np.asarray(log['target'], dtype=np.int64)
        self.date = np.asarray(log['date'], dtype=np.int64)
        self.layer = np.asarray(log['layer'], dtype=np.int64)
        self.variant = np.asarray(log['variant'], dtype=np.int64)
        self.date_exposed, self.date_recovered = date_exposed, date_recovered
        truthy = self.source > 0                                   # the reference's `if source:` (None and person 0 are both falsy)
        self.sources = np.full(self.pop_size, -1, dtype=np.int64)
        self.source_dates = np.full(self.pop_size, np.nan)
        order = np.argsort(self.date[truthy], kind='stable')       # a later infection of the same person overwrites the earlier one
        self.sources[self.target[truthy][order]] = self.source[truthy][order]
        self.source_dates[self.target[truthy][order]] = self.date[truthy][order]
        self._n_given = np.bincount(self.source[truthy], minlength=self.pop_size)     # len(targets[i])
        self.count_targets()
        self.count_transmissions()

    def __len__(self):
        return len(self.target)

    def targets_of(self, person):
        ''' The people ``person`` infected, in log order (the reference's ``tt.targets[person]``) '''
        return self.target[(self.source == person) & (self.source > 0)]

    def count_targets(self, start_day=None, end_day=None):
        ''' Number of people infected by everybody who was themselves infected (by somebody) between the two days (analysis.py:1880-1903) '''
        start_day = 0 if start_day is None else int(start_day)
        end_day = self.n_days if end_day is None else int(end_day)
        has = (self.sources >= 0) & (self.source_dates >= start_day) & (self.source_dates <= end_day)
        self.n_targets = self._n_given[has].astype(float)
        return self.n_targets

    def count_transmissions(self):
        ''' Every transmission with a source, as [source, target] pairs (analysis.py:1906-1925) '''
        has = self.source >= 0
        self.source_inds, self.target_inds = self.source[has], self.target[has]
        self.transmissions = np.stack([self.source_inds, self.target_inds], axis=1)
        return self.transmissions

    def r0(self, recovered_only=False):
        ''' Mean number of transmissions per person who was ever exposed (analysis.py:2106-2129, without NetworkX) '''
        if self.date_exposed is None:
            raise RuntimeError('r0 needs the date_exposed array (build the tree from a sim)')
        keep = ~np.isnan(self.date_exposed)
        if recovered_only:
            keep &= ~(self.date_recovered > self.n_days)
        out_degree = np.bincount(self.source[self.source >= 0], minlength=self.pop_size)
        return float(np.mean(out_degree[keep])) if keep.any() else float('nan')


__all__ += ['TransTree']
