'''
Goodness of fit between simulated and observed time series (reference covasim/analysis.py:991-1222 Fit, misc.py:707-795
compute_gof): what a calibration sweep (BASELINE config 5) evaluates for every ensemble member.  Pure host arithmetic on the
result series the device produced; ``fit_members`` evaluates all members of an ensemble at once.

Terminology follows the reference: difference (sim - data per matched day), goodness of fit (normalised absolute difference by
default), loss (gof x weight), mismatch (sum of the losses over days and keys -- the number a calibration minimises).
'''
import datetime as dt

import numpy as np

__all__ = ['compute_gof', 'Fit', 'fit_members']


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
