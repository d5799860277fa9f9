'''
Variants and the host-side (init-time) part of immunity: cross-immunity matrix and the per-day NAb
kinetics table (reference covasim/immunity.py:18-130, 269-300, 356-508).  The per-agent parts --
check_immunity, update_nab, update_peak_nab (immunity.py:138-350) -- run on the device
(csrc/people_kernels.cu, csrc/infect.cu, csrc/interventions.cu).
'''
import numpy as np

from . import defaults as cvd
from . import parameters as cvpar

__all__ = ['variant', 'init_immunity', 'precompute_waning', 'nab_growth_decay', 'calc_VE', 'calc_VE_symp']


class variant:
    ''' A variant introduced by importation on the given day(s) (reference immunity.py:18-130) '''

    def __init__(self, variant, days, label=None, n_imports=1, rescale=True):
        self.days = days
        self.n_imports = int(n_imports)
        self.rescale = rescale
        self.index = None
        self.initialized = False
        if isinstance(variant, str):
            choices, mapping = cvpar.get_variant_choices()
            key = variant.lower()
            for txt in ['.', ' ', 'variant', 'voc']:
                key = key.replace(txt, '')
            if key not in mapping:
                raise NotImplementedError(f'The selected variant "{variant}" is not implemented; choices are: {choices}')
            self.label = mapping[key]
            self.p = dict(cvpar.get_variant_pars(variant=self.label))
        elif isinstance(variant, dict):
            p = dict(variant)
            label = p.pop('label', label)
            bad = [k for k in p if k not in cvd.variant_par_keys]
            if bad:
                raise KeyError(f'Could not parse variant keys "{bad}"; valid keys are: {cvd.variant_par_keys}')
            full = dict(cvpar.get_variant_pars(default=True))
            full.update(p)
            self.p = full
            self.label = label or 'custom'
        else:
            raise ValueError(f'Could not understand {type(variant)}, please specify as a dict or a predefined variant')

    def initialize(self, sim):
        self.days = np.sort(np.atleast_1d(np.array([sim.day(d) for d in np.atleast_1d(self.days)])))
        sim['variant_pars'][self.label] = self.p
        self.index = list(sim['variant_pars'].keys()).index(self.label)
        sim['variant_map'][self.index] = self.label
        self.initialized = True

    def apply(self, sim):
        ''' Import infections of this variant (reference immunity.py:117-130); host NumPy-stream draws, device infect '''
        if np.any(self.days == sim.t):
            scale = sim.rescale_vec[sim.t] if self.rescale else 1.0
            n_imports = int(np.floor(self.n_imports / scale + sim.rng.np_.random_sample()))      # sc.randround
            if sim.rng_mode == 'mt':                   # replay mode: the reference's O(N) permutation, same stream use
                import torch
                sus = torch.nonzero(sim.people.susceptible).flatten().cpu().numpy()
                who = sim.rng.np_.choice(sus, n_imports, replace=False)
            else:
                who, _ = sim._choose_true('susceptible', sim.rng.np_, n_imports)
            sim.people.infect(who, layer='importation', variant=self.index)
            sim._host_add('n_imports', sim.t, n_imports)


def nab_growth_decay(length, growth_time, decay_rate1, decay_time1, decay_rate2, decay_time2):
    ''' Per-day NAb increments: linear growth, then exponential decay whose rate itself decays (immunity.py:404-448) '''
    if decay_time2 < decay_time1:
        raise ValueError(f'Decay time 2 must be larger than decay time 1, but you supplied {decay_time2} which is smaller than {decay_time1}.')
    length = length + 1
    t2 = np.arange(length - growth_time, dtype=np.int32)
    growth = np.arange(growth_time, dtype=np.int32) / growth_time
    rate = np.full(len(t2), decay_rate1, dtype=float)
    rate[t2 > decay_time2] = decay_rate2
    mid = np.nonzero((t2 > decay_time1) * (t2 <= decay_time2))[0]
    slowing = (1 / (decay_time2 - decay_time1)) * (decay_rate1 - decay_rate2)
    rate[mid] = decay_rate1 - slowing * np.arange(len(mid), dtype=np.int32)
    titre = np.zeros(len(t2))
    for i in range(1, len(t2)):
        titre[i] = titre[i - 1] + rate[i]
    y = np.concatenate([growth, np.exp(-titre)])
    return np.diff(y)[0:length]


def precompute_waning(length, pars=None):
    pars = dict(pars)
    form = pars.pop('form')
    if form is None or form == 'nab_growth_decay':
        return nab_growth_decay(length, **pars)
    if callable(form):
        return form(length, **pars)
    raise NotImplementedError(f'The selected functional form "{form}" is not built; choices are: nab_growth_decay or a callable')


def init_immunity(sim, create=False):
    ''' Cross-immunity matrix and NAb kinetics table (reference immunity.py:269-300) '''
    if not sim['use_waning']:
        return
    nv = sim['n_variants']
    if sim['immunity'] is None or create:
        imm = np.ones((nv, nv), dtype=cvd.default_float)
        cross = cvpar.get_cross_immunity()
        for i in range(nv):
            li = sim['variant_map'][i]
            for j in range(nv):
                lj = sim['variant_map'][j]
                if li in cross and lj in cross:
                    imm[j][i] = cross[lj][li]
        sim['immunity'] = imm
    sim['nab_kin'] = precompute_waning(length=sim.npts, pars=sim['nab_decay'])


def calc_VE(nab, ax, pars):
    ''' NAb level -> protection on one axis (reference immunity.py:216-247); host version for analysis code '''
    key = dict(sus=('alpha_inf', 'beta_inf'), symp=('alpha_symp_inf', 'beta_symp_inf'), sev=('alpha_sev_symp', 'beta_sev_symp'))
    if ax not in key:
        raise ValueError(f'Choice {ax} not in list of choices: sus, symp, sev')
    a, b = key[ax]
    lo = np.exp(pars[a]) * nab ** pars[b]
    return lo / (1 + lo)


def calc_VE_symp(nab, pars):
    ''' Marginal protection against symptomatic disease (reference immunity.py:250-262) '''
    inf = calc_VE(nab, 'sus', pars)
    symp = calc_VE(nab, 'symp', pars)
    return 1 - ((1 - inf) * (1 - symp))
