'''
TEST INFRASTRUCTURE -- the oracle's OWN copy of covasim_b200/parameters.py, so that oracle/cvoracle.py never imports the
product package (which loads the CUDA library).  Both copies are checked against the tables recorded from the unmodified
reference (tests/golden/ref_config.json, oracle/gen_config_golden.py) by tests/test_oracle_golden.py::test_config_tables.

Simulation parameters: the values the kernels are fed with.

Parameter *names and default values* follow the reference (covasim/parameters.py:15-151 for the
main dict, :155-227 for per-layer values, :230-296 for prognoses, :375-608 for variant, cross-
immunity and vaccine tables) because they are the model's inputs, not implementation.  Only the
parameters the hot path reads are kept; plotting / versioning / location data are out of scope.
'''
import copy
import numpy as np
from . import ref_defaults as cvd

__all__ = ['make_pars', 'reset_layer_pars', 'get_prognoses', 'get_variant_pars', 'get_cross_immunity',
           'get_vaccine_variant_pars', 'get_vaccine_dose_pars', 'get_variant_choices', 'get_vaccine_choices',
           'layer_pars']

layer_pars = ['beta_layer', 'contacts', 'dynam_layer', 'iso_factor', 'quar_factor']

_LN2 = float(np.log(2))


def _dist(dist, par1, par2, **kw):
    return dict(dist=dist, par1=par1, par2=par2, **kw)


def make_pars(set_prognoses=False, prog_by_age=True, **kwargs):
    ''' Build the default parameter dictionary, then overlay ``kwargs`` (reference parameters.py:15-151) '''
    p = dict(
        # population
        pop_size=20e3, pop_infected=20, pop_type='random', location=None,
        # time
        start_day='2020-03-01', end_day=None, n_days=60, rand_seed=1, verbose=0,
        # rescaling
        pop_scale=1, scaled_pop=None, rescale=True, rescale_threshold=0.05, rescale_factor=1.2,
        frac_susceptible=1.0,
        # network (filled by reset_layer_pars)
        contacts=None, dynam_layer=None, beta_layer=None,
        # transmission
        beta_dist=_dist('neg_binomial', 1.0, 0.45, step=0.01),
        viral_dist=dict(frac_time=0.3, load_ratio=2, high_cap=4),
        beta=0.016, asymp_factor=1.0,
        n_imports=0, n_variants=1,
        # immunity
        use_waning=True,
        nab_init=_dist('normal', 0, 2),
        nab_decay=dict(form='nab_growth_decay', growth_time=21, decay_rate1=_LN2 / 50, decay_time1=150,
                       decay_rate2=_LN2 / 250, decay_time2=365),
        nab_kin=None, nab_boost=1.5,
        nab_eff=dict(alpha_inf=1.08, alpha_inf_diff=1.812, beta_inf=0.967, alpha_symp_inf=-0.739,
                     beta_symp_inf=0.038, alpha_sev_symp=-0.014, beta_sev_symp=0.079),
        rel_imm_symp=dict(asymp=0.85, mild=1, severe=1.5),
        immunity=None, trans_redux=0.59,
        rel_beta=1.0,
        # durations
        dur=dict(
            exp2inf=_dist('lognormal_int', 4.5, 1.5), inf2sym=_dist('lognormal_int', 1.1, 0.9),
            sym2sev=_dist('lognormal_int', 6.6, 4.9), sev2crit=_dist('lognormal_int', 1.5, 2.0),
            asym2rec=_dist('lognormal_int', 8.0, 2.0), mild2rec=_dist('lognormal_int', 8.0, 2.0),
            sev2rec=_dist('lognormal_int', 18.1, 6.3), crit2rec=_dist('lognormal_int', 18.1, 6.3),
            crit2die=_dist('lognormal_int', 10.7, 4.8)),
        # severity
        rel_symp_prob=1.0, rel_severe_prob=1.0, rel_crit_prob=1.0, rel_death_prob=1.0,
        prog_by_age=prog_by_age, prognoses=None,
        # protection measures
        iso_factor=None, quar_factor=None, quar_period=14,
        # events
        interventions=[], analyzers=[], timelimit=None, stopping_func=None,
        # health system
        n_beds_hosp=None, n_beds_icu=None, no_hosp_factor=2.0, no_icu_factor=2.0,
        # vaccines / variants
        vaccine_pars={}, vaccine_map={}, variants=[], variant_map={0: 'wild'}, variant_pars=dict(wild={}),
    )
    for key in cvd.variant_par_keys:
        p['variant_pars']['wild'][key] = p[key]
    p.update(kwargs)
    reset_layer_pars(p)
    if set_prognoses:
        p['prognoses'] = get_prognoses(p['prog_by_age'])
    return p


_layer_defaults = dict(
    random=dict(beta_layer=dict(a=1.0), contacts=dict(a=20), dynam_layer=dict(a=0),
                iso_factor=dict(a=0.2), quar_factor=dict(a=0.3)),
    hybrid=dict(beta_layer=dict(h=3.0, s=0.6, w=0.6, c=0.3), contacts=dict(h=2.0, s=20, w=16, c=20),
                dynam_layer=dict(h=0, s=0, w=0, c=0), iso_factor=dict(h=0.3, s=0.1, w=0.1, c=0.1),
                quar_factor=dict(h=0.6, s=0.2, w=0.2, c=0.2)),
)


def reset_layer_pars(pars, layer_keys=None, force=False):
    ''' Fill the per-layer parameter dicts for the population type (reference parameters.py:158-227) '''
    pop_type = pars['pop_type']
    if pop_type not in _layer_defaults:
        raise ValueError(f'Cannot load defaults for population type "{pop_type}": must be hybrid or random')
    defaults = _layer_defaults[pop_type]
    default_keys = list(defaults['beta_layer'].keys())
    for pkey in layer_pars:
        fallback = _layer_defaults['random'][pkey]['a']
        given = dict(defaults[pkey])
        if not force and pars.get(pkey):
            given.update(pars[pkey])
        if layer_keys:
            keys = list(layer_keys)
        else:
            keys = list(dict.fromkeys(default_keys + list(given.keys())))
        pars[pkey] = {lk: given.get(lk, fallback) for lk in keys}
    return


def get_prognoses(by_age=True):
    ''' Age-banded prognosis probabilities, converted to conditional form (reference parameters.py:230-296) '''
    if not by_age:
        prog = dict(age_cutoffs=[0], sus_ORs=[1.0], trans_ORs=[1.0], symp_probs=[0.75], comorbidities=[1.0],
                    severe_probs=[0.10], crit_probs=[0.04], death_probs=[0.01])
    else:
        prog = dict(
            age_cutoffs=[0, 10, 20, 30, 40, 50, 60, 70, 80, 90],
            sus_ORs=[0.34, 0.67, 1.00, 1.00, 1.00, 1.00, 1.24, 1.47, 1.47, 1.47],
            trans_ORs=[1.0] * 10,
            comorbidities=[1.0] * 10,
            symp_probs=[0.50, 0.55, 0.60, 0.65, 0.70, 0.75, 0.80, 0.85, 0.90, 0.90],
            severe_probs=[0.00050, 0.00165, 0.00720, 0.02080, 0.03430, 0.07650, 0.13280, 0.20655, 0.24570, 0.24570],
            crit_probs=[0.00003, 0.00008, 0.00036, 0.00104, 0.00216, 0.00933, 0.03639, 0.08923, 0.17420, 0.17420],
            death_probs=[0.00002, 0.00002, 0.00010, 0.00032, 0.00098, 0.00265, 0.00766, 0.02439, 0.08292, 0.16190],
        )
    prog = {k: np.array(v, dtype=float if k != 'age_cutoffs' else int) for k, v in prog.items()}
    # Absolute -> conditional, in this order (death|crit, crit|severe, severe|symp)
    prog['death_probs'] /= prog['crit_probs']
    prog['crit_probs'] /= prog['severe_probs']
    prog['severe_probs'] /= prog['symp_probs']
    return prog


def get_variant_choices():
    choices = dict(
        wild=['wild', 'default', 'pre-existing', 'original'],
        alpha=['alpha', 'b117', 'uk', 'united kingdom', 'kent'],
        beta=['beta', 'b1351', 'sa', 'south africa'],
        gamma=['gamma', 'p1', 'b11248', 'brazil'],
        delta=['delta', 'b16172', 'india'])
    mapping = {name: key for key, names in choices.items() for name in names}
    return choices, mapping


def get_vaccine_choices():
    choices = dict(
        default=['default', None],
        pfizer=['pfizer', 'biontech', 'pfizer-biontech', 'pf', 'pfz', 'pz', 'bnt162b2', 'comirnaty'],
        moderna=['moderna', 'md', 'spikevax'],
        novavax=['novavax', 'nova', 'covovax', 'nvx', 'nv'],
        az=['astrazeneca', 'az', 'covishield', 'oxford', 'vaxzevria'],
        jj=['jnj', 'johnson & johnson', 'janssen', 'jj'],
        sinovac=['sinovac', 'coronavac'],
        sinopharm=['sinopharm'])
    mapping = {name: key for key, names in choices.items() for name in names}
    return choices, mapping


def _pick(table, default, key, defaultkey):
    if isinstance(default, str):
        key, default = default, key
    if key is not None:
        if key not in table:
            raise KeyError(f'Key "{key}" not found; choices are: {", ".join(table.keys())}')
        return table[key]
    return table[defaultkey] if default else table


def get_variant_pars(default=False, variant=None):
    ''' (rel_beta, rel_symp_prob, rel_severe_prob, rel_crit_prob, rel_death_prob) per variant; ref parameters.py:375-424 '''
    rows = dict(wild=(1.0, 1.0, 1.0, 1.0, 1.0), alpha=(1.67, 1.0, 1.64, 1.0, 1.0), beta=(1.0, 1.0, 3.6, 1.0, 1.0),
                gamma=(2.05, 1.0, 2.6, 1.0, 1.0), delta=(2.2, 1.0, 3.2, 1.0, 1.0))
    table = {k: dict(zip(cvd.variant_par_keys, v)) for k, v in rows.items()}
    return _pick(table, default, variant, 'wild')


def get_cross_immunity(default=False, variant=None):
    ''' cross[a][b]: protection against b given prior infection with a ... ref parameters.py:427-475 '''
    cols = ('wild', 'alpha', 'beta', 'gamma', 'delta')
    rows = dict(wild=(1.0, 0.5, 0.5, 0.34, 0.374), alpha=(0.5, 1.0, 0.8, 0.8, 0.689),
                beta=(0.066, 0.5, 1.0, 0.5, 0.086), gamma=(0.34, 0.4, 0.4, 1.0, 0.088),
                delta=(0.374, 0.689, 0.086, 0.088, 1.0))
    table = {k: dict(zip(cols, v)) for k, v in rows.items()}
    return _pick(table, default, variant, 'wild')


def get_vaccine_variant_pars(default=False, vaccine=None):
    ''' Relative NAb efficacy of each vaccine against each variant; ref parameters.py:478-551 '''
    cols = ('wild', 'alpha', 'beta', 'gamma', 'delta')
    rows = dict(
        default=(1.0, 1.0, 1.0, 1.0, 1.0),
        pfizer=(1.0, 1 / 2.0, 1 / 10.3, 1 / 6.7, 1 / 2.9),
        moderna=(1.0, 1 / 1.8, 1 / 4.5, 1 / 8.6, 1 / 2.9),
        az=(1.0, 1 / 2.3, 1 / 9, 1 / 2.9, 1 / 6.2),
        jj=(1.0, 1.0, 1 / 3.6, 1 / 3.4, 1 / 1.6),
        novavax=(1.0, 1 / 1.12, 1 / 4.7, 1 / 8.6, 1 / 6.2),
        sinovac=(1.0, 1 / 1.12, 1 / 4.7, 1 / 8.6, 1 / 6.2),
        sinopharm=(1.0, 1 / 1.12, 1 / 4.7, 1 / 8.6, 1 / 6.2))
    table = {k: dict(zip(cols, v)) for k, v in rows.items()}
    return _pick(table, default, vaccine, 'default')


def get_vaccine_dose_pars(default=False, vaccine=None):
    ''' (nab_init mean, nab_boost, doses, interval) per vaccine; ref parameters.py:554-608 '''
    rows = dict(default=(0, 2, 1, None), pfizer=(-1, 4, 2, 21), moderna=(-1, 8, 2, 28), az=(-1.5, 2, 2, 21),
                jj=(1, 3, 1, None), novavax=(-0.9, 3, 2, 21), sinovac=(-2, 2, 2, 14), sinopharm=(-1, 2, 2, 21))
    table = {k: dict(nab_init=_dist('normal', m, 2), nab_boost=b, doses=d, interval=i)
             for k, (m, b, d, i) in rows.items()}
    return copy.deepcopy(_pick(table, default, vaccine, 'default'))
