'''
TEST INFRASTRUCTURE -- records the parameter tables of the UNMODIFIED reference (/root/reference,
Covasim 3.1.7) as tests/golden/ref_config.json.  Run from the repo root:  python -m oracle.gen_config_golden

Both copies of the configuration tables -- covasim_b200/{defaults,parameters}.py (product) and
oracle/{ref_defaults,ref_parameters}.py (the oracle's own, so that the oracle never imports the product
package) -- are checked against this fixture by tests/test_oracle_golden.py::test_config_tables, so a wrong
constant cannot be "wrong on both sides" unnoticed.
'''
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refenv  # noqa: E402


def plain(x):
    ''' JSON-able form of nested dicts / arrays / scalars (functions and objects dropped) '''
    if isinstance(x, dict):
        return {str(k): plain(v) for k, v in x.items() if plain(v) is not None or v is None}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, float)):
        return float(x)
    if isinstance(x, (np.integer, int, bool, np.bool_)):
        return int(x) if not isinstance(x, (bool, np.bool_)) else bool(x)
    if isinstance(x, str) or x is None:
        return x
    return None


def collect(cvpar, cvd):
    ''' The same dictionary from any module pair exposing the reference's function names '''
    out = {}
    pars = cvpar.make_pars()
    skip = {'interventions', 'analyzers', 'variants', 'prognoses', 'version', 'timelimit', 'stopping_func', 'location', 'verbose',
            'variant_map', 'variant_pars', 'vaccine_pars', 'vaccine_map', 'immunity', 'nab_kin'}
    out['pars'] = {k: plain(v) for k, v in pars.items() if k not in skip}
    for by_age in (True, False):
        out[f'prognoses_{int(by_age)}'] = plain(dict(cvpar.get_prognoses(by_age)))
    for pop_type in ('random', 'hybrid'):
        p = cvpar.make_pars(pop_type=pop_type)
        cvpar.reset_layer_pars(p)
        out[f'layers_{pop_type}'] = {k: plain(p[k]) for k in ('beta_layer', 'contacts', 'dynam_layer', 'iso_factor', 'quar_factor')}
    vchoices, vmap = cvpar.get_variant_choices()
    out['variant_choices'] = plain(vchoices)
    out['variant_pars'] = {v: plain(cvpar.get_variant_pars(variant=v)) for v in vchoices}
    out['cross_immunity'] = {v: plain(cvpar.get_cross_immunity(variant=v)) for v in vchoices}
    xchoices, xmap = cvpar.get_vaccine_choices()
    out['vaccine_choices'] = plain(xchoices)
    out['vaccine_variant_pars'] = {v: plain(cvpar.get_vaccine_variant_pars(vaccine=v)) for v in xchoices if v != 'default'}
    out['vaccine_dose_pars'] = {v: plain(cvpar.get_vaccine_dose_pars(vaccine=v)) for v in xchoices if v != 'default'}
    out['vaccine_default'] = dict(variant=plain(cvpar.get_vaccine_variant_pars(default=True)), dose=plain(cvpar.get_vaccine_dose_pars(default=True)))
    for name in ('result_stocks', 'result_stocks_by_variant', 'result_flows', 'result_flows_by_variant', 'default_age_data'):
        v = getattr(cvd, name)
        out[name] = plain(list(v.keys()) if isinstance(v, dict) else v)     # the reference keeps {key: label} dicts
    return out


def reference_tables():
    cv = refenv.import_reference()
    import covasim.parameters as rpar
    import covasim.defaults as rd
    out = collect(rpar, rd)
    meta = rd.PeopleMeta()
    out['people_fields'] = dict(person=list(meta.person), states=list(meta.states), variant_states=list(meta.variant_states),
                                by_variant_states=list(meta.by_variant_states), imm_states=list(meta.imm_states), nab_states=list(meta.nab_states),
                                vacc_states=list(meta.vacc_states), dates=list(meta.dates), durs=list(meta.durs))
    out['version'] = cv.__version__
    return out


if __name__ == '__main__':
    tables = reference_tables()
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_config.json')
    with open(path, 'w') as f:
        json.dump(tables, f, indent=1, sort_keys=True)
    print('wrote', path, os.path.getsize(path), 'bytes')
