'''
TEST INFRASTRUCTURE -- the oracle's OWN copy of covasim_b200/defaults.py, so that oracle/cvoracle.py never imports the
product package (which loads the CUDA library).  Both copies are checked against the tables recorded from the unmodified
reference (tests/golden/ref_config.json, oracle/gen_config_golden.py) by tests/test_oracle_golden.py::test_config_tables.

Dtypes, per-agent field inventory and result-key inventory.

This is the structure-of-arrays contract between the Python host and the CUDA
kernels: every per-agent field listed here is one contiguous device array, and
the integer ids in ``FIELD_IDS`` are the ``cvb_field`` enum of
``include/covasim_b200.h``.  The names and dtypes follow the reference's
``PeopleMeta`` (reference covasim/defaults.py:37-138) and result lists
(defaults.py:145-201) so that ``sim.people.<name>`` / ``sim.results[<key>]``
mean the same thing as in the reference.
'''
import numpy as np

default_float = np.float32   # reference defaults.py:20-25 (precision=32)
default_int = np.int32
result_float = np.float64

# ---- per-agent fields -------------------------------------------------------------------------
person_fields = ('uid', 'age', 'sex', 'symp_prob', 'severe_prob', 'crit_prob', 'death_prob',
                 'rel_trans', 'rel_sus', 'n_infections', 'n_breakthroughs')
person_int_fields = ('uid', 'n_infections', 'n_breakthroughs')

states = ('susceptible', 'naive', 'exposed', 'infectious', 'symptomatic', 'severe', 'critical',
          'tested', 'diagnosed', 'recovered', 'known_dead', 'dead', 'known_contact',
          'quarantined', 'isolated', 'vaccinated')

variant_states = ('exposed_variant', 'infectious_variant', 'recovered_variant')   # f32, NaN = none
by_variant_states = ('exposed_by_variant', 'infectious_by_variant')              # bool [nv, N]
imm_states = ('sus_imm', 'symp_imm', 'sev_imm')                                   # f32 [nv, N]
nab_states = ('peak_nab', 'nab', 't_nab_event')
vacc_states = ('doses', 'vaccine_source')

dates = tuple(f'date_{s}' for s in states) + ('date_pos_test', 'date_end_quarantine', 'date_end_isolation')
durs = ('dur_exp2inf', 'dur_inf2sym', 'dur_sym2sev', 'dur_sev2crit', 'dur_disease')

all_states = (person_fields + states + variant_states + by_variant_states + imm_states
              + nab_states + vacc_states + dates + durs)

# Device-only per-agent scratch (the pending-quarantine ring that replaces the host dict
# People._pending_quarantine, reference people.py:116, 335-346, 638; and the 64-bit "winning
# transmission" key that the edge pass atomicMin's into) is owned by the library, not bound from here.
device_fields = ()


def field_dtype(name):
    ''' NumPy dtype of a per-agent field (bool fields are stored as one byte) '''
    if name in states or name in by_variant_states:
        return np.bool_
    if name in person_int_fields or name in vacc_states or name == 't_nab_event':
        return default_int
    return default_float


def field_is_2d(name):
    ''' Fields shaped [n_variants, N] '''
    return name in by_variant_states or name in imm_states


FIELD_IDS = {name: i for i, name in enumerate(all_states + device_fields)}

# ---- results ------------------------------------------------------------------------------------
result_stocks = ('susceptible', 'exposed', 'infectious', 'symptomatic', 'severe', 'critical',
                 'recovered', 'dead', 'diagnosed', 'known_dead', 'quarantined', 'isolated', 'vaccinated')
result_stocks_by_variant = ('exposed_by_variant', 'infectious_by_variant')
result_flows = ('infections', 'reinfections', 'infectious', 'symptomatic', 'severe', 'critical',
                'recoveries', 'deaths', 'tests', 'diagnoses', 'known_deaths', 'quarantined',
                'isolated', 'doses', 'vaccinated')
result_flows_by_variant = ('infections_by_variant', 'symptomatic_by_variant', 'severe_by_variant',
                           'infectious_by_variant')
new_result_flows = tuple(f'new_{k}' for k in result_flows)
cum_result_flows = tuple(f'cum_{k}' for k in result_flows)
new_result_flows_by_variant = tuple(f'new_{k}' for k in result_flows_by_variant)
cum_result_flows_by_variant = tuple(f'cum_{k}' for k in result_flows_by_variant)
other_results = ('n_imports', 'n_alive', 'n_naive', 'n_preinfectious', 'n_removed', 'prevalence',
                 'incidence', 'r_eff', 'doubling_time', 'test_yield', 'rel_test_yield',
                 'frac_vaccinated', 'pop_nabs', 'pop_protection', 'pop_symp_protection')
# Results that are *not* multiplied by the rescale vector at finalize (reference sim.py:311-324)
unscaled_results = ('prevalence', 'incidence', 'r_eff', 'doubling_time', 'test_yield', 'rel_test_yield',
                    'frac_vaccinated', 'pop_nabs', 'pop_protection', 'pop_symp_protection')

# Column layout of the device-side per-day counter table (int64 [npts, N_COUNTERS]); the ids are the
# ``cvb_counter`` enum of include/covasim_b200.h.
counter_keys = new_result_flows + tuple(f'n_{k}' for k in result_stocks) + ('n_imports', 'n_alive_agents')
COUNTER_IDS = {k: i for i, k in enumerate(counter_keys)}
N_COUNTERS = len(counter_keys)
# By-variant counters are a second table int64 [npts, nv, N_VCOUNTERS]
vcounter_keys = new_result_flows_by_variant + tuple(f'n_{k}' for k in result_stocks_by_variant)
VCOUNTER_IDS = {k: i for i, k in enumerate(vcounter_keys)}
N_VCOUNTERS = len(vcounter_keys)

variant_par_keys = ('rel_beta', 'rel_symp_prob', 'rel_severe_prob', 'rel_crit_prob', 'rel_death_prob')

# Age distribution used when no location is given: (min age, max age, fraction); reference
# defaults.py:223-243 (Seattle 2018 census)
default_age_data = np.array([
    (0, 4, 0.0605), (5, 9, 0.0607), (10, 14, 0.0566), (15, 19, 0.0557), (20, 24, 0.0612),
    (25, 29, 0.0843), (30, 34, 0.0848), (35, 39, 0.0764), (40, 44, 0.0697), (45, 49, 0.0701),
    (50, 54, 0.0681), (55, 59, 0.0653), (60, 64, 0.0591), (65, 69, 0.0453), (70, 74, 0.0312),
    (75, 79, 0.02016), (80, 84, 0.01344), (85, 89, 0.01008), (90, 99, 0.00672)])
