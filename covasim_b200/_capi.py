'''
ctypes binding of libcovasim_b200.so (the C ABI declared in include/covasim_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every entry point
raises ``CvbError`` with the library's message when a call fails (including "no CUDA device").
'''
import ctypes as C
import os

from . import defaults as cvd

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcovasim_b200.so')

MAX_VARIANTS = 8
MAX_LAYERS = 8
MAX_VACCINES = 8
N_DURS = 9
DUR_ORDER = ('exp2inf', 'inf2sym', 'sym2sev', 'sev2crit', 'asym2rec', 'mild2rec', 'sev2rec', 'crit2rec', 'crit2die')
LAYER_SEED, LAYER_IMPORT = -1, -2
TIMED_KINDS = ('day_begin', 'trace', 'day_mid', 'edge_pass', 'infect', 'day_end', 'regen', 'vaccinate')    # enum cvb_timed
DIST_KINDS = dict(zero=0, normal=1, normal_pos=2, normal_int=3, lognormal=4, lognormal_int=5)


class CvbError(RuntimeError):
    pass


class cvb_dist(C.Structure):
    _fields_ = [('kind', C.c_int32), ('pad_', C.c_int32), ('a', C.c_double), ('b', C.c_double)]


class cvb_pars(C.Structure):
    _fields_ = [
        ('n_variants', C.c_int32), ('n_layers', C.c_int32), ('use_waning', C.c_int32), ('n_vaccines', C.c_int32),
        ('quar_period', C.c_int32), ('has_vaccine_pars', C.c_int32),
        ('n_beds_hosp', C.c_int64), ('n_beds_icu', C.c_int64),
        ('asymp_factor', C.c_float), ('frac_time', C.c_float), ('load_ratio', C.c_float), ('high_cap', C.c_float),
        ('trans_redux', C.c_float), ('no_hosp_factor', C.c_float), ('no_icu_factor', C.c_float), ('nab_boost', C.c_float),
        ('beta', C.c_float * MAX_VARIANTS),
        ('rel_symp', C.c_float * MAX_VARIANTS), ('rel_severe', C.c_float * MAX_VARIANTS),
        ('rel_crit', C.c_float * MAX_VARIANTS), ('rel_death', C.c_float * MAX_VARIANTS),
        ('beta_layer', C.c_float * MAX_LAYERS), ('iso_factor', C.c_float * MAX_LAYERS), ('quar_factor', C.c_float * MAX_LAYERS),
        ('immunity', (C.c_float * MAX_VARIANTS) * MAX_VARIANTS),
        ('vaccine_imm', (C.c_double * MAX_VARIANTS) * MAX_VACCINES),
        ('exp_alpha_inf', C.c_double), ('beta_inf', C.c_double), ('exp_alpha_symp_inf', C.c_double), ('beta_symp_inf', C.c_double),
        ('exp_alpha_sev_symp', C.c_double), ('beta_sev_symp', C.c_double),
        ('rel_imm_asymp', C.c_double), ('rel_imm_mild', C.c_double), ('rel_imm_severe', C.c_double), ('nab_norm', C.c_double),
        ('dur', cvb_dist * N_DURS),
        ('nab_init', cvb_dist),
    ]


class cvb_test_prob_pars(C.Structure):
    _fields_ = [('symp_prob', C.c_double), ('asymp_prob', C.c_double), ('symp_quar_prob', C.c_double), ('asymp_quar_prob', C.c_double),
                ('sensitivity', C.c_double), ('loss_prob', C.c_double), ('quar_policy', C.c_int32), ('test_delay', C.c_int32),
                ('index', C.c_int32), ('pad_', C.c_int32)]


class cvb_test_num_pars(C.Structure):
    _fields_ = [('symp_test', C.c_double), ('quar_test', C.c_double), ('quar_policy', C.c_int32), ('index', C.c_int32)]


class cvb_trace_pars(C.Structure):
    _fields_ = [('trace_prob', C.c_double * MAX_LAYERS), ('trace_time', C.c_int32 * MAX_LAYERS), ('presumptive', C.c_int32),
                ('quar_period', C.c_int32), ('index', C.c_int32), ('pad_', C.c_int32)]


class cvb_vaccinate_pars(C.Structure):
    _fields_ = [('prob', C.c_double), ('nab_init', cvb_dist), ('nab_boost', C.c_float), ('booster', C.c_int32),
                ('vaccine_index', C.c_int32), ('max_doses', C.c_int32), ('index', C.c_int32), ('first_dose_today', C.c_int32),
                ('second_dose_today', C.c_int32), ('interval', C.c_int32), ('n_days', C.c_int32),
                ('nab_boost_f64', C.c_double), ('nab_boost_is_f64', C.c_int32), ('pad_', C.c_int32)]


_P = C.c_void_p
_i32, _i64, _u64, _f32 = C.c_int32, C.c_int64, C.c_uint64, C.c_float

# name -> argument types (all return int, except the two noted below); must list every symbol of the header
PROTOTYPES = dict(
    cvb_struct_sizes=[C.POINTER(_i64)],
    cvb_create=[C.POINTER(_P), _i64, _i32, _i32, _u64],
    cvb_destroy=[_P],
    cvb_set_seed=[_P, _u64],
    cvb_reset=[_P, _P],
    cvb_get_edge_work=[_P, _P],
    cvb_set_pars=[_P, C.POINTER(cvb_pars)],
    cvb_set_nab_kin=[_P, _P, _i64],
    cvb_set_quar_horizon=[_P, _i32],
    cvb_clone_scratch=[_P, _P, _P],
    cvb_bind_field=[_P, _i32, _P],
    cvb_bind_layer=[_P, _i32, _P, _P, _P, _i64],
    cvb_bind_adjacency=[_P, _P, _P, _i64, C.c_uint32],
    cvb_set_partition=[_P, _i64, _i64, _i64, _i32, _P, _P, _P, _P, _P, _i64],
    cvb_restore_compact=[_P, _i64, _P, _i32, _i64, _P],
    cvb_peer_push=[_P, _i64, _P, _i32, _i64, _P],
    cvb_set_exchange_buffers=[_P, _P, _P],
    cvb_bind_partition_adjacency=[_P, _P, _P, _i64, C.c_uint32],
    cvb_partition_status=[_P, _P],
    cvb_bind_results=[_P, _P, _P, _P],
    cvb_bind_beds=[_P, _P],
    cvb_bind_log=[_P, _P, _P, _P, _P, _P, _i64, _P],
    cvb_keyed_uniform=[_u64, C.c_uint32, C.c_uint32, _i32, _i64, _i64, C.c_uint32, _P, _P],
    cvb_compute_viral_load=[_i32, _P, _P, _P, _f32, _f32, _f32, _P, _i64, _P],
    cvb_compute_trans_sus=[_P, _P, _P, _P, _f32, _P, _P, _P, _P, _f32, _f32, _f32, _P, _P, _P, _i64, _P],
    cvb_infections_count=[_P, _f32, _P, _P, _P, _i64, _P, _P, C.POINTER(_i64), _P],
    cvb_infections_draw=[_P, _f32, _P, _P, _P, _i64, _P, _P, _P, _P, _P, C.POINTER(_i64), _P],
    cvb_find_contacts=[_P, _P, _P, _i64, _P, _i64, _P, C.POINTER(_i64), _P],
    cvb_true_indices=[_P, _P, _i64, _P, C.POINTER(_i64), _P],
    cvb_update_states_pre=[_P, _i32, _P],
    cvb_schedule_quarantine=[_P, _P, _i64, _i32, _f32, _P],
    cvb_update_states_post=[_P, _i32, _P],
    cvb_prepare_transmission=[_P, _i32, _P],
    cvb_post_and_prepare=[_P, _i32, _P],
    cvb_edge_pass=[_P, _i32, _P],
    cvb_infect_winners=[_P, _i32, _P],
    cvb_infect_list=[_P, _P, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _P],
    cvb_infect_list_taped=[_P, _P, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _P, _P],
    cvb_update_nab_count=[_P, _i32, _P],
    cvb_step_day=[_P, _i32, _P],
    cvb_test_prob=[_P, _i32, C.POINTER(cvb_test_prob_pars), _P, _P],
    cvb_test_prob_taped=[_P, _i32, C.POINTER(cvb_test_prob_pars), _P, _P, _P],
    cvb_test_num_keys=[_P, _i32, C.POINTER(cvb_test_num_pars), _P, _P, _P],
    cvb_test_list=[_P, _i32, _P, _i64, C.c_double, C.c_double, _i32, _i32, _P],
    cvb_contact_tracing=[_P, _i32, C.POINTER(cvb_trace_pars), _P],
    cvb_contact_tracing_list=[_P, _i32, C.POINTER(cvb_trace_pars), _P, _i64, _P],
    cvb_contact_tracing_taped=[_P, _i32, C.POINTER(cvb_trace_pars), _P, _P],
    cvb_pending_quarantine=[_P, _i32, _P, _P],
    cvb_set_pending_quarantine=[_P, _i32, _P, _P],
    cvb_trace_select_cases=[_P, _i32, C.POINTER(cvb_trace_pars), _P],
    cvb_trace_notify_contacts=[_P, _i32, C.POINTER(cvb_trace_pars), _P],
    cvb_vaccinate_prob=[_P, _i32, C.POINTER(cvb_vaccinate_pars), _P, _P, _P, _P],
    cvb_vaccinate_taped=[_P, _i32, C.POINTER(cvb_vaccinate_pars), _P, _P, _P, _P, _P],
    cvb_layer_regenerate=[_P, _i32, _i32, _P],
    cvb_layer_regenerate_list=[_P, _i32, _i32, _P, _i64, _P],
    cvb_plan_clear=[_P],
    cvb_plan_test_prob=[_P, C.POINTER(cvb_test_prob_pars), _i32, _i32],
    cvb_plan_contact_tracing=[_P, C.POINTER(cvb_trace_pars), _i32, _i32],
    cvb_plan_dynamic_layers=[_P, C.c_uint32],
    cvb_plan_vaccinate=[_P, C.POINTER(cvb_vaccinate_pars), _P, _P, _P],
    cvb_run_days=[_P, _i32, _i32, _P],
    cvb_run_days_multi=[_P, _i32, _i32, _i32, _P],
    cvb_fused_phase=[_P, _i32, _i32, _i32, _P],
    cvb_state_invalidate=[_P],
    cvb_tune=[_P, _i32, _i32],
    cvb_timing_enable=[_P, _i32],
    cvb_timing_read=[_P, _P, _P],
    cvb_state_check=[_P, _i32, _P, _P],
)
OTHER_SYMBOLS = ('cvb_last_error', 'cvb_abi_version', 'cvb_launch_count')
ABI_VERSION = 2          # CVB_ABI_VERSION of include/covasim_b200.h


def load_library(path=LIB_PATH):
    # (Re)build when the sources changed and a compiler is here; otherwise use the in-tree library as is.
    try:
        from . import build as _build
        import shutil
        if shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc'):
            _build.build()
    except Exception:
        if not os.path.exists(path):
            raise
    if not os.path.exists(path):
        raise CvbError(f'{path} not found: build it with `python -m covasim_b200.build` (needs nvcc). '
                       'covasim_b200 has no CPU fallback.')
    lib = C.CDLL(path)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.cvb_last_error.restype = C.c_char_p
    lib.cvb_last_error.argtypes = []
    lib.cvb_abi_version.restype = C.c_int32
    lib.cvb_abi_version.argtypes = []
    lib.cvb_launch_count.restype = C.c_int64
    lib.cvb_launch_count.argtypes = []
    if lib.cvb_abi_version() != ABI_VERSION:
        raise CvbError(f'{path} has ABI version {lib.cvb_abi_version()}, this binding was written for {ABI_VERSION} (include/covasim_b200.h): rebuild the library')
    return lib


lib = load_library()


def check(rc):
    if rc != 0:
        raise CvbError(lib.cvb_last_error().decode('utf-8', 'replace') or f'libcovasim_b200 call failed with code {rc}')


def call(name, *args):
    check(getattr(lib, name)(*args))


def dist_struct(spec, lognormal_pars=None):
    ''' Convert a {dist, par1, par2} dict into a cvb_dist (lognormal kinds carry the underlying normal's mean/sigma) '''
    import numpy as np
    d = cvb_dist()
    if spec is None:
        d.kind = DIST_KINDS['zero']
        return d
    kind, par1, par2 = spec['dist'], spec['par1'], spec['par2']
    kind = {'norm': 'normal', 'lognorm': 'lognormal', 'lognorm_int': 'lognormal_int'}.get(kind, kind)
    if kind not in DIST_KINDS:
        raise NotImplementedError(f'distribution "{kind}" is not available on the device path')
    if kind.startswith('lognormal'):
        if par1 > 0:
            d.kind = DIST_KINDS[kind]
            d.a = float(np.log(par1 ** 2 / np.sqrt(par2 ** 2 + par1 ** 2)))      # reference utils.py:223-224
            d.b = float(np.sqrt(np.log(par2 ** 2 / par1 ** 2 + 1)))
        else:
            d.kind = DIST_KINDS['zero']
    else:
        d.kind, d.a, d.b = DIST_KINDS[kind], float(par1), float(par2)
    return d
