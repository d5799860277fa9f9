'''
covasim_b200 -- B200-native implementation of Covasim's per-timestep simulation hot path.

Usage mirrors the reference (``import covasim_b200 as cv``): ``cv.Sim(pop_size=..., pop_type='hybrid',
interventions=[cv.test_prob(...), cv.contact_tracing(...)]).run()``; People and contact layers live on the
GPU and every simulated day runs as hand-written sm_100a CUDA kernels behind a C ABI
(include/covasim_b200.h, covasim_b200/libcovasim_b200.so).  There is no CPU fallback.
'''
from .version import __version__  # noqa: F401
from . import defaults  # noqa: F401
from . import parameters  # noqa: F401
from .parameters import make_pars, get_prognoses  # noqa: F401
from . import _capi  # noqa: F401  (loads libcovasim_b200.so; raises if it has not been built)
from ._capi import CvbError  # noqa: F401
from . import utils  # noqa: F401
from .utils import *  # noqa: F401,F403
from .base import Result, Layer, Contacts, AlreadyRunError  # noqa: F401
from .devarray import DeviceArray  # noqa: F401
from .people import People  # noqa: F401
from .immunity import variant, calc_VE, calc_VE_symp, precompute_waning  # noqa: F401
from .interventions import Intervention, dynamic_pars, sequence, change_beta, clip_edges, test_num, test_prob, contact_tracing, vaccinate_prob, vaccinate_num, vaccinate  # noqa: F401
from .sim import Sim  # noqa: F401
from .run import MultiSim, multi_run  # noqa: F401
from .analysis import Analyzer, snapshot, age_histogram, compute_gof, Fit, fit_members, TransTree  # noqa: F401
from . import ops  # noqa: F401
