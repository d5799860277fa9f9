''' Stand-in for pylab (matplotlib is not installed) -- TEST INFRASTRUCTURE ONLY. '''
import numpy as np
from numpy import isfinite, arange, array, zeros, ones, nan  # noqa
rcParams = {'figure.dpi': 100, 'font.family': 'sans-serif', 'font.size': 10, 'backend': 'agg'}
def get_backend():
    return 'agg'
def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)
    if hasattr(np, name):
        return getattr(np, name)
    def _noop(*a, **k):
        return None
    return _noop
