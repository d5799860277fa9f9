'''
TEST INFRASTRUCTURE -- makes the unmodified reference importable in the BUILD CONTAINER ONLY.

/root/reference does not exist on the GPU box; nothing that runs there imports this module.  It puts
the sciris/pylab/matplotlib stand-ins (oracle/shim) and /root/reference on sys.path and points the
Numba cache at a writable directory (the reference tree is read-only).
'''
import os
import sys

REFERENCE = '/root/reference'
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'covasim'))


def import_reference():
    if not available():
        raise RuntimeError('the reference tree is not present on this machine')
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache_ref')
    for p in (REFERENCE, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import covasim as cv
    return cv
