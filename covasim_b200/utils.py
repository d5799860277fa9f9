'''
Index helpers and seeding: the part of the reference's cv.utils that interventions and analyzers call
on People arrays (reference covasim/utils.py:487-669), here for device tensors.  They return int64
index tensors on the array's device, like the reference's ``arr.nonzero()[0]``.
'''
import random

import numpy as np
import torch

__all__ = ['bind_to_device_numa', 'choose_distinct', 'true', 'false', 'defined', 'undefined', 'itrue', 'ifalse', 'idefined', 'iundefined',
           'itruei', 'ifalsei', 'idefinedi', 'iundefinedi', 'set_seed', 'HostStreams', 'n_binomial', 'binomial_arr']


def _t(arr):
    return arr if isinstance(arr, torch.Tensor) else torch.as_tensor(np.asarray(arr))


def true(arr):
    ''' Indices of non-zero entries (reference utils.py:494-506) '''
    return torch.nonzero(_t(arr)).flatten()


def false(arr):
    ''' Indices of zero entries (reference utils.py:509-520) '''
    a = _t(arr)
    return torch.nonzero(a == 0).flatten()


def defined(arr):
    ''' Indices of non-NaN entries (reference utils.py:523-534) '''
    return torch.nonzero(~torch.isnan(_t(arr))).flatten()


def undefined(arr):
    ''' Indices of NaN entries (reference utils.py:537-548) '''
    return torch.nonzero(torch.isnan(_t(arr))).flatten()


def itrue(arr, inds):
    ''' inds[arr]: arr is a boolean array the same length as inds (reference utils.py:551-563) '''
    return _t(inds)[_t(arr).bool()]


def ifalse(arr, inds):
    return _t(inds)[~_t(arr).bool()]


def idefined(arr, inds):
    return _t(inds)[~torch.isnan(_t(arr))]


def iundefined(arr, inds):
    return _t(inds)[torch.isnan(_t(arr))]


def itruei(arr, inds):
    ''' inds[arr[inds]]: arr is a full-length array (reference utils.py:611-623) '''
    inds = _t(inds)
    return inds[_t(arr)[inds].bool()]


def ifalsei(arr, inds):
    inds = _t(inds)
    return inds[~_t(arr)[inds].bool()]


def idefinedi(arr, inds):
    inds = _t(inds)
    return inds[~torch.isnan(_t(arr)[inds])]


def iundefinedi(arr, inds):
    inds = _t(inds)
    return inds[torch.isnan(_t(arr)[inds])]


class HostStreams:
    '''
    The reference's two MT19937 streams (reference utils.py:271-298): ``np_`` plays the role of NumPy's
    global stream and ``nb`` of Numba's.  Same algorithm, same seed, independent state -- verified in
    lockstep against the reference by oracle/gen_golden.py.  In native-RNG mode only rare host-side set
    choices (seed infections, importations) use them; everything per-agent / per-edge is Philox on the device.
    '''

    def __init__(self, seed=None):
        self.np_ = np.random.RandomState()
        self.nb = np.random.RandomState()
        self.seed = None
        if seed is not None:
            self.set_seed(seed)

    def set_seed(self, seed=None):
        if seed is None:
            self.np_.seed()
            seed = int(self.np_.randint(int(1e9)))
        self.seed = int(seed)
        self.np_.seed(self.seed)
        self.nb.seed(self.seed)
        random.seed(self.seed)


def choose_distinct(stream, n, k):
    '''
    k distinct integers in [0, n) in O(k) for k << n: uniform draws from ``stream`` (a RandomState), first occurrences kept
    in draw order, repeated until k are found.  Native-RNG replacement for the reference's choice(n, k, replace=False)
    (utils.py:429-443 choose), which permutes all n agents -- 30 ms per call at 2M agents, on every importation day.
    Replay mode keeps the reference's call.  oracle/cvoracle.py:choose_distinct is the same function.
    '''
    n, k = int(n), int(k)
    if k > n // 8:
        return stream.choice(n, k, replace=False)
    out = np.zeros(0, dtype=np.int64)
    while len(out) < k:
        out = np.concatenate([out, stream.randint(0, n, size=int(1.2 * (k - len(out))) + 8)])
        _, first = np.unique(out, return_index=True)
        out = out[np.sort(first)]
    return out[:k]


def bind_to_device_numa(device):
    '''
    Run this process on the CPU cores next to ``device`` (the GPU's NUMA node, read from sysfs): pinned host buffers allocated
    afterwards are first-touched there, so that several ranks restoring their simulations at once (Sim.restore: ~400 MB of
    host-to-device copies per rank at 1M agents) do not all pull from one socket's memory.  Returns the CPU list, or None when the
    topology cannot be read or the affinity cannot be changed (containers without the sysfs entries, restricted cpusets).
    '''
    import os
    try:
        props = torch.cuda.get_device_properties(device)
        addr = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{addr}/local_cpulist') as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or cpus
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def set_seed(seed=None):
    ''' Seed NumPy's global stream and Python's ``random`` (reference utils.py:271-298); sims keep their own HostStreams '''
    if seed is not None:
        seed = int(seed)
    np.random.seed(seed)
    random.seed(seed)


def n_binomial(prob, n):
    return np.random.random(n) < prob


def binomial_arr(prob_arr):
    return np.random.random(len(prob_arr)) < prob_arr
