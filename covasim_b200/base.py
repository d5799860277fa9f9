'''
Containers at the user-visible boundary: Result, Layer, Contacts (reference covasim/base.py:117-192,
1482-1876).  Layers hold their edge list as device tensors (p1:int32[E], p2:int32[E], beta:f32[E])
bound to the simulation handle; the dict-like access ``layer['p1']`` returns the tensor.
'''
import ctypes as C

import numpy as np
import torch

from . import defaults as cvd
from . import _capi

__all__ = ['Result', 'Layer', 'Contacts', 'AlreadyRunError']


class AlreadyRunError(RuntimeError):
    ''' Raised when a finished simulation is stepped or run again (reference sim.py:1550-1556) '''
    pass


class Result:
    ''' One result time series (reference base.py:117-192): ``values`` is float64[npts] (or [n_variants, npts]) '''

    def __init__(self, name=None, npts=None, scale=True, n_variants=0):
        self.name = name
        self.scale = scale
        npts = int(npts) if npts is not None else 0
        self.values = np.zeros((n_variants, npts) if n_variants > 0 else npts, dtype=cvd.result_float)
        self.low = None
        self.high = None

    def __repr__(self):
        return f'Result({self.name}, npts={self.npts})\n{self.values!r}'

    def __getitem__(self, key):
        return self.values[key]

    def __setitem__(self, key, value):
        self.values[key] = value

    def __len__(self):
        return len(self.values)

    @property
    def npts(self):
        return self.values.shape[-1]


class Layer:
    '''
    One contact layer: an undirected edge list stored once (reference base.py:1547-1876).  Self-loops and
    duplicate edges are allowed, exactly as in the reference.
    '''
    columns = ('p1', 'p2', 'beta')

    def __init__(self, p1=None, p2=None, beta=None, label=None, device=None):
        self.label = label
        self.device = torch.device(device) if device is not None else None
        p1 = np.zeros(0, dtype=np.int32) if p1 is None else p1
        p2 = np.zeros(0, dtype=np.int32) if p2 is None else p2
        if beta is None:
            beta = np.ones(len(p1), dtype=np.float32)
        self._cols = {}
        self._set('p1', p1)
        self._set('p2', p2)
        self._set('beta', beta)
        self.validate()
        self._sim = None
        self._index = None

    def _set(self, key, value):
        dtype = torch.float32 if key == 'beta' else torch.int32
        if isinstance(value, torch.Tensor):
            t = value.to(dtype=dtype)
        else:
            t = torch.as_tensor(np.ascontiguousarray(value), dtype=dtype)
        if self.device is not None:
            t = t.to(self.device)
        self._cols[key] = t.contiguous()

    def to(self, device):
        self.device = torch.device(device)
        for k in self.columns:
            self._cols[k] = self._cols[k].to(self.device).contiguous()
        self._rebind()
        return self

    def validate(self):
        n = len(self._cols['p1'])
        for k in self.columns:
            if len(self._cols[k]) != n:
                raise ValueError(f'Layer column {k} has length {len(self._cols[k])}, expected {n}')

    def __len__(self):
        return int(self._cols['p1'].shape[0])

    def keys(self):
        return list(self.columns)

    def _ready(self):
        ''' Sim.restore copies the edge lists on a side stream: anything that touches the arrays waits for that copy first '''
        if self._sim is not None:
            self._sim._sync_edges()

    def __getitem__(self, key):
        self._ready()
        return self._cols[key]

    def __setitem__(self, key, value):
        if key not in self.columns:
            raise KeyError(key)
        self._ready()
        old = self._cols[key]
        if len(value) == len(old):
            old.copy_(torch.as_tensor(value).to(device=old.device, dtype=old.dtype))      # in place: the bound pointer stays valid
            if self._sim is not None:
                self._sim._adj_dirty = True
        else:
            self._set(key, value)
            self._rebind()

    def _bind(self, sim, index):
        self._sim, self._index = sim, index
        self._rebind()

    def _rebind(self):
        if self._sim is not None:
            self._sim._adj_dirty = True             # the adjacency index no longer matches this edge list
        if self._sim is not None and self._sim._handle is not None:
            c = self._cols
            _capi.call('cvb_bind_layer', self._sim._handle, self._index, c['p1'].data_ptr(), c['p2'].data_ptr(), c['beta'].data_ptr(), len(self))

    def to_numpy(self):
        self._ready()
        return {k: self._cols[k].cpu().numpy() for k in self.columns}

    def find_contacts(self, inds, as_array=True):
        ''' Sorted unique partners of ``inds`` over both columns (reference base.py:1808-1846, utils.py:131-147) '''
        if self._sim is None:
            raise RuntimeError('Layer.find_contacts needs a layer attached to a sim')
        sim = self._sim
        inds = torch.as_tensor(inds, dtype=torch.int64, device=self.device).contiguous()
        out = torch.empty(sim.n_local, dtype=torch.int32, device=self.device)
        n_out = C.c_int64(0)
        _capi.call('cvb_find_contacts', sim._handle, self['p1'].data_ptr(), self['p2'].data_ptr(), len(self), inds.data_ptr(), len(inds),
                   out.data_ptr(), C.byref(n_out), sim._stream_ptr)
        return out[:n_out.value]

    def update(self, people, frac=1.0):
        '''
        Regenerate a dynamic layer (reference base.py:1849-1876): ``round(E * frac)`` edges, chosen without replacement, get two fresh
        uniformly random endpoints and weight 1.  frac = 1 is one device pass over the layer; frac < 1 picks the edges on the host
        (the reference's cvu.choose from the Numba stream; the O(k) sampler in native-RNG mode) and regenerates those.
        '''
        sim = self._sim
        self._ready()
        if frac == 1.0:
            _capi.call('cvb_layer_regenerate', sim._handle, self._index, sim.t, sim._stream_ptr)
            return
        n_new = int(np.round(len(self) * frac))
        if n_new <= 0:
            return
        if sim.rng_mode == 'mt':
            raise NotImplementedError('partial regeneration (frac < 1) is not built for replay mode')
        from . import utils as cvu
        inds = torch.as_tensor(cvu.choose_distinct(sim.rng.nb, len(self), n_new), dtype=torch.int64, device=self.device).contiguous()
        _capi.call('cvb_layer_regenerate_list', sim._handle, self._index, sim.t, inds.data_ptr(), n_new, sim._stream_ptr)
        sim._adj_dirty = sim._adj_dirty or bool((sim._adj_mask >> self._index) & 1)

    def pop_inds(self, inds):
        ''' Remove edges by index and return them (reference base.py:1742-1757) -- used by clip_edges-style interventions '''
        self._ready()
        inds = torch.as_tensor(inds, dtype=torch.int64, device=self.device)
        keep = torch.ones(len(self), dtype=torch.bool, device=self.device)
        keep[inds] = False
        popped = {k: self._cols[k][inds].clone() for k in self.columns}
        for k in self.columns:
            self._cols[k] = self._cols[k][keep].contiguous()
        self._rebind()
        return popped

    def append(self, contacts):
        ''' Append edges (reference base.py:1760-1771) '''
        self._ready()
        for k in self.columns:
            new = torch.as_tensor(contacts[k], dtype=self._cols[k].dtype, device=self.device)
            self._cols[k] = torch.cat([self._cols[k], new]).contiguous()
        self.validate()
        self._rebind()


class Contacts(dict):
    ''' Ordered mapping layer key -> Layer (reference base.py:1509-1544); iteration order is transmission order '''

    def __init__(self, data=None, layer_keys=None, **kwargs):
        super().__init__()
        if layer_keys is not None:
            for lk in layer_keys:
                self[lk] = Layer(label=lk)
        if data:
            for lk, l in data.items():
                self[lk] = l
        for lk, l in kwargs.items():
            self[lk] = l

    def add_layer(self, **kwargs):
        for lk, l in kwargs.items():
            self[lk] = l

    def pop_layer(self, *args):
        for lk in args:
            self.pop(lk)

    def __len__(self):
        return sum(len(l) for l in self.values())

    def n_layers(self):
        return dict.__len__(self)
