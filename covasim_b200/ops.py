'''
Drop-in device versions of the reference's four Numba kernels (covasim/utils.py:39-147), with the same
names and argument order; arguments are CUDA tensors (or anything ``torch.as_tensor`` accepts) and the
results are CUDA tensors.  These are thin wrappers over the stateless entry points of the C ABI.
'''
import ctypes as C

import numpy as np
import torch

from . import _capi

__all__ = ['compute_viral_load', 'compute_trans_sus', 'compute_infections', 'find_contacts', 'Workspace']


def _dev(x, dtype, device):
    if not torch.cuda.is_available():
        raise _capi.CvbError('covasim_b200 needs a CUDA device: there is no CPU fallback')
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    if t.dtype == torch.bool and dtype == torch.uint8:
        t = t.to(device).contiguous().view(torch.uint8)
        return t
    return t.to(device=device, dtype=dtype).contiguous()


class Workspace:
    ''' Scratch for the compaction-based operators (a cvb_sim handle sized for n_agents) '''

    def __init__(self, n_agents, device='cuda'):
        if not torch.cuda.is_available():
            raise _capi.CvbError('covasim_b200 needs a CUDA device: there is no CPU fallback')
        self.device = torch.device(device)
        self.n = int(n_agents)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _capi.call('cvb_create', C.byref(self.handle), self.n, 1, 1, 0)

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            _capi.lib.cvb_destroy(h)


def compute_viral_load(t, time_start, time_recovered, time_dead, frac_time, load_ratio, high_cap, device='cuda'):
    ''' reference utils.py:39-79 '''
    ts, tr, td = (_dev(a, torch.float32, device) for a in (time_start, time_recovered, time_dead))
    out = torch.empty_like(ts)
    _capi.call('cvb_compute_viral_load', int(t), ts.data_ptr(), tr.data_ptr(), td.data_ptr(), float(frac_time), float(load_ratio),
               float(high_cap), out.data_ptr(), ts.numel(), None)
    return out


def compute_trans_sus(rel_trans, rel_sus, inf, sus, beta_layer, viral_load, symp, iso, quar, asymp_factor, iso_factor,
                      quar_factor, immunity_factors, device='cuda'):
    ''' reference utils.py:82-90 '''
    f = lambda a: _dev(a, torch.float32, device)
    b = lambda a: _dev(a, torch.uint8, device)
    rel_trans, rel_sus, viral_load, immunity_factors = f(rel_trans), f(rel_sus), f(viral_load), f(immunity_factors)
    inf, sus, symp, iso, quar = b(inf), b(sus), b(symp), b(iso), b(quar)
    ot, os_ = torch.empty_like(rel_trans), torch.empty_like(rel_sus)
    _capi.call('cvb_compute_trans_sus', rel_trans.data_ptr(), rel_sus.data_ptr(), inf.data_ptr(), sus.data_ptr(), float(beta_layer),
               viral_load.data_ptr(), symp.data_ptr(), iso.data_ptr(), quar.data_ptr(), float(asymp_factor), float(iso_factor),
               float(quar_factor), immunity_factors.data_ptr(), ot.data_ptr(), os_.data_ptr(), rel_trans.numel(), None)
    return ot, os_


def compute_infections(beta, p1, p2, layer_betas, rel_trans, rel_sus, draw, workspace=None, device='cuda'):
    '''
    reference utils.py:93-128, replay form.  ``draw(n)`` must return the next ``n`` float64 uniforms of the
    stream the reference would consume (it is called once, with the total for both directions, direction
    p1->p2 first).  Returns the ordered (source, target) int32 tensors.
    '''
    p1, p2 = _dev(p1, torch.int32, device), _dev(p2, torch.int32, device)
    lb, rt, rs = _dev(layer_betas, torch.float32, device), _dev(rel_trans, torch.float32, device), _dev(rel_sus, torch.float32, device)
    ws = workspace or Workspace(rt.numel(), device)
    n_draws = (C.c_int64 * 2)()
    _capi.call('cvb_infections_count', ws.handle, float(beta), p1.data_ptr(), p2.data_ptr(), lb.data_ptr(), p1.numel(), rt.data_ptr(),
               rs.data_ptr(), n_draws, None)
    total = int(n_draws[0] + n_draws[1])
    u = torch.as_tensor(np.ascontiguousarray(draw(total), dtype=np.float64)).to(device)
    src = torch.empty(max(total, 1), dtype=torch.int32, device=device)
    tgt = torch.empty(max(total, 1), dtype=torch.int32, device=device)
    n_out = C.c_int64(0)
    _capi.call('cvb_infections_draw', ws.handle, float(beta), p1.data_ptr(), p2.data_ptr(), lb.data_ptr(), p1.numel(), rt.data_ptr(),
               rs.data_ptr(), u.data_ptr(), src.data_ptr(), tgt.data_ptr(), C.byref(n_out), None)
    return src[:n_out.value], tgt[:n_out.value], (int(n_draws[0]), int(n_draws[1]))


def find_contacts(p1, p2, inds, n_agents=None, workspace=None, device='cuda'):
    ''' reference utils.py:131-147 + base.py:1842-1844: sorted unique partners of ``inds`` '''
    p1, p2 = _dev(p1, torch.int32, device), _dev(p2, torch.int32, device)
    inds = _dev(inds, torch.int64, device)
    if n_agents is None:
        n_agents = int(max(int(p1.max()) if p1.numel() else -1, int(p2.max()) if p2.numel() else -1, int(inds.max()) if inds.numel() else -1)) + 1
    ws = workspace or Workspace(max(n_agents, 1), device)
    out = torch.empty(max(ws.n, 1), dtype=torch.int32, device=device)
    n_out = C.c_int64(0)
    _capi.call('cvb_find_contacts', ws.handle, p1.data_ptr(), p2.data_ptr(), p1.numel(), inds.data_ptr(), inds.numel(), out.data_ptr(),
               C.byref(n_out), None)
    return out[:n_out.value]
