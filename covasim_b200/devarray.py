'''
DeviceArray: the People arrays as the reference's plug-ins expect them ("array-API hazard").

``sim.people.<field>`` is device memory the kernels are bound to, handed out as a ``torch.Tensor`` subclass that also
answers to the NumPy idioms the reference's interventions, analyzers, examples and tests use on People arrays
(covasim/utils.py:487-669, examples/t05_custom_intervention.py, t08_boosters.py:52, tests/test_immunity.py:137, 263, 283):

    arr.copy()                      a new array (on the device)
    arr.nonzero()[0]                NumPy's tuple-of-index-arrays form (torch's own nonzero() is a 2-D table)
    np.isfinite(arr), np.isnan(arr), np.logical_and(a, b), arr1 + np.float32(1) ...   NumPy ufuncs: evaluated on the
                                    device when torch has the same function, else on a host copy
    np.asarray(arr), np.sum(arr), pd.DataFrame(arr) ...   ``__array__``: a host copy (a plain CUDA tensor refuses this)
    arr.astype(bool), arr.tolist(), arr.mean(), arr.sum(), arr[inds] = value, arr[mask]

Everything else is torch: the array is zero-copy device memory (``__dlpack__``, ``data_ptr()``), and operations on it return
DeviceArrays.  Writing through it (``arr[inds] = x``) writes the memory the kernels read.
'''
import numpy as np
import torch

__all__ = ['DeviceArray', 'wrap']

_NP_TO_TORCH_DTYPE = {np.dtype(np.bool_): torch.bool, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                      np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}

# NumPy ufunc -> the torch function with the same meaning (evaluated on the device)
_UFUNCS = {
    np.isfinite: torch.isfinite, np.isnan: torch.isnan, np.isinf: torch.isinf, np.logical_not: torch.logical_not,
    np.logical_and: torch.logical_and, np.logical_or: torch.logical_or, np.logical_xor: torch.logical_xor,
    np.add: torch.add, np.subtract: torch.subtract, np.multiply: torch.multiply, np.true_divide: torch.true_divide,
    np.maximum: torch.maximum, np.minimum: torch.minimum, np.fmax: torch.fmax, np.fmin: torch.fmin,
    np.equal: torch.eq, np.not_equal: torch.ne, np.less: torch.lt, np.less_equal: torch.le, np.greater: torch.gt,
    np.greater_equal: torch.ge, np.absolute: torch.abs, np.negative: torch.neg, np.floor: torch.floor, np.ceil: torch.ceil,
    np.sqrt: torch.sqrt, np.exp: torch.exp, np.log: torch.log, np.invert: torch.bitwise_not, np.bitwise_and: torch.bitwise_and,
    np.bitwise_or: torch.bitwise_or,
}


class DeviceArray(torch.Tensor):

    @staticmethod
    def __new__(cls, data):
        t = data if isinstance(data, torch.Tensor) else torch.as_tensor(data)
        return t.as_subclass(cls)

    # ---- NumPy protocol ----------------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        out = self.detach().as_subclass(torch.Tensor).cpu().numpy()
        return out if dtype is None else out.astype(dtype, copy=False)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        fn = _UFUNCS.get(ufunc)
        if method == '__call__' and fn is not None and not kwargs:
            dev = self.device
            args = []
            for x in inputs:
                if isinstance(x, torch.Tensor):
                    args.append(x.as_subclass(torch.Tensor))
                elif isinstance(x, np.ndarray):
                    args.append(torch.as_tensor(x, device=dev))
                else:
                    args.append(x)
            if not any(isinstance(a, torch.Tensor) for a in args[1:]) and len(args) == 2 and not isinstance(args[1], torch.Tensor):
                args[1] = torch.as_tensor(args[1], device=dev)            # torch.maximum & co. want tensors on both sides
            return fn(*args).as_subclass(DeviceArray)
        host = [np.asarray(x) if isinstance(x, torch.Tensor) else x for x in inputs]      # anything else: on a host copy
        return getattr(ufunc, method)(*host, **kwargs)

    # ---- ndarray methods that differ from (or are missing in) torch.Tensor ------------------------------------
    def copy(self):
        return self.clone()

    def nonzero(self, *args, **kwargs):
        if args or kwargs:
            return super().nonzero(*args, **kwargs)
        return torch.nonzero(self.as_subclass(torch.Tensor), as_tuple=True)

    def astype(self, dtype, copy=True):
        if not isinstance(dtype, torch.dtype):
            dtype = _NP_TO_TORCH_DTYPE[np.dtype(dtype)]
        out = self.to(dtype)
        return out.clone() if (copy and out.data_ptr() == self.data_ptr()) else out

    def get(self):
        ''' A host (NumPy) copy '''
        return self.__array__()

    # reductions: NumPy calls them as arr.sum(axis=None, dtype=None, out=None), torch as t.sum(dim, keepdim) -- accept both
    def _reduce(self, name, args, kwargs):
        t = self.as_subclass(torch.Tensor)
        kwargs = {k: v for k, v in kwargs.items() if not (k in ('out', 'where', 'initial') and v is None)}
        axis = kwargs.pop('axis', kwargs.pop('dim', args[0] if args else None))
        keep = kwargs.pop('keepdims', kwargs.pop('keepdim', args[1] if len(args) > 1 else False))
        dtype = kwargs.pop('dtype', None)
        if kwargs:
            raise TypeError(f'{name}() got unexpected arguments {sorted(kwargs)}')
        if dtype is not None and not isinstance(dtype, torch.dtype):
            dtype = _NP_TO_TORCH_DTYPE[np.dtype(dtype)]
        if name in ('sum', 'mean', 'prod') and dtype is not None:
            t = t.to(dtype)
        elif name == 'mean' and not t.is_floating_point():
            t = t.to(torch.float64)                        # NumPy: the mean of integers / bools is a float64
        fn = getattr(torch, name)
        if axis is None:
            out = fn(t)
        else:
            out = fn(t, dim=axis, keepdim=bool(keep))
            if name in ('max', 'min'):
                out = out.values
        return out.as_subclass(DeviceArray)

    def sum(self, *args, **kwargs):
        return self._reduce('sum', args, kwargs)

    def mean(self, *args, **kwargs):
        return self._reduce('mean', args, kwargs)

    def prod(self, *args, **kwargs):
        return self._reduce('prod', args, kwargs)

    def any(self, *args, **kwargs):
        return self._reduce('any', args, kwargs)

    def all(self, *args, **kwargs):
        return self._reduce('all', args, kwargs)

    def max(self, *args, **kwargs):
        if args and isinstance(args[0], torch.Tensor):     # torch's elementwise max(other)
            return torch.maximum(self.as_subclass(torch.Tensor), args[0].as_subclass(torch.Tensor)).as_subclass(DeviceArray)
        return self._reduce('max', args, kwargs)

    def min(self, *args, **kwargs):
        if args and isinstance(args[0], torch.Tensor):
            return torch.minimum(self.as_subclass(torch.Tensor), args[0].as_subclass(torch.Tensor)).as_subclass(DeviceArray)
        return self._reduce('min', args, kwargs)


def wrap(t):
    return t.as_subclass(DeviceArray)
