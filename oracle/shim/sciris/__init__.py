'''
Minimal stand-in for the `sciris` package -- TEST INFRASTRUCTURE ONLY.

The reference (Covasim 3.1.7) imports sciris at module top, and sciris is not
installed in this image (no network).  This shim provides just enough of its
surface for `import covasim` + `cv.Sim(...).run()` to work so that the reference
can be executed in the build container to generate golden vectors
(oracle/gen_golden.py).  Nothing in covasim_b200/ imports this.
'''
import copy as _copy
import datetime as _dt
import json as _json
import os as _os
import time as _time
import gzip as _gzip
import pickle as _pickle
import numbers as _numbers
import collections as _co
import numpy as np

__version__ = '3.1.0'

class KeyNotFoundError(KeyError):
    def __str__(self):
        return Exception.__str__(self)

class prettyobj(object):
    def __repr__(self):
        return f'<{self.__class__.__module__}.{self.__class__.__name__} at {hex(id(self))}>'

class odict(_co.OrderedDict):
    ''' Ordered dict that also supports integer indexing '''
    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)) and not _co.OrderedDict.__contains__(self, key):
            return list(self.values())[key]
        if isinstance(key, slice):                     # sciris: a slice returns the values as an array
            return np.array(list(self.values())[key])
        return _co.OrderedDict.__getitem__(self, key)
    def __setitem__(self, key, value):
        if isinstance(key, (int, np.integer)) and not _co.OrderedDict.__contains__(self, key) and len(self) > key >= 0:
            key = list(self.keys())[key]
        return _co.OrderedDict.__setitem__(self, key, value)
    def enumitems(self):
        return [(i, k, v) for i, (k, v) in enumerate(self.items())]
    def enumvals(self):
        return list(enumerate(self.values()))
    def sort(self, *a, **k):
        return self
    @staticmethod
    def fromkeys(keys, value=None):
        return odict((k, value) for k in keys)

class objdict(odict):
    ''' odict with attribute access '''
    def __getattribute__(self, attr):
        try:
            return odict.__getattribute__(self, attr)
        except AttributeError as E:
            try:
                return odict.__getitem__(self, attr)
            except KeyError:
                raise E
    def __setattr__(self, name, value):
        if name.startswith('_OrderedDict') or name in self.__dict__:
            odict.__setattr__(self, name, value)
        else:
            odict.__setitem__(self, name, value)
    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            odict.__delattr__(self, name)
    def setattribute(self, name, value):
        odict.__setattr__(self, name, value)
    def getattribute(self, name):
        return odict.__getattribute__(self, name)
    def delattribute(self, name):
        odict.__delattr__(self, name)

ddict = _co.defaultdict

def dcp(obj, *a, **k):
    return _copy.deepcopy(obj)

def cp(obj, *a, **k):
    return _copy.copy(obj)

def mergedicts(*args, _copy_=False, **kwargs):
    out = None
    for a in args:
        if a is None:
            continue
        if out is None:
            out = type(a)() if isinstance(a, dict) else {}
        out.update(a)
    if out is None:
        out = {}
    return out

def isnumber(x, isnan=None):
    return isinstance(x, _numbers.Number)

def isstring(x):
    return isinstance(x, str)

def isiterable(x):
    try:
        iter(x)
        return True
    except TypeError:
        return False

def checktype(obj=None, objtype=None, subtype=None, die=False):
    if objtype in ['arr', 'array', 'arraylike', 'listlike']:
        ok = isinstance(obj, (list, tuple, np.ndarray))
    elif objtype in ['num', 'number']:
        ok = isnumber(obj)
    elif objtype in ['str', 'string']:
        ok = isstring(obj)
    else:
        ok = isinstance(obj, objtype)
    if die and not ok:
        raise TypeError(f'{obj} is not {objtype}')
    return ok

def tolist(obj=None, objtype=None, keepnone=False, coerce='default'):
    if obj is None:
        return [None] if keepnone else []
    if isinstance(obj, list):
        return obj
    if isinstance(obj, (tuple, set, range, np.ndarray)) or type(obj).__name__ in ('dict_keys', 'dict_values'):
        return list(obj)
    return [obj]
promotetolist = tolist

def mergelists(*args, **kwargs):
    out = []
    for a in args:
        out.extend(tolist(a, keepnone=kwargs.get('keepnone', False)))
    return out

def toarray(x, keepnone=False, asobject=True, **kwargs):
    if isnumber(x) or (isinstance(x, np.ndarray) and not x.shape):
        return np.array([x], **kwargs)
    if isinstance(x, np.ndarray) and not kwargs:
        return x
    return np.array(x, **kwargs)
promotetoarray = toarray

def strjoin(*args, sep=', '):
    items = []
    for a in args:
        items.extend([str(i) for i in tolist(a)])
    return sep.join(items)

def newlinejoin(*args):
    return strjoin(*args, sep='\n')

def findinds(arr=None, val=None, *args, **kwargs):
    arr = np.asarray(arr)
    if val is None:
        return np.nonzero(arr)[0]
    if isnumber(val) and not isinstance(val, (int, np.integer, bool)):
        return np.nonzero(np.isclose(arr, val))[0]
    return np.nonzero(arr == val)[0]

def findlast(arr, val=None, **k):
    return findinds(arr, val)[-1]

def compareversions(v1, v2):
    import re
    def parse(v):
        return tuple(int(p) for p in re.findall(r'\d+', str(v))[:3])
    if isinstance(v1, str) is False and hasattr(v1, '__version__'):
        v1 = v1.__version__
    m = re.match(r'^\s*(<=|>=|==|<|>|=|!=|~=)?\s*(.*)$', str(v2))
    op, vs = m.group(1), m.group(2)
    a, b = parse(v1), parse(vs)
    cmp = (a > b) - (a < b)
    if op is None:
        return cmp
    return {'<': cmp < 0, '<=': cmp <= 0, '>': cmp > 0, '>=': cmp >= 0, '==': cmp == 0, '=': cmp == 0, '!=': cmp != 0, '~=': cmp != 0}[op]

def _todate(x):
    import pandas as pd
    if isinstance(x, _dt.datetime):
        return x.date()
    if isinstance(x, _dt.date):
        return x
    if isinstance(x, str):
        return _dt.datetime.strptime(x[:10], '%Y-%m-%d').date()
    if isinstance(x, pd.Timestamp):
        return x.date()
    if isinstance(x, np.datetime64):
        return pd.Timestamp(x).date()
    raise TypeError(f'Cannot convert {x!r} to a date')

def date(obj=None, *args, start_date=None, as_date=True, dateformat=None, **kwargs):
    if obj is None:
        return None
    if isinstance(obj, (list, tuple, np.ndarray)) :
        return [date(o, start_date=start_date, as_date=as_date, dateformat=dateformat) for o in obj]
    if isnumber(obj):
        if start_date is None:
            raise ValueError('start_date required')
        d = _todate(start_date) + _dt.timedelta(days=int(obj))
    else:
        d = _todate(obj)
    if not as_date:
        return d.strftime(dateformat or '%Y-%m-%d')
    return d
readdate = date

def getdate(obj=None, *a, **k):
    return _dt.datetime.now().strftime('%Y-%b-%d %H:%M:%S')

def now(*a, **k):
    return _dt.datetime.now()

def day(obj, *args, start_date=None, **kwargs):
    if obj is None:
        return None
    if isinstance(obj, (list, tuple, np.ndarray)):
        return [day(o, start_date=start_date) for o in obj]
    if isnumber(obj):
        return int(obj)
    if start_date is None:
        start_date = _dt.date(_todate(obj).year, 1, 1)
    return (_todate(obj) - _todate(start_date)).days

def daydiff(*args):
    days = [(_todate(b) - _todate(a)).days for a, b in zip(args[:-1], args[1:])]
    return days[0] if len(days) == 1 else days

def daterange(start_date, end_date, inclusive=True, as_date=False, dateformat=None):
    s, e = _todate(start_date), _todate(end_date)
    n = (e - s).days + (1 if inclusive else 0)
    out = [s + _dt.timedelta(days=i) for i in range(n)]
    if not as_date:
        out = [d.strftime(dateformat or '%Y-%m-%d') for d in out]
    return out

def printv(string, thisverbose=1, verbose=2, **kwargs):
    if verbose and verbose >= thisverbose:
        print(string)

def heading(string='', *a, **k):
    print(string)
    return string

def colorize(*a, **k):
    return ''

def indent(prefix='', text='', **k):
    return prefix + str(text)

def pp(obj, *a, doprint=True, **k):
    import pprint
    s = pprint.pformat(obj)
    if doprint:
        print(s)
    return s

def prepr(obj, *a, **k):
    return f'<{type(obj).__name__} at {hex(id(obj))}>\n' + '\n'.join(f'  {k}: {str(v)[:80]}' for k, v in getattr(obj, '__dict__', {}).items())

def pr(obj, *a, **k):
    print(prepr(obj))

def objectid(obj):
    return f'<{obj.__class__.__module__}.{obj.__class__.__name__} at {hex(id(obj))}>'

def randround(x):
    ''' Stochastic rounding: consumes exactly one draw of the NumPy global stream '''
    if isinstance(x, np.ndarray):
        return np.array(np.floor(x + np.random.random(x.size).reshape(x.shape)), dtype=int)
    return int(np.floor(x + np.random.random()))

def smooth(data, repeats=None, kernel=None, **k):
    if repeats is None:
        repeats = int(np.floor(len(data) / 5))
    if kernel is None:
        kernel = [0.25, 0.5, 0.25]
    kernel = np.array(kernel)
    output = np.array(data, dtype=float)
    pad = len(kernel) // 2
    for _ in range(repeats):
        padded = np.concatenate([np.full(pad, output[0]), output, np.full(pad, output[-1])])
        output = np.convolve(padded, kernel, mode='valid')
    return output

class timer(object):
    def __init__(self, *a, **k):
        self.t0 = _time.time()
    def tic(self):
        self.t0 = _time.time()
    def toc(self, *a, output=False, **k):
        el = _time.time() - self.t0
        return el
    def start(self): self.tic()
    def stop(self): return self.toc()
    @property
    def elapsed(self):
        return _time.time() - self.t0
Timer = timer
_tic = [0.0]
def tic():
    _tic[0] = _time.time(); return _tic[0]
def toc(start=None, output=False, **k):
    el = _time.time() - (start or _tic[0])
    if output: return el
    print(f'Elapsed time: {el:0.3f} s')

def progressbar(*a, **k):
    return

def gitinfo(*a, **k):
    return dict(branch='n/a', hash='n/a', date='n/a')

def getcaller(*a, **k):
    return dict(filename='n/a', lineno=0)

def thisdir(file=None, *args, aspath=False, **k):
    d = _os.path.dirname(_os.path.abspath(file)) if file else _os.getcwd()
    out = _os.path.join(d, *args)
    if aspath:
        import pathlib
        return pathlib.Path(out)
    return out

def makefilepath(filename=None, folder=None, ext=None, default=None, **k):
    if filename is None:
        filename = default if isinstance(default, str) else 'default'
    if ext and not filename.endswith('.' + ext.lstrip('.')):
        filename = filename + '.' + ext.lstrip('.')
    if folder:
        filename = _os.path.join(folder, filename)
    return _os.path.abspath(filename)

def loadjson(filename=None, folder=None, **k):
    if folder: filename = _os.path.join(folder, filename)
    with open(filename) as f:
        return _json.load(f)

def jsonify(obj, *a, **k):
    if isinstance(obj, dict):
        return {str(k_): jsonify(v) for k_, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [jsonify(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, (_dt.date, _dt.datetime)):
        return str(obj)
    if isinstance(obj, (str, int, float, bool)) or obj is None:
        return obj
    return str(obj)
sanitizejson = jsonify

def savejson(filename=None, obj=None, folder=None, **k):
    with open(makefilepath(filename, folder), 'w') as f:
        _json.dump(jsonify(obj), f, indent=2)
    return filename

def saveobj(filename=None, obj=None, folder=None, **k):
    fn = makefilepath(filename, folder)
    with _gzip.open(fn, 'wb') as f:
        _pickle.dump(obj, f, protocol=4)
    return fn
save = saveobj

def loadobj(filename=None, folder=None, **k):
    fn = makefilepath(filename, folder)
    with _gzip.open(fn, 'rb') as f:
        return _pickle.load(f)
load = loadobj

def traceback(*a, **k):
    import traceback as tb
    return tb.format_exc()

def flattendict(d, sep=None, _prefix=None):
    out = {}
    for k, v in d.items():
        key = (k,) if _prefix is None else _prefix + (k,)
        if isinstance(v, dict):
            out.update(flattendict(v, _prefix=key))
        else:
            out[key] = v
    return out

def parallelize(func, iterarg=None, iterkwargs=None, args=None, kwargs=None, ncpus=None, serial=True, **k):
    ''' Serial stand-in: deep-copies arguments like a process pool would '''
    kwargs = kwargs or {}
    args = args or ()
    out = []
    if iterkwargs is not None:
        if isinstance(iterkwargs, dict):
            keys = list(iterkwargs.keys())
            n = len(iterkwargs[keys[0]])
            iterkwargs = [{k_: iterkwargs[k_][i] for k_ in keys} for i in range(n)]
        for ikw in iterkwargs:
            out.append(func(*_copy.deepcopy(args), **_copy.deepcopy(kwargs), **_copy.deepcopy(ikw)))
    else:
        if isinstance(iterarg, (int, np.integer)):
            iterarg = range(iterarg)
        for ia in iterarg:
            out.append(func(_copy.deepcopy(ia), *_copy.deepcopy(args), **_copy.deepcopy(kwargs)))
    return out

def fonts(*a, **k):
    return []

class _Options(objdict):
    pass
options = _Options()

def __getattr__(name):
    ''' Anything not provided (plotting helpers etc.) becomes a no-op '''
    if name.startswith('__'):
        raise AttributeError(name)
    def _noop(*a, **k):
        return None
    _noop.__name__ = name
    return _noop
