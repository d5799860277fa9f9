'''
CPU checks of the drop-in boundary: libcovasim_b200.so builds and loads without a GPU, exports every
symbol include/covasim_b200.h declares, the ctypes structs match the C layouts, the generated field
enum is up to date, and the product refuses to run without a device (no CPU fallback).
'''
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cv():
    import covasim_b200
    return covasim_b200


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'covasim_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cvb_[a-z0-9_]+)\s*\(', text)))


def test_every_declared_symbol_is_exported(cv):
    names = declared_functions()
    assert len(names) >= 30
    lib = C.CDLL(cv._capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in include/covasim_b200.h but not exported: {missing}'
    unbound = [n for n in names if n not in cv._capi.PROTOTYPES and n not in cv._capi.OTHER_SYMBOLS]
    assert not unbound, f'declared but not bound in covasim_b200/_capi.py: {unbound}'


def test_struct_layouts_match(cv):
    sizes = (C.c_int64 * 5)()
    cv._capi.call('cvb_struct_sizes', sizes)
    want = [C.sizeof(x) for x in (cv._capi.cvb_pars, cv._capi.cvb_dist, cv._capi.cvb_test_prob_pars, cv._capi.cvb_trace_pars,
                                  cv._capi.cvb_vaccinate_pars)]
    assert list(sizes) == want
    import re
    header = open(os.path.join(ROOT, 'include', 'covasim_b200.h')).read()
    assert cv._capi.lib.cvb_abi_version() == cv._capi.ABI_VERSION == int(re.search(r'#define CVB_ABI_VERSION\s+(\d+)', header).group(1)) == 2


def test_generated_field_header_is_current(cv):
    from covasim_b200 import gen_headers
    assert open(gen_headers.HEADER).read() == gen_headers.render(), 'run: python -m covasim_b200.gen_headers'


def test_no_cpu_fallback(cv):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(cv.CvbError):
        cv.Sim(pop_size=100).initialize()
    h = C.c_void_p()
    with pytest.raises(cv.CvbError, match='no CUDA device'):
        cv._capi.call('cvb_create', C.byref(h), 100, 1, 10, 1)
    with pytest.raises(cv.CvbError):
        cv.ops.compute_viral_load(0, [1.0], [2.0], [3.0], 0.3, 2.0, 4.0)


def test_error_reporting(cv):
    with pytest.raises(cv.CvbError, match='NULL'):
        cv._capi.call('cvb_set_seed', None, 1)
    with pytest.raises(cv.CvbError, match='n_agents'):
        cv._capi.call('cvb_create', C.byref(C.c_void_p()), 0, 1, 10, 1)


def test_population_generator_matches_oracle(cv):
    ''' The product's exact population builder consumes the MT streams like the reference (pinned through the oracle) '''
    import numpy as np
    from oracle import cvoracle as cvo
    from covasim_b200 import population as cvpop, parameters as cvpar, utils as cvu
    for pop_type, n in (('hybrid', 3000), ('random', 2000)):
        pars = cvpar.make_pars(pop_type=pop_type, pop_size=n, rand_seed=11)
        a_rng, b_rng = cvu.HostStreams(11), cvo.MTStreams(11)
        a = cvpop.make_randpop(pars, a_rng, exact=True)
        b = cvo.make_population(pars, b_rng)
        assert np.array_equal(a['age'], b['age']) and np.array_equal(a['sex'], b['sex'])
        for lk in b['contacts']:
            for c in ('p1', 'p2', 'beta'):
                assert np.array_equal(a['contacts'][lk][c], b['contacts'][lk][c]), (lk, c)
        assert a_rng.nb.random_sample() == b_rng.nb.random_sample()        # both streams left in the same state
        assert a_rng.np_.random_sample() == b_rng.np_.random_sample()
    # the fast builder draws the same distributions (same sizes, same edge multiset up to order inside households)
    pars = cvpar.make_pars(pop_type='hybrid', pop_size=3000, rand_seed=11)
    f = cvpop.make_randpop(pars, cvu.HostStreams(11), exact=False)
    e = cvpop.make_randpop(pars, cvu.HostStreams(11), exact=True)
    key = lambda d: np.sort(d['p1'].astype(np.int64) * 10**6 + d['p2'])
    assert np.array_equal(key(f['contacts']['h']), key(e['contacts']['h']))     # same households, edges in another order
    for lk in ('s', 'w', 'c'):                                                   # later layers: the streams have diverged, same statistics
        assert abs(len(f['contacts'][lk]['p1']) / len(e['contacts'][lk]['p1']) - 1) < 0.05


def test_binding_arity_matches_header(cv):
    ''' Every ctypes prototype has as many arguments as the C declaration (a mismatch would corrupt the call silently) '''
    text = open(os.path.join(ROOT, 'include', 'covasim_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = dict(re.findall(r'\b(cvb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', text))
    bad = {}
    for name, argtypes in cv._capi.PROTOTYPES.items():
        args = decls[name].strip()
        n = 0 if args in ('', 'void') else len(args.split(','))
        if n != len(argtypes):
            bad[name] = (n, len(argtypes))
    assert not bad, f'(header, ctypes) argument counts differ: {bad}'
    # the by-value structs added after the first ABI revision have their ctypes mirrors checked field by field here
    tn = cv._capi.cvb_test_num_pars
    assert [f[0] for f in tn._fields_] == ['symp_test', 'quar_test', 'quar_policy', 'index'] and C.sizeof(tn) == 24


def test_concurrent_builds_compile_once_and_never_expose_a_partial_library(tmp_path, monkeypatch):
    '''
    covasim_b200/build.py under several importers at once (one rank per GPU under torchrun): the builders are serialised by a file lock, the
    first one compiles into a private file and moves it into place, the others find the stamp.  The compiler is replaced by a slow writer.
    '''
    import importlib.util
    import threading
    import time
    spec = importlib.util.spec_from_file_location('cvb_build_under_test', os.path.join(ROOT, 'covasim_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = str(tmp_path / 'libx.so')
    monkeypatch.setattr(mod, 'LIB', lib)
    monkeypatch.setattr(mod, 'STAMP', lib + '.stamp')
    monkeypatch.setattr(mod, 'find_nvcc', lambda: 'nvcc')
    compiles, partial = [], []

    def fake_run(cmd, capture_output=True, text=True):
        out = cmd[cmd.index('-o') + 1]
        compiles.append(out)
        with open(out, 'wb') as f:
            for _ in range(10):
                f.write(b'x' * 1000)
                f.flush()
                time.sleep(0.02)
                if os.path.exists(lib) and os.path.getsize(lib) != 10000:
                    partial.append(os.path.getsize(lib))
        class R:
            returncode, stdout, stderr = 0, '', ''
        return R()
    monkeypatch.setattr(mod.subprocess, 'run', fake_run)
    results = []
    threads = [threading.Thread(target=lambda: results.append(mod.build())) for _ in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert results == [lib] * 6
    assert len(compiles) == 1 and compiles[0] != lib and not os.path.exists(compiles[0])
    assert os.path.getsize(lib) == 10000 and not partial
    assert open(lib + '.stamp').read().strip() == mod.source_digest()
    # the digest does not depend on where the tree lives (the built library travels with it to other machines)
    import shutil
    here = mod.source_digest()
    other = tmp_path / 'elsewhere'
    shutil.copytree(os.path.join(ROOT, 'covasim_b200', 'csrc'), other / 'covasim_b200' / 'csrc')
    shutil.copytree(os.path.join(ROOT, 'include'), other / 'include')
    monkeypatch.setattr(mod, 'ROOT', str(other))
    monkeypatch.setattr(mod, 'CSRC', str(other / 'covasim_b200' / 'csrc'))
    assert mod.source_digest() == here
