'''
CPU check of the per-element arithmetic the CUDA kernels use (covasim_b200/csrc/cvb_device.cuh is
written as __host__ __device__ functions; tests/hostcheck compiles it for the host).  Compared with
the oracle's restatements and with the golden kernel vectors recorded from the reference.  This
catches float32/float64-recipe and Philox-keying mistakes before any GPU time is spent; the GPU
parity tests (tests/test_gpu_*.py) are the ones that count.
'''
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cvoracle as cvo
from oracle import philox as ph

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, 'hostcheck', '_hostcheck.so')


@pytest.fixture(scope='module')
def hc():
    src = os.path.join(HERE, 'hostcheck', 'hostcheck.cpp')
    cmd = ['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-x', 'c++', src,
           '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(ROOT, 'covasim_b200', 'csrc'), '-o', SO]
    subprocess.run(cmd, check=True)
    return C.CDLL(SO)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize('name,day', [('hybrid3k', 12), ('hybrid3k', 25), ('variants4k', 20)])
def test_viral_load_and_trans_sus_vs_reference_vectors(hc, golden, name, day):
    g = golden(name)
    pre = f'k{day}/'
    di, dr, dd = g[pre + 'vl/date_inf'], g[pre + 'vl/date_rec'], g[pre + 'vl/date_dead']
    out = np.empty_like(di)
    hc.hc_viral_load(C.c_int(int(g[pre + 'vl/t'])), ptr(di), ptr(dr), ptr(dd), C.c_float(float(g[pre + 'vl/frac_time'])),
                     C.c_float(float(g[pre + 'vl/load_ratio'])), C.c_float(float(g[pre + 'vl/high_cap'])), ptr(out), C.c_long(len(di)))
    assert np.array_equal(out, g[pre + 'vl/out'])
    for j in range(int(g[pre + 'n_calls'])):
        a = {k.split('/')[-1]: np.ascontiguousarray(g[k]) for k in g.files if k.startswith(f'{pre}ts{j}/')}
        n = len(a['rel_trans'])
        ot, os_ = np.empty(n, np.float32), np.empty(n, np.float32)
        u8 = lambda x: np.ascontiguousarray(x.astype(np.uint8))
        inf, sus, symp, iso, quar = u8(a['inf']), u8(a['sus']), u8(a['symp']), u8(a['iso']), u8(a['quar'])
        hc.hc_trans_sus(ptr(a['rel_trans']), ptr(a['rel_sus']), ptr(inf), ptr(sus), C.c_float(float(a['beta_layer'])), ptr(a['viral_load']),
                        ptr(symp), ptr(iso), ptr(quar), C.c_float(float(a['asymp_factor'])), C.c_float(float(a['iso_factor'])),
                        C.c_float(float(a['quar_factor'])), ptr(a['immunity_factors']), ptr(ot), ptr(os_), C.c_long(n))
        assert np.array_equal(ot, a['out_trans'])
        assert np.array_equal(os_, a['out_sus'])


def test_philox_keying_matches_oracle(hc):
    idx = np.concatenate([np.arange(0, 5000, dtype=np.int64), np.array([2**31 - 1, 2**32 + 5, 2**39 + 17], dtype=np.int64)])
    for seed, purpose, sub, day, slot in [(1, ph.P_EDGE, 3, 17, 0), (2**40 + 12345, ph.P_INFECT, 0, 0, 9), (7, ph.P_TRACE, (2 << 8) | 1, 59, 0),
                                          (99, ph.P_TEST, 1, -3, 0)]:
        u1, u2, z = np.empty(len(idx)), np.empty(len(idx)), np.empty(len(idx))
        hc.hc_keyed_uniform2(C.c_ulonglong(seed), C.c_uint(purpose), C.c_uint(sub), C.c_int(day), ptr(idx), C.c_uint(slot), ptr(u1), ptr(u2), C.c_long(len(idx)))
        o1, o2 = ph.keyed_uniform2(seed, purpose, sub, day, idx, slot)
        assert np.array_equal(u1, o1) and np.array_equal(u2, o2)
        hc.hc_keyed_normal(C.c_ulonglong(seed), C.c_uint(purpose), C.c_uint(sub), C.c_int(day), ptr(idx), C.c_uint(slot), ptr(z), C.c_long(len(idx)))
        np.testing.assert_allclose(z, ph.keyed_normal(seed, purpose, sub, day, idx, slot), rtol=1e-13, atol=1e-15)


def test_distributions_and_immunity_math(hc):
    rng = np.random.RandomState(0)
    z = rng.standard_normal(4000)
    mean, sigma = cvo.lognormal_pars(4.5, 1.5)
    out = np.empty_like(z)
    hc.hc_dist(C.c_int(5), C.c_double(mean), C.c_double(sigma), ptr(z), ptr(out), C.c_long(len(z)))
    assert np.array_equal(out, np.round(np.exp(mean + sigma * z)))
    hc.hc_dist(C.c_int(1), C.c_double(-1.0), C.c_double(2.0), ptr(z), ptr(out), C.c_long(len(z)))
    assert np.array_equal(out, -1.0 + 2.0 * z)
    enab = np.abs(rng.standard_normal(4000)) * 3
    enab[::7] = 0.0
    ve = np.empty(len(enab), np.float32)
    for alpha, beta in [(1.08, 0.967), (-0.739, 0.038), (-0.014, 0.079)]:
        hc.hc_calc_ve(ptr(enab), C.c_double(np.exp(alpha)), C.c_double(beta), ptr(ve), C.c_long(len(enab)))
        np.testing.assert_allclose(ve, cvo.calc_VE(enab, alpha, beta).astype(np.float32), rtol=1e-6, atol=0)
    o = [np.empty(len(enab), np.float32) for _ in range(3)]
    ab = [(1.08, 0.967), (-0.739, 0.038), (-0.014, 0.079)]
    hc.hc_calc_ve3(ptr(enab), *[C.c_double(x) for a, b in ab for x in (np.exp(a), b)], ptr(o[0]), ptr(o[1]), ptr(o[2]), C.c_long(len(enab)))
    for (a, b), got in zip(ab, o):
        np.testing.assert_allclose(got, cvo.calc_VE(enab, a, b).astype(np.float32), rtol=2e-7, atol=0)      # at most one float32 ulp
    nab = np.abs(rng.standard_normal(4000)).astype(np.float32)
    peak = (nab + np.abs(rng.standard_normal(4000)) * 0.1).astype(np.float32)
    kin = rng.standard_normal(4000) * 0.05
    got = np.empty(4000, np.float32)
    hc.hc_nab_step(ptr(nab), ptr(peak), ptr(kin), ptr(got), C.c_long(4000))
    want = (nab.astype(np.float64) + kin * peak.astype(np.float64)).astype(np.float32)
    want = np.where(want < 0, np.float32(0), want)
    want = np.where(want > peak, peak, want)
    assert np.array_equal(got, want)
    base = rng.random_sample(4000).astype(np.float32)
    imm = rng.random_sample(4000).astype(np.float32)
    oi, of = np.empty(4000, np.float32), np.empty(4000, np.float32)
    hc.hc_prog_prob(C.c_float(1.64), ptr(base), ptr(imm), C.c_float(2.0), ptr(oi), ptr(of), C.c_long(4000))
    assert np.array_equal(oi, ((np.float32(1.64) * base).astype(np.float32) * (np.float32(1) - imm).astype(np.float32)).astype(np.float32))
    assert np.array_equal(of, ((np.float32(1.64) * base).astype(np.float32) * np.float32(2.0)).astype(np.float32))


def test_edge_probability_matches_oracle(hc):
    rng = np.random.RandomState(1)
    n = 5000
    lb = rng.random_sample(n).astype(np.float32)
    ts = (rng.random_sample(n) * 3).astype(np.float32)
    st = rng.random_sample(n).astype(np.float32)
    out = np.empty(n, np.float32)
    hc.hc_edge_prob(C.c_float(0.016), ptr(lb), ptr(ts), ptr(st), ptr(out), C.c_long(n))
    want = (((np.float32(0.016) * lb).astype(np.float32) * ts).astype(np.float32) * st).astype(np.float32)
    assert np.array_equal(out, want)


def test_agent_record_path_equals_per_layer_tables(hc):
    '''
    prepare_transmission writes ONE 16-byte record per agent and the edge pass applies the per-layer factors to the edges it
    evaluates; for every combination of flags this must give exactly the per-layer {rel_trans, rel_sus} tables of the
    reference's compute_trans_sus (utils.py:82-90), which hc_trans_sus restates and the vectors test above pins.
    '''
    rng = np.random.RandomState(3)
    n = 20000
    f32 = np.float32
    rt = (rng.gamma(0.45, 2.2, n) * (rng.random_sample(n) < 0.8)).astype(f32)
    rs = rng.choice([0.34, 0.67, 1.0, 1.24, 1.47], n).astype(f32)
    flags = {k: (rng.random_sample(n) < p).astype(np.uint8) for k, p in dict(inf=0.5, sus=0.6, symp=0.5, iso=0.3, quar=0.3, early=0.5).items()}
    imm = (rng.random_sample(n) * (rng.random_sample(n) < 0.5)).astype(f32)
    frac_time, load_ratio = f32(0.3), f32(2.0)
    vl = np.where(flags['early'], f32(load_ratio) / (f32(1) + frac_time * (load_ratio - f32(1))), f32(1) / (f32(1) + frac_time * (load_ratio - f32(1)))).astype(f32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for beta_layer, af, isf, qf in [(3.0, 1.0, 0.3, 0.6), (0.6, 0.5, 0.1, 0.2), (0.3, 1.0, 0.0, 1.0)]:
        ot, os_ = np.zeros(n, f32), np.zeros(n, f32)
        hc.hc_trans_sus(ptr(rt), ptr(rs), ptr(flags['inf']), ptr(flags['sus']), C.c_float(beta_layer), ptr(vl), ptr(flags['symp']), ptr(flags['iso']),
                        ptr(flags['quar']), C.c_float(af), C.c_float(isf), C.c_float(qf), ptr(imm), ptr(ot), ptr(os_), C.c_long(n))
        rt2, rs2 = np.zeros(n, f32), np.zeros(n, f32)
        hc.hc_record_trans_sus(ptr(rt), ptr(rs), ptr(flags['inf']), ptr(flags['sus']), C.c_float(beta_layer), ptr(flags['early']), ptr(flags['symp']),
                               ptr(flags['iso']), ptr(flags['quar']), C.c_float(af), C.c_float(isf), C.c_float(qf), ptr(imm), C.c_float(frac_time),
                               C.c_float(load_ratio), ptr(rt2), ptr(rs2), C.c_long(n))
        assert np.array_equal(ot, rt2) and np.array_equal(os_, rs2)
        assert np.count_nonzero(ot) > 1000 and np.count_nonzero(os_) > 1000
