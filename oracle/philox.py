'''
TEST INFRASTRUCTURE -- NumPy restatement of the Philox4x32-10 counter-based generator and of the
keyed draw functions the native-RNG mode of covasim_b200 uses (covasim_b200/csrc/philox.cuh).

Philox4x32-10 is the published algorithm of Salmon et al., "Parallel random numbers: as easy as
1, 2, 3" (SC'11); the constants below are the paper's.  It is checked against the known-answer
vectors of the Random123 distribution in tests/test_philox.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
'''
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

# Purpose tags: which draw of the model a counter belongs to (mirrors enum cvb_purpose in philox.cuh)
P_EDGE, P_INFECT, P_TEST, P_TEST_SENS, P_TEST_LOSS, P_TRACE, P_VACC, P_NAB_VACC, P_DYNLAYER = range(1, 10)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    ''' Vectorised Philox4x32; all inputs broadcastable uint32 arrays; returns 4 uint32 arrays '''
    with np.errstate(over='ignore'):
        c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)])
        c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
        k0 = np.uint32(k0)
        k1 = np.uint32(k1)
        for r in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(a, b):
    ''' Two uint32 words -> double in [0,1) with 53 random bits (same recipe as MT19937 random_sample) '''
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def keyed_words(seed, purpose, sub, day, index, slot=0):
    '''
    The one place that defines how (seed, purpose, sub, day, index, slot) map onto a Philox call:
        key     = (seed & 0xffffffff, (seed >> 32) ^ (purpose << 24) ^ sub)
        counter = (index & 0xffffffff, index >> 32, day (two's complement), slot)
    '''
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) ^ ((int(purpose) & 0xFF) << 24) ^ (int(sub) & 0xFFFFFF)) & 0xFFFFFFFF
    index = np.asarray(index, dtype=np.int64).astype(np.uint64)
    c0 = (index & _MASK).astype(np.uint32)
    c1 = (index >> _S32).astype(np.uint32)
    c2 = np.uint32(int(day) & 0xFFFFFFFF)
    c3 = np.uint32(int(slot) & 0xFFFFFFFF)
    return philox4x32(c0, c1, c2, c3, k0, k1)


def keyed_uniform2(seed, purpose, sub, day, index, slot=0):
    ''' Two independent uniforms per index from one Philox call '''
    w0, w1, w2, w3 = keyed_words(seed, purpose, sub, day, index, slot)
    return u53(w0, w1), u53(w2, w3)


def keyed_uniform(seed, purpose, sub, day, index, slot=0):
    return keyed_uniform2(seed, purpose, sub, day, index, slot)[0]


def keyed_normal(seed, purpose, sub, day, index, slot=0):
    ''' Standard normal by Box-Muller (cosine branch) from the two uniforms of one call '''
    u1, u2 = keyed_uniform2(seed, purpose, sub, day, index, slot)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(6.283185307179586 * u2)
