'''
Builds covasim_b200/libcovasim_b200.so from covasim_b200/csrc/*.cu with nvcc for sm_100a.

No torch, no JIT cache: an explicit nvcc command, output kept in-tree so it travels to the GPU box.
-fmad=false is part of the numerical contract (float32 products must not be
contracted into FMAs, or per-edge probabilities differ from the reference in the last bit).
'''
import glob
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcovasim_b200.so')
STAMP = LIB + '.stamp'

# CVB_NVCC_EXTRA: extra flags for tuning builds (e.g. -DCVB_BEGIN_MINB=4), part of the source digest
NVCC_FLAGS = os.environ.get('CVB_NVCC_EXTRA', '').split() + ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-fmad=false',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fno-strict-aliasing', '--shared', '-cudart', 'static']


def find_nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: covasim_b200 needs the CUDA toolkit to build its kernels')
    return nvcc


def source_digest():
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, '*')) + glob.glob(os.path.join(ROOT, 'include', '*.h')))
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())         # (relative: the tree is copied to other machines with its built library)
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _up_to_date(digest):
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest


def build(force=False, verbose=False):
    '''
    Compile the library if the sources (or flags) changed since the stamp was written.  Safe when several processes import the
    package at once (one rank per GPU under torchrun): builders are serialised by a file lock, the compiler writes to a private
    file and the result is moved into place atomically, so nobody ever maps a half-written library.
    '''
    import fcntl
    digest = source_digest()
    if not force and _up_to_date(digest):
        return LIB
    with open(LIB + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):          # another process built it while this one waited
                return LIB
            srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
            tmp = f'{LIB}.tmp{os.getpid()}'
            cmd = [find_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC] + (['-Xptxas', '-v'] if verbose else []) + srcs + ['-o', tmp]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
            os.replace(tmp, LIB)
            with open(STAMP + '.tmp', 'w') as f:
                f.write(digest)
            os.replace(STAMP + '.tmp', STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
