"""Build libb200sp.so (the C-ABI CUDA library, include/b200sp.h) in-tree with nvcc for sm_100a."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
SUFFIX = os.environ.get('B200SP_LIB_SUFFIX', '')       # experiment builds (e.g. _lean with B200SP_NVCC_EXTRA) live beside the default library
LIB = os.path.join(PKG, 'libb200sp%s.so' % SUFFIX)
NVCC_FLAGS = (os.environ.get('B200SP_NVCC_EXTRA', '').split()) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-DB200SP_LEAN_TCG',
              '-Xcompiler', '-fPIC', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.exists(c) or c == 'nvcc'):
            return c
    raise RuntimeError('nvcc not found')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, '*')) + [os.path.join(ROOT, 'include', 'b200sp.h')]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    """Compile + link under an exclusive file lock: under torchrun every rank imports the package at the same time, and
    only one of them may run nvcc; the others block on the lock and then find the library up to date.  Objects go to a
    per-process directory and the .so is moved into place atomically, so nobody can dlopen a half-written file."""
    if not force and not needs_build():
        return LIB
    import fcntl
    os.makedirs(os.path.join(ROOT, 'build'), exist_ok=True)
    with open(os.path.join(ROOT, 'build', '.lock'), 'w') as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = _nvcc()
    objdir = os.path.join(ROOT, 'build', 'obj' + SUFFIX)
    os.makedirs(objdir, exist_ok=True)
    for stale in glob.glob(os.path.join(objdir, '*.o')):      # objects of sources that no longer exist must not linger
        if not os.path.exists(os.path.join(CSRC, os.path.basename(stale)[:-2] + '.cu')):
            os.remove(stale)
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    tmp = LIB + '.tmp.%d' % os.getpid()
    cmd = [nvcc, '-shared', '-o', tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
