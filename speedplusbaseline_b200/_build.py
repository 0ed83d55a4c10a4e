"""Build libb200sp.so (the C-ABI CUDA library, include/b200sp.h) in-tree with nvcc for sm_100a."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libb200sp.so')
NVCC_FLAGS = (os.environ.get('B200SP_NVCC_EXTRA', '').split()) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
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
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(ROOT, 'build', 'obj')
    os.makedirs(objdir, exist_ok=True)
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
    cmd = [nvcc, '-shared', '-o', LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
