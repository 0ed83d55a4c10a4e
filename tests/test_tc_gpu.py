"""tcgen05 plumbing self-test (b200sp_tc_probe): pins the UMMA descriptor encodings, the 128B swizzle
and the TMEM read-back used by every tensor-core kernel, against float64 matmul."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from speedplusbaseline_b200 import _lib as L
from kutil import rel, sp

pytestmark = pytest.mark.gpu


def run_probe(N, nkb, mode, a_major, b_major, variant=0, seed=0):
    g = torch.Generator().manual_seed(seed + 17 * N + mode)
    kd = (64 if mode == 2 else 32) * nkb
    A = torch.randn(128, kd, generator=g)
    B = torch.randn(N, kd, generator=g)
    if mode == 2:
        A, B = A.bfloat16(), B.bfloat16()
    ref = A.double() @ B.double().t()
    Ad = (A if a_major == 0 else A.t().contiguous()).cuda()
    Bd = (B if b_major == 0 else B.t().contiguous()).cuda()
    D = torch.full((128, N), float('nan'), device='cuda')
    L.call('b200sp_tc_probe', Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, nkb, mode, a_major, b_major, variant, sp())
    torch.cuda.synchronize()
    return rel(D, ref)


@pytest.mark.parametrize('a_major', [0, 1])
@pytest.mark.parametrize('b_major', [0, 1])
@pytest.mark.parametrize('mode,tol', [(0, 2e-3), (1, 2e-6), (2, 1e-6)])
@pytest.mark.parametrize('N,nkb', [(64, 1), (16, 2), (256, 2), (96, 1)])
def test_umma_formats(mode, tol, a_major, b_major, N, nkb):
    assert run_probe(N, nkb, mode, a_major, b_major) < tol


if __name__ == '__main__':
    # exploratory sweep: prints the error of every (mode, majors, variant) combination
    for mode in (0, 1, 2):
        for am in (0, 1):
            for bm in (0, 1):
                for variant in (0,):
                    for N, nkb in ((64, 1), (256, 2), (16, 2), (96, 1)):
                        try:
                            e = run_probe(N, nkb, mode, am, bm, variant)
                        except Exception as ex:        # noqa
                            e = repr(ex)
                        print('mode %d a_major %d b_major %d variant %d N %3d nkb %d -> %s' % (mode, am, bm, variant, N, nkb, e), flush=True)
