"""GPU tests of the tcgen05 building blocks and the bf16x3 tensor-core point kernel."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("passes", [1, 2])
def test_tcgen05_cta_pair_gemm_layout(passes):
    """D = A.B^T through cta_group::2 lands in TMEM in the 2x2 datapath layout the point kernel assumes:
    CTA c, lane l, column j  <->  row c*64 + l%64, output column (l//64)*128 + j."""
    from disn_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(42)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = rng.standard_normal((256, 64)).astype(np.float32)
    D = np.empty((2, 128, 128), np.float32)
    _lib.check(lib.disn_tc_selftest(0, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), passes,
                                    D.ctypes.data_as(C.c_void_p)))
    ref = passes * (_bf16_round(A).astype(np.float64) @ _bf16_round(B).astype(np.float64).T)   # [128,256]
    got = np.empty((128, 256), np.float64)
    for c in range(2):
        for h in range(2):
            got[c * 64:(c + 1) * 64, h * 128:(h + 1) * 128] = D[c, h * 64:(h + 1) * 64, :]
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-3 * np.abs(ref).max()
