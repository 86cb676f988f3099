import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disn_b200 import _lib
from disn_b200.engine import Engine
eng = Engine(device=0, precision="fp32")
lib = _lib.load_test()
def run(M, N, K, mode):
    # structured inputs: A = ones, W[k, n] = indicator of K-slice -> output counts which slices contributed
    A = np.ones((M, K), np.float32)
    Wt = np.zeros((K, N), np.float32)
    if mode == "count":
        Wt[:] = 1.0 / 64            # every slice contributes exactly 1.0 to every output
    o32 = np.empty((M, N), np.float32); otc = np.empty((M, N), np.float32)
    lib.disn_debug_gemm(eng._h, A.ctypes.data_as(C.c_void_p), Wt.ctypes.data_as(C.c_void_p), None, M, N, K, 0, 0, 0, 0,
                        o32.ctypes.data_as(C.c_void_p), otc.ctypes.data_as(C.c_void_p))
    print("M=%d N=%d K=%d: fp32 unique %s ; tc unique values %s" % (M, N, K, np.unique(o32)[:5], np.unique(otc)[:10]))
    bad = np.argwhere(otc != o32)
    if len(bad):
        print("   bad count", len(bad), "rows", bad[:, 0].min(), "-", bad[:, 0].max(), "cols", bad[:, 1].min(), "-", bad[:, 1].max())
        for nb in range(N // 128):
            blk = otc[:, nb * 128:(nb + 1) * 128]
            print("   nb", nb, "values", np.unique(blk)[:8])
for M, K in ((196, 512), (784, 512), (196, 192), (196, 64 * 3)):
    run(M, 512, K, "count")
