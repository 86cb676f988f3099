import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disn_b200 import _lib
from disn_b200.engine import Engine
eng = Engine(device=0, precision="fp32")
lib = _lib.load_test()
rng = np.random.default_rng(0)
def run(M, N, K, H=0, W=0, Cin=0, bias=True, relu=1):
    A = rng.standard_normal((M, Cin if H else K)).astype(np.float32)
    Wt = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) if bias else None
    o32 = np.empty((M, N), np.float32); otc = np.empty((M, N), np.float32)
    rc = lib.disn_debug_gemm(eng._h, A.ctypes.data_as(C.c_void_p), Wt.ctypes.data_as(C.c_void_p),
                             b.ctypes.data_as(C.c_void_p) if bias else None, M, N, K, H, W, Cin, relu,
                             o32.ctypes.data_as(C.c_void_p), otc.ctypes.data_as(C.c_void_p))
    if rc: print("ERR", lib.disn_last_error()); return
    err = np.abs(o32 - otc); scale = np.abs(o32).max()
    bad_rows = np.unique(np.nonzero(err > 1e-3 * scale)[0]); bad_cols = np.unique(np.nonzero(err > 1e-3 * scale)[1])
    print("M=%6d N=%4d K=%5d H=%3d Cin=%3d : rel err %.2e  bad rows %d (%s..) bad cols %d (%s..)" % (
        M, N, K, H, Cin, err.max() / scale, len(bad_rows), bad_rows[:4], len(bad_cols), bad_cols[:4]))
# plain (projection) shapes
for M, K in ((50176, 64), (12544, 128), (3136, 256), (784, 512), (196, 512), (256, 64), (128, 64)):
    run(M, 512, K, bias=False, relu=0)
# conv shapes (B=1 and B=2)
for B in (1, 2):
    for hw, Cin, Cout in ((56, 128, 256), (56, 256, 256), (28, 256, 512), (14, 512, 512), (112, 64, 128), (224, 64, 64)):
        run(B * hw * hw, Cout, 9 * Cin, hw, hw, Cin)
