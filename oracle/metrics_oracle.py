"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the mesh-evaluation metrics next to the hot path.

nn_distance restates the reference's CPU kernel models/tf_ops/nn_distance/tf_nndistance.cpp:21-43 (`nnsearch`):
for each point of set 1 the minimum over set 2 of the float32 value (dx*dx + dy*dy) + dz*dz, first minimum wins.
It is PINNED against the reference itself: oracle/_ref/libref_nndistance.so is that file compiled in place
(oracle/Makefile, stub TF headers in oracle/ref_stubs) and tests/test_oracle_cpu.py compares the two bit for bit.
Chamfer: test/test_cd_emd.py:300-301; precision/recall/F: test/test_f_score.py:159-186,231-236.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libref_nndistance.so")


def nn_distance(xyz1, xyz2):
    """xyz1 [B,N,3], xyz2 [B,M,3] float32 -> dist1 [B,N] f32, idx1 [B,N] i32, dist2 [B,M], idx2 [B,M]."""
    a = np.ascontiguousarray(xyz1, np.float32)
    b = np.ascontiguousarray(xyz2, np.float32)

    def one_way(p, q):
        d = (p[:, :, None, :] - q[:, None, :, :]).astype(np.float32)        # q - p has the same squares
        s = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + (d[..., 2] * d[..., 2]).astype(np.float32)
        s = s.astype(np.float32)
        idx = np.argmin(s, axis=2).astype(np.int32)                         # first minimum, like `d < best`
        return np.take_along_axis(s, idx[..., None].astype(np.int64), axis=2)[..., 0], idx

    d1, i1 = one_way(a, b)
    d2, i2 = one_way(b, a)
    return d1, i1, d2, i2


def ref_nn_distance(xyz1, xyz2):
    """The reference's own compiled CPU op (oracle/_ref); raises FileNotFoundError when it was not built."""
    if not os.path.exists(REF_LIB):
        raise FileNotFoundError(REF_LIB)
    lib = C.CDLL(REF_LIB)
    a = np.ascontiguousarray(xyz1, np.float32)
    b = np.ascontiguousarray(xyz2, np.float32)
    B, N, _ = a.shape
    M = b.shape[1]
    d1, i1 = np.empty((B, N), np.float32), np.empty((B, N), np.int32)
    d2, i2 = np.empty((B, M), np.float32), np.empty((B, M), np.int32)
    err = C.create_string_buffer(256)
    rc = lib.ref_nn_distance(B, N, M, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                             d1.ctypes.data_as(C.c_void_p), i1.ctypes.data_as(C.c_void_p),
                             d2.ctypes.data_as(C.c_void_p), i2.ctypes.data_as(C.c_void_p), err, 256)
    if rc:
        raise ValueError(err.value.decode())
    return d1, i1, d2, i2


def chamfer_x1000(pred, src):
    """test/test_cd_emd.py:300-301 -- (mean fwd + mean bwd squared NN distance) * 1000 per batch item."""
    df, _, db, _ = nn_distance(pred, src)
    return (df.mean(axis=1) + db.mean(axis=1)) * np.float32(1000)


def precision_recall_f(pred, src, thresholds):
    """test/test_f_score.py:231-236,185: fractions of sqrt-distances below each threshold, F = 2PR/(P+R)."""
    df, _, db, _ = nn_distance(pred, src)
    th = np.asarray(thresholds, np.float32)[:, None]
    p = (np.sqrt(df).reshape(1, -1) < th).mean(axis=1)
    r = (np.sqrt(db).reshape(1, -1) < th).mean(axis=1)
    return p, r, 2 * p * r / np.maximum(p + r, 1e-30)
