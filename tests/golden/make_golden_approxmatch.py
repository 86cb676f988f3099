"""Generate tests/golden/approxmatch_ref.npz from the REFERENCE's own CPU ops compiled in place
(oracle/_ref/libref_approxmatch.so <- models/tf_ops/approxmatch/tf_approxmatch.cpp via oracle/Makefile).

Run in the build container only (it needs /root/reference to build oracle/_ref):
    make -C oracle && python tests/golden/make_golden_approxmatch.py
The fixture lets the GPU box (no /root/reference) pin both the CPU oracle and the CUDA kernels to outputs of the reference
itself: equal-size clouds, N > M and N < M (the integer capacity factors max(N,M)/N, max(N,M)/M differ), duplicated points,
a one-point set, and two identical clouds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import metrics_oracle as mo  # noqa: E402


def cases():
    rng = np.random.default_rng(2025)
    u = lambda *s: rng.uniform(-0.5, 0.5, s).astype(np.float32)
    out = {"square": (u(2, 128, 3), u(2, 128, 3)), "wide": (u(2, 32, 3), u(2, 100, 3)), "tall": (u(1, 96, 3), u(1, 32, 3)),
           "single": (u(1, 33, 3), u(1, 1, 3))}
    a, b = u(1, 64, 3), u(1, 64, 3)
    a[0, 5:9] = a[0, 4]                      # duplicated points
    b[0, 20] = a[0, 4]                       # and an exact coincidence between the sets
    out["dups"] = (a, b)
    c = u(1, 48, 3)
    out["same"] = (c, c.copy())
    return out


def main():
    out = {}
    for name, (x, y) in cases().items():
        m = mo.ref_approx_match(x, y)
        out.update({name + "_xyz1": x, name + "_xyz2": y, name + "_match": m, name + "_cost": mo.ref_match_cost(x, y, m)})
    np.savez_compressed(os.path.join(HERE, "approxmatch_ref.npz"), **out)
    print("approxmatch_ref.npz:", {k: (v.shape, v.tolist()) for k, v in out.items() if k.endswith("cost")})


if __name__ == "__main__":
    main()
