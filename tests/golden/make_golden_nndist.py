"""Generate tests/golden/nn_distance_ref.npz from the REFERENCE's own CPU op compiled in place
(oracle/_ref/libref_nndistance.so <- models/tf_ops/nn_distance/tf_nndistance.cpp via oracle/Makefile).

Run in the build container only (it needs /root/reference to build oracle/_ref):
    make -C oracle && python tests/golden/make_golden_nndist.py
The fixture lets the GPU box (no /root/reference) pin both the CPU oracle and the CUDA kernel to outputs of the reference
itself: random clouds, an exact tie (first index must win), duplicated points and a 1-point set.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import metrics_oracle as mo  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    cases = {}
    a = rng.uniform(-1, 1, (3, 257, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (3, 190, 3)).astype(np.float32)
    b[1, 7] = b[1, 150]                     # exact tie: the lower index must be reported
    a[2, 5:9] = a[2, 4]                     # duplicated query points
    cases["rand"] = (a, b)
    cases["single"] = (rng.uniform(-1, 1, (1, 33, 3)).astype(np.float32), rng.uniform(-1, 1, (1, 1, 3)).astype(np.float32))
    g = np.stack(np.meshgrid(*[np.linspace(-0.5, 0.5, 5, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    cases["lattice"] = (g, g[:, ::-1].copy())      # every distance ties many ways
    out = {}
    for name, (x, y) in cases.items():
        d1, i1, d2, i2 = mo.ref_nn_distance(x, y)
        out.update({name + "_xyz1": x, name + "_xyz2": y, name + "_dist1": d1, name + "_idx1": i1,
                    name + "_dist2": d2, name + "_idx2": i2})
    np.savez(os.path.join(HERE, "nn_distance_ref.npz"), **out)
    print("nn_distance_ref.npz:", {k: v.shape for k, v in out.items() if k.endswith("dist1")})


if __name__ == "__main__":
    main()
