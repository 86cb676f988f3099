"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): encode + 300 points on both kernels + MC."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disn_b200 import synth
from disn_b200.engine import Engine

W = synth.make_weights(seed=7, init="he")
for prec in (sys.argv[1:] or ["fp32", "bf16x3", "f16f8"]):
    eng = Engine(device=0, precision=prec)
    eng.load_weights(W)
    eng.encode(synth.synthetic_images(1))
    pts = np.random.default_rng(0).uniform(-1, 1, (1, 300, 3)).astype(np.float32)   # 3 pair-tiles: ring wrap-around, tail tile
    out = eng.eval_points(pts, synth.DEMO_TRANS_MAT)
    g = eng.eval_grid(synth.DEMO_SDF_PARAMS, synth.DEMO_TRANS_MAT, 6)
    v, f = eng.marching_cubes(g[0], [-1, -1, -1, 1, 1, 1], float(np.median(g)))
    print(prec, float(out.mean()), g.shape, v.shape, f.shape)
    eng.close()
