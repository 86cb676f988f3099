"""Where does a bench step go?  CUDA events between the encoder and the point kernel inside the step loop."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from disn_b200 import synth
from disn_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(device=0, precision="f16f8")
eng.load_weights(synth.make_weights(seed=7, init="he"))
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
img = torch.from_numpy(synth.synthetic_images(1)).to(dev)
tm = torch.from_numpy(synth.DEMO_TRANS_MAT.copy()).to(dev)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = res + 1
out = torch.empty((R, R, R), dtype=torch.float32, device=dev)
sp = synth.DEMO_SDF_PARAMS
for mode in ("enc+grid", "grid only", "enc+grid", "enc, idle 2 ms, grid"):
    for _ in range(3):
        eng.encode_device(img.data_ptr(), 1, 137, 137, 3)
        eng.eval_grid_device(sp, tm.data_ptr(), res, 0, R, out.data_ptr())
    torch.cuda.synchronize()
    K = 8
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K + 1)]
    for i in range(K):
        ev[i][0].record(stream)
        if mode != "grid only":
            eng.encode_device(img.data_ptr(), 1, 137, 137, 3)
        ev[i][1].record(stream)
        if mode.startswith("enc, idle"):
            torch.cuda._sleep(int(2e-3 * 1.9e9))
        eng.eval_grid_device(sp, tm.data_ptr(), res, 0, R, out.data_ptr())
        ev[i][2].record(stream)
    ev[K][0].record(stream)
    torch.cuda.synchronize()
    enc = [ev[i][0].elapsed_time(ev[i][1]) for i in range(K)]
    grid = [ev[i][1].elapsed_time(ev[i][2]) for i in range(K)]
    gap = [ev[i][2].elapsed_time(ev[i + 1][0]) for i in range(K)]
    tot = ev[0][0].elapsed_time(ev[K][0]) / K
    print("%-22s step %.3f ms | enc %.3f  grid %.3f (min %.3f max %.3f)  gap %.3f" % (
        mode, tot, np.mean(enc), np.mean(grid), min(grid), max(grid), np.mean(gap)), flush=True)
eng.close()
