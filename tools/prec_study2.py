"""Which e5m2 correction products can be dropped per layer?  (CPU emulation, operand rounding only.)
mask per layer: 'b' both corrections, '1' only (a-h(a)).w, '2' only a.(w-h(w)), 'n' none.  layers: L0 64->256, L1 256->512, L2 512->512, L3 512->256"""
import sys, itertools
import numpy as np, torch
sys.path.insert(0, ".")
from disn_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
W = synth.make_weights(7, "he")
rng = np.random.default_rng(0)
pts = torch.from_numpy(rng.uniform(-1, 1, (N, 3))).double()
e5 = torch.float8_e5m2
q = lambda x, dt: x.float().to(dt).double()
def mm(a, w, mode, s1=6, s2=8):
    ah, wh = q(a, torch.float16), q(w, torch.float16)
    out = ah @ wh
    if mode in "b1": out = out + q((a - ah) * 2.0**s1, e5) @ q(w * 2.0**-s1, e5)
    if mode in "b2": out = out + q(a * 2.0**-s2, e5) @ q((w - wh) * 2.0**s2, e5)
    return out
def run(masks, exact=False):
    tot = 0
    for scope in ("sdfprediction", "sdfprediction_imgfeat"):
        g = lambda n: torch.from_numpy(np.asarray(W[f"{scope}/{n}"], np.float64))
        sq = lambda n: g(n).reshape(-1, g(n).shape[-1])
        M = (lambda a, w, m: a @ w) if exact else mm
        net = torch.relu(pts @ sq("fold1/conv1/weights") + g("fold1/conv1/biases"))
        net = torch.relu(M(net, sq("fold1/conv2/weights"), masks[0]) + g("fold1/conv2/biases"))
        net = torch.relu(M(net, sq("fold1/conv3/weights"), masks[1]) + g("fold1/conv3/biases"))
        extra = torch.from_numpy(np.random.default_rng(5).standard_normal((N, 512)) * 0.7)
        net = torch.relu(M(net, sq("fold2/conv1/weights")[:512], masks[2]) + extra)
        net = torch.relu(M(net, sq("fold2/conv2/weights"), masks[3]) + g("fold2/conv2/biases"))
        tot = tot + net @ sq("fold2/conv5/weights") + g("fold2/conv5/biases")
    return tot / 10.0
ref = run("bbbb", exact=True)
stages = [1, 8, 16, 8]
cost = {"b": 2.0, "1": 1.5, "2": 1.5, "n": 1.0}
for masks in ["bbbb", "1111", "2222", "nnnn", "bb1b", "bb2b", "bbnb", "b1b1", "b2b2", "1b1b", "b11b", "b22b", "nbbn", "n1bn", "bnnb"]:
    e = (run(masks) - ref).abs()
    c = sum(s * cost[m] for s, m in zip(stages, masks)) / sum(stages)
    print(f"{masks}  units/product {c:.3f}   max {e.max().item():.3e}  rms {e.pow(2).mean().sqrt().item():.3e}", flush=True)
