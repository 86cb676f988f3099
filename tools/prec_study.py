"""Error-budget study for operand formats of the point MLP (CPU emulation, operand rounding only; fp64 accumulate).
Usage: python tools/prec_study.py [N]"""
import sys, itertools
import numpy as np, torch
sys.path.insert(0, ".")
from disn_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
W = synth.make_weights(7, "he")
rng = np.random.default_rng(0)
pts = torch.from_numpy(rng.uniform(-1, 1, (N, 3))).double()

def q(x, dt):
    return x.float().to(dt).double()

def mm_exact(a, w): return a @ w
def mm_fmt(dt):
    return lambda a, w: q(a, dt) @ q(w, dt)
def mm_split3(dt):
    def f(a, w):
        ah, wh = q(a, dt), q(w, dt); al, wl = q(a - ah, dt), q(w - wh, dt)
        return ah @ wh + al @ wh + ah @ wl
    return f
def mm_f16_f8(fa_lo, fw_hi8, fa_hi8, fw_lo, s1, s2, main=torch.float16):
    def f(a, w):
        ah, wh = q(a, main), q(w, main)
        al, wl = a - ah, w - wh
        c1 = q(al * 2.0**s1, fa_lo) @ q(w * 2.0**-s1, fw_hi8)
        c2 = q(a * 2.0**-s2, fa_hi8) @ q(wl * 2.0**s2, fw_lo)
        return ah @ wh + c1 + c2
    return f

def run(mm):
    tot = 0
    for scope in ("sdfprediction", "sdfprediction_imgfeat"):
        g = lambda n: torch.from_numpy(np.asarray(W[f"{scope}/{n}"], np.float64))
        sq = lambda n: g(n).reshape(-1, g(n).shape[-1])
        net = torch.relu(pts @ sq("fold1/conv1/weights") + g("fold1/conv1/biases"))      # fp32 CUDA-core layer
        net = torch.relu(mm(net, sq("fold1/conv2/weights")) + g("fold1/conv2/biases"))
        net = torch.relu(mm(net, sq("fold1/conv3/weights")) + g("fold1/conv3/biases"))
        extra = torch.from_numpy(np.random.default_rng(5).standard_normal((N, 512)) * 0.7)  # gathered map / global bias
        net = torch.relu(mm(net, sq("fold2/conv1/weights")[:512]) + extra)
        net = torch.relu(mm(net, sq("fold2/conv2/weights")) + g("fold2/conv2/biases"))
        tot = tot + net @ sq("fold2/conv5/weights") + g("fold2/conv5/biases")
    return tot / 10.0

ref = run(mm_exact)
print("rms sdf", ref.std().item())
def rep(name, mm):
    e = (run(mm) - ref).abs()
    print(f"{name:48s} max {e.max().item():.3e}  rms {e.pow(2).mean().sqrt().item():.3e}", flush=True)
rep("bf16 single", mm_fmt(torch.bfloat16))
rep("fp16 single", mm_fmt(torch.float16))
rep("bf16 x3", mm_split3(torch.bfloat16))
e5, e4 = torch.float8_e5m2, torch.float8_e4m3fn
for fmts in [(e5, e5, e5, e5), (e4, e4, e4, e4), (e5, e4, e5, e4), (e4, e5, e4, e5)]:
    for s1, s2 in [(0, 0), (4, 4), (8, 8), (6, 10), (10, 6), (12, 12)]:
        nm = "f16+f8 " + ",".join(str(f).split("_")[-1] for f in fmts) + f" s1={s1} s2={s2}"
        rep(nm, mm_f16_f8(*fmts, s1, s2))
