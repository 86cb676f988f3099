"""CPU emulation of the tensor-core operand schemes (operand rounding only, fp64 accumulate): pins the error-budget claims of
DESIGN.md section 3 without a GPU.  The scale rule is the one `tc_pack_weights` applies (s1 = 10 + L, s2 = 12 + L, 2^L ~ rms w)."""
import numpy as np
import pytest
import torch

from disn_b200 import synth

N = 6000


def _q(x, dt):
    return x.float().to(dt).double()


def _mm_exact(a, w):
    return a @ w


def _mm_fp16(a, w):
    return _q(a, torch.float16) @ _q(w, torch.float16)


def _mm_bf16x3(a, w):
    ah, wh = _q(a, torch.bfloat16), _q(w, torch.bfloat16)
    al, wl = _q(a - ah, torch.bfloat16), _q(w - wh, torch.bfloat16)
    return ah @ wh + al @ wh + ah @ wl


def _mm_f16f8(a, w):
    e5 = torch.float8_e5m2
    L = int(np.round(np.log2(float(w.pow(2).mean().sqrt()))))
    s1, s2 = 10 + L, 12 + L
    ah, wh = _q(a, torch.float16), _q(w, torch.float16)
    c1 = _q((a - ah) * 2.0 ** s1, e5) @ _q(w * 2.0 ** -s1, e5)
    c2 = _q(_q(a, torch.float16) * 2.0 ** -s2, e5) @ _q((w - wh) * 2.0 ** s2, e5)   # the kernel scales the fp16 copy
    return ah @ wh + c1 + c2


def _sdf(mm, W, pts):
    tot = 0
    for scope in ("sdfprediction", "sdfprediction_imgfeat"):
        g = lambda n: torch.from_numpy(np.asarray(W[f"{scope}/{n}"], np.float64))
        sq = lambda n: g(n).reshape(-1, g(n).shape[-1])
        net = torch.relu(pts @ sq("fold1/conv1/weights") + g("fold1/conv1/biases"))          # CUDA-core fp32 layer
        net = torch.relu(mm(net, sq("fold1/conv2/weights")) + g("fold1/conv2/biases"))
        net = torch.relu(mm(net, sq("fold1/conv3/weights")) + g("fold1/conv3/biases"))
        extra = torch.from_numpy(np.random.default_rng(5).standard_normal((pts.shape[0], 512)) * 0.7)
        net = torch.relu(mm(net, sq("fold2/conv1/weights")[:512]) + extra)                  # + global bias / gathered map
        net = torch.relu(mm(net, sq("fold2/conv2/weights")) + g("fold2/conv2/biases"))
        tot = tot + net @ sq("fold2/conv5/weights") + g("fold2/conv5/biases")                # fp32 epilogue dot product
    return tot / 10.0


@pytest.fixture(scope="module")
def study(he_weights):
    pts = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (N, 3))).double()
    ref = _sdf(_mm_exact, he_weights, pts)
    err = lambda mm: float((_sdf(mm, he_weights, pts) - ref).abs().max())
    return {"rms": float(ref.std()), "fp16": err(_mm_fp16), "bf16x3": err(_mm_bf16x3), "f16f8": err(_mm_f16f8)}


def test_field_is_order_one(study):
    assert study["rms"] > 0.02            # sdf = pred / 10 with |pred| ~ 0.5: the 1e-4 bar is a ~2e-3 relative bar


def test_split_operand_schemes_meet_the_tolerance_with_margin(study):
    assert study["bf16x3"] < 1e-5, study
    assert study["f16f8"] < 4e-5, study   # 1.2e-5 on 50k points; tolerance 1e-4


def test_single_pass_fp16_does_not(study):
    assert study["fp16"] > 4 * study["f16f8"], study
    assert study["fp16"] > 5e-5, study
