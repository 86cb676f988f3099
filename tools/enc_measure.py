"""Per-role blocked time of every tcgen05 conv / projection GEMM of one encode (DISN_CONV_MEASURE=1, DISN_NO_GRAPH=1).
    DISN_CONV_MEASURE=1 DISN_NO_GRAPH=1 python tools/enc_measure.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disn_b200 import synth
from disn_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = Engine(device=0, precision="bf16x3", max_batch=B)
eng.load_weights(synth.make_weights(seed=7, init="he"))
imgs = synth.synthetic_images(B)
eng.encode(imgs)
sys.stderr.write("---- second encode (weights packed, caches warm) ----\n")
eng.encode(imgs)
eng.close()
