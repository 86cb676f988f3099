import numpy as np, sys
sys.path.insert(0, '.')
from disn_b200 import synth
from disn_b200.engine import Engine
W = synth.make_weights(seed=7, init="he")
for B in (1, 2):
    imgs = synth.synthetic_images(B, seed=4321)
    outs = {}
    for prec in ("fp32", "bf16x3"):
        eng = Engine(device=0, precision=prec, max_batch=2)
        eng.load_weights(W); eng.encode(imgs)
        outs[prec] = [eng.get_encoded(k) for k in (1, 2, 3, 4, 5, 0, 7, 6)]
        eng.close()
    for k, (a, b) in enumerate(zip(outs["fp32"], outs["bf16x3"])):
        rel = np.abs(a - b).max() / np.abs(a).max()
        per_img = [float(np.abs(a[i] - b[i]).max() / np.abs(a).max()) for i in range(B)]
        print("B=%d what[%d] shape %s rel err %.3e per image %s" % (B, k, a.shape, rel, ["%.1e" % v for v in per_img]))
