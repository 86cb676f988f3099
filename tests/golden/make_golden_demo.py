"""tests/golden/demo_input.npz: the reference's shipped demo render decoded exactly as demo/demo.py:262-265 does
(cv2.imread(IMREAD_UNCHANGED) -> uint8 -> [:, :, :3]), stored as the uint8 array so the GPU box (no /root/reference)
can push the real input through the CUDA path.  Also the mesh statistics of the shipped demo/result.obj (vertex and face
counts, bounding box) as an orientation pin for the res-256 demo output.

    python tests/golden/make_golden_demo.py        # build container only
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/demo"


def main():
    img = cv2.imread(os.path.join(REF, "03001627_17e916fc863540ee3def89b32cef8e45_20.png"), cv2.IMREAD_UNCHANGED)
    img_arr = img.astype(np.uint8)[:, :, :3]
    nv = nf = 0
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    with open(os.path.join(REF, "result.obj")) as f:
        for line in f:
            if line.startswith("v "):
                v = np.array(line.split()[1:4], dtype=np.float64)
                lo, hi = np.minimum(lo, v), np.maximum(hi, v)
                nv += 1
            elif line.startswith("f "):
                nf += 1
    np.savez_compressed(os.path.join(HERE, "demo_input.npz"), img_u8=img_arr, channels_in_file=np.int32(img.shape[2]),
                        result_obj_counts=np.array([nv, nf], np.int64), result_obj_bbox=np.stack([lo, hi]))
    print(img.shape, img_arr.mean(), nv, nf, lo, hi)


if __name__ == "__main__":
    main()
