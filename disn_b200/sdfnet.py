"""Mirror of the two reference MLP heads on the hot path (models/sdfnet.py:69-92 get_sdf_basic2,
:171-190 get_sdf_basic2_imgfeat_twostream).  In the reference these build 1x1-conv sub-graphs inside
get_model; on the B200 path both heads live inside the fused point kernel, so the functions here return
the symbolic outputs get_model wires together and document the layer contract the kernel implements.

Layer contract (utils/tf_util.py:119-184 conv2d with kernel [1,1], VALID, bn=False: y = relu(x.W + b)):
    fold1/conv1 3->64, fold1/conv2 64->256, fold1/conv3 256->512,
    concat [point512 | global1024]  (get_sdf_basic2)         -> fold2/conv1 1536->512
    concat [point512 | point_feat1472] (..._imgfeat_twostream) -> fold2/conv1 1984->512
    fold2/conv2 512->256, fold2/conv5 256->1 (linear)
"""
from __future__ import annotations

from .model_normalization import Tensor

LAYERS = (("fold1/conv1", 3, 64, True), ("fold1/conv2", 64, 256, True), ("fold1/conv3", 256, 512, True),
          ("fold2/conv1", None, 512, True), ("fold2/conv2", 512, 256, True), ("fold2/conv5", 256, 1, False))


def get_sdf_basic2(src_pc, globalfeats, is_training, batch_size, num_point, bn, bn_decay, wd=None):
    """models/sdfnet.py:69-92 -- global stream: [B,N,3], [B,1024] -> [B,N,1]."""
    if bn:
        raise NotImplementedError("bn=True is not used at inference (test/create_sdf.py:169)")
    return Tensor("pred_sdf_value_global", getattr(src_pc, "graph", None), (batch_size, None, 1))


def get_sdf_basic2_imgfeat_twostream(src_pc, point_feat, is_training, batch_size, num_point, bn, bn_decay, wd=None):
    """models/sdfnet.py:171-190 -- local stream: [B,N,3], [B,N,1,1472] -> [B,N,1]."""
    if bn:
        raise NotImplementedError("bn=True is not used at inference (test/create_sdf.py:169)")
    return Tensor("pred_sdf_value_local", getattr(src_pc, "graph", None), (batch_size, None, 1))
