"""Python handle over the C-ABI context: one Engine = one CUDA device + one stream (not thread-safe)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DISN_DEVICE_PTR, PREC_BF16X3, PREC_F16F8, PREC_FP32, DisnConfig, check

_PREC = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "f16f8": PREC_F16F8}
TAP_HW = (224, 112, 56, 28, 14)
TAP_C = (64, 128, 256, 512, 512)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Engine:
    def __init__(self, device: int = 0, precision: str = "fp32", max_batch: int = 1, tanh: bool = False,
                 img_h: int = 137, img_w: int = 137, num_classes: int = 1024, sdf_weight: float = 10.0):
        self.lib = _lib.load()
        cfg = DisnConfig()
        self.lib.disn_default_config(C.byref(cfg))
        cfg.device = device
        cfg.precision = _PREC[precision]
        cfg.max_batch = max_batch
        cfg.tanh_out = int(bool(tanh))
        cfg.img_h, cfg.img_w, cfg.num_classes = img_h, img_w, num_classes
        cfg.sdf_weight = sdf_weight
        cfg.clamp_max = float(img_h - 1)      # models/model_normalization.py:250 (136 for 137x137)
        self.cfg = cfg
        self._h = C.c_void_p()
        check(self.lib.disn_create(C.byref(cfg), C.byref(self._h)))
        self.batch = 0

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.disn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        check(self.lib.disn_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(self.lib.disn_synchronize(self._h))

    def set_precision(self, precision: str):
        check(self.lib.disn_set_precision(self._h, _PREC[precision]))

    @property
    def launch_count(self) -> int:
        return int(self.lib.disn_launch_count(self._h))

    # -- weights ------------------------------------------------------------------------------
    def load_weights(self, weights: dict):
        """weights: TF variable name -> array (HWIO).  Unknown names are stored but unused."""
        for name, arr in weights.items():
            a = _f32(arr)
            shp = (C.c_int64 * a.ndim)(*a.shape)
            check(self.lib.disn_load_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shp, a.ndim))
        check(self.lib.disn_finalize_weights(self._h))

    # -- encoder ------------------------------------------------------------------------------
    def encode(self, imgs):
        a = _f32(imgs)
        if a.ndim != 4:
            raise ValueError("imgs must be [B,H,W,C]")
        B, H, W, Cc = a.shape
        check(self.lib.disn_encode(self._h, a.ctypes.data_as(C.c_void_p), B, H, W, Cc, 0))
        self.batch = B

    def encode_device(self, imgs_ptr: int, B: int, H: int, W: int, Cc: int = 3):
        """Device-pointer, asynchronous variant."""
        check(self.lib.disn_encode(self._h, C.c_void_p(imgs_ptr), B, H, W, Cc, DISN_DEVICE_PTR))
        self.batch = B

    def get_encoded(self, what: int) -> np.ndarray:
        B = self.batch
        cfg = self.cfg
        if what == 0:
            shape = (B, cfg.num_classes)
        elif 1 <= what <= 5:
            shape = (B, TAP_HW[what - 1], TAP_HW[what - 1], TAP_C[what - 1])
        elif what == 6:
            shape = (B, cfg.img_h, cfg.img_w, 512)
        elif what == 7:
            shape = (B, 512)
        elif what == 8:
            shape = (B, cfg.vgg_in, cfg.vgg_in, 3)
        else:
            raise ValueError(what)
        out = np.empty(shape, dtype=np.float32)
        check(self.lib.disn_get_encoded(self._h, what, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    # -- points -------------------------------------------------------------------------------
    def eval_points(self, pts, trans_mat, pts_rot=None, want_uv: bool = False):
        """One `sess.run`: pts [B,N,3], trans_mat [B,4,3] -> pred_sdf [B,N,1] (and uv [B,N,2])."""
        p = _f32(pts)
        t = _f32(trans_mat)
        B, N, _ = p.shape
        pr = None if pts_rot is None else _f32(pts_rot)
        out = np.empty((B, N, 1), dtype=np.float32)
        uv = np.empty((B, N, 2), dtype=np.float32) if want_uv else None
        check(self.lib.disn_eval_points(
            self._h, p.ctypes.data_as(C.c_void_p), None if pr is None else pr.ctypes.data_as(C.c_void_p),
            t.ctypes.data_as(C.c_void_p), B, N, out.ctypes.data_as(C.c_void_p),
            None if uv is None else uv.ctypes.data_as(C.c_void_p), 0))
        return (out, uv) if want_uv else out

    def eval_points_ex(self, pts, trans_mat, pts_rot=None):
        """pred_sdf, sample_img_points, pred_sdf_value_global, pred_sdf_value_local (model_normalization.py:194-206)."""
        p, t = _f32(pts), _f32(trans_mat)
        B, N, _ = p.shape
        pr = None if pts_rot is None else _f32(pts_rot)
        out, g, l = (np.empty((B, N, 1), np.float32) for _ in range(3))
        uv = np.empty((B, N, 2), np.float32)
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        check(self.lib.disn_eval_points_ex(self._h, ptr(p), ptr(pr), ptr(t), B, N, ptr(out), ptr(uv), ptr(g), ptr(l)))
        return out, uv, g, l

    def point_img_feat(self, pts, trans_mat):
        """end_points['point_img_feat'] [B,N,1,1472] and sample_img_points [B,N,2] (model_normalization.py:170-190)."""
        p, t = _f32(pts), _f32(trans_mat)
        B, N, _ = p.shape
        feat = np.empty((B, N, 1, 1472), np.float32)
        uv = np.empty((B, N, 2), np.float32)
        check(self.lib.disn_point_img_feat(self._h, p.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), B, N,
                                           feat.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p)))
        return feat, uv

    def eval_features(self, pts_rot, global_feat, point_feat):
        """get_decoder (model_normalization.py:223-238): explicit [B,1,1,1024] / [B,N,1,1472] features ->
        (multi_pred_sdf, pred_sdf_value_global, pred_sdf_value_local), each [B,N,1]."""
        p = _f32(pts_rot)
        B, N, _ = p.shape
        g = _f32(global_feat).reshape(B, -1)
        f = _f32(point_feat).reshape(B, N, 1472)
        out, og, ol = (np.empty((B, N, 1), np.float32) for _ in range(3))
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        check(self.lib.disn_eval_features(self._h, ptr(p), ptr(g), ptr(f), B, N, ptr(out), ptr(og), ptr(ol), 0))
        return out, og, ol

    def eval_points_device(self, pts_ptr: int, trans_mat_ptr: int, B: int, N: int, out_ptr: int,
                           uv_ptr: int = 0, pts_rot_ptr: int = 0):
        """Device-pointer, asynchronous variant (pointers from torch tensors' data_ptr())."""
        check(self.lib.disn_eval_points(self._h, C.c_void_p(pts_ptr), C.c_void_p(pts_rot_ptr or 0),
                                        C.c_void_p(trans_mat_ptr), B, N, C.c_void_p(out_ptr),
                                        C.c_void_p(uv_ptr or 0), DISN_DEVICE_PTR))

    def eval_grid(self, sdf_params, trans_mat, sdf_res: int, z0: int = 0, z1: int | None = None, out=None):
        """Dense grid slab -> [B, z1-z0, R, R] float32 = pred/sdf_weight (reference `result`, reshaped)."""
        sp = np.ascontiguousarray(sdf_params, dtype=np.float64).reshape(-1, 6)
        t = _f32(trans_mat)
        B = sp.shape[0]
        R = sdf_res + 1
        z1 = R if z1 is None else z1
        if out is None:
            out = np.empty((B, z1 - z0, R, R), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == B * (z1 - z0) * R * R
        check(self.lib.disn_eval_grid(self._h, sp.ctypes.data_as(C.POINTER(C.c_double)),
                                      t.ctypes.data_as(C.c_void_p), B, sdf_res, z0, z1,
                                      out.ctypes.data_as(C.c_void_p), 0))
        return out

    def eval_grid_device(self, sdf_params, trans_mat_ptr: int, sdf_res: int, z0: int, z1: int, out_ptr: int):
        sp = np.ascontiguousarray(sdf_params, dtype=np.float64).reshape(-1, 6)
        check(self.lib.disn_eval_grid(self._h, sp.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(trans_mat_ptr),
                                      sp.shape[0], sdf_res, z0, z1, C.c_void_p(out_ptr), DISN_DEVICE_PTR))

    # -- estimated camera ------------------------------------------------------------------------
    def cam_estimate(self, imgs, K=None, want_rt: bool = False):
        """demo/demo.py:195-258 cam_evl: imgs [B,H,W,3] -> pred_trans_mat [B,4,3] (this engine must hold the camera
        checkpoint's variables: vgg_16/* and cameraprediction/*).  load_weights_raw() skips the SDF-head checks."""
        a = _f32(imgs)
        B, H, W, Cc = a.shape
        tm = np.empty((B, 4, 3), np.float32)
        rt = np.empty((B, 4, 3), np.float32) if want_rt else None
        kk = None if K is None else _f32(K).reshape(9)
        check(self.lib.disn_cam_estimate(self._h, a.ctypes.data_as(C.c_void_p), B, H, W, Cc,
                                         None if kk is None else kk.ctypes.data_as(C.c_void_p),
                                         None if rt is None else rt.ctypes.data_as(C.c_void_p),
                                         tm.ctypes.data_as(C.c_void_p)))
        return (tm, rt) if want_rt else tm

    def load_weights_raw(self, weights: dict):
        """Upload variables without finalising the SDF heads (camera-net contexts have no sdfprediction/*)."""
        for name, arr in weights.items():
            a = _f32(arr)
            shp = (C.c_int64 * a.ndim)(*a.shape)
            check(self.lib.disn_load_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), shp, a.ndim))

    # -- mesh metrics -------------------------------------------------------------------------
    def nn_distance(self, xyz1, xyz2):
        """The reference's tf_nndistance.nn_distance(xyz1, xyz2): squared NN distances + indices, both ways."""
        a, b = _f32(xyz1), _f32(xyz2)
        B, N, _ = a.shape
        M = b.shape[1]
        d1, i1 = np.empty((B, N), np.float32), np.empty((B, N), np.int32)
        d2, i2 = np.empty((B, M), np.float32), np.empty((B, M), np.int32)
        check(self.lib.disn_nn_distance(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), B, N, M,
                                        d1.ctypes.data_as(C.c_void_p), i1.ctypes.data_as(C.c_void_p),
                                        d2.ctypes.data_as(C.c_void_p), i2.ctypes.data_as(C.c_void_p)))
        return d1, i1, d2, i2

    def chamfer_x1000(self, pred, src):
        """test/test_cd_emd.py:300-301: (mean forward + mean backward squared NN distance) * 1000, per batch item."""
        df, _, db, _ = self.nn_distance(pred, src)
        return (df.mean(axis=1) + db.mean(axis=1)) * np.float32(1000)

    def approx_match(self, xyz1, xyz2, cost=False):
        """The reference's tf_approxmatch.approx_match(xyz1, xyz2) (models/tf_ops/approxmatch/tf_approxmatch.py:12-20):
        match [B,N,M] float32, element (k,l) = mass moved from point k of xyz1 to point l of xyz2.  cost=True also returns
        match_cost(xyz1, xyz2, match) [B] from the same call."""
        a, b = _f32(xyz1), _f32(xyz2)
        B, N, _ = a.shape
        M = b.shape[1]
        m = np.empty((B, N, M), np.float32)
        cst = np.empty(B, np.float32) if cost else None
        check(self.lib.disn_approx_match(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), B, N, M,
                                         m.ctypes.data_as(C.c_void_p), cst.ctypes.data_as(C.c_void_p) if cost else None))
        return (m, cst) if cost else m

    def match_cost(self, xyz1, xyz2, match):
        """tf_approxmatch.match_cost(xyz1, xyz2, match) (tf_approxmatch.py:28-37): cost [B]."""
        a, b, m = _f32(xyz1), _f32(xyz2), _f32(match)
        B, N, _ = a.shape
        M = b.shape[1]
        assert m.shape == (B, N, M), m.shape
        cst = np.empty(B, np.float32)
        check(self.lib.disn_match_cost(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                       m.ctypes.data_as(C.c_void_p), B, N, M, cst.ctypes.data_as(C.c_void_p)))
        return cst

    def emd(self, src, pred):
        """test/test_cd_emd.py:307-308: match_cost(src, pred, approx_match(src, pred)) * 0.01 per batch item; the match
        matrix stays in HBM."""
        a, b = _f32(src), _f32(pred)
        B, N, _ = a.shape
        cst = np.empty(B, np.float32)
        check(self.lib.disn_approx_match(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), B, N, b.shape[1],
                                         None, cst.ctypes.data_as(C.c_void_p)))
        return cst * np.float32(0.01)

    def points_loss(self, sampled_pc):
        """get_points_loss of test/test_cd_emd.py:291-315: sampled_pc [1+V,N,3] = the ground-truth cloud followed by the
        clouds of V predicted views.  Returns (avg_cf, min_cf, arg_min_cf, avg_em, min_em, arg_min_em) over the views:
        Chamfer x1000 and approximate EMD x0.01 of every view against the ground truth."""
        pc = _f32(sampled_pc)
        pred = pc[1:]
        src = np.ascontiguousarray(np.broadcast_to(pc[:1], pred.shape))
        cf = self.chamfer_x1000(pred, src)
        em = self.emd(src, pred)
        return (np.float32(cf.mean()), np.float32(cf.min()), int(cf.argmin()),
                np.float32(em.mean()), np.float32(em.min()), int(em.argmin()))

    def f_score(self, pred, src, thresholds):
        """test/test_f_score.py:231-236: precision / recall = fraction of sqrt NN distances (pred->src / src->pred)
        below each threshold; F = 2PR/(P+R).  pred, src: [1,N,3] / [1,M,3].  Distances come from the CUDA NN kernel."""
        df, _, db, _ = self.nn_distance(pred, src)
        th = np.asarray(thresholds, np.float32)[:, None]
        p = (np.sqrt(df).reshape(1, -1) < th).mean(axis=1)
        r = (np.sqrt(db).reshape(1, -1) < th).mean(axis=1)
        return p, r, 2 * p * r / np.maximum(p + r, 1e-30)

    # -- marching cubes -----------------------------------------------------------------------
    def marching_cubes(self, sdf, bbox, iso: float = 0.0, device_ptr: int | None = None, R: int | None = None,
                       fetch: bool = True):
        """sdf [R,R,R] (z,y,x) host array, or a device pointer -> (verts [V,3] float32, faces [F,3] int32 0-based).
        The welded mesh also stays in HBM (write_mesh_obj); fetch=False returns only the counts."""
        bb = (C.c_double * 6)(*[float(v) for v in bbox])
        if device_ptr is None:
            a = _f32(sdf)
            R = a.shape[0]
            assert a.shape == (R, R, R)
            ptr, flags = a.ctypes.data_as(C.c_void_p), 0
        else:
            ptr, flags = C.c_void_p(device_ptr), DISN_DEVICE_PTR
        nv, nf = C.c_int64(0), C.c_int64(0)
        check(self.lib.disn_mc_run(self._h, ptr, R, bb, float(iso), flags, C.byref(nv), C.byref(nf)))
        if not fetch:
            return nv.value, nf.value
        verts = np.empty((nv.value, 3), dtype=np.float32)
        faces = np.empty((nf.value, 3), dtype=np.int32)
        if nv.value and nf.value:
            check(self.lib.disn_mc_fetch(self._h, verts.ctypes.data_as(C.c_void_p), faces.ctypes.data_as(C.c_void_p)))
        return verts, faces

    def write_mesh_obj(self, path: str):
        """OBJ of the mesh left in HBM by the last marching_cubes call (reference mesher's output conventions)."""
        check(self.lib.disn_mc_write_obj(self._h, path.encode()))

    def eval_grid_resident(self, sdf_params, trans_mat, sdf_res: int) -> int:
        """Whole [B,R,R,R] grid evaluated into the context's HBM buffer; returns its device address."""
        sp = np.ascontiguousarray(sdf_params, dtype=np.float64).reshape(-1, 6)
        t = _f32(trans_mat)
        out = C.c_void_p()
        check(self.lib.disn_eval_grid_resident(self._h, sp.ctypes.data_as(C.POINTER(C.c_double)),
                                               t.ctypes.data_as(C.c_void_p), sp.shape[0], sdf_res, C.byref(out)))
        return int(out.value)

    def fetch(self, dev_ptr: int, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        check(self.lib.disn_fetch(self._h, C.c_void_p(dev_ptr), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    # -- cross-process device buffer (peer-store gather) --------------------------------------------
    def shared_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle): rank 0's whole-grid buffer the other ranks' kernels store into."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        check(self.lib.disn_shared_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), handle.raw

    def shared_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        check(self.lib.disn_shared_open(self._h, C.create_string_buffer(handle, 64), C.byref(ptr)))
        return int(ptr.value)

    def shared_close(self, ptr: int, owner: bool):
        check(self.lib.disn_shared_close(self._h, C.c_void_p(ptr), int(owner)))

    def iou(self, verts1, faces1, verts2, faces2, dim: int = 110, want_grids: bool = False):
        """test/test_iou.py:208-233 iou_pymesh on two triangle meshes -> IoU (and the two occupancy grids)."""
        v1, v2 = _f32(verts1), _f32(verts2)
        f1, f2 = np.ascontiguousarray(faces1, np.int32), np.ascontiguousarray(faces2, np.int32)
        inter, uni = C.c_int64(0), C.c_int64(0)
        o1 = np.empty((dim, dim, dim), np.uint8) if want_grids else None
        o2 = np.empty((dim, dim, dim), np.uint8) if want_grids else None
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        check(self.lib.disn_iou(self._h, ptr(v1), len(v1), ptr(f1), len(f1), ptr(v2), len(v2), ptr(f2), len(f2), dim,
                                C.byref(inter), C.byref(uni), ptr(o1), ptr(o2)))
        val = inter.value / uni.value if uni.value else float("nan")
        return (val, inter.value, uni.value, o1, o2) if want_grids else val


def write_dist(path: str, res: int, bbox, values):
    """C-ABI .dist writer (test/create_sdf.py:292-303 layout)."""
    v = _f32(values).reshape(-1)
    if v.size != (res + 1) ** 3:
        raise ValueError("values must hold (res+1)^3 samples")
    bb = (C.c_double * 6)(*[float(x) for x in bbox])
    check(_lib.load().disn_write_dist(path.encode(), res, bb, v.ctypes.data_as(C.c_void_p)))
