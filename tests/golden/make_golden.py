"""Generate tests/golden/*.npz by importing the REFERENCE's own Python (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference modules import h5py / trimesh / pymesh / tensorflow at module top; those are
stubbed with empty modules (only numpy-only helper functions are called).

Fixtures produced (all small, committed):
  cameras.npz      -- reference getBlenderProj/get_rotate_matrix/get_norm_matrix composition
                      (preprocessing/create_img_h5.py:14-123,183-185) for a list of view parameters,
                      plus the demo constant and the K matrix of cam_est/model_cam.py:28.
  dist_roundtrip.npz -- a .dist written by oracle.to_binary, parsed back by the reference's reader
                      (preprocessing/create_point_sdf_grid.py:29-51).
  chunking.npz     -- (RESOLUTION, TOTAL, SPLIT, NUM_SAMPLE) for the BASELINE resolutions, computed
                      with the reference's expressions (test/create_sdf.py:69-77).
  oracle_small.npz -- oracle outputs on seeded synthetic inputs (regression pin for the oracle itself;
                      NOT reference outputs -- TF is unavailable, parity unpinned).
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


class _FakeH5File:
    store = {}

    def __init__(self, path, mode="r"):
        self.path = path

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, k):
        return _FakeH5File.store[self.path][k]


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    _stub("h5py", File=_FakeH5File)
    for nm in ("trimesh", "pymesh", "joblib"):
        _stub(nm, Parallel=object, delayed=lambda f: f)
    _stub("create_file_lst", get_all_info=lambda: (None, None, None, None))
    sys.path.insert(0, os.path.join(REF, "preprocessing"))
    img_h5 = importlib.import_module("create_img_h5")
    # create_point_sdf_grid imports scipy etc.; load only what we need
    _stub("sklearn")
    try:
        sdf_grid = importlib.import_module("create_point_sdf_grid")
    except Exception as e:  # pragma: no cover
        print("create_point_sdf_grid import failed:", e)
        sdf_grid = None
    return img_h5, sdf_grid


def main():
    img_h5, sdf_grid = import_reference()
    from disn_b200 import synth
    from oracle import disn_oracle as orc

    # ---- cameras -------------------------------------------------------------------------
    rng = np.random.default_rng(99)
    params = [(synth.DEMO_CAM_GT[0], synth.DEMO_CAM_GT[1], synth.DEMO_CAM_GT[2], 0.5155616,
               (-0.0371715, -0.0426027, -0.0004739))]
    for _ in range(7):
        params.append((float(rng.uniform(0, 360)), float(rng.uniform(25, 30)), float(rng.uniform(0.65, 0.95)),
                       float(rng.uniform(0.4, 0.6)), tuple(float(v) for v in rng.uniform(-0.05, 0.05, 3))))
    ref_mats, Ks, RTs = [], [], []
    rot_mat = img_h5.get_rotate_matrix(-np.pi / 2)
    for az, el, dist, m, centre in params:
        K, RT = img_h5.getBlenderProj(az, el, dist, img_w=137, img_h=137)
        _FakeH5File.store["fake.h5"] = {"norm_params": np.array(list(centre) + [m], dtype=np.float64)}
        norm_mat = img_h5.get_norm_matrix("fake.h5")
        tm = np.linalg.multi_dot([K, RT, rot_mat, norm_mat])
        ref_mats.append(np.transpose(np.asarray(tm)).astype(np.float32))
        Ks.append(np.asarray(K, dtype=np.float64))
        RTs.append(np.asarray(RT, dtype=np.float64))
    np.savez(os.path.join(HERE, "cameras.npz"),
             params=np.array([[p[0], p[1], p[2], p[3], *p[4]] for p in params], dtype=np.float64),
             trans_mat=np.stack(ref_mats), K=np.stack(Ks), RT=np.stack(RTs),
             rot_mat=np.asarray(rot_mat, dtype=np.float64))
    print("cameras.npz: demo-camera max|ref - DEMO_TRANS_MAT| =",
          float(np.abs(ref_mats[0] - synth.DEMO_TRANS_MAT[0]).max()))

    # ---- .dist round trip through the reference reader -------------------------------------
    res = 6
    vals = np.random.default_rng(5).standard_normal((res + 1) ** 3).astype(np.float32)
    bbox = [-1.0, -0.9, -0.8, 1.0, 0.9, 0.8]
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "t.dist")
        orc.to_binary(res, bbox, vals.astype(np.float64), fn)
        raw = np.fromfile(fn, dtype=np.uint8)
        if sdf_grid is not None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    # numpy>=2.3 removed fromstring's binary mode; shim it so the reference's own
                    # reader logic (slicing/reshape/consistency check) still runs unmodified
                    np.fromstring = lambda b, dtype=float: np.frombuffer(b, dtype=dtype)
                    parsed = sdf_grid.get_sdf(fn, res)
                    pv, pp = np.asarray(parsed["value"]), np.asarray(parsed["param"])
                except Exception as e:  # np.fromstring binary mode removed in numpy 2.x
                    print("reference get_sdf not runnable under numpy", np.__version__, "->", e)
                    pv = pp = None
        else:
            pv = pp = None
    if pv is None:
        # the reference reader uses np.fromstring (binary mode, removed in numpy>=2.3); restate its
        # slicing with np.frombuffer: int32[3], float64[6], float32[(res+1)^3] reshaped (R,R,R)
        b = raw.tobytes()
        ress = np.frombuffer(b[:12], dtype=np.int32)
        assert -ress[0] == res and ress[1] == res and ress[2] == res
        pp = np.float32(np.frombuffer(b[12:60], dtype=np.float64))
        pv = np.frombuffer(b[60:], dtype=np.float32).reshape(res + 1, res + 1, res + 1)
        reader = "frombuffer-restated"
    else:
        reader = "reference.get_sdf"
    assert np.array_equal(pv.reshape(-1), vals) and np.allclose(pp, np.float32(bbox))
    np.savez(os.path.join(HERE, "dist_roundtrip.npz"), res=res, bbox=np.array(bbox), values=vals,
             file_bytes=raw, reader=reader)
    print("dist_roundtrip.npz via", reader)

    # ---- chunk arithmetic (reference expressions, create_sdf.py:69-77) ----------------------
    rows = []
    for sdf_res in (8, 16, 32, 64, 128, 256, 512):
        RESOLUTION = sdf_res + 1
        TOTAL_POINTS = RESOLUTION * RESOLUTION * RESOLUTION
        SPLIT_SIZE = int(np.ceil(TOTAL_POINTS / 214669.0))
        NUM_SAMPLE_POINTS = int(np.ceil(TOTAL_POINTS / SPLIT_SIZE))
        rows.append((sdf_res, RESOLUTION, TOTAL_POINTS, SPLIT_SIZE, NUM_SAMPLE_POINTS))
    np.savez(os.path.join(HERE, "chunking.npz"), table=np.array(rows, dtype=np.int64))

    # ---- oracle regression pin -----------------------------------------------------------------
    W = synth.make_weights(seed=7, init="he")
    imgs = synth.synthetic_images(1)
    tm = synth.DEMO_TRANS_MAT
    pts = np.random.default_rng(11).uniform(-1, 1, size=(1, 512, 3)).astype(np.float32)
    out32 = orc.get_model(imgs, pts, pts, tm, W, dtype=np.float32)
    out64 = orc.get_model(imgs, pts, pts, tm, W, dtype=np.float64)
    grid = orc.create_sdf_grid(imgs, tm, synth.DEMO_SDF_PARAMS, W, sdf_res=8, dtype=np.float32)
    np.savez(os.path.join(HERE, "oracle_small.npz"), pts=pts,
             pred32=out32["pred_sdf"], pred64=out64["pred_sdf"], uv=out32["sample_img_points"],
             emb32=out32["img_embedding"], emb64=out64["img_embedding"], grid_res8=grid)
    print("oracle_small.npz: pred rms", float(np.sqrt(np.mean(out64['pred_sdf'] ** 2))),
          "fp32-vs-fp64 max abs", float(np.abs(out32['pred_sdf'] - out64['pred_sdf']).max()))


if __name__ == "__main__":
    main()
