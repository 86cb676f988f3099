"""Drop-in mirror of the reference's inference driver ``test/create_sdf.py`` (and ``demo/demo.py``) for the
SDF hot path: same module-level constants, same function names and argument meaning
(``create``, ``test_one_epoch``, ``to_binary``, ``create_obj``, ``create_one_cube_obj``), running on the
B200 library.  Dataset loading (``TEST_DATASET``: h5 files on disk) is out of scope: ``create`` takes an
iterable of ``batch_data`` dicts in the loader's layout (data/data_sdf_h5_queue.py:291-303).

    import disn_b200.create_sdf as cs
    cs.configure(FLAGS)                       # the reference does this at import time from argparse
    cs.create(weights, batches)               # -> one .obj per (object, view) under RESULT_OBJ_PATH
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from types import SimpleNamespace

import numpy as np

from . import _lib
from . import model_normalization as model
from .engine import write_dist

# module state, same names as test/create_sdf.py:66-98
FLAGS = None
BATCH_SIZE = RESOLUTION = TOTAL_POINTS = SPLIT_SIZE = NUM_SAMPLE_POINTS = NUM_POINTS = None
SDF_WEIGHT = 10.0
LOG_DIR = RESULT_OBJ_PATH = None
IMG_SIZE = 137
LOG_FOUT = None
_ENGINE = None          # engine used by create_one_cube_obj for the marching-cubes post-pass
_ENGINE_LOCK = threading.Lock()   # a context is not thread-safe; create_obj runs on a 4-worker pool


def default_flags(**kw):
    """argparse defaults of test/create_sdf.py:26-63 / demo/demo.py:27-65."""
    d = dict(gpu="0", img_h=137, img_w=137, batch_size=1, num_classes=1024, num_points=1, sdf_res=64, alpha=False,
             rot=False, tanh=False, multi_view=False, num_sample_points=1, log_dir="checkpoint/SDF_DISN",
             iso=0.0, threedcnn=False, img_feat_onestream=False, img_feat_twostream=True, binary=False,
             cam_est=False, view_num=24, category="all", precision="f16f8")
    d.update(kw)
    return SimpleNamespace(**d)


def configure(flags):
    """Derive the module constants exactly as test/create_sdf.py:66-98 does."""
    global FLAGS, BATCH_SIZE, RESOLUTION, TOTAL_POINTS, SPLIT_SIZE, NUM_SAMPLE_POINTS, NUM_POINTS
    global LOG_DIR, RESULT_OBJ_PATH, IMG_SIZE, LOG_FOUT
    FLAGS = flags
    NUM_POINTS = FLAGS.num_points
    BATCH_SIZE = FLAGS.batch_size
    RESOLUTION = FLAGS.sdf_res + 1
    TOTAL_POINTS = RESOLUTION * RESOLUTION * RESOLUTION
    if FLAGS.img_feat_twostream:
        SPLIT_SIZE = int(np.ceil(TOTAL_POINTS / 214669.0))
    elif FLAGS.threedcnn:
        SPLIT_SIZE = 1
    else:
        SPLIT_SIZE = int(np.ceil(TOTAL_POINTS / 274625.0))
    NUM_SAMPLE_POINTS = int(np.ceil(TOTAL_POINTS / SPLIT_SIZE))
    LOG_DIR = FLAGS.log_dir
    os.makedirs(LOG_DIR, exist_ok=True)
    tag = ("camest_" if FLAGS.cam_est else "") + str(RESOLUTION) + "_" + str(FLAGS.iso)
    RESULT_OBJ_PATH = os.path.join(LOG_DIR, "test_objs", tag)
    os.makedirs(RESULT_OBJ_PATH, exist_ok=True)
    IMG_SIZE = FLAGS.img_h
    LOG_FOUT = open(os.path.join(LOG_DIR, "log_test.txt"), "w")
    LOG_FOUT.write(str(FLAGS) + "\n")


def log_string(out_str):
    LOG_FOUT.write(out_str + "\n")
    LOG_FOUT.flush()
    print(out_str)


def create(weights, batches, device=0):
    """test/create_sdf.py:152-208.  ``weights``: TF-variable-name -> array, or None to mirror the reference's
    behaviour when no checkpoint restores (it then runs on its random initialisation; here that must be
    supplied explicitly -- the encoder refuses to run on absent weights).  ``batches``: iterable of
    batch_data dicts (img [B,137,137,3], trans_mat [B,4,3], sdf_params [B,6], cat_id, obj_nm, view_id)."""
    global _ENGINE
    log_string(LOG_DIR)
    input_pls = model.placeholder_inputs(BATCH_SIZE, NUM_POINTS, (IMG_SIZE, IMG_SIZE),
                                         num_sample_pc=NUM_SAMPLE_POINTS, scope="inputs_pl", FLAGS=FLAGS)
    is_training_pl = model.Placeholder("is_training", ())
    end_points = model.get_model(input_pls, NUM_POINTS, is_training_pl, bn=False, FLAGS=FLAGS)
    loss, end_points = model.get_loss(end_points, sdf_weight=SDF_WEIGHT, num_sample_points=NUM_SAMPLE_POINTS, FLAGS=FLAGS)
    sess = model.Session(device=device, precision=getattr(FLAGS, "precision", "f16f8"), max_batch=max(1, BATCH_SIZE))
    if weights is None:
        print("Fail to load overall modelfile: %s" % LOG_DIR)       # create_sdf.py:192
        raise RuntimeError("no weights supplied: pass the checkpoint variables (or a random init) explicitly")
    sess.load_weights(weights)
    print("Model loaded in file: %s" % LOG_DIR)
    _ENGINE = sess.engine
    ops = {"input_pls": input_pls, "is_training_pl": is_training_pl, "loss": loss, "step": 0,
           "end_points": end_points}
    try:
        return test_one_epoch(sess, ops, batches)
    finally:
        _ENGINE = None
        sess.close()


def build_grid_points(sdf_params_b):
    """test/create_sdf.py:246-255 -- host float64 linspace grid, (x,y,z) float32, x fastest."""
    x_ = np.linspace(sdf_params_b[0], sdf_params_b[3], num=RESOLUTION)
    y_ = np.linspace(sdf_params_b[1], sdf_params_b[4], num=RESOLUTION)
    z_ = np.linspace(sdf_params_b[2], sdf_params_b[5], num=RESOLUTION)
    z, y, x = np.meshgrid(z_, y_, x_, indexing="ij")
    return np.stack([x, y, z], axis=3).astype(np.float32).reshape(1, -1, 3)


def obj_path(dir, cat_id, obj_nm, view_id):
    """file naming of create_obj (test/create_sdf.py:305-312)"""
    if not isinstance(view_id, str):
        view_id = "%02d" % view_id
    dir = os.path.join(dir, cat_id)
    os.makedirs(dir, exist_ok=True)
    return os.path.join(dir, cat_id + "_" + obj_nm + "_" + view_id + ".obj")


def test_one_epoch(sess, ops, batches):
    """test/create_sdf.py:224-289 on the device: per batch one encode + one dense-grid evaluation that STAYS in HBM
    (disn_eval_grid_resident), then per image the CUDA marching-cubes post-pass straight on that buffer and the OBJ writer
    (formatted on a worker thread, like the reference's 4-worker pool for its mesher, create_sdf.py:238,288).
    No .dist round trip through the file system; FLAGS.keep_dist=True also writes the reference's .dist artefact.
    The reference's literal loop (host grid, SPLIT_SIZE chunks through sess.run, reassembly, /SDF_WEIGHT) is kept as a
    verification aid in tests/reference_loop.py."""
    log_string(str(datetime.now()))
    written = []
    R = RESOLUTION
    with ThreadPoolExecutor(max_workers=4) as executor:
        futures = []
        for batch_idx, batch_data in enumerate(batches):
            with _ENGINE_LOCK:
                sess.engine.encode(batch_data["img"])
                sess._img_key = None
                grid_ptr = sess.engine.eval_grid_resident(batch_data["sdf_params"], batch_data["trans_mat"], FLAGS.sdf_res)
                for b in range(BATCH_SIZE):
                    print("{}/{}, submit create_obj {}, {}, {}".format(batch_idx, "?", batch_data["cat_id"][b],
                                                                       batch_data["obj_nm"][b], batch_data["view_id"][b]))
                    path = obj_path(RESULT_OBJ_PATH, batch_data["cat_id"][b], batch_data["obj_nm"][b], batch_data["view_id"][b])
                    dev = grid_ptr + b * R * R * R * 4
                    verts, faces = sess.engine.marching_cubes(None, batch_data["sdf_params"][b], float(FLAGS.iso),
                                                              device_ptr=dev, R=R)
                    if getattr(FLAGS, "keep_dist", False):
                        to_binary(R - 1, batch_data["sdf_params"][b], sess.engine.fetch(dev, (R, R, R)), path[:-4] + ".dist")
                    futures.append(executor.submit(_write_and_return, path, verts, faces))
        for f in futures:
            written.append(f.result())
    return written


def _write_and_return(path, verts, faces):
    write_obj(path, verts, faces)
    return path


def to_binary(res, pos, pred_sdf_val_all, sdf_file):
    """test/create_sdf.py:292-303 -- .dist: int32 -res,res,res; 6 float64; (res+1)^3 float32.  Written by the
    C-ABI writer (the reference struct.packs R^3 Python floats)."""
    write_dist(sdf_file, res, pos, np.asarray(pred_sdf_val_all, dtype=np.float32).reshape(-1))


def create_obj(pred_sdf_val, sdf_params, dir, cat_id, obj_nm, view_id, i):
    """test/create_sdf.py:305-317."""
    cube_obj_file = obj_path(dir, cat_id, obj_nm, view_id)
    sdf_file = cube_obj_file[:-4] + ".dist"
    to_binary((RESOLUTION - 1), sdf_params, pred_sdf_val, sdf_file)
    create_one_cube_obj("./isosurface/computeMarchingCubes", i, sdf_file, cube_obj_file)
    if os.path.exists(sdf_file):
        os.remove(sdf_file)                   # the reference shells out to `rm -rf`
    return cube_obj_file


def read_dist(sdf_file):
    """Parse a .dist (layout pinned by preprocessing/create_point_sdf_grid.py:29-51)."""
    with open(sdf_file, "rb") as f:
        hdr = np.frombuffer(f.read(12), dtype=np.int32)
        if not (hdr[0] < 0 and hdr[1] == -hdr[0] and hdr[2] == -hdr[0]):
            raise ValueError("%s: not a float32 cubic .dist file" % sdf_file)
        res = int(-hdr[0])
        bbox = np.frombuffer(f.read(48), dtype=np.float64)
        vals = np.frombuffer(f.read(), dtype=np.float32)
    R = res + 1
    if vals.size != R ** 3:
        raise ValueError("%s: expected %d samples, found %d" % (sdf_file, R ** 3, vals.size))
    return res, bbox, vals.reshape(R, R, R)


def write_obj(path, verts, faces):
    """OBJ with the reference output's conventions (demo/result.obj: `v %g %g %g`, 1-based `f`), via the C ABI."""
    v = np.ascontiguousarray(verts, np.float32)
    f = np.ascontiguousarray(faces, np.int32)
    _lib.check(_lib.load().disn_write_obj(path.encode(), v.ctypes.data_as(C.c_void_p), len(v),
                                          f.ctypes.data_as(C.c_void_p), len(f)))


def create_one_cube_obj(marching_cube_command, i, sdf_file, cube_obj_file):
    """test/create_sdf.py:319-323.  The reference runs `<marching_cube_command> <dist> <obj> -i <iso>` (a
    closed-source CPU binary); here the .dist is meshed by the CUDA marching-cubes post-pass.
    ``marching_cube_command`` is accepted for signature compatibility and ignored."""
    from .engine import Engine
    res, bbox, sdf = read_dist(sdf_file)
    eng = _ENGINE
    own = eng is None
    if own:
        eng = Engine(device=0, precision="fp32")
    try:
        with _ENGINE_LOCK:
            verts, faces = eng.marching_cubes(sdf, bbox, float(i))
    finally:
        if own:
            eng.close()
    write_obj(cube_obj_file, verts, faces)
    return cube_obj_file
