"""GPU tests of the reference-compatible surface end to end: Session.run == oracle, the literal chunk loop ==
the fused grid call, and create() writes a mesh whose topology equals the CPU marching-cubes oracle's."""
import os

import numpy as np
import pytest

from disn_b200 import synth
from oracle import disn_oracle as orc
from oracle import mc_oracle as mco

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_session_run_matches_oracle_like_the_reference_driver(he_weights, precision):
    from disn_b200 import create_sdf as cs
    from disn_b200 import model_normalization as model
    F = cs.default_flags(sdf_res=8)
    pls = model.placeholder_inputs(1, 1, (137, 137), num_sample_pc=400, scope="inputs_pl", FLAGS=F)
    is_training = model.Placeholder("is_training", ())
    ep = model.get_model(pls, 1, is_training, bn=False, FLAGS=F)
    loss, ep = model.get_loss(ep, sdf_weight=10., num_sample_points=400, FLAGS=F)
    sess = model.Session(weights=he_weights, precision=precision, max_batch=1)
    try:
        imgs = synth.synthetic_images(1, seed=21)
        pts = np.random.default_rng(4).uniform(-1, 1, size=(1, 400, 3)).astype(np.float32)
        gt = np.random.default_rng(5).uniform(-0.2, 0.2, size=(1, 400, 1)).astype(np.float32)
        feed = {is_training: False, pls["sample_pc"]: pts, pls["sample_pc_rot"]: pts, pls["imgs"]: imgs,
                pls["trans_mat"]: synth.DEMO_TRANS_MAT, pls["sdf"]: gt}
        pred, ref_img, uv, emb, acc, real = sess.run(
            [ep["pred_sdf"], ep["ref_img"], ep["sample_img_points"], ep["img_embedding"],
             ep["losses"]["accuracy"], ep["losses"]["sdf_loss_realvalue"]], feed_dict=feed)
        ref = orc.get_model(imgs, pts, pts, synth.DEMO_TRANS_MAT, he_weights, dtype=np.float64)
        assert np.abs(pred - ref["pred_sdf"]).max() / 10.0 <= 1e-4
        np.testing.assert_allclose(uv, ref["sample_img_points"], atol=2e-4)
        np.testing.assert_array_equal(ref_img, imgs)
        assert np.abs(emb - ref["img_embedding"]).max() <= 2e-5 * np.abs(ref["img_embedding"]).max()
        m = orc.get_loss(ref["pred_sdf"], gt)
        assert abs(float(acc) - m["accuracy"]) < 0.01 and abs(float(real) - m["sdf_loss_realvalue"]) < 1e-4
    finally:
        sess.close()


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
def test_create_writes_meshes_and_literal_loop_equals_fused(he_weights, tmp_path, precision):
    from disn_b200 import create_sdf as cs
    from disn_b200 import model_normalization as model
    F = cs.default_flags(sdf_res=20, log_dir=str(tmp_path / "log"), iso=0.0, batch_size=1, precision=precision)
    cs.configure(F)
    imgs = synth.synthetic_images(1, seed=33)
    batch = {"img": imgs, "trans_mat": synth.DEMO_TRANS_MAT, "sdf_params": synth.DEMO_SDF_PARAMS.copy(),
             "cat_id": ["03001627"], "obj_nm": ["synthetic0"], "view_id": [7]}
    # the field of random weights need not cross zero: pick iso = median like SURVEY.md 8d
    sess = model.Session(weights=he_weights, precision=precision, max_batch=1)
    try:
        sess.engine.encode(imgs)
        grid = sess.engine.eval_grid(batch["sdf_params"], batch["trans_mat"], 20)[0]
        # literal reference loop (host grid, chunks, reassembly, /10) == fused device grid
        pls = model.placeholder_inputs(1, 1, (137, 137), num_sample_pc=cs.NUM_SAMPLE_POINTS, FLAGS=F)
        itp = model.Placeholder("is_training", ())
        ep = model.get_model(pls, 1, itp, FLAGS=F)
        pts = cs.build_grid_points(batch["sdf_params"][0])
        lit = sess.run(ep["pred_sdf"], {itp: False, pls["sample_pc"]: pts, pls["sample_pc_rot"]: pts,
                                        pls["imgs"]: imgs, pls["trans_mat"]: batch["trans_mat"]})
        # the reference divides in float64 and packs float32 (create_sdf.py:285,299); the kernel's correctly rounded
        # fp32 division gives the same bits
        np.testing.assert_array_equal((lit.reshape(-1).astype(np.float64) / 10.0).astype(np.float32), grid.reshape(-1))
    finally:
        sess.close()
    F.iso = float(np.median(grid))
    cs.configure(F)
    written = cs.create(he_weights, [batch])
    assert len(written) == 1 and written[0].endswith(os.path.join("03001627", "03001627_synthetic0_07.obj"))
    assert not os.path.exists(written[0][:-4] + ".dist")           # the reference rm's the .dist
    V, Fc = [], []
    for line in open(written[0]):
        if line.startswith("v "):
            V.append([float(x) for x in line.split()[1:]])
        elif line.startswith("f "):
            Fc.append([int(x) - 1 for x in line.split()[1:]])
    rv, rf = mco.marching_cubes(grid, batch["sdf_params"][0], F.iso)
    assert len(V) == len(rv) and len(Fc) == len(rf) > 50
    np.testing.assert_array_equal(np.array(Fc, np.int32), rf)       # integer topology bit-exact
    np.testing.assert_allclose(np.array(V), rv, rtol=2e-6, atol=1e-6)   # %g keeps 6 significant digits
