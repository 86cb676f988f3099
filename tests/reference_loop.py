"""Verification aid (test infrastructure): the reference driver's inner loop replayed literally against the Session shim --
host float64 linspace grid, zero padding, SPLIT_SIZE chunks of NUM_SAMPLE_POINTS through sess.run with both point
placeholders fed the same array, swapaxes / reshape / trim, division by SDF_WEIGHT in float64
(test/create_sdf.py:241-285).  The product driver (disn_b200/create_sdf.py) evaluates the grid with one device call."""
import numpy as np


def run_literal(cs, sess, ops, batch_data):
    """cs: the configured disn_b200.create_sdf module.  Returns result [B, TOTAL_POINTS, 1] float64 (= pred / 10)."""
    extra_pts = np.zeros((1, cs.SPLIT_SIZE * cs.NUM_SAMPLE_POINTS - cs.TOTAL_POINTS, 3), dtype=np.float32)
    batch_points = np.zeros((cs.SPLIT_SIZE, 0, cs.NUM_SAMPLE_POINTS, 3), dtype=np.float32)
    for b in range(cs.BATCH_SIZE):
        all_pts = cs.build_grid_points(batch_data["sdf_params"][b])
        all_pts = np.concatenate((all_pts, extra_pts), axis=1).reshape(cs.SPLIT_SIZE, 1, -1, 3)
        batch_points = np.concatenate((batch_points, all_pts), axis=1)
    pred_sdf_val_all = np.zeros((cs.SPLIT_SIZE, cs.BATCH_SIZE, cs.NUM_SAMPLE_POINTS, 1))
    for sp in range(cs.SPLIT_SIZE):
        pc = batch_points[sp, ...].reshape(cs.BATCH_SIZE, -1, 3)
        feed_dict = {ops["is_training_pl"]: False, ops["input_pls"]["sample_pc"]: pc, ops["input_pls"]["sample_pc_rot"]: pc,
                     ops["input_pls"]["imgs"]: batch_data["img"], ops["input_pls"]["trans_mat"]: batch_data["trans_mat"]}
        pred_sdf_val, ref_img_val, sample_img_points_val = sess.run(
            [ops["end_points"]["pred_sdf"], ops["end_points"]["ref_img"], ops["end_points"]["sample_img_points"]],
            feed_dict=feed_dict)
        pred_sdf_val_all[sp, :, :, :] = pred_sdf_val
    pred_sdf_val_all = np.swapaxes(pred_sdf_val_all, 0, 1)
    pred_sdf_val_all = pred_sdf_val_all.reshape((cs.BATCH_SIZE, -1, 1))[:, :cs.TOTAL_POINTS, :]
    return pred_sdf_val_all / cs.SDF_WEIGHT
