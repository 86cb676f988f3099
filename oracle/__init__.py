"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DISN SDF-inference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference's arithmetic lives in TensorFlow 1.x (tf.contrib.slim
vgg_16, tf.image.resize_bilinear, tf.contrib.resampler) which is not vendored in
/root/reference and cannot be installed here; the reference ships no tests or golden
outputs for this path.  What *is* pinned against the reference (tests/golden/, made by
tests/golden/make_golden.py importing the reference's own Python where it is importable):
camera matrices (preprocessing/create_img_h5.py), the demo trans_mat constant
(demo/demo.py:272-276), the .dist reader (preprocessing/create_point_sdf_grid.py:29-51)
and the chunking constants (test/create_sdf.py:69-77).
"""
