"""Host logic of the calibration tool (anakin_b200/calibrate.py; the reference's calibrator.h / entropy_calibrator.cpp):
KL threshold search, fused-chain bookkeeping, table file format. No GPU: the tensor collection itself is covered by
tests/test_net_gpu.py::test_calibrator_*."""
import os
import tempfile

import numpy as np


def _hist_from(samples, bins=2048):
    mx = np.abs(samples).max()
    ids = np.minimum((np.abs(samples) / (mx / bins)).astype(np.int64), bins - 1)
    return np.bincount(ids, minlength=bins)


def test_kl_threshold_clips_outliers_and_keeps_a_flat_range():
    from anakin_b200 import calibrate as Cal
    rng = np.random.default_rng(0)
    body = rng.normal(0, 1.0, 200000)
    spiky = np.concatenate([body, [40.0, -38.0, 35.0]])          # three far outliers stretch the range 10x
    t_spiky = Cal.kl_threshold(_hist_from(spiky))
    flat = rng.uniform(-1, 1, 200000)
    t_flat = Cal.kl_threshold(_hist_from(flat))
    assert 129 <= t_spiky < Cal.BIN_NUM and 129 <= t_flat < Cal.BIN_NUM
    # the gaussian body ends near |x| = 4.5 of 40 (bin ~230): clipping there beats keeping the empty tail ...
    assert t_spiky < 600, t_spiky
    # ... while a flat distribution must keep (almost) its whole range
    assert t_flat > 1800, t_flat
    assert t_spiky < t_flat


def test_kl_pieces_are_consistent():
    from anakin_b200 import calibrate as Cal
    rng = np.random.default_rng(1)
    ref_p = rng.integers(0, 50, 640).astype(np.int64)
    ref_p[100:110] = 0
    ref_q = Cal._get_ref_q(ref_p, 128)
    assert abs(float(ref_q.sum()) - float(ref_p.sum())) <= 128          # int truncation per quantised bin only
    q = Cal._expand_to_q(ref_p, ref_q)
    assert q.shape == ref_p.shape and (q[100:110] == 0).all()             # empty source bins stay empty
    np.testing.assert_allclose(q.sum(), ref_q.sum(), rtol=1e-3)
    # a distribution is at zero distance from itself (last kept bin empty: that bin is what the tail term spreads)
    ref_p[-1] = 0
    hist = np.concatenate([ref_p, np.zeros(2048 - 640, np.int64)])
    assert abs(Cal._kl_divergence(hist, ref_p.astype(np.float32))) < 1e-6


def test_fused_chains_and_table_file():
    from anakin_b200 import calibrate as Cal
    from anakin_b200 import modelzoo
    g = modelzoo.build("tiny_resnet", batch=1)
    ops = {"input_0": "Input", "conv1": "ConvBatchnormScaleReluPool", "res2a_branch1": "ConvBatchnormScale",
           "res2a_branch2a": "ConvBatchnormScaleRelu", "res2a_branch2b": "ConvBatchnormScaleRelu",
           "res2a_branch2c": "ConvEltwise"}
    ch = Cal.fused_chains(g, ops)
    assert ch["conv1"] == ["conv1", "bn_conv1", "scale_conv1", "conv1_relu", "pool1"]
    assert ch["res2a_branch1"] == ["res2a_branch1", "bn_res2a_branch1", "scale_res2a_branch1"]   # not the eltwise
    assert ch["res2a_branch2c"][:3] == ["res2a_branch2c", "bn_res2a_branch2c", "scale_res2a_branch2c"]
    assert ch["res2a_branch2c"][-2:] == ["res2a", "res2a_relu"]
    c = Cal.Calibrator(g, Cal.BatchStream([]))
    c.scale_map = {"conv1": 0.0243040379, "pool1": 0.0243040379, "fc": 1.5}
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "calib.txt")
        c.write_calibrator(p)
        lines = open(p).read().splitlines()
        assert lines[0].startswith("conv1 ") and len(lines) == 3          # "<name> <scale>", sorted by name
        back = Cal.Calibrator.read_calibrator(p)
    assert back.keys() == c.scale_map.keys()
    for k in back:
        assert abs(back[k] - c.scale_map[k]) <= 1e-8 * max(1.0, c.scale_map[k])
