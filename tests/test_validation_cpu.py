"""Validation glue (unidepth_b200/validation.py) against outputs of the reference's own functions stored in
tests/golden/validation_glue.npz (made by oracle/make_golden.py from unidepth/utils/misc.py:596-690 and
unidepth/utils/evaluation_depth.py), and the network-only geometry of the C engine."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _z():
    return np.load(os.path.join(ROOT, "tests", "golden", "validation_glue.npz"))


def test_match_gt_and_intrinsics_match_reference_outputs():
    from unidepth_b200.validation import match_gt, match_intrinsics
    z = _z()
    pred, gt, K = (torch.from_numpy(z[k]) for k in ("pred", "gt", "K"))
    p1 = [tuple(r) for r in z["pads1"]]
    p2 = [tuple(r) for r in z["pads2"]]
    img = torch.zeros(3, 3, 28, 42)
    for got, ref in ((match_gt(pred, gt, None, None), z["m_none"]), (match_gt(pred, gt, p1, None), z["m_p1"]),
                     (match_gt(pred, gt, p1, p2), z["m_p12"])):
        assert got.shape == ref.shape and torch.allclose(got, torch.from_numpy(ref), atol=1e-6, rtol=0)
    for got, ref in ((match_intrinsics(K, img, gt, p1, None), z["k_p1"]), (match_intrinsics(K, img, gt, p1, p2), z["k_p12"])):
        assert torch.allclose(got, torch.from_numpy(ref), atol=1e-5, rtol=1e-6)


def test_depth_metrics_match_reference_outputs():
    from unidepth_b200.validation import depth_metrics
    z = _z()
    gt, pred = torch.from_numpy(z["met_gt"]), torch.from_numpy(z["met_pred"])
    got = depth_metrics(gt, pred)
    for name in ("d1", "d2", "d3", "rmse", "rmselog", "arel", "sqrel", "log10", "silog"):
        ref = float(z["met_" + name])
        assert abs(got[name] - ref) <= 2e-5 * max(1.0, abs(ref)), (name, got[name], ref)


def test_network_only_geometry_is_identity():
    from unidepth_b200 import _cabi
    lib = _cabi.lib()
    cfg = _cabi.Config()
    cfg.embed_dim, cfg.depth, cfg.enc_heads, cfg.pos_grid = 1024, 24, 16, 37
    cfg.hidden, cfg.dec_heads, cfg.expansion, cfg.out_dim, cfg.n_stages = 512, 8, 4, 64, 3
    cfg.ratio_min, cfg.ratio_max, cfg.pixels_min, cfg.pixels_max = 0.5, 2.5, 200000.0, 600000.0
    h = C.c_void_p()
    assert lib.udb_create(C.byref(cfg), C.byref(h)) == 0
    try:
        g = _cabi.Geometry()
        # far outside the pixel / ratio bounds on purpose: nothing is padded or resized in this mode
        assert lib.udb_geometry(h, 14 * 9, 14 * 70, -2, C.byref(g)) == 0
        assert (g.pad_l, g.pad_r, g.pad_t, g.pad_b) == (0, 0, 0, 0) and g.factor == 1.0
        assert (g.padded_h, g.padded_w, g.net_h, g.net_w, g.gh, g.gw) == (126, 980, 126, 980, 9, 70)
        assert lib.udb_geometry(h, 125, 980, -2, C.byref(g)) != 0 and b"multiple of 14" in lib.udb_last_error()
    finally:
        lib.udb_destroy(h)


def test_forward_dispatch_has_no_cpu_or_training_path():
    import json
    from unidepth_b200 import UniDepthV2
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "config_v2_vits14.json")))
    m = UniDepthV2(cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m({"image": torch.zeros(1, 3, 28, 42), "depth": torch.ones(1, 1, 28, 42)}, [])
    with pytest.raises(RuntimeError, match="no CPU"):
        m.network_forward(torch.zeros(1, 3, 28, 42))
    with pytest.raises(ValueError, match="multiple of 14"):
        m.network_forward(torch.zeros(1, 3, 30, 42))
    m.train()
    with pytest.raises(NotImplementedError):
        m({"image": torch.zeros(1, 3, 28, 42)}, [])
