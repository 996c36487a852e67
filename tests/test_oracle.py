"""CPU: the oracle restatement against the golden fixtures generated from the REAL reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the CUDA
path with the oracle."""
import glob
import os

import pytest
import torch

from oracle import theia_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _sl(t):
    f = t.detach().flatten()
    step = max(1, f.numel() // 4096)
    return f[::step][:4096]


def _cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "tiny_*.pt")) + glob.glob(os.path.join(GOLDEN, "base_*.pt")))


def test_golden_present():
    assert len(_cases()) >= 3
    assert os.path.exists(os.path.join(GOLDEN, "readme_zeros.pt"))


@pytest.mark.parametrize("path", _cases(), ids=lambda p: os.path.basename(p))
def test_oracle_matches_reference_golden(path):
    fx = torch.load(path, weights_only=False)
    cfg = O.make_config(fx["backbone"], fx["teachers"])
    P = O.init_params(cfg, seed=fx["seed"])
    images, targets = O.synthetic_batch(cfg, fx["B"], seed=fx["seed"])
    feat = O.forward_feature(P, images, cfg, **fx["kwargs"])
    assert tuple(feat.shape) == fx["feature"]["shape"]
    torch.testing.assert_close(_sl(feat), fx["feature"]["sample"], rtol=1e-4, atol=1e-5)
    preds, losses, grads = O.distill_step(P, images, targets, cfg, **fx["kwargs"])
    for t, g in fx["pred"].items():
        assert tuple(preds[t].shape) == g["shape"]
        torch.testing.assert_close(_sl(preds[t]), g["sample"], rtol=1e-4, atol=1e-4)
    for k, v in fx["losses"].items():
        assert abs(float(losses[k]) - v) <= 1e-5 * max(1.0, abs(v)), k
    ml = float(O.main_loss(losses))
    assert abs(ml - fx["main_loss"]) <= 1e-5
    gmax = max(fx["grad_l2"].values())
    for k, v in fx["grad_l2"].items():
        mine = grads[k].double().norm().item()
        assert abs(mine - v) <= (1e-2 if "base" in path else 5e-3) * v + 1e-6 * gmax, (k, mine, v)


def test_readme_zeros_quickstart():
    """BASELINE config #1: forward_feature on a zeros image, deit-tiny -> [1,196,192]."""
    fx = torch.load(os.path.join(GOLDEN, "readme_zeros.pt"), weights_only=False)
    cfg = O.make_config("facebook/deit-tiny-patch16-224", "dinov2")
    P = O.init_params(cfg, seed=0)
    f = O.forward_feature(P, torch.zeros((1, 224, 224, 3), dtype=torch.uint8), cfg)
    assert tuple(f.shape) == (1, 196, 192)
    torch.testing.assert_close(_sl(f), fx["feature"]["sample"], rtol=1e-4, atol=1e-5)


def test_any_image_extent_against_reference_golden():
    """the reference's processor takes images of any extent: resize to 256 x 256 and / or centre crop / zero pad to
    224 x 224 (oracle/make_golden.py::anysize ran these through the real reference)"""
    fx = torch.load(os.path.join(GOLDEN, "anysize_tiny.pt"), weights_only=False)
    cfg = O.make_config("facebook/deit-tiny-patch16-224", "dinov2")
    P = O.init_params(cfg, seed=0)
    assert len(fx["cases"]) >= 6
    for c in fx["cases"]:
        H, W = c["H"], c["W"]
        x = torch.randint(0, 256, (2, H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(H * 1000 + W))
        f = O.forward_feature(P, x, cfg, do_resize=c["do_resize"])
        assert tuple(f.shape) == c["feature"]["shape"] == (2, 196, 192)
        torch.testing.assert_close(_sl(f), c["feature"]["sample"], rtol=1e-4, atol=1e-5)


def test_loss_restatement_against_torch_modules():
    """get_loss restated by hand == nn.MSELoss / SmoothL1Loss / CosineEmbeddingLoss (rvfm.py:71-74,153-168)."""
    import torch.nn as nn
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    p = {"a": torch.randn(3, 16, 8, generator=g), "b": 3 * torch.randn(3, 4, 32, generator=g)}
    y = {"a": torch.randn(3, 16, 8, generator=g), "b": torch.randn(3, 4, 32, generator=g)}
    out = O.get_loss(p, y)
    mse, l1, cos = 0, 0, 0
    for t in p:
        mse = mse + nn.MSELoss()(p[t], y[t]) / 2
        l1 = l1 + nn.SmoothL1Loss()(p[t], y[t]) / 2
        pn = F.normalize(p[t].flatten(1), dim=1)
        tn = F.normalize(y[t].flatten(1), dim=1)
        cos = cos + nn.CosineEmbeddingLoss()(pn, tn, torch.ones(3, dtype=torch.int)) / 2
    torch.testing.assert_close(out["mse_loss"], mse)
    torch.testing.assert_close(out["l1_loss"], l1)
    torch.testing.assert_close(out["cos_loss"], cos)


def test_handle_feature_output_errors():
    x = torch.zeros(1, 197, 8)
    assert O.handle_feature_output(x, "cls").shape == (1, 8)
    assert O.handle_feature_output(x, "identity").shape == (1, 197, 8)
    assert O.handle_feature_output(x, "mean_pooling").shape == (1, 8)
    with pytest.raises(NotImplementedError):
        O.handle_feature_output(x, "bogus")
