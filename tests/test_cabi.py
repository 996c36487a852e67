"""CPU: the C-ABI shared library loads without a GPU and exports every symbol the public header
declares; the host-only entry points (layout queries) work; compute entry points are NOT called."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "theia_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(theia_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from theia_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/theia_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in theia_b200/_lib.py"
    assert lib.theia_version() >= 1


def test_missing_library_fails_loudly(monkeypatch):
    from theia_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtheia_b200.so")
    with pytest.raises(_lib.TheiaError):
        _lib.lib()


def test_model_layout_matches_reference_state_dict(lib):
    """state_dict keys / shapes == the reference's, as restated by the oracle."""
    from oracle import theia_oracle as O
    from theia_b200 import RobotVisionFM
    for backbone, tset in (("facebook/deit-tiny-patch16-224", "cdiv"), ("facebook/deit-base-patch16-224", "dinov2"),
                           ("facebook/deit-tiny-patch16-224", "cddsv"), ("nocls-facebook/deit-tiny-patch16-224", "dinov2"),
                           ("reg-facebook/deit-small-patch16-224", "cdiv")):
        cfg = O.make_config(backbone, tset)
        m = RobotVisionFM(backbone=backbone, translator="lconv", target_feature_sizes=dict(cfg.teachers),
                          translator_kwargs={"hidden_size_factor": 1.0})
        sd = m.state_dict()
        want = O.param_shapes(cfg)
        assert set(sd.keys()) == set(want.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(want[k]), k
        # weight-decay grouping of optimizers/utils.py:26-33 depends on ndim / '.bias'
        nd = {k: p.ndim for k, p in m.named_parameters()}
        assert nd["backbone.model.embeddings.position_embeddings"] == 3
        assert ("backbone.model.embeddings.cls_token" in nd) == (not backbone.startswith("nocls-"))
        assert ("backbone.model.embeddings.reg_token" in nd) == backbone.startswith("reg-")
        # q/k/v weights are adjacent in the flat buffer (single fused QKV GEMM, no copies)
        lay = {n: o for n, _, o in m._param_meta}
        p = "backbone.model.encoder.layer.0.attention.attention."
        D = cfg.hidden
        assert lay[p + "key.weight"] - lay[p + "query.weight"] == D * D
        assert lay[p + "value.weight"] - lay[p + "key.weight"] == D * D
        assert lay[p + "key.bias"] - lay[p + "query.bias"] == D


def test_state_dict_roundtrip_and_flat_views():
    import torch
    from oracle import theia_oracle as O
    from theia_b200 import RobotVisionFM
    cfg = O.make_config("facebook/deit-tiny-patch16-224", "dinov2")
    m = RobotVisionFM(backbone="facebook/deit-tiny-patch16-224", target_feature_sizes=dict(cfg.teachers))
    P = O.init_params(cfg, seed=1)
    m.load_state_dict(P)
    for k, v in m.state_dict().items():
        assert torch.equal(v, P[k]), k
    # parameters alias the flat buffer
    n, shape, off = m._param_meta[5]
    assert m._param_list[5].data_ptr() == m._flat.data_ptr() + 4 * off


def test_reference_error_conventions():
    from theia_b200 import RobotVisionFM
    with pytest.raises(NotImplementedError):
        RobotVisionFM(backbone="facebook/not-a-model")
    with pytest.raises(NotImplementedError):
        RobotVisionFM(backbone="facebook/deit-tiny-patch16-224", translator="bogus",
                      target_feature_sizes={"facebook/dinov2-large": (1024, 16, 16)})


def test_cpu_model_refuses_to_compute():
    """No CPU fallback: a forward on a CPU-resident model must raise, not silently run in PyTorch."""
    import torch
    from theia_b200 import RobotVisionFM, _lib
    m = RobotVisionFM(backbone="facebook/deit-tiny-patch16-224",
                      target_feature_sizes={"facebook/dinov2-large": (1024, 16, 16)})
    with pytest.raises(_lib.TheiaError):
        m.forward_feature(torch.zeros((1, 224, 224, 3), dtype=torch.uint8), do_resize=False)


def test_header_is_plain_c():
    """include/theia_b200.h is the boundary a C / cgo / JNI binding would include: it must compile as C99."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "theia_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_teacher_descriptor_host_logic(lib):
    """theia_vit_* host-side checks (no GPU work): workspace carving and the unsupported-geometry errors"""
    from theia_b200 import _lib as L
    layers = (L.VitLayer * 2)()
    d = L.VitDesc()
    d.hidden, d.heads, d.layers, d.mlp = 1024, 16, 2, 4096
    d.tokens, d.patch_off, d.patch_tokens, d.patch_k = 257, 1, 256, 592
    d.ln_eps, d.act, d.final_ln_mode = 1e-6, 0, 1
    d.layer = layers
    B, M = 8, 8 * 257
    lean = lib.theia_vit_workspace_bytes(C.byref(d), B)
    assert 2 * M * (4 * 1024 + 3 * 1024 + 4096) <= lean < 2 * M * (4 * 1024 + 3 * 1024 + 4096) + (1 << 20)  # x0, x1, ln, attn | qkv | act (bf16) + cls rows
    d.residual_f32 = 1
    wide = lib.theia_vit_workspace_bytes(C.byref(d), B)
    assert wide - lean >= 2 * M * 1024 * 2  # the two stream buffers double
    assert lib.theia_vit_workspace_bytes(C.byref(d), 0) < 0
    # geometry the kernels do not cover: rejected before anything is launched
    d.w_patch, d.tok_table = 1, 1  # non-null placeholders: the checks below come first
    for field, value, msg in (("heads", 8, b"head dim"), ("tokens", 300, b"tokens"), ("hidden", 2048, b"")):
        old = getattr(d, field)
        setattr(d, field, value)
        rc = lib.theia_vit_forward(C.byref(d), 1, 1, 1, 1, 0, 0)
        assert rc == 3, (field, rc)  # THEIA_ERR_UNSUPPORTED
        assert msg in lib.theia_last_error()
        setattr(d, field, old)


def test_image_input_host_logic():
    """what `RobotVisionFM._prep_images` accepts (backbones.py:337-339 hands the same things to the HF processor):
    uint8 / float tensors, numpy arrays and PIL lists, HWC or CHW, any extent; host logic only, nothing is computed"""
    import numpy as np
    import torch
    from theia_b200 import RobotVisionFM
    m = RobotVisionFM(backbone="facebook/deit-tiny-patch16-224",
                      target_feature_sizes={"facebook/dinov2-large": (1024, 16, 16)})
    x, chw = m._prep_images(torch.zeros((2, 224, 224, 3), dtype=torch.uint8), True)
    assert tuple(x.shape) == (2, 224, 224, 3) and chw == 0 and x.dtype == torch.uint8
    x, chw = m._prep_images(torch.zeros((2, 3, 300, 240), dtype=torch.uint8), True)
    assert chw == 1 and x.is_contiguous()
    x, chw = m._prep_images(np.zeros((160, 200, 3), np.uint8), False)  # a single image gets a batch axis
    assert tuple(x.shape) == (1, 160, 200, 3) and chw == 0
    try:
        from PIL import Image
        imgs = [Image.fromarray(np.full((64, 48, 3), 7, np.uint8)) for _ in range(3)]
        x, chw = m._prep_images(imgs, True)
        assert tuple(x.shape) == (3, 64, 48, 3) and chw == 0
    except ImportError:
        pass
    x, chw = m._prep_images(torch.rand(2, 224, 224, 3, dtype=torch.float64), False)  # floats: fp32, no resize
    assert x.dtype == torch.float32
    with pytest.raises(NotImplementedError):
        m._prep_images(torch.rand(2, 224, 224, 3), True)
    with pytest.raises(NotImplementedError):
        m._prep_images(torch.zeros((2, 224, 224, 3), dtype=torch.int32), False)
    with pytest.raises(NotImplementedError):
        m._prep_images(torch.zeros((2, 224, 224, 4), dtype=torch.uint8), False)
