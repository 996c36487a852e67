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
