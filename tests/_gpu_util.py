"""Test helpers: read the CUDA path's internal activations through the bring-up accessor."""
import ctypes as C

import torch

from theia_b200 import _lib as L


def from_ptr(p, shape, dtype=torch.bfloat16):
    n = torch.Size(shape).numel()

    class W:
        pass

    w = W()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u2", "data": (p, False), "version": 2}
    return torch.as_tensor(w, device="cuda").view(dtype).view(shape).clone()


def fetch(m, name, i, shape):
    ptr, n, f32 = C.c_void_p(), C.c_longlong(), C.c_int()
    L.check(L.lib().theia_model_debug_ptr(m._handle, name.encode(), i, C.byref(ptr), C.byref(n), C.byref(f32)), name)
    assert n.value == torch.Size(shape).numel(), (name, n.value, shape)
    return from_ptr(ptr.value, shape)


def fetch_all(m, cfg, B):
    """every stored forward activation of the last forward, keyed like the oracle's taps"""
    D, NT = cfg.hidden, cfg.seq
    out = {("x", 0): fetch(m, "x", 0, (B, NT, D)), ("tokens", 0): fetch(m, "tokens", 0, (B, NT, D))}
    for l in range(cfg.layers):
        # ("h" holds gelu'(pre-activation) in the CUDA path, not the pre-activation: it is not forced)
        for name, w in (("ln1", D), ("qkv", 3 * D), ("attn", D), ("xmid", D), ("ln2", D), ("a", 4 * D)):
            out[(name, l)] = fetch(m, name, l, (B, NT, w))
        out[("x", l + 1)] = fetch(m, "x", l + 1, (B, NT, D))
    for i, (t, size) in enumerate(cfg.teachers.items()):
        if len(size) == 1:
            continue  # CLS head: a single Linear, nothing stored
        ht = size[1]
        v1, p1, v2 = (16, 16, 16) if ht == 16 else (31, 32, 64)
        for name in ("padout", "hln0"):
            out[(name, t)] = fetch(m, name, i, (B, 16, 16, D))
        for name in ("c1", "hln1"):  # the 31x31 stage of the 64x64 heads is stored at pitch 32 (zero padded)
            out[(name, t)] = fetch(m, name, i, (B, p1, p1, D))[:, :v1, :v1].contiguous()
        for name in ("c2", "hln2"):
            out[(name, t)] = fetch(m, name, i, (B, v2, v2, D))
    return out
