"""Target ingest on the GPU.

The reference's dataloader workers decode each teacher embedding from safetensors as `[C,H,W]` bf16, rearrange it
to `(h w) c` and z-score it with the per-channel ImageNet statistics (`dataset/data_utils.py:152-153,342-355`) on
the CPU, then `train_rvfm.py:112-114` uploads it and upcasts to fp32.  `ingest_targets` does the transpose and the
normalisation in one kernel on the device (bit-exact with the reference's bf16 arithmetic) and keeps the result in
bf16, which `RobotVisionFM.get_loss` accepts directly."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


def ingest_targets(embedding_chw: torch.Tensor, mean: Optional[torch.Tensor] = None,
                   std: Optional[torch.Tensor] = None) -> torch.Tensor:
    """embedding_chw: bf16 [B, C, H, W] (or [B, C, H*W]) on the GPU; mean/std: [C].  Returns bf16 [B, H*W, C]."""
    if not embedding_chw.is_cuda:
        raise L.TheiaError("ingest_targets needs CUDA tensors (no CPU fallback)")
    x = embedding_chw.to(torch.bfloat16).contiguous()
    B, C = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    out = torch.empty((B, HW, C), dtype=torch.bfloat16, device=x.device)
    m = None if mean is None else mean.to(device=x.device, dtype=torch.bfloat16).contiguous()
    s = None if std is None else std.to(device=x.device, dtype=torch.bfloat16).contiguous()
    L.check(L.lib().theia_target_ingest(x.data_ptr(), L.ptr(m), L.ptr(s), out.data_ptr(), B, C, HW, L.stream_ptr()),
            "theia_target_ingest")
    return out
