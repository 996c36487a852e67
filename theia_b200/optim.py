"""Fused optimizer tail for `RobotVisionFM`.

`train_rvfm.py:126-133` runs `clip_grad_norm_` (optional), `torch.optim.AdamW.step()` over two parameter groups
(`optimizers/utils.py:8-35`: no decay for ndim <= 1 or `.bias`) and a LR scheduler.  All parameters of the CUDA
path are views of ONE flat fp32 buffer and so are their gradients, so the whole tail is one HBM-bound kernel
(16 B read + 12 B written per parameter) instead of a multi-tensor sweep.  Arithmetic = torch.optim.AdamW
(decoupled decay, bias correction, eps outside the sqrt), verified against it in tests/test_model_gpu.py."""
from __future__ import annotations

import math

import torch

from . import _lib as L


class FlatAdamW:
    """Drop-in for the AdamW that train_rvfm.py instantiates; `lr` may be changed between steps (schedulers)."""

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 0.0):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        flat = model._flat
        if not flat.is_cuda:
            raise L.TheiaError("FlatAdamW needs the model on a CUDA device")
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.scratch = torch.zeros(2, dtype=torch.float32, device=flat.device)
        flags = torch.zeros((flat.numel() + 63) // 64, dtype=torch.uint8)
        for (name, shape, off), p in zip(model._param_meta, model._param_list):
            decay = not (p.ndim <= 1 or name.endswith(".bias"))  # optimizers/utils.py:26-33
            if decay:
                flags[off // 64:(off + math.prod(shape) + 63) // 64] = 1
        self.flags = flags.to(flat.device)
        self._gbuf = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model._param_list:
            p.grad = None

    def _flat_grads(self) -> torch.Tensor:
        """Gradients as one flat tensor: the views autograd stored as .grad alias one buffer (the module returned
        slices of it); if something re-materialised them, gather (slow path)."""
        plist, meta = self.model._param_list, self.model._param_meta
        g0 = plist[0].grad
        if g0 is None:
            raise RuntimeError("FlatAdamW.step() before backward()")
        base = g0.data_ptr() - 4 * meta[0][2]
        aliased = all(p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * off
                      for p, (_, _, off) in zip(plist, meta))
        n = self.model._flat.numel()
        if aliased:
            st = g0.untyped_storage()
            flat = torch.empty(0, dtype=torch.float32, device=g0.device).set_(st, 0, (st.nbytes() // 4,))
            start = (base - st.data_ptr()) // 4
            return flat[start:start + n]
        if self._gbuf is None:
            self._gbuf = torch.zeros_like(self.model._flat)
        for p, (_, shape, off) in zip(plist, meta):
            self._gbuf[off:off + math.prod(shape)].view(shape).copy_(p.grad if p.grad is not None else 0)
        return self._gbuf

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        g = self._flat_grads()
        flat = self.model._flat
        b1, b2 = self.betas
        L.check(L.lib().theia_adamw_flat(flat.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                         self.flags.data_ptr(), flat.numel(), float(self.lr), float(b1), float(b2),
                                         float(self.eps), float(self.weight_decay), int(self.step_count),
                                         float(self.max_grad_norm), self.scratch.data_ptr(), L.stream_ptr()),
                "theia_adamw_flat")
        self.model._packed_version = None  # the bf16 operand copies are re-packed at the next forward
