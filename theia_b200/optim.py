"""Fused optimizer tail for `RobotVisionFM`.

`train_rvfm.py:126-133` runs `clip_grad_norm_` (optional), `torch.optim.AdamW.step()` over two parameter groups
(`optimizers/utils.py:8-35`: no decay for ndim <= 1 or `.bias`) and a LR scheduler.  All parameters of the CUDA
path are views of ONE flat fp32 buffer and so are their gradients, so the whole tail is one HBM-bound kernel
(16 B read + 12 B written per parameter, + 2 B for the bf16 GEMM-operand copy it refreshes in the same pass)
instead of a multi-tensor sweep followed by ~50 cast kernels.  Arithmetic = torch.optim.AdamW (decoupled decay,
bias correction, eps outside the sqrt), verified against it in tests/test_model_gpu.py.

`FlatAdamW` IS a `torch.optim.Optimizer`: one param group holding every model parameter (`lr`, `betas`, `eps`,
`weight_decay` are read from the group at each step, so the reference's LR schedulers -- `lr_schedulers.py:41-77`
wraps `torch.optim.lr_scheduler.*` around the optimizer -- drive it unchanged), `state_dict()` /
`load_state_dict()` carry the moments and the step count.  Like torch.optim.AdamW it leaves a parameter
untouched (no decay, no moment update) when its gradient is None or it does not require a gradient
(`RobotVisionFM.freeze_translator()`, heads not selected through `target_model_names`)."""
from __future__ import annotations

import math

import torch

from . import _lib as L

FLAG_DECAY, FLAG_SKIP = 1, 2


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 0.0):
        flat = model._flat
        if not flat.is_cuda:
            raise L.TheiaError("FlatAdamW needs the model on a CUDA device (call .cuda()/.to(device) first)")
        self.model = model
        super().__init__(list(model._param_list), dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.scratch = torch.zeros(2, dtype=torch.float32, device=flat.device)
        nblk = (flat.numel() + 63) // 64
        base = torch.zeros(nblk, dtype=torch.uint8)
        self._blocks = []  # per parameter: (first block, last block + 1)
        for (name, shape, off), p in zip(model._param_meta, model._param_list):
            b0, b1 = off // 64, (off + math.prod(shape) + 63) // 64
            self._blocks.append((b0, b1))
            if not (p.ndim <= 1 or name.endswith(".bias")):  # optimizers/utils.py:26-33
                base[b0:b1] = FLAG_DECAY
        self._base_flags = base
        self._flags = base.to(flat.device)
        self._skip_key = ()
        self._gbuf = None

    # ------------------------------------------------------------------ gradients as one flat tensor
    def _flat_grads(self) -> torch.Tensor:
        """The views autograd stored as .grad alias one buffer (the module returned slices of it); if something
        re-materialised them, gather (slow path)."""
        plist, meta = self.model._param_list, self.model._param_meta
        first = next((i for i, p in enumerate(plist) if p.grad is not None), None)
        if first is None:
            raise RuntimeError("FlatAdamW.step() before backward()")
        g0 = plist[first].grad
        base = g0.data_ptr() - 4 * meta[first][2]
        aliased = all(p.grad is None or (p.grad.is_contiguous() and p.grad.data_ptr() == base + 4 * off)
                      for p, (_, _, off) in zip(plist, meta))
        n = self.model._flat.numel()
        if aliased:
            st = g0.untyped_storage()
            start = (base - st.data_ptr()) // 4
            if start >= 0 and (start + n) * 4 <= st.nbytes():
                flat = torch.empty(0, dtype=torch.float32, device=g0.device).set_(st, 0, (st.nbytes() // 4,))
                return flat[start:start + n]
        if self._gbuf is None:
            self._gbuf = torch.zeros_like(self.model._flat)
        for p, (_, shape, off) in zip(plist, meta):
            if p.grad is not None:
                self._gbuf[off:off + math.prod(shape)].view(shape).copy_(p.grad)
        return self._gbuf

    def _refresh_flags(self):
        """Blocks of parameters without a gradient are skipped, as torch.optim.AdamW skips `p.grad is None`."""
        key = tuple(i for i, p in enumerate(self.model._param_list) if p.grad is None or not p.requires_grad)
        if key != self._skip_key:
            f = self._base_flags.clone()
            for i in key:
                b0, b1 = self._blocks[i]
                f[b0:b1] |= FLAG_SKIP
            self._flags = f.to(self.model._flat.device)
            self._skip_key = key

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        model = self.model
        flat = model._flat
        with torch.cuda.device(flat.device):
            self._refresh_flags()
            g = self._flat_grads()
            self.step_count += 1
            b1, b2 = group["betas"]
            # the same pass refreshes the bf16 GEMM-operand copies of the weights it updates (model bound: the
            # pack table lives in the model's workspace); otherwise the next forward re-packs
            fused = model._adamw_pack_args()
            L.check(L.lib().theia_adamw_flat(flat.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                             self._flags.data_ptr(), flat.numel(), float(group["lr"]), float(b1),
                                             float(b2), float(group["eps"]), float(group["weight_decay"]),
                                             int(self.step_count), float(self.max_grad_norm), self.scratch.data_ptr(),
                                             fused[0], fused[1], L.stream_ptr()), "theia_adamw_flat")
            model._after_optimizer_step(fused_cast=fused[0] != 0)
        return loss

    # ------------------------------------------------------------------ checkpointing (torch.optim.Optimizer API)
    def state_dict(self):
        return {"state": {"step": self.step_count, "exp_avg": self.m, "exp_avg_sq": self.v,
                          "max_grad_norm": self.max_grad_norm},
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        st = sd["state"]
        self.step_count = int(st["step"])
        self.m.copy_(st["exp_avg"])
        self.v.copy_(st["exp_avg_sq"])
        self.max_grad_norm = float(st.get("max_grad_norm", self.max_grad_norm))
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in s.items() if k != "params"})

    # attribute shorthands kept from round 1
    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v
