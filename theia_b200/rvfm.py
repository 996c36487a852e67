"""`RobotVisionFM` -- drop-in for the reference's model class (src/theia/models/rvfm.py:15-185) whose
forward / backward run as hand-written sm_100a kernels behind the C ABI (include/theia_b200.h).

Same constructor signature, methods (`forward`, `forward_feature`, `get_loss`, `freeze_translator`,
`load_pretrained_weights`), attributes and `state_dict` keys / shapes as the reference, so
`train_rvfm.py` (DDP wrap, AdamW param groups, checkpoint save) can use it unchanged.

Storage: every parameter is an fp32 view into ONE flat buffer (`self._flat`, layout defined by
csrc/model.cu); the library packs bf16 operand copies from it after each optimizer step and writes
gradients into a flat fp32 buffer of the same layout.  There is no PyTorch compute fallback: a
missing library or a CPU tensor model raises.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import Any, Optional

import torch
import torch.nn as nn

from . import _lib as L

# hub configs of the DeiT backbones the reference builds (backbones.py:506-526); model_type "vit"
BACKBONES = {
    "facebook/deit-tiny-patch16-224": (192, 3),
    "facebook/deit-small-patch16-224": (384, 6),
    "facebook/deit-base-patch16-224": (768, 12),
}
IMAGE_MEAN = (0.485, 0.456, 0.406)  # hub preprocessor_config.json of facebook/deit-*-patch16-224
IMAGE_STD = (0.229, 0.224, 0.225)


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's state_dict key hierarchy."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("container module; call RobotVisionFM")


def _descend(root: nn.Module, path: list[str]) -> nn.Module:
    cur = root
    for name in path:
        nxt = cur._modules.get(name)
        if nxt is None:
            nxt = _Node()
            cur.add_module(name, nxt)
        cur = nxt
    return cur


class _ForwardFn(torch.autograd.Function):
    """images -> per-teacher predictions; parameters are explicit inputs so autograd (and DDP's
    reducer hooks) see them."""

    @staticmethod
    def forward(ctx, module, images, names, kw, *params):
        preds = module._run_forward(images, names, kw)
        ctx.module = module
        ctx.names = names
        ctx.n_params = len(params)
        ctx.step_id = module._fwd_id
        return tuple(preds)

    @staticmethod
    def backward(ctx, *dpreds):
        m = ctx.module
        if ctx.step_id != m._fwd_id:
            raise L.TheiaError("backward() called for a stale forward: theia_b200 keeps the activations of the "
                               "most recent forward only")
        grads = m._run_backward(ctx.names, dpreds)
        used = {m._teachers.index(t) for t in ctx.names}
        out = []
        for i, h in enumerate(m._param_head):
            # parameters of heads that were not run have no gradient at all (the reference's autograd never reaches
            # them: grad stays None and torch.optim.AdamW skips them), not a zero gradient
            out.append(grads[i] if (ctx.needs_input_grad[4 + i] and (h < 0 or h in used)) else None)
        return (None, None, None, None, *out)


# id(fp32 loss gradient tensor, as autograd carries it) -> (weakref to it, the bf16 copy the loss kernel wrote in the
# same pass, its version).  Keyed by id with an identity check (tensors cannot be dict keys: == is elementwise).
_DPRED_BF16 = {}


def _stash_dpred(dpred, dbf):
    if len(_DPRED_BF16) > 64:  # entries whose gradient tensor died without reaching _run_backward
        for k in [k for k, v in _DPRED_BF16.items() if v[0]() is None]:
            del _DPRED_BF16[k]
    _DPRED_BF16[id(dpred)] = (weakref.ref(dpred), dbf, dpred._version)


def _take_dpred(g):
    ent = _DPRED_BF16.pop(id(g), None)
    if ent is not None and ent[0]() is g and ent[2] == g._version and ent[1].shape == g.shape:
        return ent[1]
    return None


class _LossFn(torch.autograd.Function):
    """(pred, target) -> tensor [3] = (mse, cos, l1) as nn.MSELoss / CosineEmbeddingLoss /
    SmoothL1Loss return them (rvfm.py:153-168)."""

    @staticmethod
    def forward(ctx, pred, target):
        if not pred.is_cuda:
            raise L.TheiaError("theia_b200 losses need CUDA tensors (no CPU fallback)")
        pred = pred.contiguous()
        if pred.dtype != torch.float32:
            pred = pred.float()
        if target.dtype not in (torch.float32, torch.bfloat16):
            target = target.float()
        target = target.to(pred.device).contiguous()
        B = pred.shape[0]
        n = pred[0].numel()
        if target.numel() != pred.numel():
            raise ValueError(f"target shape {tuple(target.shape)} does not match prediction {tuple(pred.shape)}")
        acc = torch.empty((B, 5), dtype=torch.float32, device=pred.device)
        out = torch.empty((3,), dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            L.check(L.lib().theia_loss_fwd(pred.data_ptr(), target.data_ptr(), int(target.dtype == torch.bfloat16),
                                           acc.data_ptr(), out.data_ptr(), B, n, L.stream_ptr()), "theia_loss_fwd")
        ctx.save_for_backward(pred, target, acc)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target, acc = ctx.saved_tensors
        B = pred.shape[0]
        n = pred[0].numel()
        coef = g.contiguous().float()
        dpred = torch.empty_like(pred)
        # autograd needs the fp32 gradient (dtype of `pred`); the head-Linear backward GEMMs need bf16: the same pass
        # writes both, and the copy is handed to RobotVisionFM._run_backward through a weak table keyed by the
        # gradient tensor itself (used only if that very tensor, unmodified, arrives there)
        dbf = torch.empty(pred.shape, dtype=torch.bfloat16, device=pred.device)
        with torch.cuda.device(pred.device):
            L.check(L.lib().theia_loss_bwd(pred.data_ptr(), target.data_ptr(), int(target.dtype == torch.bfloat16),
                                           acc.data_ptr(), coef.data_ptr(), dpred.data_ptr(), 1, dbf.data_ptr(), B, n,
                                           L.stream_ptr()), "theia_loss_bwd")
        _stash_dpred(dpred, dbf)
        return dpred, None


class RobotVisionFM(nn.Module):
    """Robot Vision Foundation Model -- B200-native drop-in (reference rvfm.py:15-75)."""

    def __init__(
        self,
        backbone: str | nn.Module = "facebook/deit-small-patch16-224",
        pretrained: bool = False,
        translator: str | nn.Module = "lconv",
        target_feature_sizes: Optional[dict[str, torch.Size | tuple[int, ...]]] = None,
        translator_kwargs: Optional[dict[str, Any]] = None,
        target_loss_weights: Optional[dict[str, float]] = None,
        checkpoint_path: Optional[str] = None,
        feature_reduce_method: Optional[str] = None,
        image_size: int = 224,
        **kwargs: Any,
    ) -> None:
        super().__init__()
        # build_backbone (backbones.py:506-526): "reg" -> DeiTReg, "nocls" -> DeiTNoCLS, "deit" -> DeiT
        self._variant, base_name = 0, backbone
        if isinstance(backbone, str) and backbone.startswith("reg-"):
            self._variant, base_name = 2, backbone[len("reg-"):]
        elif isinstance(backbone, str) and backbone.startswith("nocls-"):
            self._variant, base_name = 1, backbone[len("nocls-"):]
        if not isinstance(backbone, str) or base_name not in BACKBONES:
            raise NotImplementedError(f"Requested {backbone} is not implemented.")  # backbones.py:526
        self._num_reg = int(kwargs.pop("num_reg_tokens", 7)) if self._variant == 2 else 0
        if pretrained:
            raise NotImplementedError("pretrained hub weights need network access; use load_pretrained_weights()")
        if translator != "lconv":
            raise NotImplementedError(f"Requested {translator} is not implemented yet.")  # feature_translators.py:313
        hsf = 1.0
        if translator_kwargs is not None:
            hsf = float(dict(translator_kwargs).get("hidden_size_factor", 1.0))
        if hsf != 1.0:
            raise NotImplementedError("only hidden_size_factor 1.0 (configs/model/translator/lconv.yaml:3)")
        # `image_size` is stored, not used to build the model: the reference's DeiT wrapper does the same
        # (backbones.py:267-285 always instantiates the 224-pixel hub config)
        self.target_feature_sizes = target_feature_sizes
        self.preprocessor = None
        self.pretrained = pretrained
        self.image_size = image_size
        self.final_spatial = None
        self.feature_reduce_method = feature_reduce_method
        self.no_cls = self._variant == 1            # rvfm.py:59 hasattr(backbone, "no_cls")
        self.num_reg_tokens = self._num_reg         # rvfm.py:60
        self._seq = 196 + (0 if self._variant == 1 else 1) + self._num_reg
        self.target_loss_weights = target_loss_weights
        self.backbone_name = backbone
        self.hidden, self.heads = BACKBONES[base_name]
        self.image_mean, self.image_std = IMAGE_MEAN, IMAGE_STD
        self._teachers = list(target_feature_sizes.keys()) if target_feature_sizes else []
        self._max_batch = int(kwargs.pop("max_batch", 0))
        self._handle = None
        self._input_hw = (224, 224)
        self._input_f32 = False
        self._handle_batch = 0
        self._workspace = None
        self._gbufs = [None, None]
        self._gcur = 0
        self._pack_table = (0, 0)
        self._packed_version = None
        self._fwd_id = 0
        self._dpred_bf16 = {}
        self._grad_sync = None  # (process group,) when the module all-reduces its own flat gradient buffer

        # layout comes from the library (host-only call; works without a GPU)
        h = self._create_handle(1)
        lib = L.lib()
        nfl = lib.theia_model_param_floats(h)
        self._flat = torch.zeros(nfl, dtype=torch.float32)
        self._param_meta = []
        name_buf = C.create_string_buffer(512)
        dims = (C.c_longlong * 4)()
        ndim = C.c_int()
        off = C.c_longlong()
        for i in range(lib.theia_model_num_params(h)):
            L.check(lib.theia_model_param_info(h, i, name_buf, 512, dims, C.byref(ndim), C.byref(off)), "param_info")
            shape = tuple(int(dims[k]) for k in range(ndim.value))
            self._param_meta.append((name_buf.value.decode(), shape, int(off.value)))
        lib.theia_model_destroy(h)
        self._param_list = []
        for name, shape, o in self._param_meta:
            parts = name.split(".")
            # teacher names may contain '.' in the original key? no: '.' -> '_' (feature_translators.py:46)
            owner = _descend(self, parts[:-1])
            p = nn.Parameter(self._flat[o:o + math.prod(shape)].view(shape))
            owner.register_parameter(parts[-1], p)
            self._param_list.append(p)
        prefixes = ["translator.translator_heads." + t.replace(".", "_") + "." for t in self._teachers]
        self._param_head = [next((i for i, pre in enumerate(prefixes) if name.startswith(pre)), -1)
                            for name, _, _ in self._param_meta]
        self.reset_parameters()
        if checkpoint_path:
            self.load_pretrained_weights(checkpoint_path)

    # ------------------------------------------------------------------ construction helpers
    def _make_config(self, max_batch: int) -> L.ModelConfig:
        cfg = L.ModelConfig()
        cfg.hidden, cfg.heads, cfg.layers, cfg.image, cfg.patch = self.hidden, self.heads, 12, 224, 16
        cfg.max_batch = max_batch
        cfg.ln_eps = 1e-12
        cfg.variant, cfg.num_reg_tokens = self._variant, self._num_reg
        cfg.num_teachers = len(self._teachers)
        if cfg.num_teachers > L.MAX_TEACHERS:
            raise NotImplementedError("too many teachers")
        self._name_keepalive = [t.encode() for t in self._teachers]
        for i, t in enumerate(self._teachers):
            size = tuple(self.target_feature_sizes[t])
            if len(size) == 1 or "_cls" in t:  # LinearAdapterHead on the CLS token (feature_translators.py:195-199)
                c, hh, ww = size[0], 1, 1
            else:
                c, hh, ww = size
            if hh != ww:
                raise NotImplementedError("Currently does not support non-square feature maps")
            cfg.teacher_names[i] = self._name_keepalive[i]
            cfg.teacher_c[i] = int(c)
            cfg.teacher_hw[i] = int(hh)
        return cfg

    def _create_handle(self, max_batch: int):
        cfg = self._make_config(max_batch)
        h = C.c_void_p()
        rc = L.lib().theia_model_create(C.byref(cfg), C.byref(h))
        if rc == 3:
            raise NotImplementedError(L.lib().theia_last_error().decode())
        L.check(rc, "theia_model_create")
        return h

    @torch.no_grad()
    def reset_parameters(self) -> None:
        """hf:modeling_vit.py:385-399 for the backbone; torch defaults for the translator."""
        for (name, shape, _), p in zip(self._param_meta, self._param_list):
            if name.startswith("backbone"):
                if "layernorm" in name:
                    p.fill_(1.0 if name.endswith("weight") else 0.0)
                elif name.endswith(".bias"):
                    p.zero_()
                else:
                    nn.init.trunc_normal_(p, mean=0.0, std=0.02)
            else:
                leaf = name.split(".")[-2]
                if leaf in ("0", "3", "6"):  # LayerNorm([C,H,W])
                    p.fill_(1.0 if name.endswith("weight") else 0.0)
                else:
                    fan_in = (shape[1] * 9) if len(shape) == 4 else (shape[1] if len(shape) == 2 else None)
                    if fan_in is None:  # bias: fan_in of its weight
                        fan_in = self.hidden * (9 if leaf in ("1", "4") else 1)
                    bound = 1.0 / math.sqrt(fan_in)
                    p.uniform_(-bound, bound)

    def _apply(self, fn, recurse=True):
        """Keep every parameter a view of one flat buffer across .to()/.cuda()."""
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise NotImplementedError("theia_b200 keeps fp32 master weights (bf16 compute copies are internal)")
        if new_flat is not self._flat:
            new_flat = new_flat.contiguous()
            for (name, shape, o), p in zip(self._param_meta, self._param_list):
                p.data = new_flat[o:o + math.prod(shape)].view(shape)
                if p.grad is not None:
                    p.grad = None
            self._flat = new_flat
            self._release_handle()
        return self

    def _release_handle(self):
        if self._handle is not None:
            L.lib().theia_model_destroy(self._handle)
        self._handle = None
        self._input_hw = (224, 224)  # a new context starts at the default extent
        self._input_f32 = False      # ... and at uint8 pixels
        self._workspace = None
        self._gbufs = [None, None]
        self._pack_table = (0, 0)
        self._packed_version = None

    def __del__(self):
        try:
            self._release_handle()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    def load_pretrained_weights(self, checkpoint_path: str):
        """rvfm.py:77-87."""
        if checkpoint_path:
            weights_dict = torch.load(checkpoint_path, map_location="cpu")
            pretrained_dict = {k: v for k, v in weights_dict.items() if k in self.state_dict()}
            self.load_state_dict(pretrained_dict, strict=False)

    def sync_gradients(self, enabled: bool = True, group=None) -> None:
        """Data-parallel alternative to wrapping the module in DistributedDataParallel (train_rvfm.py:258): every
        gradient lives in ONE flat fp32 buffer, so backward() finishes with a single NCCL all-reduce (average) over it
        -- no bucket copies, one collective.  Do not combine with a DDP wrapper (the reduction would happen twice).
        Parameters must already be identical on all ranks (same seed / checkpoint, or broadcast `self._flat`)."""
        self._grad_sync = (group,) if enabled else None

    def freeze_translator(self) -> None:
        """rvfm.py:89-92."""
        for param in self.translator.parameters():
            param.requires_grad = False

    def _ensure(self, B: int):
        if not self._flat.is_cuda:
            raise L.TheiaError("theia_b200.RobotVisionFM runs on CUDA only: call .cuda()/.to(device) first "
                               "(there is no CPU fallback)")
        if self._handle is None or B > self._handle_batch:
            self._release_handle()
            mb = max(B, self._max_batch)
            self._handle = self._create_handle(mb)
            self._handle_batch = mb
            lib = L.lib()
            nbytes = lib.theia_model_workspace_bytes(self._handle)
            dev = self._flat.device
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._gbufs = [None, None]  # gradient buffers are allocated by the first backward (none for inference)
            self._gcur = 0
            L.check(lib.theia_model_bind(self._handle, self._flat.data_ptr(), 0, self._workspace.data_ptr()),
                    "theia_model_bind")
            tab, pk = C.c_void_p(), C.c_void_p()
            L.check(lib.theia_model_pack_table(self._handle, C.byref(tab), C.byref(pk)), "theia_model_pack_table")
            self._pack_table = (tab.value, pk.value)
            self._packed_version = None
        v = self._weights_version()
        if self._packed_version != v:
            L.check(L.lib().theia_model_pack(self._handle, 0, L.stream_ptr()), "theia_model_pack")
            self._packed_version = v

    def _adamw_pack_args(self):
        """(pack table, bf16 pack buffer) device pointers for theia_adamw_flat, or (0, 0) before the first forward."""
        return self._pack_table if self._handle is not None else (0, 0)

    def _after_optimizer_step(self, fused_cast: bool) -> None:
        """FlatAdamW updated the flat buffer through its raw pointer (no version counter moves): refresh the packs
        that are not plain casts (token table, conv-weight gathers, LayerNorm[C,H,W] affines) right away."""
        if self._handle is None:
            return
        L.check(L.lib().theia_model_pack(self._handle, int(fused_cast), L.stream_ptr()), "theia_model_pack")
        self._packed_version = self._weights_version()

    def _next_grad_buffer(self) -> torch.Tensor:
        """Flat gradient buffer the next backward writes.  Two buffers alternate so that the views handed to autograd
        by the previous backward (which it may keep as `.grad`) are never overwritten; if the candidate is STILL
        referenced by some `.grad` (gradient accumulation over 3+ backwards), a fresh one replaces it."""
        k = self._gcur ^ 1
        buf = self._gbufs[k]
        if buf is not None:
            lo, hi = buf.data_ptr(), buf.data_ptr() + buf.numel() * 4
            if any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self._param_list):
                buf = None
        if buf is None:
            buf = torch.empty_like(self._flat)
            self._gbufs[k] = buf
        self._gcur = k
        return buf

    def _weights_version(self):
        """Changes whenever any parameter is written in place.  After `.to()/.cuda()` every Parameter keeps its OWN
        version counter (`p.data = view` does not share the new flat buffer's), so the key sums all of them: an
        in-place update through any view (torch optimizers, `load_state_dict`, `p.mul_()`) is seen."""
        return (self._flat._version, sum(p._version for p in self._param_list))

    def _prep_images(self, x, do_resize: bool) -> tuple[torch.Tensor, int]:
        """uint8 [B,224,224,3] or [B,3,224,224] on the model's device (backbones.py:337-339 accepts tensors,
        numpy arrays and PIL lists)."""
        if isinstance(x, (list, tuple)):
            import numpy as np
            x = torch.from_numpy(np.stack([np.asarray(im) for im in x]))
        elif not torch.is_tensor(x):
            import numpy as np
            x = torch.from_numpy(np.asarray(x))
        if x.dim() == 3:
            x = x[None]
        if x.dtype != torch.uint8:
            # float images (values in [0, 255], or [0, 1] with do_rescale=False): same fused rescale / normalise; the
            # bicubic resize of float tensors is not restated on this path
            if not x.dtype.is_floating_point:
                raise NotImplementedError(f"theia_b200 takes uint8 or floating-point images, not {x.dtype}")
            if do_resize:
                raise NotImplementedError("float images are not resized on this path: pass do_resize=False "
                                          "(uint8 images of any extent are resized like the reference does)")
            x = x.to(torch.float32)
        chw = 1 if (x.shape[1] in (1, 3) and x.shape[-1] not in (1, 3)) else 0
        if x.shape[1 if chw else 3] != 3:
            raise NotImplementedError("theia_b200 takes 3-channel images")
        # any extent is accepted, as by the reference's processor: it resizes to 256 x 256 (do_resize) and ALWAYS
        # centre-crops / zero-pads to 224 x 224 (hf:image_processing_backends.py center_crop), so the ViT sees 196 patches
        # whatever came in (and `interpolate_pos_encoding` is the identity: hf:modeling_vit.py:74-76)
        return x.to(self._flat.device, non_blocking=True).contiguous(), chw

    def _run_backbone(self, x, kw, run_heads: bool, names, tokens_out=None):
        do_resize = kw.get("do_resize", True)
        # interpolate_pos_encoding is the identity on this path: the processor always hands the ViT 224 x 224 pixels
        # (see _prep_images), and hf:modeling_vit.py:74-76 returns the stored table when the patch grid matches
        # the reference's processor resizes CPU uint8 tensors (and PIL / numpy inputs) with torchvision's fixed-point
        # path and CUDA tensors with the float path: reproduce whichever the caller would have got (2 / 1)
        on_cpu = not (torch.is_tensor(x) and x.is_cuda)
        images, chw = self._prep_images(x, do_resize)
        B = images.shape[0]
        self._ensure(B)
        mean = (C.c_float * 3)(*self.image_mean)
        std = (C.c_float * 3)(*self.image_std)
        preds, ptrs = [], (C.c_void_p * L.MAX_TEACHERS)()
        if run_heads:
            for i, t in enumerate(self._teachers):
                if t in names:
                    size = tuple(self.target_feature_sizes[t])
                    shape = (B, size[0]) if (len(size) == 1 or "_cls" in t) else (B, size[1] * size[2], size[0])
                    p = torch.empty(shape, dtype=torch.float32, device=images.device)
                    preds.append(p)
                    ptrs[i] = p.data_ptr()
        self._fwd_id += 1
        H, W = (images.shape[2], images.shape[3]) if chw else (images.shape[1], images.shape[2])
        if (H, W) != self._input_hw:
            L.check(L.lib().theia_model_set_input_size(self._handle, H, W), "theia_model_set_input_size")
            self._input_hw = (H, W)
        is_f32 = images.dtype == torch.float32
        if is_f32 != self._input_f32:
            L.check(L.lib().theia_model_set_input_dtype(self._handle, int(is_f32)), "theia_model_set_input_dtype")
            self._input_f32 = is_f32
        L.check(L.lib().theia_model_forward(
            self._handle, images.data_ptr(), B, chw, (2 if on_cpu else 1) if do_resize else 0, int(kw.get("do_rescale", True)),
            int(kw.get("do_normalize", True)), mean, std, int(run_heads), ptrs,
            0 if tokens_out is None else tokens_out.data_ptr(), L.stream_ptr()), "theia_model_forward")
        self._last_B = B
        return preds

    def _run_forward(self, images, names, kw):
        with torch.cuda.device(self._flat.device):  # kernels launch on the model's device whatever the current one is
            return self._run_backbone(images, kw, True, names)

    def _run_backward(self, names, dpreds):
        with torch.cuda.device(self._flat.device):
            return self._run_backward_impl(names, dpreds)

    def _run_backward_impl(self, names, dpreds):
        lib = L.lib()
        B = self._last_B
        ptrs = (C.c_void_p * L.MAX_TEACHERS)()
        k = 0
        keep = []  # bf16 gradients stay alive until the backward kernels have been enqueued
        for i, t in enumerate(self._teachers):
            if t not in names:
                continue
            g = dpreds[k]
            k += 1
            if g is None:
                continue
            hit = _take_dpred(g)
            if hit is not None:
                ptrs[i] = hit.data_ptr()  # produced by theia_loss_bwd together with g: no cast
                keep.append(hit)
                continue
            g = g.contiguous()
            if g.dtype != torch.float32:
                g = g.float()
            buf = self._dpred_bf16.get(t)
            if buf is None or buf.numel() != g.numel() or buf.device != g.device:
                buf = torch.empty(g.shape, dtype=torch.bfloat16, device=g.device)
                self._dpred_bf16[t] = buf
            L.check(lib.theia_cast_bf16(g.data_ptr(), buf.data_ptr(), g.numel(), L.stream_ptr()), "theia_cast_bf16")
            ptrs[i] = buf.data_ptr()
        flat_g = self._next_grad_buffer()  # autograd may keep what we return as .grad: buffers alternate, no copy
        L.check(lib.theia_model_set_grads(self._handle, flat_g.data_ptr()), "theia_model_set_grads")
        L.check(lib.theia_model_backward(self._handle, ptrs, L.stream_ptr()), "theia_model_backward")
        if self._grad_sync is not None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self._grad_sync[0]) > 1:
                dist.all_reduce(flat_g, op=dist.ReduceOp.AVG, group=self._grad_sync[0])
        return [flat_g[o:o + math.prod(shape)].view(shape) for (_, shape, o) in self._param_meta]

    def forward_feature(self, x: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """rvfm.py:94-113.  Returns fp32 like the reference; not differentiable (inference API)."""
        B = len(x) if isinstance(x, (list, tuple)) else (1 if getattr(x, "ndim", 4) == 3 else x.shape[0])
        if not self._flat.is_cuda:
            raise L.TheiaError("theia_b200.RobotVisionFM runs on CUDA only: call .cuda()/.to(device) first "
                               "(there is no CPU fallback)")
        with torch.cuda.device(self._flat.device):
            tok = torch.empty((B, self._seq, self.hidden), dtype=torch.bfloat16, device=self._flat.device)
            self._run_backbone(x, kwargs, False, (), tokens_out=tok)
        return handle_feature_output(tok.float(), self.feature_reduce_method, self.num_reg_tokens)

    def forward(self, x: torch.Tensor, target_model_names: Optional[list[str]] = None,
                **kwargs: Any) -> dict[str, torch.Tensor]:
        """rvfm.py:115-136: dict teacher -> [B, H*W, C] fp32 predictions."""
        names = tuple(target_model_names) if target_model_names is not None else tuple(self._teachers)
        for t in names:
            if t not in self._teachers:
                raise KeyError(t)
        ordered = tuple(t for t in self._teachers if t in names)
        preds = _ForwardFn.apply(self, x, ordered, dict(kwargs), *self._param_list)
        by_name = dict(zip(ordered, preds))
        return {t: by_name[t] for t in names}

    def get_loss(self, pred_features: dict[str, torch.Tensor], y: dict[str, torch.Tensor]) -> dict[str, Any]:
        """rvfm.py:138-185; one fused reduction per teacher and ONE device->host copy for the per-model floats."""
        T = len(pred_features)
        mse_avg, cos_avg, l1_avg = 0, 0, 0
        per = []
        for t in pred_features:
            out = _LossFn.apply(pred_features[t], y[t])
            weight = self.target_loss_weights if self.target_loss_weights else 1.0 / T
            mse_avg = mse_avg + out[0] * weight
            cos_avg = cos_avg + out[1] / T
            l1_avg = l1_avg + out[2] * weight
            per.append(out.detach())
        vals = torch.stack(per).tolist() if per else []
        names = list(pred_features.keys())
        return {
            "mse_loss": mse_avg,
            "cos_loss": cos_avg,
            "l1_loss": l1_avg,
            "mse_losses_per_model": {t: v[0] for t, v in zip(names, vals)},
            "cos_losses_per_model": {t: v[1] for t, v in zip(names, vals)},
            "l1_losses_per_model": {t: v[2] for t, v in zip(names, vals)},
        }


def handle_feature_output(x: torch.Tensor, feature_reduce_method: Optional[str] = None,
                          num_discard_tokens: int = 0) -> torch.Tensor:
    """models/utils.py:8-43 (token selection: views / tiny reductions on the returned feature)."""
    if feature_reduce_method == "mean_pooling":
        return torch.mean(x[:, 1: x.size(1) - num_discard_tokens], dim=1)
    if feature_reduce_method == "max_pooling":
        return torch.amax(x[:, 1: x.size(1) - num_discard_tokens], dim=1)
    if feature_reduce_method == "cls":
        return x[:, 0]
    if feature_reduce_method == "identity":
        return x
    if feature_reduce_method is None:
        return x[:, 1: x.size(1) - num_discard_tokens]
    raise NotImplementedError(f"feature_reduce_method {feature_reduce_method} it not implemented.")
