"""Teacher feature extraction on the CUDA path.

Host-side mirror of the reference's teacher wrappers -- `get_dinov2_feature` (src/theia/foundation_models/
vision_models/dinov2.py:8-41), `get_clip_feature` (vision_language_models/clip.py:8-41), `get_vit_feature`
(vision_models/vit.py:8-33) and the dispatcher `get_feature_outputs` (src/theia/preprocessing/
feature_extraction_core/models.py:55-95): same names, arguments and returned tensors.  The transformer forward runs
through `theia_vit_forward` (csrc/vit_infer.cu: tcgen05 GEMMs, LayerNorm, TMEM attention); the image processor stays
the HF one the reference uses (CPU, PIL), its `pixel_values` are patchified on the GPU.

`TeacherViT.from_hf(model)` converts a loaded `Dinov2Model` / `CLIPVisionModel` / `ViTModel` once: bf16 GEMM operands
(q | k | v fused), fp32 biases and LayerNorm affines, DINOv2's LayerScale folded into the output projections, the
position table interpolated to the working resolution by the HF module's own `interpolate_pos_encoding`.
Supported: head dim 64 or 80, hidden <= 1280, <= 272 tokens (DINOv2-S/B/L, CLIP ViT-B/L, ViT-B/L/H at 224 px).
Anything else (SAM, LLaVA, Depth-Anything) raises -- there is no PyTorch fallback.  Inference only."""
from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from . import _lib as L


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _bf16(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).to(torch.bfloat16).contiguous()


class TeacherViT:
    """A frozen ViT teacher bound to one CUDA device.  `__call__(pixel_values)` returns
    (last_hidden_state [B, tokens, D], pooler_output [B, D]) as the HF model's output of the same names."""

    def __init__(self, kind: str, cfg: dict, tensors: dict, layers: list[dict], device, residual_fp32: bool = True):
        self.kind, self.cfg, self.device = kind, dict(cfg), torch.device(device)
        self._t, self._layers = tensors, layers  # keep every device buffer alive
        self.lib = L.lib()
        n = cfg["layers"]
        self._layer_arr = (L.VitLayer * n)()
        for i, ly in enumerate(layers):
            for name, _ in L.VitLayer._fields_:
                setattr(self._layer_arr[i], name, ly[name].data_ptr())
        d = L.VitDesc()
        for k in ("hidden", "heads", "layers", "mlp", "tokens", "patch_off", "patch_tokens", "patch_k", "act", "final_ln_mode"):
            setattr(d, k, int(cfg[k]))
        d.ln_eps = float(cfg["ln_eps"])
        for k in ("w_patch", "b_patch", "tok_table", "pre_ln_w", "pre_ln_b", "final_ln_w", "final_ln_b"):
            setattr(d, k, L.ptr(tensors.get(k)))
        d.layer = self._layer_arr
        # fp32 residual stream between the blocks: these features are distillation TARGETS; with a bf16 stream the
        # stream's rounding alone is 8e-3 of the 1e-2 total error of a 24-layer forward (tools / DESIGN section 6)
        d.residual_f32 = 1 if residual_fp32 else 0
        self.residual_fp32 = bool(residual_fp32)
        self._desc = d
        self._ws = None
        # what the reference reads off the HF model
        self.config = cfg.get("hf_config")

    # ---- conversion -------------------------------------------------------------------------------------------
    @classmethod
    def from_hf(cls, model: torch.nn.Module, device: Any = "cuda", image_size: int | None = None,
                _convert_only: bool = False, residual_fp32: bool = True) -> "TeacherViT":
        """_convert_only: build the converted buffers on a non-CUDA device (weight-conversion unit tests; such an
        object cannot run a forward)"""
        name = type(model).__name__
        dev = torch.device(device)
        if dev.type != "cuda" and not _convert_only:
            raise L.TheiaError("TeacherViT runs on a CUDA device only (no CPU / PyTorch fallback)")
        hc = model.config
        sd = {k: v for k, v in model.state_dict().items()}
        D, H, nl, p = hc.hidden_size, hc.num_attention_heads, hc.num_hidden_layers, hc.patch_size
        if name == "Dinov2Model":
            size = image_size or 224  # the DINOv2 processor centre-crops to 224 (dinov2.py:26)
            if getattr(hc, "use_swiglu_ffn", False):
                raise L.TheiaError("DINOv2 with a SwiGLU FFN (giant) is not supported")
            mlp, act, eps = int(D * hc.mlp_ratio), 0, hc.layer_norm_eps
            with torch.no_grad():  # the HF module's own bicubic interpolation of the 37x37 table (+ class position)
                pos = model.embeddings.interpolate_pos_encoding(torch.zeros(1, 1 + (size // p) ** 2, D), size, size)[0].float()
            table = pos.clone()
            table[0] += sd["embeddings.cls_token"].reshape(D).float()
            wp, bp = sd["embeddings.patch_embeddings.projection.weight"], sd["embeddings.patch_embeddings.projection.bias"]
            pre = None
            fin = (sd["layernorm.weight"], sd["layernorm.bias"], 1)

            def layer(i):
                b = f"encoder.layer.{i}."
                g1, g2 = sd[b + "layer_scale1.lambda1"].float(), sd[b + "layer_scale2.lambda1"].float()
                a = b + "attention.attention."
                return dict(
                    ln1=(sd[b + "norm1.weight"], sd[b + "norm1.bias"]),
                    qkv=(torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]]),
                         torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]])),
                    # LayerScale: x + g * (W y + b) = x + (diag(g) W) y + g b
                    o=(g1[:, None] * sd[b + "attention.output.dense.weight"].float(), g1 * sd[b + "attention.output.dense.bias"].float()),
                    ln2=(sd[b + "norm2.weight"], sd[b + "norm2.bias"]),
                    fc1=(sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]),
                    fc2=(g2[:, None] * sd[b + "mlp.fc2.weight"].float(), g2 * sd[b + "mlp.fc2.bias"].float()))
        elif name == "CLIPVisionModel":
            size = image_size or hc.image_size
            if size != hc.image_size:
                raise L.TheiaError("CLIP: only the native resolution is supported")
            mlp, eps = hc.intermediate_size, hc.layer_norm_eps
            act = {"quick_gelu": 1, "gelu": 0}.get(hc.hidden_act)
            if act is None:
                raise L.TheiaError(f"CLIP activation {hc.hidden_act!r} is not supported")
            v = "vision_model."
            table = sd[v + "embeddings.position_embedding.weight"].float().clone()
            table[0] += sd[v + "embeddings.class_embedding"].float()
            wp, bp = sd[v + "embeddings.patch_embedding.weight"], None
            pre = (sd[v + "pre_layrnorm.weight"], sd[v + "pre_layrnorm.bias"])
            fin = (sd[v + "post_layernorm.weight"], sd[v + "post_layernorm.bias"], 2)

            def layer(i):
                b = f"{v}encoder.layers.{i}."
                a = b + "self_attn."
                return dict(
                    ln1=(sd[b + "layer_norm1.weight"], sd[b + "layer_norm1.bias"]),
                    qkv=(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]]),
                         torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]])),
                    o=(sd[a + "out_proj.weight"], sd[a + "out_proj.bias"]),
                    ln2=(sd[b + "layer_norm2.weight"], sd[b + "layer_norm2.bias"]),
                    fc1=(sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]),
                    fc2=(sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"]))
        elif name == "ViTModel":
            size = image_size or hc.image_size
            if size != hc.image_size:
                raise L.TheiaError("ViT: only the native resolution is supported")
            if hc.hidden_act != "gelu":
                raise L.TheiaError(f"ViT activation {hc.hidden_act!r} is not supported")
            mlp, act, eps = hc.intermediate_size, 0, hc.layer_norm_eps
            table = sd["embeddings.position_embeddings"][0].float().clone()
            table[0] += sd["embeddings.cls_token"].reshape(D).float()
            wp, bp = sd["embeddings.patch_embeddings.projection.weight"], sd["embeddings.patch_embeddings.projection.bias"]
            pre = None
            fin = (sd["layernorm.weight"], sd["layernorm.bias"], 1)

            def layer(i):
                b = f"encoder.layer.{i}."
                a = b + "attention.attention."
                return dict(
                    ln1=(sd[b + "layernorm_before.weight"], sd[b + "layernorm_before.bias"]),
                    qkv=(torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]]),
                         torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]])),
                    o=(sd[b + "attention.output.dense.weight"], sd[b + "attention.output.dense.bias"]),
                    ln2=(sd[b + "layernorm_after.weight"], sd[b + "layernorm_after.bias"]),
                    fc1=(sd[b + "intermediate.dense.weight"], sd[b + "intermediate.dense.bias"]),
                    fc2=(sd[b + "output.dense.weight"], sd[b + "output.dense.bias"]))
        else:
            raise L.TheiaError(f"{name}: no CUDA teacher path (supported: Dinov2Model, CLIPVisionModel, ViTModel)")
        if D % H != 0 or D // H not in (64, 80) or D > 1280:
            raise L.TheiaError(f"{name}: hidden {D} / heads {H} -- the attention kernels are built for head dim 64 and 80, "
                               "hidden <= 1280")
        g = size // p
        tokens = 1 + g * g
        if tokens > 272 or table.shape[0] != tokens:
            raise L.TheiaError(f"{name}: {tokens} tokens per image (position table {tuple(table.shape)}); at most 272 are supported")
        k_real = 3 * p * p
        k_pad = (k_real + 7) // 8 * 8
        wpk = torch.zeros(D, k_pad, dtype=torch.float32)
        wpk[:, :k_real] = wp.detach().float().reshape(D, k_real).cpu()
        tensors = {"w_patch": _bf16(wpk, dev), "b_patch": None if bp is None else _f32(bp, dev), "tok_table": _f32(table, dev),
                   "pre_ln_w": None if pre is None else _f32(pre[0], dev), "pre_ln_b": None if pre is None else _f32(pre[1], dev),
                   "final_ln_w": _f32(fin[0], dev), "final_ln_b": _f32(fin[1], dev)}
        layers = []
        for i in range(nl):
            ly = layer(i)
            layers.append({"ln1_w": _f32(ly["ln1"][0], dev), "ln1_b": _f32(ly["ln1"][1], dev),
                           "w_qkv": _bf16(ly["qkv"][0], dev), "b_qkv": _f32(ly["qkv"][1], dev),
                           "w_o": _bf16(ly["o"][0], dev), "b_o": _f32(ly["o"][1], dev),
                           "ln2_w": _f32(ly["ln2"][0], dev), "ln2_b": _f32(ly["ln2"][1], dev),
                           "w_fc1": _bf16(ly["fc1"][0], dev), "b_fc1": _f32(ly["fc1"][1], dev),
                           "w_fc2": _bf16(ly["fc2"][0], dev), "b_fc2": _f32(ly["fc2"][1], dev)})
        cfg = dict(hidden=D, heads=H, layers=nl, mlp=mlp, tokens=tokens, patch_off=1, patch_tokens=g * g, patch_k=k_pad,
                   ln_eps=eps, act=act, final_ln_mode=fin[2], patch=p, image=size, hf_config=hc)
        return cls(name, cfg, tensors, layers, dev, residual_fp32=residual_fp32)

    # ---- forward ----------------------------------------------------------------------------------------------
    def _workspace(self, B: int) -> torch.Tensor:
        need = self.lib.theia_vit_workspace_bytes(C.byref(self._desc), B)
        if need < 0:
            raise L.TheiaError("theia_vit_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @torch.no_grad()
    def __call__(self, pixel_values: torch.Tensor, out_dtype: torch.dtype = torch.float32):
        c = self.cfg
        if pixel_values.ndim != 4 or pixel_values.shape[1] != 3 or tuple(pixel_values.shape[2:]) != (c["image"], c["image"]):
            raise ValueError(f"pixel_values must be [B, 3, {c['image']}, {c['image']}], got {tuple(pixel_values.shape)}")
        B = int(pixel_values.shape[0])
        with torch.cuda.device(self.device):
            pv = pixel_values.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
            s = L.stream_ptr()
            patches = torch.empty(B * c["tokens"], c["patch_k"], dtype=torch.bfloat16, device=self.device)
            L.check(self.lib.theia_patchify_f32(pv.data_ptr(), patches.data_ptr(), B, 3, c["image"], c["image"], c["patch"],
                                                c["tokens"], c["patch_off"], c["patch_k"], s), "theia_patchify_f32")
            hid = torch.empty(B, c["tokens"], c["hidden"], dtype=torch.bfloat16, device=self.device)
            pooled = torch.empty(B, c["hidden"], dtype=torch.bfloat16, device=self.device)
            L.check(self.lib.theia_vit_forward(C.byref(self._desc), patches.data_ptr(), B, self._workspace(B).data_ptr(),
                                               hid.data_ptr(), pooled.data_ptr(), s), "theia_vit_forward")
        return hid.to(out_dtype), pooled.to(out_dtype)

    def to(self, *a, **k):  # the reference calls `.to(device)` on the HF model (dinov2.py:56); already bound
        return self

    def eval(self):
        return self


def _as_teacher(model, device=None) -> TeacherViT:
    """a loaded HF model is converted on first use and the converted teacher kept on the module (teachers are frozen:
    `feature_extraction.py` never updates them)"""
    if isinstance(model, TeacherViT):
        return model
    cached = getattr(model, "_theia_b200_teacher", None)
    dev = torch.device(device) if device is not None else next(model.parameters()).device
    if cached is None or cached.device != dev:
        cached = TeacherViT.from_hf(model, device=dev)
        object.__setattr__(model, "_theia_b200_teacher", cached)
    return cached


def _tokens_to_bchw(visual_tokens: torch.Tensor) -> torch.Tensor:
    b, n, ch = visual_tokens.size()
    g = int(np.sqrt(n))
    return visual_tokens.transpose(1, 2).reshape(b, ch, g, g)


def _forward(model, inputs, requires_grad: bool):
    if requires_grad:
        raise NotImplementedError("the CUDA teacher path is inference only (requires_grad=False)")
    return _as_teacher(model)(inputs["pixel_values"])


def get_dinov2_feature(model: TeacherViT, processor, images: list[np.ndarray], requires_grad: bool = False):
    """dinov2.py:8-41 -> (cls_token [B,1,D], visual_tokens [B,D,16,16], pooled_cls_token [B,1,D])"""
    hid, pooled = _forward(model, processor(images, return_tensors="pt"), requires_grad)
    return hid[:, :1], _tokens_to_bchw(hid[:, 1:]), pooled.unsqueeze(1)


def get_clip_feature(model: TeacherViT, processor, images: list[np.ndarray], requires_grad: bool = False):
    """clip.py:8-41 -> (cls_token, visual_tokens BCHW, pooled_cls_token = post_layernorm(cls))"""
    hid, pooled = _forward(model, processor(images=images, return_tensors="pt"), requires_grad)
    return hid[:, :1], _tokens_to_bchw(hid[:, 1:]), pooled.unsqueeze(1)


def get_vit_feature(model: TeacherViT, processor, images: list[np.ndarray], requires_grad: bool = False):
    """vit.py:8-33 -> (cls_token [B,D], last_hidden_state BCHW)"""
    hid, _ = _forward(model, processor(images, return_tensors="pt"), requires_grad)
    return hid[:, 0], _tokens_to_bchw(hid[:, 1:])


def _load(hf_cls_name: str, proc_cls_name: str, model_name: str, device):
    import transformers
    processor = getattr(transformers, proc_cls_name).from_pretrained(model_name)
    model = getattr(transformers, hf_cls_name).from_pretrained(model_name)
    return TeacherViT.from_hf(model, device=device), processor


def get_dinov2_model(model_name: str = "facebook/dinov2-large", device: Any = "cuda"):
    """dinov2.py:44-58"""
    return _load("Dinov2Model", "AutoImageProcessor", model_name, device)


def get_clip_model(model_name: str = "openai/clip-vit-large-patch14", device: Any = "cuda"):
    """clip.py:44-58"""
    return _load("CLIPVisionModel", "AutoProcessor", model_name, device)


def get_vit_model(model_name: str = "google/vit-huge-patch14-224-in21k", device: Any = "cuda"):
    """vit.py:36-50"""
    return _load("ViTModel", "AutoImageProcessor", model_name, device)


def get_model(model_name: str, device: Any = "cuda"):
    """feature_extraction_core/models.py:25-40, for the teachers this path covers"""
    if "google/vit" in model_name:
        return get_vit_model(model_name, device=device)
    if "openai/clip" in model_name:
        return get_clip_model(model_name, device=device)
    if "facebook/dinov2" in model_name:
        return get_dinov2_model(model_name, device=device)
    raise NotImplementedError(f"{model_name}: no CUDA teacher path (SAM / LLaVA / Depth-Anything are out of scope)")


def get_feature_outputs(model_name: str, model, processor, batch_images: list, dtype: torch.dtype = torch.bfloat16):
    """feature_extraction_core/models.py:55-95: the dict the extraction script writes to the feature shards"""
    features: dict[str, dict[str, torch.Tensor]] = {model_name: {}}
    fin = lambda t: t.detach().cpu().to(dtype).contiguous()
    if "google_vit" in model_name:
        cls_token, feature = get_vit_feature(model, processor, batch_images)
        features[model_name] = {"cls_token": fin(cls_token), "embedding": fin(feature)}
    elif "openai_clip" in model_name:
        cls_token, visual_tokens, pooled = get_clip_feature(model, processor, batch_images)
        features[model_name] = {"embedding": fin(visual_tokens), "cls_token": fin(cls_token), "pooled_cls_token": fin(pooled)}
    elif "facebook_dinov2" in model_name:
        cls_token, visual_tokens, pooled = get_dinov2_feature(model, processor, batch_images)
        features[model_name] = {"embedding": fin(visual_tokens), "cls_token": fin(cls_token), "pooled_cls_token": fin(pooled)}
    else:
        raise NotImplementedError(f"model {model_name} is not supported")
    return features
