"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Plain torch fp32 (CPU or GPU eager) restatement of the Theia distillation hot path
(student DeiT/ViT forward -> lconv feature-translator heads -> mse / cos / smooth-l1
losses), written from the behaviour of the reference and its third-party dependency:

  reference          /root/reference/src/theia/models/rvfm.py:94-185
                     /root/reference/src/theia/models/backbones.py:255-341 (DeiT wrapper)
                     /root/reference/src/theia/models/adapter_heads.py:232-359 (lconv head)
                     /root/reference/src/theia/models/feature_translators.py:68-88,159-205
                     /root/reference/src/theia/models/utils.py:8-43
  transformers 5.5.0 models/vit/modeling_vit.py:43-458 (un-pinned dependency carrying
                     the ViT arithmetic), models/deit/image_processing_deit.py and
                     image_processing_backends.py:200-414 (DeiT image processor)

Pinning: the reference ships no tests / golden vectors for this path ("parity unpinned" by
the reference's own tests).  This restatement is therefore pinned
against the reference modules themselves, imported from /root/reference in the build
container by oracle/make_golden.py (which asserts equality and writes the fixtures under
tests/golden/).  tests/test_oracle.py re-checks the restatement against those fixtures.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

# The oracle is an fp32 restatement: when it runs on a GPU (tests, smoke, bench parity check) cuDNN / cuBLAS must not
# silently drop to TF32 (cuDNN convolutions allow it by default).
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# (C, H, W) per teacher -- reference foundation_models/common.py:18-25
MODEL_FEATURE_SIZES = {
    "facebook/dinov2-large": (1024, 16, 16),
    "facebook/sam-vit-huge": (256, 64, 64),
    "google/vit-huge-patch14-224-in21k": (1280, 16, 16),
    "openai/clip-vit-large-patch14": (1024, 16, 16),
    "LiheYoung/depth-anything-large-hf": (32, 64, 64),
}
# reference configs/training/target_models/{cdiv,cddsv}.yaml
TEACHER_SETS = {
    "dinov2": ["facebook/dinov2-large"],
    "cdiv": ["google/vit-huge-patch14-224-in21k", "facebook/dinov2-large", "openai/clip-vit-large-patch14"],
    "cddsv": [
        "google/vit-huge-patch14-224-in21k",
        "facebook/dinov2-large",
        "openai/clip-vit-large-patch14",
        "facebook/sam-vit-huge",
        "LiheYoung/depth-anything-large-hf",
    ],
}
# hub configs of facebook/deit-{tiny,small,base}-patch16-224 (model_type "vit")
BACKBONES = {
    "facebook/deit-tiny-patch16-224": (192, 3),
    "facebook/deit-small-patch16-224": (384, 6),
    "facebook/deit-base-patch16-224": (768, 12),
}
# hub preprocessor_config.json of facebook/deit-*-patch16-224
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


@dataclass
class OracleConfig:
    hidden: int = 192
    heads: int = 3
    layers: int = 12
    patch: int = 16
    image: int = 224
    ln_eps: float = 1e-12  # ViTConfig.layer_norm_eps
    teachers: dict = field(default_factory=dict)  # name -> (C,H,W)
    # When True, activations / GEMM weights are rounded to bf16 at exactly the points where the CUDA
    # path stores them (fp32 accumulation everywhere).  Still the same algorithm; used by the GPU
    # tests to separate kernel bugs from the precision effect of bf16 storage (ReLU-mask flips make
    # the gradient a discontinuous function of the forward activations).
    emulate_bf16: bool = False
    # backbone variant (backbones.py:506-526): "deit" (CLS + 196 patches), "nocls" (DeiTNoCLS: 196 patches),
    # "reg" (DeiTReg: CLS + 196 patches + num_reg register tokens, stripped before the translator)
    variant: str = "deit"
    num_reg: int = 0

    @property
    def tokens(self) -> int:
        """rows of position_embeddings (always 197: ViTEmbeddings keeps the CLS slot even in the no-CLS variant)"""
        return (self.image // self.patch) ** 2 + 1

    @property
    def seq(self) -> int:
        n = (self.image // self.patch) ** 2
        return n + (0 if self.variant == "nocls" else 1) + self.num_reg


def make_config(backbone: str, teachers, distill_cls: bool = False) -> OracleConfig:
    """distill_cls: train_rvfm.py:239-246 -- ViT / DINOv2 / CLIP teachers also get a '<name>_cls' target of size
    (C_t,) predicted by a LinearAdapterHead on the CLS token."""
    variant, num_reg = "deit", 0
    if backbone.startswith("nocls-"):
        variant, backbone = "nocls", backbone[len("nocls-"):]
    elif backbone.startswith("reg-"):
        variant, num_reg, backbone = "reg", 7, backbone[len("reg-"):]
    d, h = BACKBONES[backbone]
    if isinstance(teachers, str):
        teachers = TEACHER_SETS[teachers]
    sizes = {}
    for t in teachers:
        sizes[t] = MODEL_FEATURE_SIZES[t[:-4]][:1] if t.endswith("_cls") else MODEL_FEATURE_SIZES[t]
    if distill_cls:
        for t in list(teachers):
            if "google/vit" in t or "facebook/dino" in t or "openai/clip" in t:
                sizes[t + "_cls"] = MODEL_FEATURE_SIZES[t][:1]
    return OracleConfig(hidden=d, heads=h, teachers=sizes, variant=variant, num_reg=num_reg)


def _r(x: torch.Tensor, cfg: "OracleConfig") -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if cfg.emulate_bf16 else x


def head_key(t: str) -> str:
    """feature_translators.py:46 -- ModuleDict keys cannot hold '.'"""
    return t.replace(".", "_")


# --------------------------------------------------------------------------------------
# parameter construction (deterministic, independent of the reference's RNG consumption
# order so the same tensors can be rebuilt on the GPU box without the reference)
# --------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> dict:
    """state_dict key -> shape, in the reference's layout."""
    D = cfg.hidden
    s = {}
    e = "backbone.model.embeddings."
    if cfg.variant != "nocls":
        s[e + "cls_token"] = (1, 1, D)
    s[e + "position_embeddings"] = (1, cfg.tokens, D)
    if cfg.variant == "reg":
        s[e + "reg_token"] = (1, cfg.num_reg, D)
        s[e + "reg_pos_embed"] = (1, cfg.num_reg, D)
    s[e + "patch_embeddings.projection.weight"] = (D, 3, cfg.patch, cfg.patch)
    s[e + "patch_embeddings.projection.bias"] = (D,)
    for l in range(cfg.layers):
        p = f"backbone.model.encoder.layer.{l}."
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (D, D)
            s[p + f"attention.attention.{n}.bias"] = (D,)
        s[p + "attention.output.dense.weight"] = (D, D)
        s[p + "attention.output.dense.bias"] = (D,)
        s[p + "intermediate.dense.weight"] = (4 * D, D)
        s[p + "intermediate.dense.bias"] = (4 * D,)
        s[p + "output.dense.weight"] = (D, 4 * D)
        s[p + "output.dense.bias"] = (D,)
        for n in ("layernorm_before", "layernorm_after"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
    s["backbone.model.layernorm.weight"] = (D,)
    s["backbone.model.layernorm.bias"] = (D,)
    C = D  # hidden_size_factor 1.0 (configs/model/translator/lconv.yaml:3)
    for t, size in cfg.teachers.items():
        p = f"translator.translator_heads.{head_key(t)}."
        if len(size) == 1:  # LinearAdapterHead (adapter_heads.py:28-58)
            s[p + "adapter.0.weight"] = (size[0], C)
            s[p + "adapter.0.bias"] = (size[0],)
            continue
        ct, ht, wt = size
        s[p + "pad.1.weight"] = (C, C, 3, 3)  # ConvTranspose2d: [Cin, Cout, kh, kw]
        s[p + "pad.1.bias"] = (C,)
        if ht == 16:
            sp = [(16, 16), (16, 16), (16, 16)]
        elif ht == 64:
            sp = [(16, 16), (31, 31), (64, 64)]
        else:
            raise NotImplementedError(ht)
        for i, k in enumerate((0, 3, 6)):
            s[p + f"adapter.{k}.weight"] = (C, *sp[i])
            s[p + f"adapter.{k}.bias"] = (C, *sp[i])
        for k in (1, 4):
            s[p + f"adapter.{k}.weight"] = (C, C, 3, 3)
            s[p + f"adapter.{k}.bias"] = (C,)
        s[p + "adapter.8.weight"] = (ct, C)
        s[p + "adapter.8.bias"] = (ct,)
    return s


def init_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32) -> dict:
    """Deterministic synthetic weights.  Statistics follow the reference's init
    (modeling_vit.py:385-399 trunc_normal std 0.02 / zeros / ones; torch defaults for the
    translator) but every tensor -- biases and LN affines included -- is made non-trivial so
    parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in param_shapes(cfg).items():
        if k.startswith("backbone"):
            if k.endswith("layernorm.weight") or "layernorm_" in k and k.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(shp, generator=g)
            elif k.endswith(".bias"):
                v = 0.02 * torch.randn(shp, generator=g)
            else:
                v = 0.02 * torch.randn(shp, generator=g).clamp_(-2, 2)
        else:
            if len(shp) == 3:  # LN[C,H,W] affine
                v = (1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(shp, generator=g)
            else:
                fan_in = shp[1] * (9 if len(shp) == 4 else 1) if len(shp) > 1 else cfg.hidden * 9
                bound = 1.0 / math.sqrt(fan_in)
                v = (torch.rand(shp, generator=g) * 2 - 1) * bound
        out[k] = v.to(dtype)
    return out


# --------------------------------------------------------------------------------------
# forward restatement
# --------------------------------------------------------------------------------------
def preprocess(x, do_resize=True, do_rescale=True, do_normalize=True,
               mean=IMAGE_MEAN, std=IMAGE_STD) -> torch.Tensor:
    """DeiT image processor inside DeiT.forward (backbones.py:337-339 ->
    image_processing_backends.py:361-414): channels-first, resize 256 bicubic antialias on
    uint8, centre crop 224, fused (x - 255*mean) / (255*std) -> fp32."""
    if x.dim() == 3:
        x = x[None]
    if x.shape[1] not in (1, 3) and x.shape[-1] in (1, 3):  # infer_channel_dimension_format
        x = x.permute(0, 3, 1, 2)
    if do_resize:
        from torchvision.transforms.v2 import functional as tvF
        x = tvF.resize(x, [256, 256], interpolation=tvF.InterpolationMode.BICUBIC, antialias=True)
    H, W = x.shape[-2:]
    if H != 224 or W != 224:  # center_crop is applied regardless of do_resize (image_processing_backends.py center_crop)
        if H < 224 or W < 224:  # smaller than the crop: zero padding, split (n // 2, (n + 1) // 2) per axis
            ph, pw = max(224 - H, 0), max(224 - W, 0)
            x = torch.nn.functional.pad(x, (pw // 2, (pw + 1) // 2, ph // 2, (ph + 1) // 2), value=0)
            H, W = x.shape[-2:]
        top, left = int((H - 224) / 2.0), int((W - 224) / 2.0)
        x = x[..., top:top + 224, left:left + 224]
    if do_normalize:
        m = torch.tensor(mean, dtype=torch.float32, device=x.device)
        s = torch.tensor(std, dtype=torch.float32, device=x.device)
        if do_rescale:
            m, s = m * 255.0, s * 255.0
        x = (x.to(torch.float32) - m[:, None, None]) / s[:, None, None]
    elif do_rescale:
        x = x * (1.0 / 255.0)
    return x.contiguous()


def vit_forward(P: dict, pix: torch.Tensor, cfg: OracleConfig, taps: Optional[dict] = None,
                force: Optional[dict] = None) -> torch.Tensor:
    """ViTModel.forward (modeling_vit.py:428-458), pooler = Identity."""
    D, nh = cfg.hidden, cfg.heads
    e = "backbone.model.embeddings."
    if pix.shape[-1] != cfg.image or pix.shape[-2] != cfg.image:
        raise ValueError(f"Input image size ({pix.shape[-2]}*{pix.shape[-1]}) doesn't match model "
                         f"({cfg.image}*{cfg.image}).")  # modeling_vit.py:160-165
    x = F.conv2d(_r(pix, cfg), _r(P[e + "patch_embeddings.projection.weight"], cfg),
                 P[e + "patch_embeddings.projection.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)  # [B,196,D]
    B = x.shape[0]
    if cfg.variant == "nocls":  # ViTEmbeddingsNoCLS.forward (backbones.py:70-93)
        x = _r(x + P[e + "position_embeddings"][:, 1:], cfg)
    elif cfg.variant == "reg":  # ViTEmbeddingsReg.forward (backbones.py:186-217)
        x = torch.cat([P[e + "cls_token"].expand(B, -1, -1), x, P[e + "reg_token"].expand(B, -1, -1)], dim=1)
        x = _r(x + torch.cat([P[e + "position_embeddings"], P[e + "reg_pos_embed"]], dim=1), cfg)
    else:
        x = _r(torch.cat([P[e + "cls_token"].expand(B, -1, -1), x], dim=1) + P[e + "position_embeddings"], cfg)

    def tap(name, l, v):
        """record an intermediate (taps) and/or teacher-force it (force): the VALUE becomes the given tensor,
        the gradient still flows through the oracle's own op (tests: backward parity on identical activations)"""
        if force is not None and (name, l) in force:
            v = v + (force[(name, l)].to(v.dtype).view_as(v) - v).detach()
        if taps is not None:
            taps[(name, l)] = v.detach()
        return v

    x = tap("x", 0, x)
    for l in range(cfg.layers):
        p = f"backbone.model.encoder.layer.{l}."
        h = _r(F.layer_norm(x, (D,), P[p + "layernorm_before.weight"], P[p + "layernorm_before.bias"], cfg.ln_eps), cfg)
        h = tap("ln1", l, h)
        q = _r(F.linear(h, _r(P[p + "attention.attention.query.weight"], cfg), P[p + "attention.attention.query.bias"]), cfg)
        k = _r(F.linear(h, _r(P[p + "attention.attention.key.weight"], cfg), P[p + "attention.attention.key.bias"]), cfg)
        v = _r(F.linear(h, _r(P[p + "attention.attention.value.weight"], cfg), P[p + "attention.attention.value.bias"]), cfg)
        N = x.shape[1]
        q, k, v = tap("qkv", l, torch.cat([q, k, v], dim=-1)).split(D, dim=-1)
        q, k, v = (t.view(B, N, nh, D // nh).transpose(1, 2) for t in (q, k, v))
        sc = (q @ k.transpose(2, 3)) * ((D // nh) ** -0.5)
        if cfg.emulate_bf16:  # unnormalised probabilities rounded to bf16 before P.V, fp32 row sum
            pe = torch.exp(sc - sc.amax(-1, keepdim=True))
            a = (_r(pe, cfg) @ v) / pe.sum(-1, keepdim=True)
        else:
            a = torch.softmax(sc, dim=-1) @ v
        a = _r(a.transpose(1, 2).reshape(B, N, D), cfg)
        a = tap("attn", l, a)
        x = _r(F.linear(a, _r(P[p + "attention.output.dense.weight"], cfg), P[p + "attention.output.dense.bias"]) + x, cfg)
        x = tap("xmid", l, x)
        h = _r(F.layer_norm(x, (D,), P[p + "layernorm_after.weight"], P[p + "layernorm_after.bias"], cfg.ln_eps), cfg)
        h = tap("ln2", l, h)
        pre = F.linear(h, _r(P[p + "intermediate.dense.weight"], cfg), P[p + "intermediate.dense.bias"])
        pre_t = tap("h", l, _r(pre, cfg))
        h = _r(F.gelu(pre_t if force is not None else pre), cfg)
        h = tap("a", l, h)
        x = _r(F.linear(h, _r(P[p + "output.dense.weight"], cfg), P[p + "output.dense.bias"]) + x, cfg)
        x = tap("x", l + 1, x)
    return _r(F.layer_norm(x, (D,), P["backbone.model.layernorm.weight"], P["backbone.model.layernorm.bias"],
                           cfg.ln_eps), cfg)


def lconv_head_forward(P: dict, t: str, x: torch.Tensor, cfg: OracleConfig, taps: Optional[dict] = None,
                       force: Optional[dict] = None) -> torch.Tensor:
    """LightConvAdapterHead.forward (adapter_heads.py:352-359) for a 14x14 source; LinearAdapterHead.forward
    (adapter_heads.py:51-58) for '<name>_cls' targets."""
    p = f"translator.translator_heads.{head_key(t)}."
    if len(cfg.teachers[t]) == 1:
        return F.linear(x[:, 0], _r(P[p + "adapter.0.weight"], cfg), P[p + "adapter.0.bias"])
    ct, ht, wt = cfg.teachers[t]
    B, C = x.shape[0], x.shape[2]
    g = cfg.image // cfg.patch
    if cfg.variant != "nocls":
        x = x[:, 1:]  # drop CLS (adapter_heads.py:355-356, backbone_no_cls False)
    y = x.reshape(B, g, g, C).permute(0, 3, 1, 2)  # b (h w) c -> b c h w
    y = _r(F.conv_transpose2d(y, _r(P[p + "pad.1.weight"], cfg), P[p + "pad.1.bias"], stride=1), cfg)  # 14 -> 16

    def tap(name, v):  # NHWC like the CUDA path stores it
        if force is not None and (name, t) in force:
            v = v + (force[(name, t)].to(v.dtype).permute(0, 3, 1, 2) - v).detach()
        if taps is not None:
            taps[(name, t)] = v.detach().permute(0, 2, 3, 1).contiguous()
        return v

    y = tap("padout", y)

    def ln(y, k):
        return _r(F.layer_norm(y, y.shape[1:], P[p + f"adapter.{k}.weight"], P[p + f"adapter.{k}.bias"], 1e-5), cfg)

    def w(k):
        return _r(P[p + f"adapter.{k}.weight"], cfg)

    y = ln(y, 0)
    y = tap("hln0", y)
    if ht == 16:
        y = _r(F.relu(F.conv2d(y, w(1), P[p + "adapter.1.bias"], padding=1)), cfg)
        y = tap("c1", y)
        y = ln(y, 3)
        y = tap("hln1", y)
        y = _r(F.relu(F.conv2d(y, w(4), P[p + "adapter.4.bias"], padding=1)), cfg)
        y = tap("c2", y)
        y = ln(y, 6)
        y = tap("hln2", y)
    else:
        y = _r(F.relu(F.conv_transpose2d(y, w(1), P[p + "adapter.1.bias"], stride=2, padding=1)), cfg)
        y = ln(y, 3)
        y = _r(F.relu(F.conv_transpose2d(y, w(4), P[p + "adapter.4.bias"], stride=2, output_padding=1)), cfg)
        y = ln(y, 6)
    y = y.flatten(2).transpose(1, 2)  # b c h w -> b (h w) c
    return F.linear(y, w(8), P[p + "adapter.8.bias"])


def handle_feature_output(x, feature_reduce_method=None, num_discard_tokens=0):
    """models/utils.py:31-43."""
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


def backbone_forward(P, images, cfg, taps=None, force=None, **kw):
    return vit_forward(P, preprocess(images, **kw).to(next(iter(P.values())).device), cfg, taps, force)


def forward_feature(P, images, cfg, feature_reduce_method=None, **kw):
    """RobotVisionFM.forward_feature (rvfm.py:94-113)."""
    return handle_feature_output(backbone_forward(P, images, cfg, **kw), feature_reduce_method, cfg.num_reg)


def forward(P, images, cfg, target_model_names=None, taps=None, force=None, **kw) -> dict:
    """RobotVisionFM.forward (rvfm.py:115-136)."""
    x = backbone_forward(P, images, cfg, taps, force, **kw)
    if force is not None and ("tokens", 0) in force:
        x = x + (force[("tokens", 0)].to(x.dtype).view_as(x) - x).detach()
    if taps is not None:
        taps[("tokens", 0)] = x.detach()
    if cfg.num_reg > 0:
        x = x[:, :-cfg.num_reg]  # rvfm.py:133-134
    names = target_model_names if target_model_names is not None else list(cfg.teachers)
    return {t: lconv_head_forward(P, t, x, cfg, taps, force) for t in names}


def get_loss(pred: dict, y: dict, target_loss_weights=None) -> dict:
    """RobotVisionFM.get_loss (rvfm.py:138-185): nn.MSELoss, nn.SmoothL1Loss(beta=1),
    F.normalize(flatten) + nn.CosineEmbeddingLoss(target=+1), each weighted 1/T."""
    T = len(pred)
    mse_avg, cos_avg, l1_avg = 0, 0, 0
    per = {"mse": {}, "cos": {}, "l1": {}}
    for t in pred:
        p, tg = pred[t], y[t]
        d = p - tg
        mse = (d * d).mean()
        ad = d.abs()
        l1 = torch.where(ad < 1.0, 0.5 * d * d, ad - 0.5).mean()
        pn = p.flatten(1)
        tn = tg.flatten(1)
        pn = pn / pn.norm(dim=1, keepdim=True).clamp_min(1e-12)
        tn = tn / tn.norm(dim=1, keepdim=True).clamp_min(1e-12)
        cosv = (pn * tn).sum(1) / torch.sqrt(((pn * pn).sum(1) + 1e-8) * ((tn * tn).sum(1) + 1e-8))
        cos = (1.0 - cosv).mean()
        w = target_loss_weights if target_loss_weights else 1.0 / T
        mse_avg = mse_avg + mse * w
        cos_avg = cos_avg + cos / T
        l1_avg = l1_avg + l1 * w
        per["mse"][t], per["cos"][t], per["l1"][t] = mse.item(), cos.item(), l1.item()
    return {"mse_loss": mse_avg, "cos_loss": cos_avg, "l1_loss": l1_avg,
            "mse_losses_per_model": per["mse"], "cos_losses_per_model": per["cos"],
            "l1_losses_per_model": per["l1"]}


def main_loss(losses: dict, kind: str = "cos_l1"):
    """train_rvfm.py:119-122."""
    if kind in ("mse", None):
        return losses["mse_loss"]
    return 0.9 * losses["cos_loss"] + 0.1 * losses["l1_loss"]


def distill_step(P: dict, images, targets: dict, cfg: OracleConfig, kind="cos_l1", force=None, **kw):
    """Replay of train_rvfm.py:116-125 (forward, get_loss, main loss, backward) under autograd.
    Returns (preds, losses, grads)."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    preds = forward(Pg, images, cfg, force=force, **kw)
    losses = get_loss(preds, targets)
    ml = main_loss(losses, kind)
    ml.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return {k: v.detach() for k, v in preds.items()}, losses, grads


def synthetic_batch(cfg: OracleConfig, B: int, seed: int = 0, device="cpu"):
    """SURVEY section 8d inputs: uniform uint8 images, N(0,1) targets rounded through bf16."""
    g = torch.Generator().manual_seed(1000 + seed)
    images = torch.randint(0, 256, (B, cfg.image, cfg.image, 3), dtype=torch.uint8, generator=g)
    g2 = torch.Generator().manual_seed(2000 + seed)
    targets = {t: torch.randn((B, sz[1] * sz[2], sz[0]) if len(sz) == 3 else (B, sz[0]), generator=g2)
               .to(torch.bfloat16).float() for t, sz in cfg.teachers.items()}
    return images.to(device), {k: v.to(device) for k, v in targets.items()}
