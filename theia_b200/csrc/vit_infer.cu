// Teacher ViT inference: the forward of the frozen foundation models whose features the
// student is distilled from -- hf Dinov2Model / CLIPVisionModel / ViTModel as the reference calls them in
// src/theia/foundation_models/vision_models/dinov2.py:8-41, vision_language_models/clip.py:8-41 and
// vision_models/vit.py:8-33 (driven by src/theia/preprocessing/feature_extraction_core/models.py:55-95).
// Same kernels as the student forward (tcgen05 GEMM with fused epilogues, LayerNorm, TMEM attention), patch 14 /
// 257 tokens / D = 1024; weights are caller-owned bf16 / fp32 device buffers, activations live in a ping-pong
// workspace (nothing is kept for a backward pass).  Stateless: one call = one fixed launch sequence on one stream.
#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {
namespace {

inline long long align256(long long b) { return (b + 255) & ~255LL; }

struct VitWs {
  long long x0, x1, ln, qkv, attn, act, cls, total;
};

VitWs carve(const theia_vit_desc* d, int B) {
  const long long M = static_cast<long long>(B) * d->tokens, D = d->hidden;
  VitWs w;
  long long o = 0;
  auto take = [&](long long elems) {
    const long long at = o;
    o += align256(elems * 2);
    return at;
  };
  const long long xs = d->residual_f32 ? 2 : 1;  // fp32 residual stream: 4-byte elements
  w.x0 = take(xs * M * D), w.x1 = take(xs * M * D), w.ln = take(M * D), w.qkv = take(3 * M * D), w.attn = take(M * D);
  w.act = take(M * d->mlp), w.cls = take(xs * static_cast<long long>(B) * D);
  w.total = o;
  return w;
}

int check_desc(const theia_vit_desc* d, int B) {
  if (!d || !d->layer || !d->w_patch || !d->tok_table) return set_error(THEIA_ERR_ARG, "vit: null descriptor field");
  if (B < 1) return set_error(THEIA_ERR_ARG, "vit: batch %d", B);
  const int hd = d->heads > 0 ? d->hidden / d->heads : 0;
  if (d->heads < 1 || d->hidden != d->heads * hd || (hd != 64 && hd != 80))
    return set_error(THEIA_ERR_UNSUPPORTED, "vit: head dim %d (hidden %d / heads %d); the attention kernels are built for 64 and 80",
                     hd, d->hidden, d->heads);
  if (d->hidden > 1280 || d->hidden % 64 != 0) return set_error(THEIA_ERR_UNSUPPORTED, "vit: hidden %d (<= 1280, %% 64)", d->hidden);
  if (d->tokens < 1 || d->tokens > 272) return set_error(THEIA_ERR_UNSUPPORTED, "vit: %d tokens per image (<= 272)", d->tokens);
  if (d->patch_k % 8 != 0 || d->mlp % 8 != 0) return set_error(THEIA_ERR_ARG, "vit: patch_k / mlp must be multiples of 8");
  if (d->patch_off < 0 || d->patch_off + d->patch_tokens > d->tokens) return set_error(THEIA_ERR_ARG, "vit: patch token range");
  if (d->act != 0 && d->act != 1) return set_error(THEIA_ERR_ARG, "vit: act %d", d->act);
  return THEIA_OK;
}

int lin(cudaStream_t s, const void* x, const void* w, const float* bias, void* out, int M, int N, int K, int epi,
        const void* aux = nullptr) {
  theia_gemm_desc g;
  memset(&g, 0, sizeof(g));
  g.M = M, g.N = N, g.K = K, g.a_mode = THEIA_OP_K2D, g.b_mode = THEIA_OP_K2D, g.splits = 1, g.batch_z = 1;
  g.A = x, g.lda = K, g.B = w, g.ldb = K, g.out = out, g.ldo = N, g.bias = bias, g.epi = epi, g.aux = aux;
  return theia_gemm(&g, s);
}

// pixel_values fp32 [B][C][H][W] -> bf16 patch rows [B*tokens][patch_k]: row b*tokens + patch_off + py*gw + px, column
// c*p*p + i*p + j (= the flattened Conv2d(C, D, p, p) weight, hf:modeling_dinov2.py Dinov2PatchEmbeddings);
// columns >= C*p*p and the rows of non-patch tokens are zero
__global__ void __launch_bounds__(256) patchify_f32_kernel(const float* __restrict__ pv, bf16* __restrict__ out, long long total,
                                                           int C, int H, int W, int p, int tokens, int patch_off, int patch_k) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int col = static_cast<int>(i % patch_k);
  const long long row = i / patch_k;
  const int t = static_cast<int>(row % tokens) - patch_off;
  const long long b = row / tokens;
  const int gw = W / p, gh = H / p;
  float v = 0.f;
  if (t >= 0 && t < gw * gh && col < C * p * p) {
    const int c = col / (p * p), r = col - c * p * p;
    const int ii = r / p, jj = r - ii * p;
    const int py = t / gw, px = t - py * gw;
    v = pv[((b * C + c) * H + py * p + ii) * W + px * p + jj];
  }
  out[i] = __float2bfloat16(v);
}

}  // namespace
}  // namespace theia

using namespace theia;

#define TRY(x)        \
  do {                \
    int rc__ = (x);   \
    if (rc__) return rc__; \
  } while (0)

extern "C" int theia_patchify_f32(const float* pixel_values, void* patches, int B, int C, int H, int W, int patch, int tokens,
                                  int patch_off, int patch_k, void* stream) {
  if (!pixel_values || !patches || B < 1 || patch < 1 || H % patch != 0 || W % patch != 0 || patch_k < C * patch * patch ||
      patch_off + (H / patch) * (W / patch) > tokens)
    return set_error(THEIA_ERR_ARG, "patchify: bad geometry (B %d, %dx%dx%d, patch %d, tokens %d, K %d)", B, C, H, W, patch,
                     tokens, patch_k);
  const long long total = static_cast<long long>(B) * tokens * patch_k;
  patchify_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pixel_values, static_cast<bf16*>(patches), total, C, H, W, patch, tokens, patch_off, patch_k);
  THEIA_CHECK_LAUNCH("patchify_f32");
  return THEIA_OK;
}

extern "C" long long theia_vit_workspace_bytes(const theia_vit_desc* d, int B) {
  if (!d || B < 1) return -1;
  return carve(d, B).total;
}

extern "C" int theia_vit_forward(const theia_vit_desc* d, const void* patches, int B, void* workspace, void* last_hidden,
                                 void* pooled, void* stream) {
  TRY(check_desc(d, B));
  if (!patches || !workspace || !last_hidden) return set_error(THEIA_ERR_ARG, "vit: null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const VitWs w = carve(d, B);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  void* x0 = base + w.x0;  // residual stream, bf16 or fp32 (residual_f32)
  void* x1 = base + w.x1;
  bf16* ln = reinterpret_cast<bf16*>(base + w.ln);
  bf16* qkv = reinterpret_cast<bf16*>(base + w.qkv);
  bf16* attn = reinterpret_cast<bf16*>(base + w.attn);
  bf16* act = reinterpret_cast<bf16*>(base + w.act);
  void* cls = base + w.cls;
  const int D = d->hidden, NT = d->tokens, M = B * NT;
  const bool F = d->residual_f32 != 0;  // residual stream x0 / x1 in fp32
  // LayerNorm of the residual stream into a bf16 GEMM operand (or, CLIP pre_layrnorm, into the stream itself)
  auto ln_x = [&](const void* x, const float* g, const float* bta, void* y, bool y_stream, int rows) -> int {
    if (F) return theia_layernorm_fwd_f32(static_cast<const float*>(x), g, bta, y, y_stream ? 1 : 0, rows, D, d->ln_eps, s);
    return theia_layernorm_fwd(x, g, bta, y, nullptr, nullptr, rows, D, d->ln_eps, s);
  };
  // x_out = x_in + a W^T + b
  auto resid = [&](const void* a, const void* wt, const float* bias, void* x_out, const void* x_in, int K) -> int {
    return lin(s, a, wt, bias, x_out, M, D, K, F ? (THEIA_EPI_RESID_F32 | THEIA_EPI_OUT_F32) : THEIA_EPI_RESID, x_in);
  };
  {  // patch embedding + CLS / position table (hf:modeling_dinov2.py Dinov2Embeddings.forward; modeling_clip.py
     // CLIPVisionEmbeddings.forward; modeling_vit.py:100-128)
    theia_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.M = M, g.N = D, g.K = d->patch_k, g.a_mode = THEIA_OP_K2D, g.b_mode = THEIA_OP_K2D, g.splits = 1, g.batch_z = 1;
    g.A = patches, g.lda = d->patch_k, g.B = d->w_patch, g.ldb = d->patch_k;
    g.out = d->pre_ln_w ? x1 : x0, g.ldo = D, g.bias = d->b_patch;
    g.epi = THEIA_EPI_POSCLS | (F ? THEIA_EPI_OUT_F32 : 0), g.pos = d->tok_table, g.tokens = NT, g.tok_p0 = d->patch_off,
    g.tok_p1 = d->patch_off + d->patch_tokens;
    TRY(theia_gemm(&g, s));
    if (d->pre_ln_w) TRY(ln_x(x1, d->pre_ln_w, d->pre_ln_b, x0, true, M));  // CLIPVisionTransformer.pre_layrnorm
  }
  const int act_epi = d->act == 1 ? THEIA_EPI_QUICK_GELU : THEIA_EPI_GELU_FWD;
  for (int l = 0; l < d->layers; ++l) {  // pre-norm block: x += Wo attn(LN1 x);  x += W2 act(W1 LN2 x)
    const theia_vit_layer& p = d->layer[l];
    TRY(ln_x(x0, p.ln1_w, p.ln1_b, ln, false, M));
    TRY(lin(s, ln, p.w_qkv, p.b_qkv, qkv, M, 3 * D, D, 0));
    if (d->hidden == d->heads * 80)
      TRY(theia_attention_fwd_hd80(qkv, attn, nullptr, B, NT, d->heads, s));
    else
      TRY(theia_attention_tc_fwd(qkv, attn, nullptr, B, NT, d->heads, s));
    TRY(resid(attn, p.w_o, p.b_o, x1, x0, D));
    TRY(ln_x(x1, p.ln2_w, p.ln2_b, ln, false, M));
    TRY(lin(s, ln, p.w_fc1, p.b_fc1, act, M, d->mlp, D, act_epi));
    TRY(resid(act, p.w_fc2, p.b_fc2, x0, x1, d->mlp));
  }
  const size_t row = sizeof(bf16) * D, xrow = (F ? sizeof(float) : sizeof(bf16)) * static_cast<size_t>(D);
  cudaError_t e = cudaSuccess;
  if (d->final_ln_w && d->final_ln_mode == 1) {  // Dinov2Model.layernorm / ViTModel.layernorm over every token
    TRY(ln_x(x0, d->final_ln_w, d->final_ln_b, last_hidden, false, M));
    if (pooled) e = cudaMemcpy2DAsync(pooled, row, last_hidden, row * NT, row, B, cudaMemcpyDeviceToDevice, s);
  } else {
    if (F) TRY(theia_cast_bf16(reinterpret_cast<const float*>(x0), last_hidden, static_cast<long long>(M) * D, s));
    else e = cudaMemcpyAsync(last_hidden, x0, row * M, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && pooled) {
      if (d->final_ln_w) {  // CLIPVisionTransformer: pooler_output = post_layernorm(last_hidden_state[:, 0])
        e = cudaMemcpy2DAsync(cls, xrow, x0, xrow * NT, xrow, B, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) TRY(ln_x(cls, d->final_ln_w, d->final_ln_b, pooled, false, B));
      } else {
        e = cudaMemcpy2DAsync(pooled, row, last_hidden, row * NT, row, B, cudaMemcpyDeviceToDevice, s);
      }
    }
  }
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "vit: memcpy: %s", cudaGetErrorString(e));
  return THEIA_OK;
}
