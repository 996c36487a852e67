// Orchestration of the distillation step: student ViT forward -> lconv translator heads
// (forward), and the full backward, as a fixed sequence of kernel launches on one stream.
// The context owns NO device memory: the caller binds a flat fp32 parameter buffer, a flat
// fp32 gradient buffer of the same layout and one workspace; this file carves the workspace.
//
// Mirrors: src/theia/models/rvfm.py:115-136 (forward), hf:models/vit/modeling_vit.py:100-458
// (ViTModel), src/theia/models/adapter_heads.py:279-359 (LightConvAdapterHead, 16x16 targets),
// and their autograd graphs (train_rvfm.py:125).
#include <stdlib.h>
#include <string>
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

struct PInfo {
  std::string name;
  int ndim;
  long long dims[4];
  long long off;    // element offset in the flat fp32 buffers
  long long numel;
};

struct LayerP {
  int ln1w, ln1b, qw, kw, vw, qb, kb, vb, ow, ob, ln2w, ln2b, f1w, f1b, f2w, f2b;
};
struct HeadP {
  int padw, padb, g0, b0, c1w, c1b, g1, b1, c2w, c2b, g2, b2, lw, lb;
  int ct;
  int hw;          // target H = W: 16 (Conv2d 3x3 stack) or 64 (two stride-2 ConvTranspose2d, adapter_heads.py:304-315)
  int v1, p1;      // stage-1 map: valid extent / storage pitch (16/16 or 31/32)
  int v2, p2;      // stage-2 map (16/16 or 64/64)
};

// bf16 packed weights (offsets in elements of the bf16 pack buffer)
// Linear weights keep ONE bf16 copy [out][in]: the forward reads it as a K-major B operand, the dgrad as an
// MN-major one (no transposed copies).
struct LayerW {
  long long wqkv, wo, w1, w2;
};
struct HeadW {
  long long padF, padD, c1F, c1D, c2F, c2D, l;
  long long gb[3][2];  // fp32 pack: gamma/beta HWC for the three LNs (offsets in floats)
};

struct LayerA {
  long long ln1, qkv, attn, xmid, ln2, h, a;  // bf16 element offsets (h holds gelu'(pre-activation))
  long long mean1, rstd1, mean2, rstd2, lse;  // fp32 element offsets
};
struct HeadA {
  long long padout, ln0, c1, ln1, c2, ln2;  // bf16
  long long stats;                          // fp32 [3][B][2]
  long long wsc[3];                         // fp32 backward scratch: conv weight gradients [tap][..][..] per conv
  long long gbs[3];                         // fp32 backward scratch: LN dgamma | dbeta (HWC) per LayerNorm
};

// one strided layout conversion of the pack / unpack tables (pointers are resolved at bind time)
enum { SEG_IN_MASTER = 0, SEG_IN_ACTF32 = 1 };
enum { SEG_OUT_PACKBF = 0, SEG_OUT_PACKF32 = 1, SEG_OUT_GRADS = 2 };
// one conv weight of the tiled pack / unpack tables (theia_conv_pack / theia_conv_unpack)
struct ConvSpec {
  long long w_off;                 // parameter offset (floats) in the flat master / gradient buffers
  long long pack0, pack1;          // pack: bf16 pack offsets; unpack: pack0 = fp32 scratch offset (act f32 region)
  int flags0, flags1;
};
struct SegSpec {
  int in_base, out_base;
  long long in_off, out_off;
  int n0, n1, n2, n3;
  long long s0, s1, s2, s3, base;
  int lim1, lim2;
};

}  // namespace theia

using namespace theia;

struct theia_model {
  theia_model_config cfg;
  int D, H, L, T, Bmax;
  int N, p0, R;            // tokens per image, index of the first patch token, register tokens
  int regt, regp;          // DeiTReg parameters (indices), -1 otherwise
  long long tokt, tgrad;   // token table (fp32 pack) / its gradient scratch
  std::vector<PInfo> params;
  long long n_params_total;  // floats in the flat buffer
  int cls, pos, pew, peb, lnfw, lnfb;
  std::vector<LayerP> lp;
  std::vector<HeadP> hp;
  // bound buffers
  float* master = nullptr;
  float* grads = nullptr;
  uint8_t* ws = nullptr;
  // workspace carving (byte offsets)
  long long ws_bytes = 0;
  long long o_packbf = 0, o_packf32 = 0, o_actbf = 0, o_actf32 = 0;
  long long n_packbf = 0, n_packf32 = 0, n_actbf = 0, n_actf32 = 0;
  long long wpe;
  std::vector<LayerW> lw;
  std::vector<HeadW> hw;
  // activations
  long long patches, tokens, meanf, rstdf;
  std::vector<long long> x;  // L+1
  std::vector<LayerA> la;
  std::vector<HeadA> ha;
  // backward scratch
  long long dtok, dx0, dx1, dln, dqkv, dh, dattn, dA0, dA1;  // bf16
  long long red;                                              // fp32
  long long bscr = 0, n_bscr = 0;                             // fp32 backward scratch (zeroed once per backward)
  // pack table (one int per 64 floats of the flat buffer: destination 64-block in the bf16 pack, or -1) and the
  // segment tables of the permuted packs / gradient unpacks; uploaded into the workspace by theia_model_bind
  std::vector<int> h_table;
  std::vector<SegSpec> pack_specs, unpack_specs;
  std::vector<ConvSpec> conv_pack, conv_unpack;
  long long o_table = 0, o_pack_segs = 0, o_unpack_segs = 0;  // byte offsets in the workspace
  long long o_conv_pack = 0, o_conv_unpack = 0;
  long long pack_blocks = 0, unpack_blocks = 0;
  int last_B = 0;
  int in_h = 224, in_w = 224;  // extent of the images theia_model_forward receives (theia_model_set_input_size)
  int in_f32 = 0;              // pixel type of those images: 0 = uint8, 1 = fp32 (theia_model_set_input_dtype)
};

namespace {

long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

int add_param(theia_model* m, const std::string& name, std::initializer_list<long long> dims) {
  PInfo p;
  p.name = name;
  p.ndim = static_cast<int>(dims.size());
  p.numel = 1;
  int i = 0;
  for (long long d : dims) {
    p.dims[i++] = d;
    p.numel *= d;
  }
  for (; i < 4; ++i) p.dims[i] = 1;
  p.off = m->n_params_total;
  m->n_params_total = align_up(m->n_params_total + p.numel, 64);
  m->params.push_back(p);
  return static_cast<int>(m->params.size()) - 1;
}

std::string legit(const char* t) {
  std::string s(t);
  for (auto& c : s)
    if (c == '.') c = '_';
  return s;
}

struct Carver {
  long long n = 0;
  long long take(long long elems, long long align_elems) {
    n = align_up(n, align_elems);
    const long long o = n;
    n += elems;
    return o;
  }
};

}  // namespace

extern "C" int theia_model_create(const theia_model_config* cfg, theia_model** out) {
  if (!cfg || !out) return set_error(THEIA_ERR_ARG, "model_create: null");
  if (cfg->hidden % 64 != 0 || cfg->heads * 64 != cfg->hidden)
    return set_error(THEIA_ERR_UNSUPPORTED, "hidden must be heads*64");
  if (cfg->image != 224 || cfg->patch != 16) return set_error(THEIA_ERR_UNSUPPORTED, "only 224/16 ViT geometry");
  if (cfg->num_teachers < 0 || cfg->num_teachers > THEIA_MAX_TEACHERS) return set_error(THEIA_ERR_ARG, "num_teachers");
  if (cfg->variant < 0 || cfg->variant > 2) return set_error(THEIA_ERR_ARG, "variant must be 0 (DeiT), 1 (NoCLS) or 2 (Reg)");
  if (cfg->variant == 2 && (cfg->num_reg_tokens < 1 || cfg->num_reg_tokens > 11))
    return set_error(THEIA_ERR_UNSUPPORTED, "DeiTReg: 1..11 register tokens (sequence <= 208)");
  for (int t = 0; t < cfg->num_teachers; ++t) {
    if (cfg->teacher_hw[t] == 1 && cfg->variant == 1)
      return set_error(THEIA_ERR_UNSUPPORTED, "CLS-token heads need a backbone with a CLS token");
    if (cfg->teacher_hw[t] != 16 && cfg->teacher_hw[t] != 64 && cfg->teacher_hw[t] != 1)
      return set_error(THEIA_ERR_UNSUPPORTED, "teacher %s: target maps must be 16x16, 64x64 or a CLS vector (got %d)",
                       cfg->teacher_names[t], cfg->teacher_hw[t]);
    if (cfg->teacher_c[t] % 8 != 0) return set_error(THEIA_ERR_UNSUPPORTED, "teacher channels %% 8 != 0");
  }
  theia_model* m = new theia_model();
  m->cfg = *cfg;
  const int D = m->D = cfg->hidden;
  m->H = cfg->heads;
  const int L = m->L = cfg->layers;
  const int T = m->T = cfg->num_teachers;
  const int B = m->Bmax = cfg->max_batch;
  m->n_params_total = 0;
  m->R = cfg->variant == 2 ? cfg->num_reg_tokens : 0;
  m->p0 = cfg->variant == 1 ? 0 : 1;
  m->N = 196 + m->p0 + m->R;
  const std::string e = "backbone.model.embeddings.";
  m->cls = m->regt = m->regp = -1;
  if (cfg->variant != 1) m->cls = add_param(m, e + "cls_token", {1, 1, D});
  m->pos = add_param(m, e + "position_embeddings", {1, 197, D});
  if (cfg->variant == 2) {
    m->regt = add_param(m, e + "reg_token", {1, m->R, D});
    m->regp = add_param(m, e + "reg_pos_embed", {1, m->R, D});
  }
  m->pew = add_param(m, e + "patch_embeddings.projection.weight", {D, 3, 16, 16});
  m->peb = add_param(m, e + "patch_embeddings.projection.bias", {D});
  for (int l = 0; l < L; ++l) {
    const std::string p = "backbone.model.encoder.layer." + std::to_string(l) + ".";
    LayerP q;
    q.ln1w = add_param(m, p + "layernorm_before.weight", {D});
    q.ln1b = add_param(m, p + "layernorm_before.bias", {D});
    q.qw = add_param(m, p + "attention.attention.query.weight", {D, D});
    q.kw = add_param(m, p + "attention.attention.key.weight", {D, D});
    q.vw = add_param(m, p + "attention.attention.value.weight", {D, D});
    q.qb = add_param(m, p + "attention.attention.query.bias", {D});
    q.kb = add_param(m, p + "attention.attention.key.bias", {D});
    q.vb = add_param(m, p + "attention.attention.value.bias", {D});
    q.ow = add_param(m, p + "attention.output.dense.weight", {D, D});
    q.ob = add_param(m, p + "attention.output.dense.bias", {D});
    q.ln2w = add_param(m, p + "layernorm_after.weight", {D});
    q.ln2b = add_param(m, p + "layernorm_after.bias", {D});
    q.f1w = add_param(m, p + "intermediate.dense.weight", {4 * D, D});
    q.f1b = add_param(m, p + "intermediate.dense.bias", {4 * D});
    q.f2w = add_param(m, p + "output.dense.weight", {D, 4 * D});
    q.f2b = add_param(m, p + "output.dense.bias", {D});
    m->lp.push_back(q);
  }
  m->lnfw = add_param(m, "backbone.model.layernorm.weight", {D});
  m->lnfb = add_param(m, "backbone.model.layernorm.bias", {D});
  const int C = D;
  for (int t = 0; t < T; ++t) {
    const std::string p = "translator.translator_heads." + legit(cfg->teacher_names[t]) + ".";
    HeadP q;
    q.ct = cfg->teacher_c[t];
    q.hw = cfg->teacher_hw[t];
    if (q.hw == 1) {  // LinearAdapterHead on the CLS token (adapter_heads.py:28-58; train_rvfm.py:239-246)
      q.v1 = q.p1 = q.v2 = q.p2 = 0;
      q.padw = q.padb = q.g0 = q.b0 = q.c1w = q.c1b = q.g1 = q.b1 = q.c2w = q.c2b = q.g2 = q.b2 = -1;
      q.lw = add_param(m, p + "adapter.0.weight", {q.ct, C});
      q.lb = add_param(m, p + "adapter.0.bias", {q.ct});
      m->hp.push_back(q);
      continue;
    }
    q.v1 = q.hw == 16 ? 16 : 31, q.p1 = q.hw == 16 ? 16 : 32;
    q.v2 = q.hw == 16 ? 16 : 64, q.p2 = q.v2;
    q.padw = add_param(m, p + "pad.1.weight", {C, C, 3, 3});
    q.padb = add_param(m, p + "pad.1.bias", {C});
    q.g0 = add_param(m, p + "adapter.0.weight", {C, 16, 16});
    q.b0 = add_param(m, p + "adapter.0.bias", {C, 16, 16});
    q.c1w = add_param(m, p + "adapter.1.weight", {C, C, 3, 3});
    q.c1b = add_param(m, p + "adapter.1.bias", {C});
    q.g1 = add_param(m, p + "adapter.3.weight", {C, q.v1, q.v1});
    q.b1 = add_param(m, p + "adapter.3.bias", {C, q.v1, q.v1});
    q.c2w = add_param(m, p + "adapter.4.weight", {C, C, 3, 3});
    q.c2b = add_param(m, p + "adapter.4.bias", {C});
    q.g2 = add_param(m, p + "adapter.6.weight", {C, q.v2, q.v2});
    q.b2 = add_param(m, p + "adapter.6.bias", {C, q.v2, q.v2});
    q.lw = add_param(m, p + "adapter.8.weight", {q.ct, C});
    q.lb = add_param(m, p + "adapter.8.bias", {q.ct});
    m->hp.push_back(q);
  }

  // ---- workspace carving ----
  const long long M = static_cast<long long>(B) * m->N, P = static_cast<long long>(B) * 256;
  Carver pb, pf, ab, af;
  const long long AL = 128;  // 256-byte alignment for bf16, 512 for fp32: fine for TMA (16 B) and vectors
  m->wpe = pb.take(static_cast<long long>(D) * 768, AL);
  m->tokt = pf.take(static_cast<long long>(m->N) * D, AL);
  m->tgrad = af.take(static_cast<long long>(m->N) * D, AL);
  m->h_table.assign(static_cast<size_t>(m->n_params_total / 64), -1);
  // pack-table entry: parameter pi (numel a multiple of 64) is cast into the bf16 pack at element offset dst
  auto table_cast = [&](int pi, long long dst) {
    const PInfo& q = m->params[pi];
    for (long long j = 0; j < (q.numel + 63) / 64; ++j) m->h_table[q.off / 64 + j] = static_cast<int>(dst / 64 + j);
  };
  table_cast(m->pew, m->wpe);
  for (int l = 0; l < L; ++l) {
    LayerW w;
    const long long DD = static_cast<long long>(D) * D;
    w.wqkv = pb.take(3 * DD, AL);
    w.wo = pb.take(DD, AL);
    w.w1 = pb.take(4 * DD, AL);
    w.w2 = pb.take(4 * DD, AL);
    m->lw.push_back(w);
    const LayerP& q = m->lp[l];
    table_cast(q.qw, w.wqkv);  // q, k, v are adjacent in the flat buffer (D*D is a multiple of 64)
    table_cast(q.kw, w.wqkv + DD);
    table_cast(q.vw, w.wqkv + 2 * DD);
    table_cast(q.ow, w.wo);
    table_cast(q.f1w, w.w1);
    table_cast(q.f2w, w.w2);
  }
  auto seg = [&](std::vector<SegSpec>& v, int ib, long long ioff, int ob, long long ooff, int n0, int n1, int n2, int n3,
                 long long s0, long long s1, long long s2, long long s3, long long base, int lim1 = 0, int lim2 = 0) {
    SegSpec q{ib, ob, ioff, ooff, n0, n1, n2, n3, s0, s1, s2, s3, base, lim1, lim2};
    v.push_back(q);
  };
  for (int t = 0; t < T; ++t) {
    HeadW w;
    memset(&w, 0, sizeof(w));
    const HeadP& hp = m->hp[t];
    if (hp.hw == 1) {
      w.l = pb.take(static_cast<long long>(hp.ct) * C, AL);
      table_cast(hp.lw, w.l);
      m->hw.push_back(w);
      continue;
    }
    const long long W9 = 9LL * C * C, C9 = 9LL * C;
    w.padF = pb.take(W9, AL);
    w.padD = pb.take(W9, AL);
    w.c1F = pb.take(W9, AL);
    w.c1D = pb.take(W9, AL);
    w.c2F = pb.take(W9, AL);
    w.c2D = pb.take(W9, AL);
    w.l = pb.take(static_cast<long long>(hp.ct) * C, AL);
    table_cast(hp.lw, w.l);
    const long long npix[3] = {256, 1LL * hp.p1 * hp.p1, 1LL * hp.p2 * hp.p2};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 2; ++j) w.gb[i][j] = pf.take(npix[i] * C, AL);
    m->hw.push_back(w);
    auto& ps = m->pack_specs;
    const auto off = [&](int pi) { return m->params[pi].off; };
    // ConvTranspose2d weight Wt[ci][co][3][3] (a = ci, b = co): fwd pack F[co][tap2][ci] = Wt[ci][co][8-tap2],
    // dgrad pack D[ci][tap][co] = Wt[ci][co][tap]
    m->conv_pack.push_back(ConvSpec{off(hp.padw), w.padF, w.padD, THEIA_CP_SWAP | THEIA_CP_FLIP, 0});
    const int cw[2] = {hp.c1w, hp.c2w};
    const long long cF[2] = {w.c1F, w.c2F}, cD[2] = {w.c1D, w.c2D};
    for (int i = 0; i < 2; ++i) {
      if (hp.hw == 16)  // Conv2d W[co][ci][3][3]: fwd F[co][tap][ci]; dgrad D[ci][tap2][co] = W[co][ci][8-tap2]
        m->conv_pack.push_back(ConvSpec{off(cw[i]), cF[i], cD[i], 0, THEIA_CP_SWAP | THEIA_CP_FLIP});
      else  // ConvTranspose2d(s2) Wt[ci][co][3][3] -> tap-major F[tap][co][ci] and D[tap][ci][co]
        m->conv_pack.push_back(ConvSpec{off(cw[i]), cF[i], cD[i], THEIA_CP_TAPMAJOR | THEIA_CP_SWAP, THEIA_CP_TAPMAJOR});
    }
    // LN affine [C][Hv][Wv] -> NHWC [Hp][Wp][C] (zero padded for the 31x31 stage)
    const int gbp[3][2] = {{hp.g0, hp.b0}, {hp.g1, hp.b1}, {hp.g2, hp.b2}};
    const int vv[3] = {16, hp.v1, hp.v2}, pp[3] = {16, hp.p1, hp.p2};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 2; ++j)
        seg(ps, SEG_IN_MASTER, off(gbp[i][j]), SEG_OUT_PACKF32, w.gb[i][j], 1, pp[i], pp[i], C, 0, vv[i], 1,
            1LL * vv[i] * vv[i], 0, vv[i], vv[i]);
  }
  m->patches = ab.take(M * 768, AL);
  m->x.resize(L + 1);
  for (int l = 0; l <= L; ++l) m->x[l] = ab.take(M * D, AL);
  for (int l = 0; l < L; ++l) {
    LayerA a;
    a.ln1 = ab.take(M * D, AL);
    a.qkv = ab.take(M * 3 * D, AL);
    a.attn = ab.take(M * D, AL);
    a.xmid = ab.take(M * D, AL);
    a.ln2 = ab.take(M * D, AL);
    a.h = ab.take(M * 2 * D, AL);  // gelu'(pre-activation) as 8-bit codes (THEIA_EPI_AUX_U8): M * 4 D bytes
    a.a = ab.take(M * 4 * D, AL);
    a.mean1 = af.take(M, AL);
    a.rstd1 = af.take(M, AL);
    a.mean2 = af.take(M, AL);
    a.rstd2 = af.take(M, AL);
    a.lse = af.take(static_cast<long long>(B) * m->H * m->N, AL);
    m->la.push_back(a);
  }
  m->tokens = ab.take(M * D, AL);
  m->meanf = af.take(M, AL);
  m->rstdf = af.take(M, AL);
  long long Pmax = P;
  for (int t = 0; t < T; ++t) {
    HeadA a;
    memset(&a, 0, sizeof(a));
    if (m->hp[t].hw == 1) {
      m->ha.push_back(a);
      continue;
    }
    const long long P1 = static_cast<long long>(B) * m->hp[t].p1 * m->hp[t].p1;
    const long long P2 = static_cast<long long>(B) * m->hp[t].p2 * m->hp[t].p2;
    if (P2 > Pmax) Pmax = P2;
    a.padout = ab.take(P * C, AL);
    a.ln0 = ab.take(P * C, AL);
    a.c1 = ab.take(P1 * C, AL);
    a.ln1 = ab.take(P1 * C, AL);
    a.c2 = ab.take(P2 * C, AL);
    a.ln2 = ab.take(P2 * C, AL);
    a.stats = af.take(3LL * B * 2, AL);
    m->ha.push_back(a);
  }
  {  // backward scratch of the heads (contiguous: one memset per backward) and the gradient unpack table
    Carver sc;
    for (int t = 0; t < T; ++t) {
      const HeadP& hp = m->hp[t];
      if (hp.hw == 1) continue;
      HeadA& a = m->ha[t];
      const long long npix[3] = {256, 1LL * hp.p1 * hp.p1, 1LL * hp.p2 * hp.p2};
      for (int i = 0; i < 3; ++i) {
        a.wsc[i] = sc.take(9LL * C * C, AL);
        a.gbs[i] = sc.take(2 * npix[i] * C, AL);
      }
    }
    m->n_bscr = align_up(sc.n, AL);
    m->bscr = af.take(m->n_bscr, AL);
    auto& us = m->unpack_specs;
    for (int t = 0; t < T; ++t) {
      const HeadP& hp = m->hp[t];
      if (hp.hw == 1) continue;
      const HeadA& a = m->ha[t];
      const auto off = [&](int pi) { return m->params[pi].off; };
      // pad (stride-1 ConvTranspose2d): grad Wt[ci][co][t] = ws[8-t][co][ci];  Conv2d: grad W[co][ci][tap] =
      // ws[tap][co][ci];  ConvTranspose2d(s2): grad Wt[ci][co][tap] = ws[tap][ci][co]
      m->conv_unpack.push_back(ConvSpec{off(hp.padw), m->bscr + a.wsc[0], 0, THEIA_CP_SWAP | THEIA_CP_FLIP, 0});
      m->conv_unpack.push_back(ConvSpec{off(hp.c1w), m->bscr + a.wsc[1], 0, 0, 0});
      m->conv_unpack.push_back(ConvSpec{off(hp.c2w), m->bscr + a.wsc[2], 0, 0, 0});
      const int gbp[3][2] = {{hp.g0, hp.b0}, {hp.g1, hp.b1}, {hp.g2, hp.b2}};
      const int vv[3] = {16, hp.v1, hp.v2}, pp[3] = {16, hp.p1, hp.p2};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j)  // NHWC (pitch pp) -> [C][Hv][Wv]
          seg(us, SEG_IN_ACTF32, m->bscr + a.gbs[i] + j * 1LL * pp[i] * pp[i] * C, SEG_OUT_GRADS, off(gbp[i][j]), 1, C,
              vv[i], vv[i], 0, 1, 1LL * pp[i] * C, C, 0);
    }
  }
  m->dtok = ab.take(M * D, AL);
  m->dx0 = ab.take(M * D, AL);
  m->dx1 = ab.take(M * D, AL);
  m->dln = ab.take(M * D, AL);
  m->dqkv = ab.take(M * 3 * D, AL);
  m->dh = ab.take(M * 4 * D, AL);
  m->dattn = ab.take(M * D, AL);
  m->dA0 = ab.take(Pmax * C, AL);
  m->dA1 = ab.take(Pmax * C, AL);
  m->red = af.take(2LL * B, AL);
  m->n_packbf = pb.n, m->n_packf32 = pf.n, m->n_actbf = ab.n, m->n_actf32 = af.n;
  long long o = 0;
  m->o_packbf = o;
  o = align_up(o + m->n_packbf * 2, 1024);
  m->o_packf32 = o;
  o = align_up(o + m->n_packf32 * 4, 1024);
  m->o_actbf = o;
  o = align_up(o + m->n_actbf * 2, 1024);
  m->o_actf32 = o;
  o = align_up(o + m->n_actf32 * 4, 1024);
  m->o_table = o;
  o = align_up(o + static_cast<long long>(m->h_table.size()) * 4, 1024);
  m->o_pack_segs = o;
  o = align_up(o + static_cast<long long>(m->pack_specs.size()) * sizeof(theia_perm_seg), 1024);
  m->o_unpack_segs = o;
  o = align_up(o + static_cast<long long>(m->unpack_specs.size()) * sizeof(theia_perm_seg), 1024);
  m->o_conv_pack = o;
  o = align_up(o + static_cast<long long>(m->conv_pack.size()) * sizeof(theia_conv_perm), 1024);
  m->o_conv_unpack = o;
  o = align_up(o + static_cast<long long>(m->conv_unpack.size()) * sizeof(theia_conv_perm), 1024);
  m->ws_bytes = o;
  *out = m;
  return THEIA_OK;
}

extern "C" void theia_model_destroy(theia_model* m) { delete m; }
extern "C" long long theia_model_param_floats(const theia_model* m) { return m->n_params_total; }
extern "C" long long theia_model_workspace_bytes(const theia_model* m) { return m->ws_bytes; }
extern "C" int theia_model_num_params(const theia_model* m) { return static_cast<int>(m->params.size()); }
extern "C" int theia_model_param_info(const theia_model* m, int i, char* name, int name_cap, long long* dims4,
                                      int* ndim, long long* offset) {
  if (i < 0 || i >= static_cast<int>(m->params.size())) return set_error(THEIA_ERR_ARG, "param index");
  const PInfo& p = m->params[i];
  snprintf(name, name_cap, "%s", p.name.c_str());
  for (int k = 0; k < 4; ++k) dims4[k] = p.dims[k];
  *ndim = p.ndim;
  *offset = p.off;
  return THEIA_OK;
}
// Bring-up / test accessor: device pointer of a named internal activation (layer or head index i).
extern "C" int theia_model_debug_ptr(theia_model* m, const char* name, int i, void** ptr, long long* elems,
                                     int* is_f32) {
  if (!m->ws) return set_error(THEIA_ERR_ARG, "model not bound");
  const std::string n(name);
  const long long M = static_cast<long long>(m->last_B) * m->N, P = static_cast<long long>(m->last_B) * 256;
  const int D = m->D;
  auto AB = [&](long long off) { return static_cast<void*>(reinterpret_cast<bf16*>(m->ws + m->o_actbf) + off); };
  *is_f32 = 0;
  if (n == "patches") { *ptr = AB(m->patches); *elems = M * 768; return 0; }
  if (n == "tokens") { *ptr = AB(m->tokens); *elems = M * D; return 0; }
  if (n == "dtok") { *ptr = AB(m->dtok); *elems = M * D; return 0; }
  if (n == "x") { if (i < 0 || i > m->L) return THEIA_ERR_ARG; *ptr = AB(m->x[i]); *elems = M * D; return 0; }
  if (i >= 0 && i < m->L) {  // bf16 operand copies of the layer's Linear weights (refreshed by pack / adamw)
    const LayerW& w = m->lw[i];
    const long long DD = 1LL * D * D;
    auto PBp = [&](long long off) { return static_cast<void*>(reinterpret_cast<bf16*>(m->ws + m->o_packbf) + off); };
    if (n == "wqkv") { *ptr = PBp(w.wqkv); *elems = 3 * DD; return 0; }
    if (n == "wo") { *ptr = PBp(w.wo); *elems = DD; return 0; }
    if (n == "w1") { *ptr = PBp(w.w1); *elems = 4 * DD; return 0; }
    if (n == "w2") { *ptr = PBp(w.w2); *elems = 4 * DD; return 0; }
  }
  if (i >= 0 && i < m->T && m->hp[i].hw != 1) {
    auto PBp = [&](long long off) { return static_cast<void*>(reinterpret_cast<bf16*>(m->ws + m->o_packbf) + off); };
    if (n == "c1F") { *ptr = PBp(m->hw[i].c1F); *elems = 9LL * D * D; return 0; }
    if (n == "padD") { *ptr = PBp(m->hw[i].padD); *elems = 9LL * D * D; return 0; }
  }
  if (i >= 0 && i < m->L) {
    const LayerA& a = m->la[i];
    if (n == "ln1") { *ptr = AB(a.ln1); *elems = M * D; return 0; }
    if (n == "qkv") { *ptr = AB(a.qkv); *elems = M * 3 * D; return 0; }
    if (n == "attn") { *ptr = AB(a.attn); *elems = M * D; return 0; }
    if (n == "xmid") { *ptr = AB(a.xmid); *elems = M * D; return 0; }
    if (n == "ln2") { *ptr = AB(a.ln2); *elems = M * D; return 0; }
    if (n == "h") { *ptr = AB(a.h); *elems = M * 2 * D; return 0; }  // uint8 codes, M * 4 D bytes
    if (n == "a") { *ptr = AB(a.a); *elems = M * 4 * D; return 0; }
  }
  if (i >= 0 && i < m->T) {
    const HeadA& a = m->ha[i];
    const long long P1 = 1LL * m->last_B * m->hp[i].p1 * m->hp[i].p1, P2 = 1LL * m->last_B * m->hp[i].p2 * m->hp[i].p2;
    if (n == "padout") { *ptr = AB(a.padout); *elems = P * D; return 0; }
    if (n == "hln0") { *ptr = AB(a.ln0); *elems = P * D; return 0; }
    if (n == "c1") { *ptr = AB(a.c1); *elems = P1 * D; return 0; }
    if (n == "hln1") { *ptr = AB(a.ln1); *elems = P1 * D; return 0; }
    if (n == "c2") { *ptr = AB(a.c2); *elems = P2 * D; return 0; }
    if (n == "hln2") { *ptr = AB(a.ln2); *elems = P2 * D; return 0; }
  }
  return set_error(THEIA_ERR_ARG, "unknown activation %s[%d]", name, i);
}

namespace {
// resolve a segment table against the bound buffers and upload it (bind time only: synchronous copy)
int upload_segs(theia_model* m, const std::vector<SegSpec>& specs, long long ws_off, long long* total_blocks) {
  std::vector<theia_perm_seg> h(specs.size());
  long long blocks = 0;
  for (size_t i = 0; i < specs.size(); ++i) {
    const SegSpec& q = specs[i];
    theia_perm_seg& g = h[i];
    memset(&g, 0, sizeof(g));
    g.in = q.in_base == SEG_IN_MASTER ? static_cast<const void*>(m->master + q.in_off)
                                      : static_cast<const void*>(reinterpret_cast<float*>(m->ws + m->o_actf32) + q.in_off);
    if (q.out_base == SEG_OUT_PACKBF) g.out = reinterpret_cast<bf16*>(m->ws + m->o_packbf) + q.out_off;
    else if (q.out_base == SEG_OUT_PACKF32) g.out = reinterpret_cast<float*>(m->ws + m->o_packf32) + q.out_off;
    else g.out = reinterpret_cast<void*>(static_cast<uintptr_t>(q.out_off) * 4);  // relative to the gradient buffer
    g.out_f32 = q.out_base != SEG_OUT_PACKBF;
    g.n0 = q.n0, g.n1 = q.n1, g.n2 = q.n2, g.n3 = q.n3;
    g.s0 = q.s0, g.s1 = q.s1, g.s2 = q.s2, g.s3 = q.s3, g.base = q.base;
    g.lim1 = q.lim1, g.lim2 = q.lim2;
    g.first_block = blocks;
    blocks += (1LL * q.n0 * q.n1 * q.n2 * q.n3 + 255) / 256;
  }
  *total_blocks = blocks;
  if (h.empty()) return THEIA_OK;
  cudaError_t e = cudaMemcpy(m->ws + ws_off, h.data(), h.size() * sizeof(theia_perm_seg), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return theia::set_error(THEIA_ERR_CUDA, "segment table upload: %s", cudaGetErrorString(e));
  return THEIA_OK;
}
}  // namespace

// grads may be NULL (inference only) or re-bound before each backward (theia_model_set_grads)
extern "C" int theia_model_bind(theia_model* m, float* master, float* grads, void* workspace) {
  if (!master || !workspace) return set_error(THEIA_ERR_ARG, "bind: null");
  m->master = master;
  m->grads = grads;
  m->ws = static_cast<uint8_t*>(workspace);
  cudaError_t e = cudaMemcpy(m->ws + m->o_table, m->h_table.data(), m->h_table.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "pack table upload: %s", cudaGetErrorString(e));
  int rc = upload_segs(m, m->pack_specs, m->o_pack_segs, &m->pack_blocks);
  if (rc) return rc;
  rc = upload_segs(m, m->unpack_specs, m->o_unpack_segs, &m->unpack_blocks);
  if (rc) return rc;
  // conv-weight tables: pack (master -> two bf16 packs), unpack (tap-major fp32 scratch -> gradient buffer offset)
  std::vector<theia_conv_perm> hp(m->conv_pack.size()), hu(m->conv_unpack.size());
  bf16* pbf = reinterpret_cast<bf16*>(m->ws + m->o_packbf);
  float* af32 = reinterpret_cast<float*>(m->ws + m->o_actf32);
  for (size_t i = 0; i < hp.size(); ++i) {
    const ConvSpec& q = m->conv_pack[i];
    hp[i] = theia_conv_perm{m->master + q.w_off, pbf + q.pack0, pbf + q.pack1, q.flags0, q.flags1, m->D};
  }
  for (size_t i = 0; i < hu.size(); ++i) {
    const ConvSpec& q = m->conv_unpack[i];
    hu[i] = theia_conv_perm{reinterpret_cast<const void*>(static_cast<uintptr_t>(q.w_off) * 4), af32 + q.pack0, nullptr,
                            q.flags0, 0, m->D};
  }
  if (!hp.empty()) {
    e = cudaMemcpy(m->ws + m->o_conv_pack, hp.data(), hp.size() * sizeof(theia_conv_perm), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
      e = cudaMemcpy(m->ws + m->o_conv_unpack, hu.data(), hu.size() * sizeof(theia_conv_perm), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "conv table upload: %s", cudaGetErrorString(e));
  }
  return THEIA_OK;
}

// Switch the gradient buffer (same layout) without touching anything else: lets the caller hand out the buffer of
// the previous backward (autograd keeps views of it as .grad) while the next backward writes another one.
extern "C" int theia_model_set_grads(theia_model* m, float* grads) {
  if (!m->master || !m->ws) return set_error(THEIA_ERR_ARG, "model not bound");
  if (!grads) return set_error(THEIA_ERR_ARG, "set_grads: null");
  m->grads = grads;
  return THEIA_OK;
}

extern "C" int theia_model_set_input_size(theia_model* m, int height, int width) {
  if (!m || height < 1 || width < 1 || height > 8192 || width > 8192)
    return set_error(THEIA_ERR_ARG, "theia_model_set_input_size: %d x %d", height, width);
  m->in_h = height, m->in_w = width;
  return THEIA_OK;
}

extern "C" int theia_model_set_input_dtype(theia_model* m, int is_f32) {
  if (!m) return set_error(THEIA_ERR_ARG, "theia_model_set_input_dtype: null model");
  m->in_f32 = is_f32 ? 1 : 0;
  return THEIA_OK;
}

extern "C" int theia_model_pack_table(theia_model* m, const int** table, void** packbf) {
  if (!m->master || !m->ws) return set_error(THEIA_ERR_ARG, "model not bound");
  *table = reinterpret_cast<const int*>(m->ws + m->o_table);
  *packbf = m->ws + m->o_packbf;
  return THEIA_OK;
}

namespace theia {
// Per-token additive table of the embedding stage (hf:modeling_vit.py:116-126; backbones.py:70-93,186-217):
//   patch token t: position embedding of its slot;  CLS: cls_token + pos[0];  register r: reg_token[r] + reg_pos[r]
__global__ void token_table_kernel(float* __restrict__ tab, const float* __restrict__ pos, const float* __restrict__ cls,
                                   const float* __restrict__ regt, const float* __restrict__ regp, int N, int D, int p0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int t = i / D, d = i - t * D;
  float v;
  if (t < p0) v = cls[d] + pos[d];
  else if (t < p0 + 196) v = pos[(t - p0 + 1) * D + d];
  else v = regt[(t - p0 - 196) * D + d] + regp[(t - p0 - 196) * D + d];
  tab[i] = v;
}
// gradient of the table -> gradients of position_embeddings / cls_token / reg_token / reg_pos_embed
__global__ void token_table_grad_kernel(const float* __restrict__ tg, float* __restrict__ gpos, float* __restrict__ gcls,
                                        float* __restrict__ gregt, float* __restrict__ gregp, int N, int D, int p0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int t = i / D, d = i - t * D;
  const float v = tg[i];
  if (t < p0) {
    gcls[d] = v;
    gpos[d] = v;
  } else if (t < p0 + 196) {
    gpos[(t - p0 + 1) * D + d] = v;
  } else {
    gregt[(t - p0 - 196) * D + d] = v;
    gregp[(t - p0 - 196) * D + d] = v;
  }
}
}  // namespace theia

namespace {

struct Ctx {
  theia_model* m;
  cudaStream_t s;
  bf16* PB(long long off) const { return reinterpret_cast<bf16*>(m->ws + m->o_packbf) + off; }
  float* PF(long long off) const { return reinterpret_cast<float*>(m->ws + m->o_packf32) + off; }
  bf16* AB(long long off) const { return reinterpret_cast<bf16*>(m->ws + m->o_actbf) + off; }
  float* AF(long long off) const { return reinterpret_cast<float*>(m->ws + m->o_actf32) + off; }
  float* W(int pi) const { return m->master + m->params[pi].off; }
  float* G(int pi) const { return m->grads + m->params[pi].off; }
};

#define TRY(x)            \
  do {                    \
    int rc__ = (x);       \
    if (rc__) return rc__; \
  } while (0)

theia_gemm_desc gemm_base(int M, int N, int K) {
  theia_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.M = M, d.N = N, d.K = K;
  d.a_mode = THEIA_OP_K2D;
  d.b_mode = THEIA_OP_K2D;
  d.splits = 1;
  d.batch_z = 1;
  return d;
}

// y[M,N] = x[M,K] * w[N,K]^T (+bias) with epilogue flags
int linear(const Ctx& c, const bf16* x, const bf16* w, const float* bias, void* out, int M, int N, int K, int epi,
           const void* aux = nullptr, void* out2 = nullptr, float* colsum = nullptr) {
  theia_gemm_desc d = gemm_base(M, N, K);
  d.A = x, d.lda = K, d.B = w, d.ldb = K;
  d.out = out, d.ldo = N, d.bias = bias, d.epi = epi, d.aux = aux, d.out2 = out2, d.colsum = colsum;
  return theia_gemm(&d, c.s);
}

// dx[M,N] = dy[M,K] * w[K,N]  with w = the forward weight [out = K][in = N] read in place as an MN-major B operand
int linear_dgrad(const Ctx& c, const bf16* dy, const bf16* w, void* out, int M, int N, int K, int epi,
                 const void* aux = nullptr, float* colsum = nullptr) {
  theia_gemm_desc d = gemm_base(M, N, K);
  d.A = dy, d.lda = K, d.B = w, d.ldb = N, d.b_mode = THEIA_OP_MN2D;
  d.out = out, d.ldo = N, d.epi = epi, d.aux = aux, d.colsum = colsum;
  return theia_gemm(&d, c.s);
}

// split-K factor for a wgrad GEMM: fill whole waves of the persistent grid (units*splits close to a multiple of
// the resident scheduling units) with at least 8 K-blocks per split; fewer splits win ties (each split adds a
// tile of atomics).  A scheduling unit is one CTA tile, or one 256-row CTA-pair tile when the launcher pairs.
int pick_splits(int m_tiles, int nz_tiles, int bn, int num_kb) {
  const int pair = gemm_pair_mode(bn, m_tiles, THEIA_OP_MN2D);  // wgrads: A = dY, MN-major
  const int tiles = ((m_tiles + pair - 1) / pair) * nz_tiles;
  const int sms = num_sms() / pair;
  const int maxs = num_kb / 8 > 0 ? num_kb / 8 : 1;
  // wave efficiency at which the search stops adding splits (THEIA_SPLITK_EFF overrides, for tuning runs)
  static const double stop_eff = getenv("THEIA_SPLITK_EFF") ? atof(getenv("THEIA_SPLITK_EFF")) : 0.9;
  int best = 1;
  double best_eff = 0.0;
  const int smax = maxs < 256 ? maxs : 256;  // the last candidate is always eligible: tiny outputs (2 tiles) still split
  for (int s = 1; s <= smax; ++s) {
    const long long items = 1LL * tiles * s;
    const long long waves = (items + sms - 1) / sms;
    const double eff = static_cast<double>(items) / (waves * sms);
    if (items * 10 < sms * 9LL && s < smax) continue;  // do not leave >10 % of the SMs idle when more splits are possible
    if (eff > best_eff + 0.02) best_eff = eff, best = s;
    if (items >= 4LL * sms && eff > stop_eff) break;
  }
  return best;
}

// launch plan of a wgrad GEMM (host logic only): N tile, CTAs per MMA, split-K factor; z = independent slices (taps)
void plan_wgrad(int Nout, int Kin, int Mtok, int z, int* bn_out, int* pair_out, int* splits_out) {
  const int bn = (Kin % 256 == 0) ? 256 : (Kin % 192 == 0 ? 192 : (Kin > 192 ? 256 : (Kin > 128 ? 192 : 128)));
  const int m_tiles = (Nout + 127) / 128;
  *bn_out = bn;
  *pair_out = gemm_pair_mode(bn, m_tiles, THEIA_OP_MN2D);
  *splits_out = pick_splits(m_tiles, z * ((Kin + bn - 1) / bn), bn, (Mtok + 63) / 64);
}

// dW[Nout,Kin] += dY[Mtok,Nout]^T * X[Mtok,Kin]   (fp32 atomics, split-K over tokens)
int wgrad(const Ctx& c, const bf16* dy, const bf16* x, float* dw, int Mtok, int Nout, int Kin) {
  theia_gemm_desc d = gemm_base(Nout, Kin, Mtok);
  d.a_mode = THEIA_OP_MN2D, d.b_mode = THEIA_OP_MN2D;
  d.A = dy, d.lda = Nout, d.B = x, d.ldb = Kin;
  d.out = dw, d.ldo = Kin, d.epi = THEIA_EPI_ATOMIC;
  int pair = 1;
  plan_wgrad(Nout, Kin, Mtok, 1, &d.bn, &pair, &d.splits);
  return theia_gemm(&d, c.s);
}

void conv_geom_16(theia_conv_geom& g, int C, int Hin, int B, long long sw, long long sh, long long sb, int shift0) {
  memset(&g, 0, sizeof(g));
  g.C = C, g.H = Hin, g.W = Hin, g.B = B;
  g.stride_w = sw, g.stride_h = sh, g.stride_b = sb;
  g.ntaps = 9;
  for (int t = 0; t < 9; ++t) g.dh[t] = t / 3 + shift0, g.dw[t] = t % 3 + shift0;
  g.tile_w = 16, g.tile_h = 8;
  g.out_h = 16, g.out_w = 16, g.out_img_rows = 256, g.out_row_off = 0, g.out_wpitch = 16;
  g.sy = g.sx = 1, g.py = g.px = 0;
}

// ---- stride-2 ConvTranspose2d(k3) of the 64x64 heads ------------------------------------------------
// Forward: the outputs of parity (py,px) form a dense stride-1 gather with 1/2/2/4 taps -> 4 GEMM launches
// over sub-grids, no zero stuffing.  in: [B, vin x vin (pitch pin), C]; out: [B, vout x vout (pitch pout), C].
// Weights: ONE tap-major pack [9][Cout][Cin]; each launch names the taps it uses (wtap).
int convT2x_fwd(const Ctx& c, const bf16* x, int vin, int pin, const bf16* w9, const float* bias, bf16* out, int vout,
                int pout, int pad, int C, int B, float* stats) {
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      theia_conv_geom g;
      memset(&g, 0, sizeof(g));
      g.C = C, g.H = vin, g.W = vin, g.B = B;
      g.stride_w = C, g.stride_h = 1LL * pin * C, g.stride_b = 1LL * pin * pin * C;
      int nt = 0;
      for (int kh = 0; kh < 3; ++kh) {
        if ((py + pad - kh) % 2 != 0) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if ((px + pad - kw) % 2 != 0) continue;
          g.dh[nt] = (py + pad - kh) / 2, g.dw[nt] = (px + pad - kw) / 2, g.wtap[nt] = kh * 3 + kw;
          ++nt;
        }
      }
      g.ntaps = nt;
      const int sub_h = (vout - 1 - py) / 2 + 1, sub_w = (vout - 1 - px) / 2 + 1;
      g.tile_w = sub_w <= 16 ? 16 : 32, g.tile_h = 128 / g.tile_w;
      g.out_h = sub_h, g.out_w = sub_w;
      g.out_img_rows = pout * pout, g.out_row_off = 0, g.out_wpitch = pout;
      g.sy = g.sx = 2, g.py = py, g.px = px;
      g.in_stride = 1, g.b_tap_rows = C;
      const int tiles = (sub_h + g.tile_h - 1) / g.tile_h;
      theia_gemm_desc d = gemm_base(B * tiles * 128, C, nt * C);
      d.a_mode = THEIA_OP_CONV_K;
      d.A = x, d.B = w9;
      d.conv = g;
      d.out = out, d.ldo = C, d.bias = bias, d.epi = THEIA_EPI_RELU | THEIA_EPI_STATS, d.stats = stats;
      TRY(theia_gemm(&d, c.s));
    }
  return THEIA_OK;
}

// geometry shared by the dgrad / wgrad of a stride-2 transposed conv: gather dY[2*i - pad + k] with TMA element
// stride 2 over the INPUT grid (vin x vin at pitch pin); dY is [B, vout x vout (pitch pout), C].
void convT2x_bwd_geom(theia_conv_geom& g, int C, int B, int vin, int pin, int vout, int pout, int pad) {
  memset(&g, 0, sizeof(g));
  g.C = C, g.H = pout > vout ? pout : vout, g.W = g.H, g.B = B;  // padded rows/cols of dY are zero
  g.stride_w = C, g.stride_h = 1LL * pout * C, g.stride_b = 1LL * pout * pout * C;
  g.ntaps = 9;
  for (int t = 0; t < 9; ++t) g.dh[t] = t / 3 - pad, g.dw[t] = t % 3 - pad, g.wtap[t] = t;
  g.tile_w = pin <= 16 ? 16 : 32, g.tile_h = 128 / g.tile_w;
  g.out_h = vin, g.out_w = vin;
  g.out_img_rows = pin * pin, g.out_row_off = 0, g.out_wpitch = pin;
  g.sy = g.sx = 1;
  g.in_stride = 2, g.b_tap_rows = C;
}

// dX[B, vin x vin (pitch pin), C] = sum_k dY[2 i - pad + k] * Wt[ci,co,k]; weights tap-major [9][Cin][Cout]
int convT2x_dgrad(const Ctx& c, const bf16* dy, const bf16* w9d, bf16* dx, int C, int B, int vin, int pin, int vout,
                  int pout, int pad) {
  theia_conv_geom g;
  convT2x_bwd_geom(g, C, B, vin, pin, vout, pout, pad);
  const int tiles = (vin + g.tile_h - 1) / g.tile_h;
  theia_gemm_desc d = gemm_base(B * tiles * 128, C, 9 * C);
  d.a_mode = THEIA_OP_CONV_K;
  d.A = dy, d.B = w9d;
  d.conv = g;
  d.out = dx, d.ldo = C;
  return theia_gemm(&d, c.s);
}

// ws[tap][Cin][Cout] += X[pix, Cin]^T * dY[2 pix - pad + tap, Cout]   (X at pitch pin, padding rows are zero)
int convT2x_wgrad(const Ctx& c, const bf16* x, const bf16* dy, float* wsout, int C, int B, int vin, int pin, int vout,
                  int pout, int pad) {
  theia_conv_geom g;
  convT2x_bwd_geom(g, C, B, vin, pin, vout, pout, pad);
  g.tile_h = 64 / g.tile_w;
  g.out_h = pin;  // K-blocks run over the pitched input grid
  const int Ppix = B * pin * pin;
  theia_gemm_desc d = gemm_base(C, C, Ppix);
  d.a_mode = THEIA_OP_MN2D, d.b_mode = THEIA_OP_CONV_MN;
  d.A = x, d.lda = C, d.B = dy;
  d.conv = g;
  d.out = wsout, d.ldo = C, d.epi = THEIA_EPI_ATOMIC;
  d.batch_z = 9, d.out_z_stride = 1LL * C * C;
  const int bn = (C % 256 == 0) ? 256 : (C % 192 == 0 ? 192 : 128);
  d.bn = bn;
  d.splits = pick_splits((C + 127) / 128, 9 * ((C + bn - 1) / bn), bn, Ppix / 64);
  return theia_gemm(&d, c.s);
}

// stride-1 3x3 conv as implicit GEMM over an NHWC tensor; weights packed [Cout][tap][Cin]
int conv3x3(const Ctx& c, const bf16* x, const theia_conv_geom& g, const bf16* w, const float* bias, void* out,
            long long ldo, int Cout, int epi, float* stats, const void* aux) {
  const int tiles_per_img = (g.out_h + g.tile_h - 1) / g.tile_h;
  theia_gemm_desc d = gemm_base(g.B * tiles_per_img * 128, Cout, 9 * g.C);
  d.a_mode = THEIA_OP_CONV_K;
  d.A = x, d.B = w, d.ldb = 9LL * g.C;
  d.conv = g;
  d.out = out, d.ldo = ldo, d.bias = bias, d.epi = epi, d.stats = stats, d.aux = aux;
  return theia_gemm(&d, c.s);
}

// conv wgrad: ws[tap][Cout][Cin] += dY[pix,Cout]^T * X[pix + tap shift, Cin]
int conv_wgrad(const Ctx& c, const bf16* dy, const bf16* x, const theia_conv_geom& g, float* wsout, int Cout) {
  const int P = g.B * 256;
  theia_gemm_desc d = gemm_base(Cout, g.C, P);
  d.a_mode = THEIA_OP_MN2D, d.b_mode = THEIA_OP_CONV_MN;
  d.A = dy, d.lda = Cout, d.B = x;
  d.conv = g;
  d.out = wsout, d.ldo = g.C, d.epi = THEIA_EPI_ATOMIC;
  d.batch_z = 9, d.out_z_stride = static_cast<long long>(Cout) * g.C;
  const int bn = (g.C % 256 == 0) ? 256 : (g.C % 192 == 0 ? 192 : 128);
  d.bn = bn;
  d.splits = pick_splits((Cout + 127) / 128, 9 * ((g.C + bn - 1) / bn), bn, P / 64);
  return theia_gemm(&d, c.s);
}

int zero_f32(const Ctx& c, float* p, long long n) {
  cudaError_t e = cudaMemsetAsync(p, 0, n * sizeof(float), c.s);
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace

// fp32 master -> bf16 / permuted operand copies: 3 launches (Linear-weight casts through the pack table -- skipped
// when theia_adamw_flat already refreshed them --, the token table, one segmented gather for the conv-weight packs
// and the LayerNorm[C,H,W] affines).  Call after every optimizer step.
extern "C" int theia_model_pack(theia_model* m, int skip_linear_cast, void* stream) {
  if (!m->master || !m->ws) return set_error(THEIA_ERR_ARG, "model not bound");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  const int D = m->D;
  if (!skip_linear_cast)
    TRY(theia_pack_cast(m->master, reinterpret_cast<const int*>(m->ws + m->o_table), m->ws + m->o_packbf,
                        m->n_params_total, c.s));
  token_table_kernel<<<(m->N * D + 255) / 256, 256, 0, c.s>>>(c.PF(m->tokt), c.W(m->pos), m->cls >= 0 ? c.W(m->cls) : nullptr,
                                                             m->regt >= 0 ? c.W(m->regt) : nullptr,
                                                             m->regp >= 0 ? c.W(m->regp) : nullptr, m->N, D, m->p0);
  THEIA_CHECK_LAUNCH("token_table");
  TRY(theia_conv_pack(reinterpret_cast<const theia_conv_perm*>(m->ws + m->o_conv_pack),
                      static_cast<int>(m->conv_pack.size()), D, c.s));
  TRY(theia_perm_segments(reinterpret_cast<const theia_perm_seg*>(m->ws + m->o_pack_segs),
                          static_cast<int>(m->pack_specs.size()), m->pack_blocks, nullptr, c.s));
  return THEIA_OK;
}

// images uint8 [B,224,224,3] (or [B,3,224,224]) -> tokens (bf16, workspace) and, when
// run_heads, preds[t] fp32 [B,256,C_t] (caller-owned).  tokens_f32 (optional) receives the
// final-LayerNorm output [B,197,D] in fp32 for forward_feature().
extern "C" int theia_model_forward(theia_model* m, const uint8_t* images, int B, int channels_first, int do_resize,
                                   int do_rescale, int do_normalize, const float* mean3, const float* std3, int run_heads,
                                   float* const* preds, void* tokens_bf16_out, void* stream) {
  if (!m->master || !m->ws) return set_error(THEIA_ERR_ARG, "model not bound");
  if (B < 1 || B > m->Bmax) return set_error(THEIA_ERR_ARG, "batch %d outside [1,%d]", B, m->Bmax);
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  const int D = m->D, C = m->D, H = m->H, L = m->L;
  const int NT = m->N;
  const int M = B * NT, P = B * 256;
  m->last_B = B;
  if (m->in_f32) {
    if (do_resize) return set_error(THEIA_ERR_UNSUPPORTED, "float images are not resized on this path (pass do_resize = 0)");
    TRY(theia_preprocess_f32(reinterpret_cast<const float*>(images), m->in_h, m->in_w, c.AB(m->patches), B, channels_first,
                             do_rescale, do_normalize, mean3, std3, NT, m->p0, c.s));
  } else {
    TRY(theia_preprocess_hw(images, m->in_h, m->in_w, c.AB(m->patches), B, channels_first, do_resize, do_rescale, do_normalize,
                            mean3, std3, NT, m->p0, c.s));
  }
  {  // patch embedding + CLS + position embeddings (hf:modeling_vit.py:100-128,153-168)
    theia_gemm_desc d = gemm_base(M, D, 768);
    d.A = c.AB(m->patches), d.lda = 768, d.B = c.PB(m->wpe), d.ldb = 768;
    d.out = c.AB(m->x[0]), d.ldo = D, d.bias = c.W(m->peb);
    d.epi = THEIA_EPI_POSCLS, d.pos = c.PF(m->tokt), d.tokens = NT, d.tok_p0 = m->p0, d.tok_p1 = m->p0 + 196;
    TRY(theia_gemm(&d, c.s));
  }
  for (int l = 0; l < L; ++l) {  // hf:modeling_vit.py:328-346
    const LayerP& p = m->lp[l];
    const LayerW& w = m->lw[l];
    const LayerA& a = m->la[l];
    TRY(theia_layernorm_fwd(c.AB(m->x[l]), c.W(p.ln1w), c.W(p.ln1b), c.AB(a.ln1), c.AF(a.mean1), c.AF(a.rstd1), M, D,
                            m->cfg.ln_eps, c.s));
    TRY(linear(c, c.AB(a.ln1), c.PB(w.wqkv), c.W(p.qb), c.AB(a.qkv), M, 3 * D, D, 0));
    TRY(theia_attention_tc_fwd(c.AB(a.qkv), c.AB(a.attn), c.AF(a.lse), B, NT, H, c.s));
    TRY(linear(c, c.AB(a.attn), c.PB(w.wo), c.W(p.ob), c.AB(a.xmid), M, D, D, THEIA_EPI_RESID, c.AB(m->x[l])));
    TRY(theia_layernorm_fwd(c.AB(a.xmid), c.W(p.ln2w), c.W(p.ln2b), c.AB(a.ln2), c.AF(a.mean2), c.AF(a.rstd2), M, D,
                            m->cfg.ln_eps, c.s));
    TRY(linear(c, c.AB(a.ln2), c.PB(w.w1), c.W(p.f1b), c.AB(a.a), M, 4 * D, D, THEIA_EPI_GELU | THEIA_EPI_AUX_U8, nullptr, c.AB(a.h)));
    TRY(linear(c, c.AB(a.a), c.PB(w.w2), c.W(p.f2b), c.AB(m->x[l + 1]), M, D, 4 * D, THEIA_EPI_RESID, c.AB(a.xmid)));
  }
  TRY(theia_layernorm_fwd(c.AB(m->x[L]), c.W(m->lnfw), c.W(m->lnfb), c.AB(m->tokens), c.AF(m->meanf), c.AF(m->rstdf), M,
                          D, m->cfg.ln_eps, c.s));
  if (tokens_bf16_out) {
    cudaError_t e = cudaMemcpyAsync(tokens_bf16_out, c.AB(m->tokens), sizeof(bf16) * M * D, cudaMemcpyDeviceToDevice, c.s);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memcpy: %s", cudaGetErrorString(e));
  }
  if (!run_heads) return THEIA_OK;
  for (int t = 0; t < m->T; ++t) {  // adapter_heads.py:352-359
    if (!preds || !preds[t]) continue;
    const HeadP& p = m->hp[t];
    const HeadW& w = m->hw[t];
    const HeadA& a = m->ha[t];
    if (p.hw == 1) {  // pred[B, C_t] = tokens[:, 0] W^T + b : rows of A are the CLS rows (pitch 197*D)
      theia_gemm_desc d = gemm_base(B, p.ct, C);
      d.A = c.AB(m->tokens), d.lda = 1LL * NT * D, d.B = c.PB(w.l), d.ldb = C;
      d.out = preds[t], d.ldo = p.ct, d.bias = c.W(p.lb), d.epi = THEIA_EPI_OUT_F32;
      TRY(theia_gemm(&d, c.s));
      continue;
    }
    float* st0 = c.AF(a.stats);
    float* st1 = st0 + 2 * B;
    float* st2 = st1 + 2 * B;
    TRY(zero_f32(c, st0, 6LL * B));
    theia_conv_geom g;
    // pad: ConvTranspose2d(3x3, s1) 14 -> 16 over the spatial tokens (CLS skipped by the base offset)
    conv_geom_16(g, C, 14, B, D, 14LL * D, 1LL * NT * D, -2);
    TRY(conv3x3(c, c.AB(m->tokens) + 1LL * m->p0 * D, g, c.PB(w.padF), c.W(p.padb), c.AB(a.padout), C, C, THEIA_EPI_STATS, st0, nullptr));
    TRY(theia_ln3d_apply(c.AB(a.padout), st0, c.PF(w.gb[0][0]), c.PF(w.gb[0][1]), c.AB(a.ln0), B, 256 * C, 1e-5f, C, 0,
                         0, c.s));
    if (p.hw == 16) {
      conv_geom_16(g, C, 16, B, C, 16LL * C, 256LL * C, -1);
      TRY(conv3x3(c, c.AB(a.ln0), g, c.PB(w.c1F), c.W(p.c1b), c.AB(a.c1), C, C, THEIA_EPI_RELU | THEIA_EPI_STATS, st1, nullptr));
      TRY(theia_ln3d_apply(c.AB(a.c1), st1, c.PF(w.gb[1][0]), c.PF(w.gb[1][1]), c.AB(a.ln1), B, 256 * C, 1e-5f, C, 0, 0, c.s));
      TRY(conv3x3(c, c.AB(a.ln1), g, c.PB(w.c2F), c.W(p.c2b), c.AB(a.c2), C, C, THEIA_EPI_RELU | THEIA_EPI_STATS, st2, nullptr));
      TRY(theia_ln3d_apply(c.AB(a.c2), st2, c.PF(w.gb[2][0]), c.PF(w.gb[2][1]), c.AB(a.ln2), B, 256 * C, 1e-5f, C, 0, 0, c.s));
      TRY(linear(c, c.AB(a.ln2), c.PB(w.l), c.W(p.lb), preds[t], P, p.ct, C, THEIA_EPI_OUT_F32));
    } else {
      // adapter_heads.py:304-315: ConvT(s2,p1) 16->31, ReLU, LN[C,31,31], ConvT(s2,op1) 31->64, ReLU, LN[C,64,64], Linear
      TRY(convT2x_fwd(c, c.AB(a.ln0), 16, 16, c.PB(w.c1F), c.W(p.c1b), c.AB(a.c1), 31, 32, 1, C, B, st1));
      TRY(theia_ln3d_apply(c.AB(a.c1), st1, c.PF(w.gb[1][0]), c.PF(w.gb[1][1]), c.AB(a.ln1), B, 1024 * C, 1e-5f, C, 32, 31, c.s));
      TRY(convT2x_fwd(c, c.AB(a.ln1), 31, 32, c.PB(w.c2F), c.W(p.c2b), c.AB(a.c2), 64, 64, 0, C, B, st2));
      TRY(theia_ln3d_apply(c.AB(a.c2), st2, c.PF(w.gb[2][0]), c.PF(w.gb[2][1]), c.AB(a.ln2), B, 4096 * C, 1e-5f, C, 0, 0, c.s));
      TRY(linear(c, c.AB(a.ln2), c.PB(w.l), c.W(p.lb), preds[t], B * 4096, p.ct, C, THEIA_EPI_OUT_F32));
    }
  }
  return THEIA_OK;
}

// Backward of theia_model_forward for the batch of the last forward call.
// dpreds[t]: bf16 [B,256,C_t] gradient w.r.t. preds[t] (NULL = head not used this step).
// Writes every parameter gradient into the bound flat `grads` buffer (overwrites).
extern "C" int theia_model_backward(theia_model* m, const void* const* dpreds, void* stream) {
  if (!m->master || !m->ws || !m->grads) return set_error(THEIA_ERR_ARG, "model not bound (grads)");
  Ctx c{m, static_cast<cudaStream_t>(stream)};
  const int B = m->last_B;
  if (B < 1) return set_error(THEIA_ERR_ARG, "backward before forward");
  const int D = m->D, C = m->D, H = m->H, L = m->L;
  const int NT = m->N;
  const int M = B * NT, P = B * 256;
  TRY(zero_f32(c, m->grads, m->n_params_total));
  cudaError_t e = cudaMemsetAsync(c.AB(m->dtok), 0, sizeof(bf16) * M * D, c.s);
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  bf16* dA0 = c.AB(m->dA0);
  bf16* dA1 = c.AB(m->dA1);
  float* red = c.AF(m->red);
  bool any_head = false;
  for (int t = 0; t < m->T; ++t) {
    if (!dpreds || !dpreds[t]) continue;
    const HeadP& p = m->hp[t];
    const HeadW& w = m->hw[t];
    const HeadA& a = m->ha[t];
    const bf16* dp = static_cast<const bf16*>(dpreds[t]);
    if (p.hw == 1) {
      {  // dW[C_t, C] = dp^T cls_tokens
        theia_gemm_desc d = gemm_base(p.ct, C, B);
        d.a_mode = THEIA_OP_MN2D, d.b_mode = THEIA_OP_MN2D;
        d.A = dp, d.lda = p.ct, d.B = c.AB(m->tokens), d.ldb = 1LL * NT * D;
        d.out = c.G(p.lw), d.ldo = C, d.epi = THEIA_EPI_ATOMIC;
        TRY(theia_gemm(&d, c.s));
      }
      TRY(theia_colsum(dp, c.G(p.lb), B, p.ct, p.ct, 0, c.s));
      {  // d tokens[:, 0] += dp W
        theia_gemm_desc d = gemm_base(B, C, p.ct);
        d.A = dp, d.lda = p.ct, d.B = c.PB(w.l), d.ldb = C, d.b_mode = THEIA_OP_MN2D;
        d.out = c.AB(m->dtok), d.ldo = 1LL * NT * D, d.epi = THEIA_EPI_RESID, d.aux = c.AB(m->dtok);
        TRY(theia_gemm(&d, c.s));
      }
      continue;
    }
    if (!any_head) {  // conv weight-gradient / LN-affine-gradient scratch of ALL heads: one memset
      TRY(zero_f32(c, c.AF(m->bscr), m->n_bscr));
      any_head = true;
    }
    float* st0 = c.AF(a.stats);
    float* st1 = st0 + 2 * B;
    float* st2 = st1 + 2 * B;
    const long long P2 = static_cast<long long>(B) * p.p2 * p.p2;  // rows of the last stage
    // Linear(C -> C_t)
    TRY(wgrad(c, dp, c.AB(a.ln2), c.G(p.lw), static_cast<int>(P2), p.ct, C));
    TRY(theia_colsum(dp, c.G(p.lb), static_cast<int>(P2), p.ct, p.ct, 0, c.s));
    TRY(linear_dgrad(c, dp, c.PB(w.l), dA0, static_cast<int>(P2), C, p.ct, 0));
    theia_conv_geom g;
    const bf16* lnin[3] = {c.AB(a.padout), c.AB(a.c1), c.AB(a.c2)};
    float* sts[3] = {st0, st1, st2};
    const bf16* convin[3] = {nullptr, c.AB(a.ln0), c.AB(a.ln1)};
    const int convb[3] = {p.padb, p.c1b, p.c2b};
    const long long convD[3] = {w.padD, w.c1D, w.c2D};
    const int vv[3] = {16, p.v1, p.v2}, pp[3] = {16, p.p1, p.p2};
    for (int i = 2; i >= 0; --i) {
      // LayerNorm([C,H,W]) backward (+ ReLU mask of the conv that produced its input)
      const long long npix = 1LL * pp[i] * pp[i];
      const int rows = static_cast<int>(B * npix);
      float* gbs = c.AF(m->bscr + a.gbs[i]);  // dgamma | dbeta in NHWC; unpacked to [C,H,W] at the end
      float* wsc = c.AF(m->bscr + a.wsc[i]);  // conv weight gradient, tap-major; unpacked at the end
      TRY(theia_ln3d_bwd(dA0, lnin[i], sts[i], c.PF(w.gb[i][0]), red, dA1, gbs, gbs + npix * C, B,
                         static_cast<int>(npix * C), 1e-5f, i > 0 ? 1 : 0, C, pp[i] != vv[i] ? pp[i] : 0, vv[i], c.s));
      // conv i backward: dA1 = gradient of its (pre-activation) output
      TRY(theia_colsum(dA1, c.G(convb[i]), rows, C, C, 0, c.s));
      if (i > 0 && p.hw == 16) {
        conv_geom_16(g, C, 16, B, C, 16LL * C, 256LL * C, -1);
        TRY(conv_wgrad(c, dA1, convin[i], g, wsc, C));
        TRY(conv3x3(c, dA1, g, c.PB(convD[i]), nullptr, dA0, C, C, 0, nullptr, nullptr));
      } else if (i > 0) {
        // stride-2 ConvTranspose2d: input map vv[i-1] (pitch pp[i-1]) -> output map vv[i] (pitch pp[i])
        const int pad = i == 1 ? 1 : 0;
        TRY(convT2x_wgrad(c, convin[i], dA1, wsc, C, B, vv[i - 1], pp[i - 1], vv[i], pp[i], pad));
        TRY(convT2x_dgrad(c, dA1, c.PB(convD[i]), dA0, C, B, vv[i - 1], pp[i - 1], vv[i], pp[i], pad));
      } else {
        conv_geom_16(g, C, 14, B, D, 14LL * D, 1LL * NT * D, -2);
        TRY(conv_wgrad(c, dA1, c.AB(m->tokens) + 1LL * m->p0 * D, g, wsc, C));
        // dgrad onto the 14x14 token grid, accumulated over heads
        conv_geom_16(g, C, 16, B, C, 16LL * C, 256LL * C, 0);
        g.out_h = 14, g.out_w = 14, g.out_img_rows = NT, g.out_row_off = m->p0, g.out_wpitch = 14;
        TRY(conv3x3(c, dA1, g, c.PB(convD[i]), nullptr, c.AB(m->dtok), D, C, THEIA_EPI_RESID, nullptr, c.AB(m->dtok)));
      }
    }
  }
  // tap-major conv weight gradients and NHWC LayerNorm-affine gradients of all heads -> reference layouts: one launch
  if (any_head) {
    TRY(theia_conv_unpack(reinterpret_cast<const theia_conv_perm*>(m->ws + m->o_conv_unpack),
                          static_cast<int>(m->conv_unpack.size()), C, m->grads, c.s));
    TRY(theia_perm_segments(reinterpret_cast<const theia_perm_seg*>(m->ws + m->o_unpack_segs),
                            static_cast<int>(m->unpack_specs.size()), m->unpack_blocks, m->grads, c.s));
  }
  // final LayerNorm
  bf16* dx = c.AB(m->dx0);
  bf16* dx2 = c.AB(m->dx1);
  // column sums of each produced dx are the bias gradients of the Linear that produced that residual
  // stream value (fc2 / attention out-proj / patch projection): fused into the LayerNorm backward
  TRY(theia_layernorm_bwd(c.AB(m->dtok), c.AB(m->x[L]), c.W(m->lnfw), c.AF(m->meanf), c.AF(m->rstdf), nullptr, dx,
                          c.G(m->lnfw), c.G(m->lnfb), c.G(m->lp[L - 1].f2b), M, D, c.s));
  for (int l = L - 1; l >= 0; --l) {
    const LayerP& p = m->lp[l];
    const LayerW& w = m->lw[l];
    const LayerA& a = m->la[l];
    // MLP
    TRY(wgrad(c, dx, c.AB(a.a), c.G(p.f2w), M, D, 4 * D));
    TRY(linear_dgrad(c, dx, c.PB(w.w2), c.AB(m->dh), M, 4 * D, D, THEIA_EPI_MUL_AUX | THEIA_EPI_COLSUM | THEIA_EPI_AUX_U8, c.AB(a.h),
                     c.G(p.f1b)));
    TRY(wgrad(c, c.AB(m->dh), c.AB(a.ln2), c.G(p.f1w), M, 4 * D, D));
    TRY(linear_dgrad(c, c.AB(m->dh), c.PB(w.w1), c.AB(m->dln), M, D, 4 * D, 0));
    TRY(theia_layernorm_bwd(c.AB(m->dln), c.AB(a.xmid), c.W(p.ln2w), c.AF(a.mean2), c.AF(a.rstd2), dx, dx2,
                            c.G(p.ln2w), c.G(p.ln2b), c.G(p.ob), M, D, c.s));
    // attention
    TRY(wgrad(c, dx2, c.AB(a.attn), c.G(p.ow), M, D, D));
    TRY(linear_dgrad(c, dx2, c.PB(w.wo), c.AB(m->dattn), M, D, D, 0));
    TRY(theia_attention_tc_bwd(c.AB(a.qkv), c.AB(a.attn), c.AB(m->dattn), c.AF(a.lse), c.AB(m->dqkv), B, NT, H, c.s));
    TRY(wgrad(c, c.AB(m->dqkv), c.AB(a.ln1), c.G(p.qw), M, 3 * D, D));
    TRY(theia_colsum(c.AB(m->dqkv), c.G(p.qb), M, 3 * D, 3 * D, 0, c.s));
    TRY(linear_dgrad(c, c.AB(m->dqkv), c.PB(w.wqkv), c.AB(m->dln), M, D, 3 * D, 0));
    TRY(theia_layernorm_bwd(c.AB(m->dln), c.AB(m->x[l]), c.W(p.ln1w), c.AF(a.mean1), c.AF(a.rstd1), dx2, dx,
                            c.G(p.ln1w), c.G(p.ln1b), l > 0 ? c.G(m->lp[l - 1].f2b) : nullptr, M, D, c.s));
  }
  // embeddings: position / cls / patch projection
  TRY(theia_batchsum(dx, c.AF(m->tgrad), B, NT * D, c.s));
  token_table_grad_kernel<<<(NT * D + 255) / 256, 256, 0, c.s>>>(c.AF(m->tgrad), c.G(m->pos), m->cls >= 0 ? c.G(m->cls) : nullptr,
                                                                 m->regt >= 0 ? c.G(m->regt) : nullptr,
                                                                 m->regp >= 0 ? c.G(m->regp) : nullptr, NT, D, m->p0);
  THEIA_CHECK_LAUNCH("token_table_grad");
  TRY(wgrad(c, dx, c.AB(m->patches), c.G(m->pew), M, D, 768));
  TRY(theia_colsum_tokens(dx, c.G(m->peb), M, D, D, NT, m->p0, m->p0 + 196, c.s));
  return THEIA_OK;
}

// Host-side launch plan of the wgrad GEMM dW[Nout,Kin] += dY[Mtok,Nout]^T X[Mtok,Kin] (z tap slices for a
// convolution): no GPU work, lets the CPU tests check that no shape of the step leaves the grid underfilled.
extern "C" int theia_plan_wgrad(int Nout, int Kin, int Mtok, int z, int* bn, int* ctas_per_mma, int* splits) {
  if (!bn || !ctas_per_mma || !splits || Nout <= 0 || Kin <= 0 || Mtok <= 0 || z <= 0)
    return set_error(THEIA_ERR_ARG, "theia_plan_wgrad: bad argument");
  plan_wgrad(Nout, Kin, Mtok, z, bn, ctas_per_mma, splits);
  return THEIA_OK;
}
