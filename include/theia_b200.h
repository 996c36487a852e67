/* theia_b200 -- C ABI of the B200-native Theia distillation hot path.
 *
 * Plain C: device pointers are `void*` / typed pointers into CUDA memory that the CALLER owns
 * (PyTorch allocates every tensor, workspace included); every entry point takes an explicit
 * `cudaStream_t` (passed as void*), never synchronises, never allocates on the hot path, and
 * returns 0 on success or a non-zero error code (`theia_last_error()` gives the text).
 *
 * The reference (bdaiinstitute/theia) has no native layer; the entry points below replace the
 * PyTorch / transformers library calls its Python makes on this path.  Each one cites the
 * reference call site it replaces (paths relative to /root/reference, or `hf:` =
 * site-packages/transformers 5.5.0).
 */
#ifndef THEIA_B200_H
#define THEIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THEIA_OK 0
#define THEIA_ERR_ARG 1
#define THEIA_ERR_CUDA 2
#define THEIA_ERR_UNSUPPORTED 3

const char* theia_last_error(void);
int theia_version(void);
/* Number of kernels this library has launched since load (all streams, this process). */
long long theia_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * tcgen05 GEMM / implicit-GEMM convolution.   D[M,N] (+)= A[M,K] * B[N,K]^T , bf16 in, fp32
 * accumulate in TMEM, fused epilogue.  Replaces: nn.Linear (hf:models/vit/modeling_vit.py:216-230,
 * 262, 290, 305; src/theia/models/adapter_heads.py:313-314,325-326), nn.Conv2d(3,D,16,16)
 * (hf:modeling_vit.py:151,166), nn.Conv2d 3x3 / nn.ConvTranspose2d (adapter_heads.py:282-288,
 * 304-327) and their autograd dgrad / wgrad.
 * ------------------------------------------------------------------------------------------ */
enum {
  THEIA_OP_K2D = 0,    /* operand stored [rows][K]  (K contiguous), 2-D                           */
  THEIA_OP_MN2D = 1,   /* operand stored [K][rows]  (rows contiguous), 2-D  (wgrad operands)       */
  THEIA_OP_CONV_K = 2, /* A only: NHWC activation gathered per tap, K = ntaps * C                  */
  THEIA_OP_CONV_MN = 3 /* B only: NHWC activation gathered for ONE tap (= z index), K = pixels     */
};

enum {
  THEIA_EPI_GELU = 1 << 1,      /* out2 = bf16(gelu'(v)) (saved for backward), v = gelu(v)          */
  THEIA_EPI_RELU = 1 << 2,      /* v = max(v, 0)                                                    */
  THEIA_EPI_RESID = 1 << 3,     /* v += aux[m,n]        (aux may alias out: accumulate)             */
  THEIA_EPI_OUT_F32 = 1 << 4,   /* store fp32 instead of bf16                                       */
  THEIA_EPI_ATOMIC = 1 << 5,    /* fp32 atomicAdd into out (split-K wgrad)                          */
  THEIA_EPI_MUL_AUX = 1 << 6,   /* v *= aux[m,n]   (backward of GELU with the saved derivative)      */
  THEIA_EPI_MUL_RELUMASK = 1 << 7, /* v = aux[m,n] > 0 ? v : 0                                      */
  THEIA_EPI_POSCLS = 1 << 8,    /* token t = m % tokens: patch token (tok_p0 <= t < tok_p1) ? v + pos[t,n]
                                   : pos[t,n]   (pos = per-token table incl. CLS / register tokens)      */
  THEIA_EPI_STATS = 1 << 9,     /* per-image sum / sum-of-squares of the stored values -> stats     */
  THEIA_EPI_COLSUM = 1 << 10,   /* colsum[n] += sum over rows of the stored (bf16) values           */
  THEIA_EPI_GELU_FWD = 1 << 11, /* v = gelu(v), no derivative output (inference: teacher ViTs)      */
  THEIA_EPI_QUICK_GELU = 1 << 12,/* v = v * sigmoid(1.702 v)  (hf:activations.py QuickGELUActivation; CLIP) */
  THEIA_EPI_RESID_F32 = 1 << 13,/* v += aux[m,n] with aux in fp32 (with OUT_F32: the fp32 residual stream of the
                                   teacher path; aux may alias out) */
  THEIA_EPI_AUX_U8 = 1 << 14    /* modifier: the saved GELU derivative is an 8-bit code, q = rint((g' + 0.129) * 255 /
                                   1.258) (g' lies in [-0.1289, 1.1289]; step 4.9e-3, below bf16's own step near 1).
                                   With GELU: out2 is uint8 [M][ldo]; with MUL_AUX: aux is that tensor.  The fc1 GEMM
                                   is bound by HBM write bandwidth: 3 instead of 4 bytes written per element.
                                   N and ldo must be multiples of 32 */
};

typedef struct theia_conv_geom {
  /* the NHWC tensor gathered by a CONV_* operand: element strides, extents */
  int C, H, W, B;
  long long stride_w, stride_h, stride_b; /* in elements; channel stride is 1 */
  int ntaps;                              /* CONV_K: taps folded into K.  CONV_MN: number of z slices */
  int dh[9], dw[9];                       /* input pixel = output pixel + (dh, dw); OOB reads are zero */
  int tile_w, tile_h;                     /* CONV_K: M-tile = tile_h x tile_w output pixels (=128)   */
  int out_h, out_w;                       /* valid output extent (rows with h>=out_h or w>=out_w are not stored) */
  /* output row of pixel (b,h,w) = b*out_img_rows + out_row_off + (h*sy+py)*out_wpitch + (w*sx+px) */
  int out_img_rows, out_row_off, out_wpitch, sy, sx, py, px;
  int in_stride;    /* gather stride of the NHWC operand (1, or 2 for the dgrad / wgrad of a stride-2 transposed
                       conv): input pixel = in_stride * output pixel + (dh, dw), via TMA element strides     */
  int b_tap_rows;   /* CONV_K: 0 = B is [N][ntaps*C]; >0 = B is tap-major [tap][b_tap_rows][C] and wtap[t] selects
                       the weight tap of GEMM tap t (one 9-tap pack serves all parity classes)                  */
  int wtap[9];
} theia_conv_geom;

typedef struct theia_gemm_desc {
  int M, N, K;      /* CONV_K: M = B * tiles_per_image * 128 virtual rows, K = ntaps*C          */
  int a_mode, b_mode;
  const void* A;    /* bf16 */
  long long lda;    /* elements between consecutive rows of the stored 2-D operand              */
  const void* B;    /* bf16 */
  long long ldb;
  theia_conv_geom conv; /* used when a_mode == CONV_K or b_mode == CONV_MN */
  int epi;          /* THEIA_EPI_* flags */
  void* out;        /* bf16 (default) or fp32 */
  long long ldo;
  void* out2;       /* GELU: gelu'(pre-activation), bf16, same shape as out */
  const float* bias;/* [N] or NULL */
  const void* aux;  /* bf16, indexed like out */
  const float* pos; /* POSCLS: [tokens, N] */
  const float* cls; /* POSCLS: [N] */
  int tokens;       /* POSCLS: tokens per image */
  float* stats;     /* STATS: [images][2] */
  int rows_per_image; /* STATS (non-conv A): rows of M per image */
  int splits;       /* split-K factor (>=1); requires ATOMIC when > 1 */
  int batch_z;      /* z slices (CONV_MN taps); out advances by out_z_stride elements per z */
  long long out_z_stride;
  int bn;           /* N tile: 0 = auto, else 128 / 192 / 256 */
  float* colsum;    /* COLSUM: [N] fp32, accumulated */
  int tok_p0, tok_p1; /* POSCLS: token range of the patch tokens */
} theia_gemm_desc;

int theia_gemm(const theia_gemm_desc* d, void* stream);


/* ------------------------------------------------------------------------------------------
 * HBM-bound fused kernels (bf16 I/O, fp32 statistics).
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm(D) rows (hf:modeling_vit.py:325-326,333,340,455).  y = (x-mean)*rstd*gamma+beta */
int theia_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                        int M, int D, float eps, void* stream);
/* the same over fp32 rows (the fp32 residual stream of the teacher path); y is bf16, or fp32 when y_is_f32;
 * D % 4 == 0, D <= 1280 */
int theia_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, void* y, int y_is_f32, int M, int D,
                            float eps, void* stream);
/* dx = LN'(dy) (+ dadd); dgamma/dbeta are ACCUMULATED (+=) */
int theia_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                        const void* dadd, void* dx, float* dgamma, float* dbeta, float* dxsum, int M, int D,
                        void* stream); /* dxsum (optional): [D] += column sums of the produced dx */
/* nn.LayerNorm([C,H,W]) per image, NHWC storage, n = H*W*C (adapter_heads.py:306-324).  stats[b] =
 * {sum, sumsq} produced by the GEMM epilogue (THEIA_EPI_STATS). gamma/beta are the HWC-permuted affine. */
/* C, Wp, Wv: Wp > 0 = the map is stored zero-padded [Wp x Wp x C] with Wv x Wv valid pixels (the 31x31 stage of
 * the 64x64 heads is kept at pitch 32); padding is written as zero and excluded from the statistics */
int theia_ln3d_apply(const void* x, const float* stats, const float* gamma_hwc, const float* beta_hwc, void* y,
                     int B, int n, float eps, int C, int Wp, int Wv, void* stream);
/* red: scratch [B][2]; dgamma/dbeta (HWC) ACCUMULATED; relu_mask: also apply the producer's ReLU mask (x>0) */
int theia_ln3d_bwd(const void* dy, const void* x, const float* stats, const float* gamma_hwc, float* red, void* dx,
                   float* dgamma_hwc, float* dbeta_hwc, int B, int n, float eps, int relu_mask, int C, int Wp, int Wv,
                   void* stream);
/* Loss terms of one teacher (src/theia/models/rvfm.py:153-176): acc scratch [B][5]; out3 = {mse, cos, l1} */
int theia_loss_fwd(const float* pred, const void* target, int target_is_bf16, float* acc, float* out3, int B, int n,
                   void* stream);
/* dpred (bf16 or fp32) = coef3[0]*d(mse)/dp + coef3[1]*d(cos)/dp + coef3[2]*d(l1)/dp ; coef3 lives on the device.
 * dpred_bf16_copy (optional, with an fp32 dpred): a bf16 copy written in the same pass (the operand of the head-Linear
 * backward GEMMs; saves a separate cast over the largest tensor of the loss path) */
int theia_loss_bwd(const float* pred, const void* target, int target_is_bf16, const float* acc, const float* coef3,
                   void* dpred, int dpred_is_f32, void* dpred_bf16_copy, int B, int n, void* stream);
/* DeiT image processor (backbones.py:337-339; hf:image_processing_backends.py:361-414): uint8 HWC/CHW
 * 224x224 -> [bicubic-antialias resize to 256 + centre crop 224 when do_resize: 1 = float arithmetic + round, what
 * torchvision does for CUDA tensors; 2 = int16 fixed-point two-pass scheme with a uint8 intermediate, what it does
 * for CPU uint8 tensors -- both bit-exact with torchvision] -> rescale/normalise ->
 * bf16 patch rows [B*197, 768] (row b*197 is the zero CLS slot; column = c*256 + i*16 + j) */
int theia_preprocess(const uint8_t* images, void* patches, int B, int channels_first, int do_resize, int do_rescale,
                     int do_normalize, const float* mean3, const float* std3, int tokens, int patch_off, void* stream);
/* the same for images of any extent (the reference's processor accepts whatever the caller has): do_resize != 0
 * resizes in_h x in_w -> 256 x 256 (bicubic antialias; 1 / 2 = float / fixed-point arithmetic as above) and crops the
 * centre 224; do_resize == 0 crops the centre 224 x 224, zero-padding images that are smaller
 * (hf:image_processing_backends.py center_crop).  224 x 224 inputs take the kernels of theia_preprocess.  The per-axis
 * tap tables are built on the first use of an extent (one cudaMalloc per new extent, cached per device). */
int theia_preprocess_hw(const uint8_t* images, int in_h, int in_w, void* patches, int B, int channels_first, int do_resize,
                        int do_rescale, int do_normalize, const float* mean3, const float* std3, int tokens, int patch_off,
                        void* stream);
/* float pixels (the processor also accepts float tensors: [0, 255] values, or [0, 1] with do_rescale = 0): centre crop /
 * zero pad to 224 x 224 and the same fused rescale / normalise; no resize on this path */
int theia_preprocess_f32(const float* images, int in_h, int in_w, void* patches, int B, int channels_first, int do_rescale,
                         int do_normalize, const float* mean3, const float* std3, int tokens, int patch_off, void* stream);
/* test hook for the byte stage: resized_u8_out != NULL -> following do_resize calls also write the resized +
 * centre-cropped uint8 image [B,224,224,3] (what tvF.resize(..., antialias=True) + center_crop give the reference) */
int theia_preprocess_debug_u8(void* resized_u8_out);
/* tokens / patch_off: rows per image of the patch matrix and the row of the first patch (197 / 1 for DeiT, 196 / 0
 * for DeiTNoCLS, 204 / 1 for DeiTReg); rows of non-patch tokens are zero */
/* attention (hf:modeling_vit.py:171-196,232-246): qkv [B*N,3*H*64] bf16 -> out [B*N,H*64]; lse [B,H,N] */
/* tcgen05 kernels: S / dP accumulators in TMEM, TMA-fed, persistent */
int theia_attention_tc_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream);
int theia_attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B,
                           int N, int H, void* stream);
/* head dim 80 (google/vit-huge-patch14-224-in21k, src/theia/foundation_models/vision_models/vit.py:36): forward only,
 * qkv [B*N, 3*H*80] -> out [B*N, H*80], scale 1/sqrt(80), N <= 272; dims 64..79 of a head travel as a second
 * 32-byte-swizzled operand tile */
int theia_attention_fwd_hd80(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream);
/* parameter packing helpers */
int theia_gather4(const void* in, void* out, int in_is_f32, int out_is_f32, int n0, int n1, int n2, int n3,
                  long long s0, long long s1, long long s2, long long s3, long long base, void* stream);
/* Fused optimizer tail (train_rvfm.py:126-133; optimizers/utils.py:8-35): torch.optim.AdamW arithmetic over the
 * flat fp32 parameter / gradient / moment buffers in one pass.  flag64[i] bit 0 = elements [64 i, 64 i + 64)
 * belong to the weight-decay group, bit 1 = skip the block entirely (parameter without a gradient: torch leaves it
 * untouched); max_grad_norm > 0 applies clip_grad_norm_ first (scratch2: 2 floats).  pack_table / packbf (optional,
 * from theia_model_pack_table): the same pass refreshes the bf16 GEMM-operand copy of each updated weight block
 * (pack_table[i] = destination 64-element block in packbf, or -1). */
int theia_adamw_flat(float* p, const float* g, float* m, float* v, const uint8_t* flag64, long long n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, float max_grad_norm,
                     float* scratch2, const int* pack_table, void* packbf, void* stream);
/* bf16 copies of all blocks the pack table names (one launch; what theia_adamw_flat fuses) */
int theia_pack_cast(const float* p, const int* pack_table, void* packbf, long long n, void* stream);
/* Conv-weight layout conversions through shared-memory tiles (coalesced on both sides), one launch for all conv
 * weights.  Reference layout W[a][b][9] (Conv2d: a = Cout, b = Cin; ConvTranspose2d: a = Cin, b = Cout) <-> GEMM
 * packs P[X][T][Y] (or tap-major P[T][X][Y]) with (X, Y) = (a, b) / (b, a) and T = t / 8 - t. */
enum { THEIA_CP_SWAP = 1, THEIA_CP_FLIP = 2, THEIA_CP_TAPMAJOR = 4 };
typedef struct theia_conv_perm {
  const void* ref;   /* pack: fp32 master weight.  unpack: byte OFFSET of the weight in the gradient buffer */
  void* pack0;       /* pack: bf16 pack (or NULL).    unpack: fp32 tap-major scratch ws[T][X][Y]             */
  void* pack1;       /* pack: second bf16 pack (or NULL) */
  int flags0, flags1;
  int C;
} theia_conv_perm;
int theia_conv_pack(const theia_conv_perm* segs_dev, int nconv, int C, void* stream);
int theia_conv_unpack(const theia_conv_perm* segs_dev, int nconv, int C, const void* out_rebase, void* stream);
/* One launch over a device table of strided layout conversions (conv-weight packs, LayerNorm[C,H,W] affine <-> NHWC,
 * conv weight-gradient scratch -> reference layout); `in` is fp32, `out` bf16 or fp32 */
typedef struct theia_perm_seg {
  const void* in;
  void* out;
  int n0, n1, n2, n3;           /* output extents, row-major */
  long long s0, s1, s2, s3;     /* input element strides */
  long long base;               /* input element offset */
  int lim1, lim2;               /* >0: output is zero where index1 >= lim1 or index2 >= lim2 (zero padding) */
  int out_f32;
  long long first_block;        /* prefix sum of ceil(n0*n1*n2*n3 / 256) over the preceding segments */
} theia_perm_seg;
/* out_rebase: byte address added to every segment's `out` (segments may store offsets relative to a buffer that is
 * chosen at launch time, e.g. the current gradient buffer); NULL = `out` is absolute */
int theia_perm_segments(const theia_perm_seg* segs_dev, int nseg, long long total_blocks, const void* out_rebase,
                        void* stream);
/* Target ingest (src/theia/dataset/data_utils.py:152-153,342-355): teacher embeddings [B,C,H*W] bf16 ->
 * [B,H*W,C] bf16, z-scored with per-channel bf16 mean/std (NULL = no normalisation); bit-exact with the
 * reference's bf16 `(x - mean) / std` */
int theia_target_ingest(const void* emb_chw, const void* mean_c, const void* std_c, void* out_hwc, int B, int C,
                        int HW, void* stream);
/* LayerNorm[C,H,W] affine layouts: reference [C][Hv][Wv] <-> NHWC zero-padded [Hp][Wp][C] (fp32) */
int theia_chw_to_hwc(const float* in, float* out, int C, int Hv, int Wv, int Hp, int Wp, void* stream);
int theia_hwc_to_chw(const float* in, float* out, int C, int Hv, int Wv, int Hp, int Wp, void* stream);
int theia_cast_bf16(const float* in, void* out, long long n, void* stream);
int theia_transpose_cast_bf16(const float* in, void* out, int R, int C, void* stream);
/* out[n] += sum_m x[m,n] (rows with m % skip_mod == 0 skipped when skip_mod > 0) */
int theia_colsum(const void* x, float* out, int M, int N, long long ld, int skip_mod, void* stream);
/* same, keeping only rows whose token index (m % period) lies in [t0, t1) */
int theia_colsum_tokens(const void* x, float* out, int M, int N, long long ld, int period, int t0, int t1, void* stream);
/* out[j] = sum_b x[b*n + j] */
int theia_batchsum(const void* x, float* out, int B, int n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-path model: replaces RobotVisionFM.forward / autograd backward
 * (src/theia/models/rvfm.py:115-136; src/theia/scripts/train/train_rvfm.py:116,125).
 * ------------------------------------------------------------------------------------------ */
#define THEIA_MAX_TEACHERS 8
typedef struct theia_model_config {
  int hidden, heads, layers, image, patch, max_batch;
  float ln_eps;
  int variant;        /* 0 = DeiT (CLS + patches), 1 = DeiTNoCLS (patches only), 2 = DeiTReg (CLS + patches + registers);
                         src/theia/models/backbones.py:506-526 */
  int num_reg_tokens; /* DeiTReg: 7 */
  int num_teachers;
  const char* teacher_names[THEIA_MAX_TEACHERS];
  int teacher_c[THEIA_MAX_TEACHERS];
  int teacher_hw[THEIA_MAX_TEACHERS]; /* target H (= W) */
} theia_model_config;
typedef struct theia_model theia_model;

int theia_model_create(const theia_model_config* cfg, theia_model** out);
void theia_model_destroy(theia_model* m);
long long theia_model_param_floats(const theia_model* m);    /* size of the flat fp32 parameter buffer */
long long theia_model_workspace_bytes(const theia_model* m);
int theia_model_num_params(const theia_model* m);
/* state_dict key, shape and element offset (in the flat buffer) of parameter i */
int theia_model_param_info(const theia_model* m, int i, char* name, int name_cap, long long* dims4, int* ndim,
                           long long* offset);
/* test / bring-up accessor to a named internal activation of the last forward (see csrc/model.cu) */
int theia_model_debug_ptr(theia_model* m, const char* name, int i, void** ptr, long long* elems, int* is_f32);
int theia_model_bind(theia_model* m, float* master, float* grads, void* workspace);
/* switch the flat gradient buffer theia_model_backward writes (same layout; no device work, no synchronisation) */
int theia_model_set_grads(theia_model* m, float* grads);
/* fp32 master -> bf16 / permuted operand copies.  skip_linear_cast != 0: the Linear-weight casts were already done
 * by theia_adamw_flat (pack table); only the token table and the permuted packs are refreshed. */
int theia_model_pack(theia_model* m, int skip_linear_cast, void* stream);
/* device pointers of the pack table (one int per 64 floats of the flat parameter buffer) and of the bf16 pack
 * buffer, valid after theia_model_bind */
int theia_model_pack_table(theia_model* m, const int** table, void** packbf);
/* extent of the image batches handed to theia_model_forward from now on (default 224 x 224; see theia_preprocess_hw) */
int theia_model_set_input_size(theia_model* m, int height, int width);
/* pixel type of those batches: 0 = uint8 (default), 1 = fp32 (theia_preprocess_f32; do_resize must then be 0) */
int theia_model_set_input_dtype(theia_model* m, int is_f32);
int theia_model_forward(theia_model* m, const uint8_t* images, int B, int channels_first, int do_resize,
                        int do_rescale, int do_normalize, const float* mean3, const float* std3, int run_heads, float* const* preds,
                        void* tokens_bf16_out, void* stream);
int theia_model_backward(theia_model* m, const void* const* dpreds, void* stream);

/* ------------------------------------------------------------------------------------------
 * Teacher ViT inference: the frozen foundation models the student is distilled from --
 * hf Dinov2Model / CLIPVisionModel / ViTModel as called by src/theia/foundation_models/vision_models/dinov2.py:26-31,
 * vision_language_models/clip.py:26-31 and vision_models/vit.py:23-28.  Forward only; weights are caller-owned
 * device buffers (bf16 GEMM operands [out][in], fp32 biases / LayerNorm affines); head dim 64 or 80, <= 272 tokens,
 * hidden <= 1280.  DINOv2's LayerScale is folded into w_o / b_o and w_fc2 / b_fc2 by the caller.
 * ------------------------------------------------------------------------------------------ */
typedef struct theia_vit_layer {
  const float *ln1_w, *ln1_b;
  const void* w_qkv;   /* bf16 [3*hidden][hidden]: query | key | value rows */
  const float* b_qkv;  /* [3*hidden] */
  const void* w_o;     /* bf16 [hidden][hidden] */
  const float* b_o;
  const float *ln2_w, *ln2_b;
  const void* w_fc1;   /* bf16 [mlp][hidden] */
  const float* b_fc1;
  const void* w_fc2;   /* bf16 [hidden][mlp] */
  const float* b_fc2;
} theia_vit_layer;
typedef struct theia_vit_desc {
  int hidden, heads, layers, mlp;
  int tokens, patch_off, patch_tokens; /* tokens per image; index of the first patch token; number of patch tokens */
  int patch_k;                         /* columns of the patch matrix: channels * patch^2 rounded up to 8 */
  float ln_eps;
  int act;                             /* 0 = gelu (erf), 1 = quick_gelu (CLIP) */
  const void* w_patch;                 /* bf16 [hidden][patch_k] (zero in the padding columns) */
  const float* b_patch;                /* [hidden] or NULL (CLIP's patch conv has no bias) */
  const float* tok_table;              /* fp32 [tokens][hidden]: position embedding (+ class embedding in row 0) */
  const float *pre_ln_w, *pre_ln_b;    /* CLIP pre_layrnorm, else NULL */
  const float *final_ln_w, *final_ln_b;/* NULL = none */
  int final_ln_mode;                   /* 1 = every token (Dinov2Model / ViTModel .layernorm): last_hidden is normalised,
                                          pooled = its CLS rows.  2 = CLIP: last_hidden is the raw encoder output,
                                          pooled = post_layernorm(CLS rows) */
  const theia_vit_layer* layer;        /* HOST array [layers] */
  int residual_f32;                    /* != 0: the residual stream x is kept in fp32 between the blocks (the bf16
                                          stream is the largest single error term of a 24-layer forward: 8e-3 of
                                          the 1e-2 total, measured); costs ~6 % time */
} theia_vit_desc;
long long theia_vit_workspace_bytes(const theia_vit_desc* d, int B);
/* patches: bf16 [B*tokens][patch_k] (theia_patchify_f32); last_hidden: bf16 [B*tokens][hidden]; pooled: bf16
 * [B][hidden] or NULL */
int theia_vit_forward(const theia_vit_desc* d, const void* patches, int B, void* workspace, void* last_hidden, void* pooled,
                      void* stream);
/* processor output pixel_values fp32 [B][C][H][W] -> the bf16 patch matrix (non-patch rows / padding columns zero) */
int theia_patchify_f32(const float* pixel_values, void* patches, int B, int C, int H, int W, int patch, int tokens,
                       int patch_off, int patch_k, void* stream);

/* Per-launch CUDA-event timing of the GEMM kernel on its launching stream (bench.py roofline). */
int theia_prof_enable(int on);
int theia_prof_collect(double* total_ms, double* total_flops, long long* launches);
int theia_prof_record(int i, double* ms, int* meta8);

/* Host-side launch plan of a weight-gradient GEMM dW[Nout,Kin] += dY[Mtok,Nout]^T X[Mtok,Kin] (z = tap slices of a
 * convolution wgrad, 1 for a Linear): N tile, CTAs per MMA (1 / 2) and split-K factor.  No GPU work. */
int theia_plan_wgrad(int Nout, int Kin, int Mtok, int z, int* bn, int* ctas_per_mma, int* splits);

/* Diagnostics of the GEMM kernel; key 0 clears all.
 *   1-6: shared-memory descriptor field overrides (bring-up of new operand layouts)
 *   7  : bit 0 = no operand loads after the first ring fill (MMA rate without operand traffic),
 *        bit 1 = skip the epilogue (mainloop rate); results are garbage, timing only
 *   8  : 1 = single-CTA kernels only (tcgen05 cta_group::1), 2 = CTA pairs wherever the tile allows, 0 = launcher's choice
 * Environment: THEIA_GEMM_SINGLE_CTA (same as key 8 = 1), THEIA_SPLITK_EFF (wave efficiency at which split-K stops). */
int theia_debug_set(int key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* THEIA_B200_H */
