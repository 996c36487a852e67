// tcgen05 self-attention for the ViT token sequence (<= 208 tokens, head_dim 64), forward and backward.
// Replaces hf:models/vit/modeling_vit.py:171-196,232-246 (SDPA / eager attention) and its autograd.
//
// One persistent CTA per SM loops over (image, head) items.  Q / K / V (and dO) of an item are TMA-loaded as
// 128-byte-swizzled [rows x 64] tiles straight from the fused qkv activation [B*N, 3*D]; score matrices live in TMEM;
// probabilities go back to shared memory as bf16 MMA operands; V (and K, Q, dO in the backward) are read as MN-major
// B operands in place, so nothing is ever transposed or copied.  The whole key row of a query is on chip at once:
// single-pass softmax, no online rescaling.  MMA-issuing warps run warp-uniform control flow with one elected lane
// issuing (descriptors stay in uniform registers); all hand-offs are mbarriers.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

constexpr int AT_HD = 64;
constexpr int AT_ROWS = 208;                  // rows per operand tile (197 padded to a multiple of 16)
constexpr int AT_TILE = AT_ROWS * 128;        // 26624 B, a multiple of the 1024-B swizzle atom
constexpr int AT_PBUF = 128 * 128 * 4;        // P tile: 4 K-blocks of [128 rows x 64 keys] = 64 KB
constexpr int AT_KSTEPS = AT_ROWS / 16;       // 13 MMA K-steps over the keys

// Descriptor words for the MMA-issuing warps (128-byte swizzle, SBO = 1024, version 1): the high word is a constant,
// the low word = address >> 4 | (LBO >> 4) << 16 advances by plain 32-bit adds (k-step of 32 bytes = +2, of 2048 = +128).
// The whole issuing warp runs the (warp-uniform) control flow and one elected lane issues, which lets the compiler
// keep all of this in uniform registers instead of wrapping every tcgen05.mma in a broadcast loop (gemm_tc.cu).
constexpr uint32_t AT_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t at_desc_lo(uint32_t saddr, uint32_t lbo) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t at_desc(uint32_t lo) { return (static_cast<uint64_t>(AT_DESC_HI) << 32) | lo; }

struct AttnFwdParams {
  bf16* out;
  float* lse;
  int B, N, H, D;
  int items;
  float scale;
};


// --------------------------------------------------------------------------------------------
// forward.  Per item two 128-query tiles u:  S_u = Q_u K^T (N = 208 keys) -> TMEM;  P_u = exp2((S_u - rowmax) c)
// (bf16, shared memory), l = rowsum;  O_u = P_u V -> TMEM -> out, lse.
//   * the 8 softmax warps work on ONE tile at a time, two warps per TMEM lane quarter, each thread holding its half
//     of a score row (112 / 96 columns) in REGISTERS: the scores are read from TMEM once, the row maximum / sum of
//     the two halves are exchanged through shared memory, and the TMEM score buffer is free for the next tile's
//     MMAs as soon as it has been loaded;
//   * two score buffers (columns 0-207 / 208-415) and a separate output accumulator (416-479): S_{u+1} = Q K^T and
//     O_{u-1} = P V run on the tensor core WHILE the softmax of tile u runs on the SIMT pipes;
//   * Q / K / V are double buffered across (image, head) items, so the loads of the next item are never exposed.
// warps: 0 TMA, 1 MMA issue, 2 TMEM alloc, 4-11 softmax (quarter = warp % 4, column half = (warp - 4) / 4).
// --------------------------------------------------------------------------------------------
constexpr int AF_THREADS = 384;
constexpr int AF_NSW = 8;
constexpr int AF_OFF_Q = 0;
constexpr int AF_OFF_K = 2 * AT_TILE;
constexpr int AF_OFF_V = 4 * AT_TILE;
constexpr int AF_OFF_P = 6 * AT_TILE;           // 4 blocks of [128 queries x 64 keys]
constexpr int AF_OFF_X = AF_OFF_P + AT_PBUF;     // exchange: max[2][128] | sum[2][128] floats
constexpr int AF_OFF_BAR = AF_OFF_X + 4 * 128 * 4;
constexpr int AF_SMEM = AF_OFF_BAR + 256 + 1024;
static_assert(AF_SMEM <= 232448, "attention forward: shared memory");


// Softmax of one 128-query score tile for the column half HH of this thread's row: scores -> registers (the TMEM
// buffer is released right after the load), row max / sum exchanged with the partner warp of the same lane quarter
// through shared memory (named barrier 1 + quarter, 64 threads), P written as the K-major bf16 A operand of P V.
// NFIX > 0: the sequence length is a compile-time constant, so only the chunk that holds padding keys is masked.
template <int HH, int NFIX>
__device__ __forceinline__ void softmax_tile(uint32_t tsb, int n_rt, float c2, uint64_t* s_free_bar, uint64_t* o_full_bar,
                                             uint32_t o_wait, uint32_t o_parity, uint32_t xs, int wq, int r, int lane,
                                             uint32_t prow_s, int sw, float& m_out, float& l_out) {
  constexpr int CH0 = HH ? 7 : 0;
  constexpr int NCH = HH ? AT_KSTEPS - 7 : 7;
  const int N = NFIX > 0 ? NFIX : n_rt;
  uint32_t v[NCH][16];
#pragma unroll
  for (int k = 0; k < NCH; ++k) tmem_ld16(tsb + (CH0 + k) * 16, v[k]);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_free_bar);
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    constexpr int dummy = 0;
    (void)dummy;
    const int c0 = (CH0 + k) * 16;
    if (NFIX > 0 ? (c0 + 16 > NFIX) : (c0 + 16 > N)) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c0 + e >= N) v[k][e] = __float_as_uint(-INFINITY);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) m = fmaxf(m, __uint_as_float(v[k][e]));
  }
  sts32(xs + (HH * 128 + r) * 4, __float_as_uint(m));
  asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
  m = fmaxf(m, lds_f32(xs + ((HH ^ 1) * 128 + r) * 4));
  const float mc = m * c2;
  float l = 0.f;
  uint32_t pk[NCH][8];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    float pv[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) pv[e] = ex2_approx_ftz(fmaf(__uint_as_float(v[k][e]), c2, -mc));
    float l4[4] = {0.f, 0.f, 0.f, 0.f};  // four partial sums: no 16-deep dependent FADD chain behind the MUFU results
#pragma unroll
    for (int e = 0; e < 16; ++e) l4[e & 3] += pv[e];
    l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
#pragma unroll
    for (int e = 0; e < 8; ++e) pk[k][e] = pack_bf16x2(pv[2 * e], pv[2 * e + 1]);
  }
  sts32(xs + (256 + HH * 128 + r) * 4, __float_as_uint(l));
  asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
  l += lds_f32(xs + (256 + (HH ^ 1) * 128 + r) * 4);
  // P_{g-1} V has finished: its accumulator is complete and the P buffer may be overwritten
  if (o_wait) {
    mbar_wait(o_full_bar, o_parity);
    tc_fence_after();
  }
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    constexpr int dummy2 = 0;
    (void)dummy2;
    const int gch = CH0 + k;
    const uint32_t blk = prow_s + (gch >> 2) * 16384;
    const int ck = (gch & 3) * 2;
    sts128(blk + ((ck ^ sw) << 4), pk[k][0], pk[k][1], pk[k][2], pk[k][3]);
    sts128(blk + (((ck + 1) ^ sw) << 4), pk[k][4], pk[k][5], pk[k][6], pk[k][7]);
  }
  m_out = m, l_out = l;
}

template <int NFIX>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw_at[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at) + 1023) & ~uintptr_t(1023));
  float* xmax = reinterpret_cast<float*>(smem + AF_OFF_X);  // [2 halves][128 rows]
  float* xsum = xmax + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AF_OFF_BAR);
  uint64_t* qk_full = bars + 0;   // [2]
  uint64_t* qk_empty = bars + 2;  // [2]
  uint64_t* v_full = bars + 4;    // [2]
  uint64_t* v_empty = bars + 6;   // [2]
  uint64_t* s_full = bars + 8;    // [2] score buffers
  uint64_t* s_free = bars + 10;   // [2]
  uint64_t* p_full = bars + 12;
  uint64_t* o_full = bars + 13;
  uint64_t* o_free = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], AF_NSW);
    }
    mbar_init(p_full, AF_NSW);
    mbar_init(o_full, 1);
    mbar_init(o_free, AF_NSW);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int D = p.D;
  const uint32_t T_O = tmem_base + 2 * AT_ROWS;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * p.N, col0 = h * AT_HD;
        const int slot = it & 1;
        if (it >= 2) mbar_wait(&qk_empty[slot], ((it >> 1) - 1) & 1);
        mbar_expect_tx(&qk_full[slot], 2 * AT_TILE);
        tma_load_2d(&tmQKV, smem + AF_OFF_Q + slot * AT_TILE, &qk_full[slot], col0, row0);
        tma_load_2d(&tmQKV, smem + AF_OFF_K + slot * AT_TILE, &qk_full[slot], D + col0, row0);
        if (it >= 2) mbar_wait(&v_empty[slot], ((it >> 1) - 1) & 1);
        mbar_expect_tx(&v_full[slot], AT_TILE);
        tma_load_2d(&tmQKV, smem + AF_OFF_V + slot * AT_TILE, &v_full[slot], 2 * D + col0, row0);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc_s = make_idesc_bf16(128, AT_ROWS, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, AT_HD, 0, 1);
      const uint32_t p_lo = at_desc_lo(smem_u32(smem + AF_OFF_P), 16);
      int n_items = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) ++n_items;
      const int n_units = 2 * n_items;
      auto issue_s = [&](int g) {  // S_g = Q_t K^T into score buffer g & 1
        const int it = g >> 1, t = g & 1, slot = it & 1, sb = g & 1;
        if (t == 0) mbar_wait(&qk_full[slot], (it >> 1) & 1);
        if (g >= 2) mbar_wait(&s_free[sb], ((g >> 1) - 1) & 1);  // tile g-2 has been pulled into registers
        tc_fence_after();
        const uint32_t q_lo = at_desc_lo(smem_u32(smem + AF_OFF_Q + slot * AT_TILE) + t * 16384, 16);
        const uint32_t k_lo = at_desc_lo(smem_u32(smem + AF_OFF_K + slot * AT_TILE), 16);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base + sb * AT_ROWS, at_desc(q_lo + 2 * k), at_desc(k_lo + 2 * k), idesc_s, k > 0 ? 1u : 0u);
          tc_commit(&s_full[sb]);
          if (t == 1) tc_commit(&qk_empty[slot]);
        }
        __syncwarp();
      };
      issue_s(0);
      for (int g = 0; g < n_units; ++g) {
        if (g + 1 < n_units) issue_s(g + 1);
        const int it = g >> 1, t = g & 1, slot = it & 1;
        if (t == 0) mbar_wait(&v_full[slot], (it >> 1) & 1);
        mbar_wait(p_full, g & 1);
        if (g > 0) mbar_wait(o_free, (g - 1) & 1);  // O_{g-1} has been read out
        tc_fence_after();
        const uint32_t v_lo = at_desc_lo(smem_u32(smem + AF_OFF_V + slot * AT_TILE), 8192);
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < AT_KSTEPS; ++ks)
            tc_mma_bf16(T_O, at_desc(p_lo + (ks >> 2) * 1024 + (ks & 3) * 2), at_desc(v_lo + ks * 128), idesc_pv,
                        ks > 0 ? 1u : 0u);
          tc_commit(o_full);
          if (t == 1) tc_commit(&v_empty[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;          // TMEM lane quarter
    const int hh = (warp - 4) >> 2;   // column half: 0 -> chunks 0..6, 1 -> chunks 7..12
    const int r = wq * 32 + lane;     // row within the tile
    const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
    const float c2 = p.scale * 1.4426950408889634f;
    const uint32_t prow_s = smem_u32(smem + AF_OFF_P + r * 128);
    const uint32_t xs = smem_u32(xmax);
    const int sw = r & 7;
    // state of the previous tile (its output is read out one tile later, while its successor's MMAs run)
    float m_prev = 0.f, l_prev = 1.f;
    int item_prev = -1, t_prev = 0;
    auto store_out = [&]() {  // O_{prev} columns hh*32 .. +31 of this lane's row
      uint32_t o[2][16];
      tmem_ld16(T_O + lane_sel + hh * 32, o[0]);
      tmem_ld16(T_O + lane_sel + hh * 32 + 16, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      const int q = t_prev * 128 + r;
      if (q < p.N) {
        const int b = item_prev / p.H, h = item_prev - b * p.H;
        const float inv = 1.f / l_prev;
        bf16* dst = p.out + (static_cast<long long>(b) * p.N + q) * D + h * AT_HD + hh * 32;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 lo, hi;
          lo.x = pack_bf16x2(__uint_as_float(o[j][0]) * inv, __uint_as_float(o[j][1]) * inv);
          lo.y = pack_bf16x2(__uint_as_float(o[j][2]) * inv, __uint_as_float(o[j][3]) * inv);
          lo.z = pack_bf16x2(__uint_as_float(o[j][4]) * inv, __uint_as_float(o[j][5]) * inv);
          lo.w = pack_bf16x2(__uint_as_float(o[j][6]) * inv, __uint_as_float(o[j][7]) * inv);
          hi.x = pack_bf16x2(__uint_as_float(o[j][8]) * inv, __uint_as_float(o[j][9]) * inv);
          hi.y = pack_bf16x2(__uint_as_float(o[j][10]) * inv, __uint_as_float(o[j][11]) * inv);
          hi.z = pack_bf16x2(__uint_as_float(o[j][12]) * inv, __uint_as_float(o[j][13]) * inv);
          hi.w = pack_bf16x2(__uint_as_float(o[j][14]) * inv, __uint_as_float(o[j][15]) * inv);
          *reinterpret_cast<uint4*>(dst + j * 16) = lo;
          *reinterpret_cast<uint4*>(dst + j * 16 + 8) = hi;
        }
        if (hh == 0 && p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = m_prev * p.scale + __logf(l_prev);
      }
    };
    int g = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      for (int t = 0; t < 2; ++t, ++g) {
        const int sb = g & 1;
        mbar_wait(&s_full[sb], (g >> 1) & 1);
        tc_fence_after();
        float m, l;
        const uint32_t tsb = tmem_base + lane_sel + sb * AT_ROWS;
        const uint32_t o_wait = (g > 0) ? 1u : 0u;
        if (hh == 0)
          softmax_tile<0, NFIX>(tsb, p.N, c2, &s_free[sb], o_full, o_wait, (g - 1) & 1, xs, wq, r, lane, prow_s, sw, m, l);
        else
          softmax_tile<1, NFIX>(tsb, p.N, c2, &s_free[sb], o_full, o_wait, (g - 1) & 1, xs, wq, r, lane, prow_s, sw, m, l);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (g > 0) store_out();
        m_prev = m, l_prev = l, item_prev = item, t_prev = t;
      }
    }
    if (g > 0) {
      mbar_wait(o_full, (g - 1) & 1);
      tc_fence_after();
      store_out();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// --------------------------------------------------------------------------------------------
// forward for longer sequences (209 .. 272 tokens: the 257-token ViT-L/14 teachers of feature_extraction.py).
// Same operand handling as above, simpler schedule: one 128-query tile at a time, a single 272-column score buffer
// (two MMAs per k-step, N = 256 + 16: the UMMA N limit is 256), Q / K / V single buffered (V on its own barrier: it is
// not needed before the first P V).  Softmax: two warps per TMEM lane quarter, each owning half of the 17 column
// chunks (9 / 8), two streaming passes over TMEM (row max, then exp / sum / bf16 P blocks to shared memory), row max
// and row sum exchanged through shared memory with a 64-thread named barrier.  The MMA warp issues S of the next
// query tile right behind P V of the current one (the score buffer is free once P is written), so the tensor pipe
// works under the output epilogue.
//   TMEM: S 0-271 | O 272-335.   warps: 0 TMA, 1 TMEM alloc + MMA issue, 2-9 softmax (quarter = warp % 4).
// --------------------------------------------------------------------------------------------
constexpr int AL_ROWS = 272;
constexpr int AL_TILE = AL_ROWS * 128;
constexpr int AL_KSTEPS = AL_ROWS / 16;
constexpr int AL_THREADS = 320;
constexpr int AL_XTILE = AL_ROWS * 32;                // head dim 80: dims 64..79 as a second [272 x 16] tile (32-byte swizzle)
constexpr int AL_TAIL_MAX = 2;  // up to this many rows beyond the last full 128-query tile go through the CUDA-core path
// per-head-dim shared-memory map: Q | K | V operand slots (main tile [+ extra tile]), P (5 blocks of [128 queries x 64
// keys]), row max / row sum exchange [2][2][128] fp32, tail-row scratch (scores [272], reductions [32], partial
// outputs [8][80]), barriers
template <int HD>
struct AlMap {
  static constexpr int SLOT = HD > 64 ? ((AL_TILE + AL_XTILE + 1023) / 1024) * 1024 : AL_TILE;
  static constexpr int OFF_P = 3 * SLOT;
  static constexpr int OFF_X = OFF_P + 5 * 16384;
  static constexpr int OFF_T = OFF_X + 2048;
  static constexpr int OFF_BAR = OFF_T + 4096;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static_assert(AL_TILE % 1024 == 0 && SLOT % 1024 == 0 && SMEM <= 232448, "attention (long): shared memory");
};
// 32-byte-swizzle descriptors of the extra tiles (SBO = 8 rows x 32 B, version 1, layout type 6)
constexpr uint32_t AL_DESC32_HI = (256u >> 4) | (1u << 14) | (6u << 29);
__device__ __forceinline__ uint64_t al_desc32(uint32_t lo) { return (static_cast<uint64_t>(AL_DESC32_HI) << 32) | lo; }

template <int NFIX, int HD>
__global__ void __launch_bounds__(AL_THREADS, 1)
attn_tc_fwd_long_kernel(const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tm16,
                        const __grid_constant__ CUtensorMap tx256, const __grid_constant__ CUtensorMap tx16,
                        const AttnFwdParams p) {
  using MAP = AlMap<HD>;
  constexpr bool XT = HD > 64;  // head dim 80 (ViT-H): the last 16 dims live in the extra tiles
  static_assert(HD == 64 || HD == 80, "head dim 64 or 80");
  extern __shared__ uint8_t smem_raw_at[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + MAP::SLOT;
  uint8_t* sV = smem + 2 * MAP::SLOT;
  uint8_t* sP = smem + MAP::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MAP::OFF_BAR);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* in_empty = bars + 2;
  uint64_t* s_full = bars + 3;
  uint64_t* p_full = bars + 4;
  uint64_t* o_full = bars + 5;
  uint64_t* t_free = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm256);
    tma_prefetch_desc(&tm16);
    if (XT) {
      tma_prefetch_desc(&tx256);
      tma_prefetch_desc(&tx16);
    }
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(in_empty, 1 + 8);  // tcgen05.commit of the last P V + the 8 softmax warps (tail rows read Q / K / V)
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(t_free, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int D = p.D;
  const int N = NFIX ? NFIX : p.N;
  // query rows beyond the last full tile: one or two (the 257-token case: 2 x 128 + 1) are computed by the softmax
  // warps on the CUDA cores from the shared-memory tiles while P V of the last tile runs -- a third tensor-core
  // tile for a single row would cost as much as a full one.  More rows than that get a regular (partial) tile.
  const int nfull = N / 128;
  const int ntail = (nfull > 0 && N - nfull * 128 <= AL_TAIL_MAX) ? N - nfull * 128 : 0;
  const int nqt = ntail ? nfull : (N + 127) / 128;
  const uint32_t T_O = tmem_base + AL_ROWS;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * N, col0 = h * HD;
        if (it > 0) mbar_wait(in_empty, (it - 1) & 1);
        mbar_expect_tx(qk_full, 2 * (AL_TILE + (XT ? AL_XTILE : 0)));
        for (int m = 0; m < 2; ++m) {  // q, k column blocks of the fused qkv activation
          tma_load_2d(&tm256, smem + m * MAP::SLOT, qk_full, m * D + col0, row0);
          tma_load_2d(&tm16, smem + m * MAP::SLOT + 256 * 128, qk_full, m * D + col0, row0 + 256);
          if (XT) {
            tma_load_2d(&tx256, smem + m * MAP::SLOT + AL_TILE, qk_full, m * D + col0 + 64, row0);
            tma_load_2d(&tx16, smem + m * MAP::SLOT + AL_TILE + 256 * 32, qk_full, m * D + col0 + 64, row0 + 256);
          }
        }
        mbar_expect_tx(v_full, AL_TILE + (XT ? AL_XTILE : 0));
        tma_load_2d(&tm256, sV, v_full, 2 * D + col0, row0);
        tma_load_2d(&tm16, sV + 256 * 128, v_full, 2 * D + col0, row0 + 256);
        if (XT) {
          tma_load_2d(&tx256, sV + AL_TILE, v_full, 2 * D + col0 + 64, row0);
          tma_load_2d(&tx16, sV + AL_TILE + 256 * 32, v_full, 2 * D + col0 + 64, row0 + 256);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s256 = make_idesc_bf16(128, 256, 0, 0), idesc_s16 = make_idesc_bf16(128, 16, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1), idesc_pvx = make_idesc_bf16(128, 16, 0, 1);
    const uint32_t k_lo = at_desc_lo(smem_u32(sK), 16), k2_lo = at_desc_lo(smem_u32(sK) + 256 * 128, 16);
    const uint32_t v_lo = at_desc_lo(smem_u32(sV), 8192), p_lo = at_desc_lo(smem_u32(sP), 16);
    const uint32_t q_lo0 = at_desc_lo(smem_u32(sQ), 16);
    // extra tiles (dims 64..79): one 16-element k-step for S, a 16-column B operand for P V
    const uint32_t qx_lo0 = at_desc_lo(smem_u32(sQ) + AL_TILE, 16), kx_lo = at_desc_lo(smem_u32(sK) + AL_TILE, 16);
    const uint32_t kx2_lo = at_desc_lo(smem_u32(sK) + AL_TILE + 256 * 32, 16), vx_lo = at_desc_lo(smem_u32(sV) + AL_TILE, 16);
    auto issue_s = [&](int t) {  // S = Q_t K^T into TMEM columns 0-271
      const uint32_t q_lo = q_lo0 + t * 1024;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma_bf16(tmem_base, at_desc(q_lo + 2 * k), at_desc(k_lo + 2 * k), idesc_s256, k > 0 ? 1u : 0u);
          tc_mma_bf16(tmem_base + 256, at_desc(q_lo + 2 * k), at_desc(k2_lo + 2 * k), idesc_s16, k > 0 ? 1u : 0u);
        }
        if (XT) {
          const uint32_t qx_lo = qx_lo0 + t * 256;  // 128 rows x 32 B
          tc_mma_bf16(tmem_base, al_desc32(qx_lo), al_desc32(kx_lo), idesc_s256, 1u);
          tc_mma_bf16(tmem_base + 256, al_desc32(qx_lo), al_desc32(kx2_lo), idesc_s16, 1u);
        }
        tc_commit(s_full);
      }
      __syncwarp();
    };
    int it = 0, g = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      mbar_wait(qk_full, it & 1);
      tc_fence_after();
      issue_s(0);  // every softmax read of the previous item's last S finished before its p_full, waited below
      for (int t = 0; t < nqt; ++t, ++g) {
        mbar_wait(p_full, g & 1);                   // P_t in shared memory, S buffer free
        if (t == 0) mbar_wait(v_full, it & 1);
        if (g > 0) mbar_wait(t_free, (g - 1) & 1);  // O of the previous tile has been read
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < AL_KSTEPS; ++ks) {
            const uint64_t pd = at_desc(p_lo + (ks >> 2) * 1024 + (ks & 3) * 2);
            tc_mma_bf16(T_O, pd, at_desc(v_lo + ks * 128), idesc_pv, ks > 0 ? 1u : 0u);
            if (XT) tc_mma_bf16(T_O + 64, pd, al_desc32(vx_lo + ks * 32), idesc_pvx, ks > 0 ? 1u : 0u);  // 16 keys x 32 B
          }
          tc_commit(o_full);
          if (t == nqt - 1) tc_commit(in_empty);
        }
        __syncwarp();
        if (t + 1 < nqt) issue_s(t + 1);
      }
    }
  } else {
    const int wq = warp & 3, hh = (warp - 2) >> 2;
    const int r = wq * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const float c2 = p.scale * 1.4426950408889634f;
    const uint32_t prow = smem_u32(sP) + r * 128;
    const int sw = r & 7;
    float* xmax = reinterpret_cast<float*>(smem + MAP::OFF_X);  // [2][128]
    float* xsum = xmax + 256;
    const int j0 = hh ? 9 : 0, j1 = hh ? AL_KSTEPS : 9;
    int g = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      for (int t = 0; t < nqt; ++t, ++g) {
        const int q = t * 128 + r;
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        float m = -INFINITY;
        {
          uint32_t v[2][16];
          tmem_ld16(trow + j0 * 16, v[0]);
#pragma unroll
          for (int jj = 0; jj < 9; ++jj) {
            const int j = j0 + jj;
            if (j < j1) {
              tmem_ld_wait();
              if (j + 1 < j1) tmem_ld16(trow + (j + 1) * 16, v[(jj + 1) & 1]);
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (j * 16 + e < N) m = fmaxf(m, __uint_as_float(v[jj & 1][e]));
            }
          }
        }
        xmax[hh * 128 + r] = m;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
        m = fmaxf(m, xmax[(hh ^ 1) * 128 + r]);
        float l = 0.f;
        const float mc = m * c2;
        {
          uint32_t v[2][16];
          tmem_ld16(trow + j0 * 16, v[0]);
#pragma unroll
          for (int jj = 0; jj < 9; ++jj) {
            const int j = j0 + jj;
            if (j < j1) {
              tmem_ld_wait();
              if (j + 1 < j1) tmem_ld16(trow + (j + 1) * 16, v[(jj + 1) & 1]);
              float pv[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                pv[e] = (j * 16 + e < N) ? ex2_approx_ftz(fmaf(__uint_as_float(v[jj & 1][e]), c2, -mc)) : 0.f;
                l += pv[e];
              }
              const uint32_t blk = prow + (j >> 2) * 16384;
              const int ck = (j & 3) * 2;
              sts128(blk + ((ck ^ sw) << 4), pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                     pack_bf16x2(pv[6], pv[7]));
              sts128(blk + (((ck + 1) ^ sw) << 4), pack_bf16x2(pv[8], pv[9]), pack_bf16x2(pv[10], pv[11]),
                     pack_bf16x2(pv[12], pv[13]), pack_bf16x2(pv[14], pv[15]));
            }
          }
        }
        xsum[hh * 128 + r] = l;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (t == nqt - 1) {  // last tile of the item: tail rows, then release Q / K / V
          for (int tr = 0; tr < ntail; ++tr) {
            const int qrow = nfull * 128 + tr;
            const int tid = threadIdx.x - 64, tw = tid >> 5;
            float* ts = reinterpret_cast<float*>(smem + MAP::OFF_T);  // [272] raw scores, then probabilities
            float* red = ts + 272;                                  // [0,8) warp maxima, [8,16) warp sums
            float* part = ts + 320;                                 // [8][HD] partial outputs
            float qf[HD];
            if (XT) {  // dims 64..79: [row][32 B], 16-byte chunk c at (c ^ ((row >> 2) & 1))
              const uint32_t qa = smem_u32(sQ) + AL_TILE + qrow * 32;
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const uint4 u = lds128(qa + ((c ^ ((qrow >> 2) & 1)) << 4));
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(w4[e]);
                  qf[64 + c * 8 + 2 * e] = f.x, qf[64 + c * 8 + 2 * e + 1] = f.y;
                }
              }
            }
            {
              const uint32_t qa = smem_u32(sQ) + qrow * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 u = lds128(qa + ((c ^ (qrow & 7)) << 4));
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(w4[e]);
                  qf[c * 8 + 2 * e] = f.x, qf[c * 8 + 2 * e + 1] = f.y;
                }
              }
            }
            auto score = [&](int j) {
              const uint32_t ka = smem_u32(sK) + j * 128;
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 u = lds128(ka + ((c ^ (j & 7)) << 4));
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16x2(w4[e]);
                  acc = fmaf(qf[c * 8 + 2 * e], f.x, acc);
                  acc = fmaf(qf[c * 8 + 2 * e + 1], f.y, acc);
                }
              }
              if (XT) {
                const uint32_t kx = smem_u32(sK) + AL_TILE + j * 32;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const uint4 u = lds128(kx + ((c ^ ((j >> 2) & 1)) << 4));
                  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf16x2(w4[e]);
                    acc = fmaf(qf[64 + c * 8 + 2 * e], f.x, acc);
                    acc = fmaf(qf[64 + c * 8 + 2 * e + 1], f.y, acc);
                  }
                }
              }
              return acc;
            };
            const bool one = tid < N;  // keys 0 .. 255 (all valid when N >= 256: the 257-token case)
            const float s0 = one ? score(tid) : -INFINITY;
            const bool two = 256 + tid < N;
            const float s1 = two ? score(256 + tid) : -INFINITY;
            const float wm = warp_max(fmaxf(s0, s1));
            if (lane == 0) red[tw] = wm;
            asm volatile("bar.sync 5, 256;" ::: "memory");
            float tm = red[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) tm = fmaxf(tm, red[k]);
            const float tmc = tm * c2;
            const float p0 = one ? ex2_approx_ftz(fmaf(s0, c2, -tmc)) : 0.f;
            const float p1 = two ? ex2_approx_ftz(fmaf(s1, c2, -tmc)) : 0.f;
            ts[tid] = p0;
            if (two) ts[256 + tid] = p1;
            const float wsum = warp_sum(p0 + p1);
            if (lane == 0) red[8 + tw] = wsum;
            asm volatile("bar.sync 5, 256;" ::: "memory");
            float tl = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) tl += red[8 + k];
            // out[d] = sum_j p_j V[j][d]: warp tw takes keys tw, tw + 8, ...; lane = dimension pair
            float a0 = 0.f, a1 = 0.f, ax0 = 0.f, ax1 = 0.f;
            const uint32_t va = smem_u32(sV) + (lane & 3) * 4;
            const uint32_t vxa = smem_u32(sV) + AL_TILE + (lane & 3) * 4;  // lanes 0..7: dims 64 + 2 lane, 65 + 2 lane
            for (int j = tw; j < N; j += 8) {
              const float2 f = unpack_bf16x2(lds32(va + j * 128 + (((lane >> 2) ^ (j & 7)) << 4)));
              const float pj = ts[j];
              a0 = fmaf(pj, f.x, a0), a1 = fmaf(pj, f.y, a1);
              if (XT && lane < 8) {
                const float2 g = unpack_bf16x2(lds32(vxa + j * 32 + ((((lane >> 2) & 1) ^ ((j >> 2) & 1)) << 4)));
                ax0 = fmaf(pj, g.x, ax0), ax1 = fmaf(pj, g.y, ax1);
              }
            }
            if (XT && lane < 8) part[tw * HD + 64 + 2 * lane] = ax0, part[tw * HD + 64 + 2 * lane + 1] = ax1;
            part[tw * HD + 2 * lane] = a0, part[tw * HD + 2 * lane + 1] = a1;
            asm volatile("bar.sync 5, 256;" ::: "memory");
            if (tid < HD) {
              float o = 0.f;
#pragma unroll
              for (int k = 0; k < 8; ++k) o += part[k * HD + tid];
              p.out[(static_cast<long long>(b) * N + qrow) * D + h * HD + tid] = __float2bfloat16(o / tl);
              if (tid == 0 && p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * N + qrow] = tm * p.scale + __logf(tl);
            }
            asm volatile("bar.sync 5, 256;" ::: "memory");  // ts / red / part are reused by the next tail row / item
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(in_empty);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");  // the other half's row sum is in xsum
        mbar_wait(o_full, g & 1);
        tc_fence_after();
        // output columns of this half-row warp: head dim 64: 32 + 32; head dim 80: 48 + 32
        constexpr int OC1 = XT ? 48 : 32;
        const int oc0 = hh ? OC1 : 0, onch = (XT && hh == 0) ? 3 : 2;
        uint32_t o[3][16];
        tmem_ld16(trow + AL_ROWS + oc0, o[0]);
        tmem_ld16(trow + AL_ROWS + oc0 + 16, o[1]);
        if (XT && hh == 0) tmem_ld16(trow + AL_ROWS + 32, o[2]);
        tmem_ld_wait();
        l += xsum[(hh ^ 1) * 128 + r];
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_free);
        if (q < N) {
          const float inv = 1.f / l;
          bf16* dst = p.out + (static_cast<long long>(b) * N + q) * D + h * HD + oc0;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < onch) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                pk[e] = pack_bf16x2(__uint_as_float(o[j][2 * e]) * inv, __uint_as_float(o[j][2 * e + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + j * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dst + j * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          if (hh == 0 && p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * N + q] = m * p.scale + __logf(l);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct AttnBwdParams {
  const bf16* out;
  const bf16* dout;
  const float* lse;
  bf16* dqkv;
  int B, N, H, D;
  int items;
  float scale;
};

__device__ __forceinline__ void store16_bf16(bf16* dst, const uint32_t (&o)[16]) {
  uint4 lo, hi;
  lo.x = pack_bf16x2(__uint_as_float(o[0]), __uint_as_float(o[1]));
  lo.y = pack_bf16x2(__uint_as_float(o[2]), __uint_as_float(o[3]));
  lo.z = pack_bf16x2(__uint_as_float(o[4]), __uint_as_float(o[5]));
  lo.w = pack_bf16x2(__uint_as_float(o[6]), __uint_as_float(o[7]));
  hi.x = pack_bf16x2(__uint_as_float(o[8]), __uint_as_float(o[9]));
  hi.y = pack_bf16x2(__uint_as_float(o[10]), __uint_as_float(o[11]));
  hi.z = pack_bf16x2(__uint_as_float(o[12]), __uint_as_float(o[13]));
  hi.w = pack_bf16x2(__uint_as_float(o[14]), __uint_as_float(o[15]));
  *reinterpret_cast<uint4*>(dst) = lo;
  *reinterpret_cast<uint4*>(dst + 8) = hi;
}


// --------------------------------------------------------------------------------------------
// backward:  dV = P^T dO ; dP = dO V^T ; dS = P o (dP - delta) * scale ; dQ = dS K ; dK = dS^T Q, with no
// recomputation: every score element is computed, exponentiated and read from TMEM exactly once.
//
// Rows = KEYS.  For key tile u (128 rows) and query block j (64 queries; the last block has 16) -- one "sub-unit":
//     S^T = K_u Q_j^T ,  dP^T = V_u dO_j^T                      -> TMEM [128 x 64] each (two score sets, alternating)
//     P^T = exp2(S^T c - lse_q) ,  dS^T = P^T (dP^T scale - delta_q scale)   (registers -> bf16 smem blocks)
//     dV_u += P^T  dO_j          dK_u += dS^T Q_j               (A = the smem block, K-major; K = queries of j)
//     dQ[queries of blocks j-1, j] += (dS^T blocks j-1, j)^T K_u    after every second block: the two blocks are
//                                                 read IN PLACE as an MN-major A operand (M = queries, K = keys)
// The compute warps hold their 16-column chunk of S^T / dP^T in registers, so a score set is free for new MMAs as soon
// as it has been loaded; the MMA warp runs an event loop (non-blocking barrier tests): output MMAs of the oldest
// published sub-unit and score MMAs of the next one are issued as soon as their barriers allow.  The read-outs of
// dV_u / dK_u (per key tile) and dQ (per item) happen one sub-unit later, after that sub-unit's scores are in registers.
//   TMEM (512 columns): score set b: S^T 128 b, dP^T 128 b + 64 | dV_u 256-319 | dK_u 320-383 | dQ 384-511
//   smem: Q, dO double buffered across items (prefetch), K / V single (tile 0 and tile 1 halves refilled as soon
//   as their tile is done), P^T 2 blocks, dS^T 2 blocks, lse / delta double buffered.
// warps: 0 TMA, 1 MMA issue, 2 TMEM alloc, 2-3 delta / lse of the NEXT item, 4-19 compute (quarter = warp % 4, 16-column
// chunk = (warp - 4) / 4).
// --------------------------------------------------------------------------------------------
constexpr int AB_THREADS = 640;
constexpr int AB_NCW = 16;
constexpr int AB_KB = 16384;  // one operand block [128 rows x 64 bf16], 128-byte swizzle
constexpr int AB_OFF_Q = 0;
constexpr int AB_OFF_DO = 2 * AT_TILE;
constexpr int AB_OFF_K = 4 * AT_TILE;
constexpr int AB_OFF_V = 5 * AT_TILE;
constexpr int AB_OFF_P = 6 * AT_TILE;
constexpr int AB_OFF_DS = AB_OFF_P + 2 * AB_KB;
constexpr int AB_OFF_VEC = AB_OFF_DS + 2 * AB_KB;
constexpr int AB_OFF_BAR = AB_OFF_VEC + 4 * AT_ROWS * 4;
constexpr int AB_SMEM = AB_OFF_BAR + 256 + 1024;
static_assert(AB_SMEM <= 232448, "attention backward: shared memory");
static_assert((6 * AT_TILE) % 1024 == 0, "operand blocks must stay 1024-byte aligned");

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                    const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw_at[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at) + 1023) & ~uintptr_t(1023));
  float* sLseAll = reinterpret_cast<float*>(smem + AB_OFF_VEC);  // [2][AT_ROWS] lse * log2(e); +inf beyond N
  float* sDelAll = sLseAll + 2 * AT_ROWS;                        // [2][AT_ROWS] rowsum(dO o O)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AB_OFF_BAR);
  uint64_t* qdo_full = bars + 0;    // [2]
  uint64_t* qdo_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;     // [2] key tile 0 / 1
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* sd_full = bars + 8;     // [2] score sets
  uint64_t* sd_free = bars + 10;    // [2]
  // [2], by sub-unit parity: the scores run two sub-units ahead, so a fast warp can publish sub-unit s+1 before a
  // slow one has published s -- with one barrier its arrival would complete the wrong phase
  uint64_t* pb_full = bars + 12;
  uint64_t* dsq_done = bars + 14;
  uint64_t* acc_full = bars + 15;
  uint64_t* acc_free = bars + 16;
  uint64_t* dq_full = bars + 17;
  uint64_t* dq_free = bars + 18;
  uint64_t* del_full = bars + 19;   // [2]
  uint64_t* del_empty = bars + 21;  // [2]
  uint64_t* slot_free = bars + 23;  // [2] the output MMAs of sub-unit n (readers of the P^T / dS^T blocks n & 1) are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmK1);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&sd_full[i], 1);
      mbar_init(&sd_free[i], AB_NCW);
      mbar_init(&pb_full[i], AB_NCW);
      mbar_init(&slot_free[i], 1);
      mbar_init(&del_full[i], 64);
      mbar_init(&del_empty[i], AB_NCW);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, AB_NCW);
    mbar_init(dsq_done, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, AB_NCW);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int D = p.D;
  // score set b: S^T at 128 b, dP^T at 128 b + 64 | dV_u 256-319 | dK_u 320-383 | dQ 384-511
  const uint32_t T_SET = tmem_base, T_DV = tmem_base + 256, T_DK = tmem_base + 320, T_DQ = tmem_base + 384;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * p.N, col0 = h * AT_HD;
        const int slot = it & 1;
        if (it >= 2) mbar_wait(&qdo_empty[slot], ((it >> 1) - 1) & 1);
        mbar_expect_tx(&qdo_full[slot], 2 * AT_TILE);
        tma_load_2d(&tmQ, smem + AB_OFF_Q + slot * AT_TILE, &qdo_full[slot], col0, row0);
        tma_load_2d(&tmDO, smem + AB_OFF_DO + slot * AT_TILE, &qdo_full[slot], col0, row0);
        if (it >= 1) mbar_wait(&kv_empty[0], (it - 1) & 1);
        mbar_expect_tx(&kv_full[0], 2 * AB_KB);
        tma_load_2d(&tmK0, smem + AB_OFF_K, &kv_full[0], D + col0, row0);
        tma_load_2d(&tmK0, smem + AB_OFF_V, &kv_full[0], 2 * D + col0, row0);
        if (it >= 1) mbar_wait(&kv_empty[1], (it - 1) & 1);
        mbar_expect_tx(&kv_full[1], 2 * (AT_TILE - AB_KB));
        tma_load_2d(&tmK1, smem + AB_OFF_K + AB_KB, &kv_full[1], D + col0, row0 + 128);
        tma_load_2d(&tmK1, smem + AB_OFF_V + AB_KB, &kv_full[1], 2 * D + col0, row0 + 128);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    {
      const uint32_t idesc_s64 = make_idesc_bf16(128, 64, 0, 0);   // scores: A (keys) K-major, B (queries) K-major
      const uint32_t idesc_s16 = make_idesc_bf16(128, 16, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, AT_HD, 0, 1);   // dV / dK: A K-major (queries), B MN-major
      const uint32_t idesc_q = make_idesc_bf16(128, AT_HD, 1, 1);   // dQ: A MN-major (dS^T blocks), B MN-major (K_u)
      const uint32_t aK = smem_u32(smem + AB_OFF_K), aV = smem_u32(smem + AB_OFF_V);
      const uint32_t aP = smem_u32(smem + AB_OFF_P), aDS = smem_u32(smem + AB_OFF_DS);
      const uint32_t ds_mn_lo = at_desc_lo(aDS, AB_KB);  // both dS^T blocks as one MN-major A operand (atoms AB_KB apart)
      int n_items = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) ++n_items;
      const int total = 8 * n_items;  // sub-unit n: item n >> 3, key tile (n >> 2) & 1, query block n & 3
      int n_sub = 0;                  // sub-units whose scores have been issued (into score set n_sub & 1)
      int n_out = 0;                  // sub-units whose output MMAs have been issued
      // non-blocking barrier test by lane 0, broadcast: the control flow stays warp-uniform
      auto ready = [&](uint64_t* bar, uint32_t parity) -> bool {
        uint32_t ok = 0;
        if (lane == 0) ok = mbar_test_wait(bar, parity) ? 1u : 0u;
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
      };
      auto issue_scores = [&](int n) {
        const int it = n >> 3, u = (n >> 2) & 1, j = n & 3, slot = it & 1, set = n & 1;
        tc_fence_after();
        const uint32_t idesc = j < 3 ? idesc_s64 : idesc_s16;
        const uint32_t k_lo = at_desc_lo(aK + u * AB_KB, 16), v_lo = at_desc_lo(aV + u * AB_KB, 16);
        const uint32_t q_lo = at_desc_lo(smem_u32(smem + AB_OFF_Q + slot * AT_TILE) + j * 8192, 16);
        const uint32_t o_lo = at_desc_lo(smem_u32(smem + AB_OFF_DO + slot * AT_TILE) + j * 8192, 16);
        const uint32_t ts = T_SET + set * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(ts, at_desc(k_lo + 2 * k), at_desc(q_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(ts + 64, at_desc(v_lo + 2 * k), at_desc(o_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
          tc_commit(&sd_full[set]);
        }
        __syncwarp();
      };
      auto issue_out = [&](int n) {
        const int it = n >> 3, u = (n >> 2) & 1, j = n & 3, slot = it & 1;
        tc_fence_after();
        const uint32_t kmn_lo = at_desc_lo(aK + u * AB_KB, 8192);
        const uint32_t omn_lo = at_desc_lo(smem_u32(smem + AB_OFF_DO + slot * AT_TILE) + j * 8192, 8192);
        const uint32_t qmn_lo = at_desc_lo(smem_u32(smem + AB_OFF_Q + slot * AT_TILE) + j * 8192, 8192);
        const uint32_t pb_lo = at_desc_lo(aP + (n & 1) * AB_KB, 16);
        const uint32_t db_lo = at_desc_lo(aDS + (n & 1) * AB_KB, 16);
        if (elect_one_sync()) {
          if (j & 1) {  // dQ[tile m] (+)= (dS^T blocks j-1, j)^T K_u : M = 128 queries, K = 128 keys
            const uint32_t dq = T_DQ + (j >> 1) * 64;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              tc_mma_bf16(dq, at_desc(ds_mn_lo + kk * 128), at_desc(kmn_lo + kk * 128), idesc_q, (u > 0 || kk > 0) ? 1u : 0u);
            tc_commit(dsq_done);
          }
          if (j < 3) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma_bf16(T_DV, at_desc(pb_lo + 2 * kk), at_desc(omn_lo + kk * 128), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma_bf16(T_DK, at_desc(db_lo + 2 * kk), at_desc(qmn_lo + kk * 128), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          } else {  // the last query block holds 16 queries: one k-step
            tc_mma_bf16(T_DV, at_desc(pb_lo), at_desc(omn_lo), idesc_o, 1u);
            tc_mma_bf16(T_DK, at_desc(db_lo), at_desc(qmn_lo), idesc_o, 1u);
            tc_commit(acc_full);
            tc_commit(&kv_empty[u]);
            if (u == 1) {
              tc_commit(dq_full);
              tc_commit(&qdo_empty[slot]);
            }
          }
          tc_commit(&slot_free[n & 1]);
        }
        __syncwarp();
      };
      // Event loop: whichever is possible next -- the output MMAs of the oldest published sub-unit (they unblock the
      // compute warps' next block writes and the dQ chain) or the scores of the next sub-unit (at most two score
      // sets in flight) -- is issued as soon as its barriers allow; nothing is waited for in a fixed order.
      while (n_out < total) {
        if (n_out < n_sub) {
          const int n = n_out, it = n >> 3, u = (n >> 2) & 1, j = n & 3;
          bool ok = ready(&pb_full[n & 1], (n >> 1) & 1);
          if (ok && u == 0 && j == 1 && it > 0) ok = ready(dq_free, (it - 1) & 1);  // previous item's dQ read out
          if (ok && j == 0 && n >= 4) ok = ready(acc_free, ((n >> 2) - 1) & 1);      // previous key tile read out
          if (ok) {
            issue_out(n);
            ++n_out;
          }
        }
        if (n_sub < total) {  // bounded by the two score sets: sub-unit n_sub - 2 must be in registers
          const int n = n_sub, it = n >> 3, u = (n >> 2) & 1, j = n & 3;
          bool ok = true;
          if (j == 0 && u == 0) ok = ready(&qdo_full[it & 1], (it >> 1) & 1);
          if (ok && j == 0) ok = ready(&kv_full[u], it & 1);
          if (ok && n >= 2) ok = ready(&sd_free[n & 1], ((n >> 1) - 1) & 1);  // sub-unit n-2 is in registers
          if (ok) {
            issue_scores(n);
            ++n_sub;
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ---- delta_i = sum_d dO[i,d] O[i,d] and lse_i of the NEXT item: two threads per row, straight from global ----
    const int t = threadIdx.x - 64;  // 0..63
    int buf = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      mbar_wait(&del_empty[buf], ph ^ 1);
      float* sLse = sLseAll + buf * AT_ROWS;
      float* sDel = sDelAll + buf * AT_ROWS;
#pragma unroll 1
      for (int w0 = 0; w0 < 2 * AT_ROWS; w0 += 64) {
        const int wi = w0 + t;
        const int row = wi >> 1, half = wi & 1;
        float acc = 0.f;
        if (wi < 2 * AT_ROWS && row < p.N) {
          const long long off = (static_cast<long long>(b) * p.N + row) * D + h * AT_HD + half * 32;
          uint4 a[4], d[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            a[c] = *reinterpret_cast<const uint4*>(p.out + off + c * 8);
            d[c] = *reinterpret_cast<const uint4*>(p.dout + off + c * 8);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a[c]);
            const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d[c]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16x2(pa[e]), y = unpack_bf16x2(pd[e]);
              acc += x.x * y.x + x.y * y.y;
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (wi < 2 * AT_ROWS && half == 0) {
          sDel[row] = acc * p.scale;  // pre-scaled: dS = p (dp scale - delta scale)
          sLse[row] = row < p.N ? p.lse[(static_cast<long long>(b) * p.H + h) * p.N + row] * 1.4426950408889634f
                                : INFINITY;
        }
      }
      mbar_arrive(&del_full[buf]);
      buf ^= 1;
      if (buf == 0) ph ^= 1;
    }
  } else {
    // ============================== compute warps ==============================
    const int cw = warp - 4;
    const int wq = cw & 3;  // TMEM lane quarter (== warp % 4)
    const int cg = cw >> 2; // 16-column chunk of the 64-column block
    const int r = wq * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
    const float c2 = p.scale * 1.4426950408889634f;
    uint8_t* prow_p = smem + AB_OFF_P + r * 128;
    uint8_t* prow_d = smem + AB_OFF_DS + r * 128;
    const int sw = r & 7;
    uint32_t dph = 0;
    int dbuf = 0;
    int n_sub = 0;
    // read-outs: dV_u / dK_u of a finished key tile, dQ of a finished item
    auto drain_acc = [&](int u, int ditem, uint32_t parity) {
      mbar_wait(acc_full, parity);
      tc_fence_after();
      uint32_t ov[16], ok_[16];
      tmem_ld16(T_DV + lane_sel + cg * 16, ov);
      tmem_ld16(T_DK + lane_sel + cg * 16, ok_);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
      const int b = ditem / p.H, h = ditem - b * p.H;
      const int key = u * 128 + r;
      if (key < p.N) {
        bf16* base = p.dqkv + (static_cast<long long>(b) * p.N + key) * 3 * D + h * AT_HD + cg * 16;
        store16_bf16(base + D, ok_);
        store16_bf16(base + 2 * D, ov);
      }
    };
    auto drain_dq = [&](int ditem, uint32_t parity) {
      mbar_wait(dq_full, parity);
      tc_fence_after();
      uint32_t o0[16], o1[16];
      tmem_ld16(T_DQ + lane_sel + cg * 16, o0);
      tmem_ld16(T_DQ + 64 + lane_sel + cg * 16, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      const int b = ditem / p.H, h = ditem - b * p.H;
      bf16* base = p.dqkv + static_cast<long long>(b) * p.N * 3 * D + h * AT_HD + cg * 16;
      if (r < p.N) store16_bf16(base + static_cast<long long>(r) * 3 * D, o0);
      if (128 + r < p.N) store16_bf16(base + static_cast<long long>(128 + r) * 3 * D, o1);
    };
    const float sc = p.scale;
    int it = 0, prev_item = -1;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      mbar_wait(&del_full[dbuf], dph);
      const uint32_t sLse = smem_u32(sLseAll + dbuf * AT_ROWS);
      const uint32_t sDel = smem_u32(sDelAll + dbuf * AT_ROWS);
      for (int u = 0; u < 2; ++u) {
        const bool kvalid = (u * 128 + r) < p.N;
        const bool rows_ok = __all_sync(0xffffffffu, kvalid);
        for (int j = 0; j < 4; ++j, ++n_sub) {
          const int set = n_sub & 1;
          mbar_wait(&sd_full[set], (n_sub >> 1) & 1);
          tc_fence_after();
          const bool active = (j < 3) || (cg == 0);
          uint32_t s[16], dp[16];
          if (active) {
            const uint32_t ts = T_SET + set * 128 + lane_sel + cg * 16;
            tmem_ld16(ts, s);
            tmem_ld16(ts + 64, dp);
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sd_free[set]);
          // the outputs of the key tile (item) that finished two sub-units ago: their last MMAs were issued when that
          // tile's last block was published, they are complete by now, and the score registers of this sub-unit are
          // already loaded -- the read-out costs no waiting (the first output MMAs of the new tile wait for it)
          if (j == 1) {
            if (u == 1) {
              drain_acc(0, item, 0u);  // key tile 2 it: even
            } else if (it > 0) {
              drain_acc(1, prev_item, 1u);
              drain_dq(prev_item, (it - 1) & 1);
            }
          }
          uint32_t pkp[8], pkd[8];  // P^T / dS^T of this thread's 16 columns, packed bf16
          const int q0 = j * 64 + cg * 16;
          if (active) {
            const bool mask = !rows_ok || j == 3;  // padding keys (rows) / padding queries (columns): exact zeros
            // p = exp2(s c - lse), dS = p (dp scale - delta scale): 4 instructions per element (+ bf16 packing)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int qq = q0 + e4 * 4;
              const uint4 l4 = lds128(sLse + qq * 4);
              const uint4 d4 = lds128(sDel + qq * 4);
              const float ls[4] = {__uint_as_float(l4.x), __uint_as_float(l4.y), __uint_as_float(l4.z), __uint_as_float(l4.w)};
              const float dl[4] = {__uint_as_float(d4.x), __uint_as_float(d4.y), __uint_as_float(d4.z), __uint_as_float(d4.w)};
              float pe[4], de[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = e4 * 4 + e;
                pe[e] = ex2_approx_ftz(fmaf(__uint_as_float(s[col]), c2, -ls[e]));  // lse = +inf beyond N: 0
                de[e] = pe[e] * fmaf(__uint_as_float(dp[col]), sc, -dl[e]);
              }
              if (mask) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const bool ok = kvalid && (qq + e) < p.N;
                  pe[e] = ok ? pe[e] : 0.f;
                  de[e] = ok ? de[e] : 0.f;
                }
              }
              pkp[2 * e4] = pack_bf16x2(pe[0], pe[1]), pkp[2 * e4 + 1] = pack_bf16x2(pe[2], pe[3]);
              pkd[2 * e4] = pack_bf16x2(de[0], de[1]), pkd[2 * e4 + 1] = pack_bf16x2(de[2], de[3]);
            }
          }
          // dS^T blocks are also the A operand of the dQ MMAs issued after every odd block: the block written at
          // an even j was last read by the group issued after j-1, which is NOT ordered before this block's score
          // MMAs -> wait for it explicitly, after the math (odd j: ordered by the score commit)
          if ((j & 1) == 0 && n_sub >= 2) mbar_wait(dsq_done, ((n_sub >> 1) - 1) & 1);  // group of sub-unit n_sub - 1
          // the blocks of this parity were last read by the output MMAs of sub-unit n_sub - 2
          if (n_sub >= 2) mbar_wait(&slot_free[set], ((n_sub >> 1) - 1) & 1);
          if (active) {
            const uint32_t bp = smem_u32(prow_p) + set * AB_KB, bd = smem_u32(prow_d) + set * AB_KB;
            const int ck = cg * 2;
            sts128(bp + ((ck ^ sw) << 4), pkp[0], pkp[1], pkp[2], pkp[3]);
            sts128(bp + (((ck + 1) ^ sw) << 4), pkp[4], pkp[5], pkp[6], pkp[7]);
            sts128(bd + ((ck ^ sw) << 4), pkd[0], pkd[1], pkd[2], pkd[3]);
            sts128(bd + (((ck + 1) ^ sw) << 4), pkd[4], pkd[5], pkd[6], pkd[7]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pb_full[set]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&del_empty[dbuf]);
      dbuf ^= 1;
      if (dbuf == 0) dph ^= 1;
      prev_item = item;
    }
    if (prev_item >= 0) {
      drain_acc(1, prev_item, 1u);
      drain_dq(prev_item, (it - 1) & 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace theia

using namespace theia;

static int encode_qkv_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, int box_rows = AT_ROWS) {
  uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
  uint64_t strides[1] = {static_cast<uint64_t>(cols) * 2};
  uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
  return encode_tensor_map(tm, ptr, 2, dims, strides, box);
}

// extra tiles of head dim 80: columns 64..79 of a head, 32-byte swizzle
static int encode_qkv_map_x(CUtensorMap* tm, const void* ptr, long long rows, long long cols, int box_rows) {
  uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
  uint64_t strides[1] = {static_cast<uint64_t>(cols) * 2};
  uint32_t box[2] = {16, static_cast<uint32_t>(box_rows)};
  return encode_tensor_map(tm, ptr, 2, dims, strides, box, nullptr, 32);
}

template <int HD>
static int launch_attn_long(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream) {
  const int D = H * HD;
  CUtensorMap t256, t16, x256, x16;
  int rc2 = encode_qkv_map(&t256, qkv, static_cast<long long>(B) * N, 3LL * D, 256);
  if (rc2) return rc2;
  rc2 = encode_qkv_map(&t16, qkv, static_cast<long long>(B) * N, 3LL * D, 16);
  if (rc2) return rc2;
  x256 = t256, x16 = t16;
  if (HD > 64) {
    rc2 = encode_qkv_map_x(&x256, qkv, static_cast<long long>(B) * N, 3LL * D, 256);
    if (rc2) return rc2;
    rc2 = encode_qkv_map_x(&x16, qkv, static_cast<long long>(B) * N, 3LL * D, 16);
    if (rc2) return rc2;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_long_kernel<257, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AlMap<HD>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_tc_fwd_long_kernel<0, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, AlMap<HD>::SMEM);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn_tc fwd long attr: %s", cudaGetErrorString(e));
  }
  AttnFwdParams pl;
  pl.out = static_cast<bf16*>(out);
  pl.lse = lse;
  pl.B = B, pl.N = N, pl.H = H, pl.D = D;
  pl.items = B * H;
  pl.scale = HD == 64 ? 0.125f : 1.0f / sqrtf(static_cast<float>(HD));
  const int gridl = pl.items < num_sms() ? pl.items : num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 257)  // CLS + 16 x 16 patches: compile-time length (the key masks fold away)
    attn_tc_fwd_long_kernel<257, HD><<<gridl, AL_THREADS, AlMap<HD>::SMEM, st>>>(t256, t16, x256, x16, pl);
  else
    attn_tc_fwd_long_kernel<0, HD><<<gridl, AL_THREADS, AlMap<HD>::SMEM, st>>>(t256, t16, x256, x16, pl);
  THEIA_CHECK_LAUNCH("attention_tc_fwd_long");
  return THEIA_OK;
}

// head dim 80 (ViT-H/14 teacher): forward only, through the long-sequence kernel whatever the length (<= 272 tokens)
extern "C" int theia_attention_fwd_hd80(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream) {
  if (N > AL_ROWS || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention (head dim 80): %d tokens outside [1, %d]", N, AL_ROWS);
  return launch_attn_long<80>(qkv, out, lse, B, N, H, stream);
}

extern "C" int theia_attention_tc_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream) {
  if (N > AL_ROWS || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention: sequence length %d > %d", N, AL_ROWS);
  const int D = H * AT_HD;
  if (N > AT_ROWS) return launch_attn_long<64>(qkv, out, lse, B, N, H, stream);  // 209 .. 272 tokens (ViT-L/14 teachers)
  CUtensorMap tm;
  int rc = encode_qkv_map(&tm, qkv, static_cast<long long>(B) * N, 3LL * D);
  if (rc) return rc;
  AttnFwdParams p;
  p.out = static_cast<bf16*>(out);
  p.lse = lse;
  p.B = B, p.N = N, p.H = H, p.D = D;
  p.items = B * H;
  p.scale = 0.125f;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  {
    static PerDeviceOnce attr_once;
    if (attr_once.first_use()) {
      cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd2_kernel<197>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM);
      if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn_tc fwd2 attr: %s", cudaGetErrorString(e));
    }
    // 197 tokens (DeiT: CLS + 196 patches) gets the compile-time sequence length; DeiTNoCLS / DeiTReg run the generic one
    if (N == 197) attn_tc_fwd2_kernel<197><<<grid, AF_THREADS, AF_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, p);
    else attn_tc_fwd2_kernel<0><<<grid, AF_THREADS, AF_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, p);
    THEIA_CHECK_LAUNCH("attention_tc_fwd2");
    return THEIA_OK;
  }
}

extern "C" int theia_attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                      int B, int N, int H, void* stream) {
  if (N > AT_ROWS || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention: sequence length %d > %d", N, AT_ROWS);
  const int D = H * AT_HD;
  CUtensorMap tmq, tmd;
  int rc = encode_qkv_map(&tmq, qkv, static_cast<long long>(B) * N, 3LL * D);
  if (rc) return rc;
  rc = encode_qkv_map(&tmd, dout, static_cast<long long>(B) * N, D);
  if (rc) return rc;
  AttnBwdParams p;
  p.out = static_cast<const bf16*>(out);
  p.dout = static_cast<const bf16*>(dout);
  p.lse = lse;
  p.dqkv = static_cast<bf16*>(dqkv);
  p.B = B, p.N = N, p.H = H, p.D = D;
  p.items = B * H;
  p.scale = 0.125f;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  {
    CUtensorMap tmk0, tmk1;
    rc = encode_qkv_map(&tmk0, qkv, static_cast<long long>(B) * N, 3LL * D, 128);
    if (rc) return rc;
    rc = encode_qkv_map(&tmk1, qkv, static_cast<long long>(B) * N, 3LL * D, AT_ROWS - 128);
    if (rc) return rc;
    static PerDeviceOnce attr_once;
    if (attr_once.first_use()) {
      cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
      if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn_tc bwd2 attr: %s", cudaGetErrorString(e));
    }
    attn_tc_bwd2_kernel<<<grid, AB_THREADS, AB_SMEM, static_cast<cudaStream_t>(stream)>>>(tmq, tmk0, tmk1, tmd, p);
    THEIA_CHECK_LAUNCH("attention_tc_bwd2");
    return THEIA_OK;
  }
}
