// tcgen05 self-attention for the 197-token ViT sequence (head_dim 64), forward and backward.
// Replaces hf:models/vit/modeling_vit.py:171-196,232-246 (SDPA / eager attention) and its autograd.
//
// One persistent CTA per SM loops over (image, head) items.  Q/K/V (and dO) of an item are TMA-loaded
// as 128-byte-swizzled [rows x 64] tiles straight from the fused qkv activation [B*N, 3*D]; the score
// matrices live in TMEM; the probabilities go back to shared memory as the K-major A operand of the
// second MMA; V (and K, Q, dO in backward) are read as MN-major B operands in place, so nothing is ever
// transposed or copied.  Softmax is single pass (the whole 197-key row is in TMEM: no online rescaling).
//
//   forward, per item:    S_t = Q_t K^T  (2 query tiles of 128 rows, N = 208 keys)   -> TMEM
//                         P_t = exp2((S_t - rowmax) * c)  (bf16, smem),  l = rowsum     (1 thread / row)
//                         O_t = P_t V                                                   -> TMEM -> out, lse
//   warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 softmax/epilogue of tile 0, 8-11 of tile 1.
#include <cuda.h>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

constexpr int AT_HD = 64;
constexpr int AT_ROWS = 208;                  // rows per operand tile (197 padded to a multiple of 16)
constexpr int AT_TILE = AT_ROWS * 128;        // 26624 B, a multiple of the 1024-B swizzle atom
constexpr int AT_PBUF = 128 * 128 * 4;        // P tile: 4 K-blocks of [128 rows x 64 keys] = 64 KB
constexpr int AT_KSTEPS = AT_ROWS / 16;       // 13 MMA K-steps over the keys
constexpr int ATF_THREADS = 384;
constexpr int ATF_SMEM = 3 * AT_TILE + 2 * AT_PBUF + 256 + 1024;

struct AttnFwdParams {
  bf16* out;
  float* lse;
  int B, N, H, D;
  int items;
  float scale;
};

__global__ void __launch_bounds__(ATF_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw_at[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 2 * AT_TILE;
  uint8_t* sP[2] = {smem + 3 * AT_TILE, smem + 3 * AT_TILE + AT_PBUF};
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * AT_TILE + 2 * AT_PBUF);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* qk_empty = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;   // [2]
  uint64_t* p_full = bars + 6;   // [2]
  uint64_t* o_full = bars + 8;   // [2]
  uint64_t* t_free = bars + 10;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(qk_empty, 1);
    mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_full[t], 1);
      mbar_init(&t_free[t], 128);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * p.N, col0 = h * AT_HD;
        if (it > 0) mbar_wait(qk_empty, (it - 1) & 1);
        mbar_expect_tx(qk_full, 2 * AT_TILE);
        tma_load_2d(&tmQKV, sQ, qk_full, col0, row0);
        tma_load_2d(&tmQKV, sK, qk_full, D + col0, row0);
        if (it > 0) mbar_wait(v_empty, (it - 1) & 1);
        mbar_expect_tx(v_full, AT_TILE);
        tma_load_2d(&tmQKV, sV, v_full, 2 * D + col0, row0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, AT_ROWS, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, AT_HD, 0, 1);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      int it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        mbar_wait(qk_full, it & 1);
        if (it > 0) {
          mbar_wait(&t_free[0], (it - 1) & 1);
          mbar_wait(&t_free[1], (it - 1) & 1);
        }
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base + t * 256, make_smem_desc(aQ + t * 16384 + k * 32, 16, 1024),
                        make_smem_desc(aK + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
          tc_commit(&s_full[t]);
        }
        tc_commit(qk_empty);
        mbar_wait(v_full, it & 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], it & 1);
          tc_fence_after();
          const uint32_t aP = smem_u32(sP[t]);
#pragma unroll
          for (int ks = 0; ks < AT_KSTEPS; ++ks)
            tc_mma_bf16(tmem_base + t * 256, make_smem_desc(aP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                        make_smem_desc(aV + ks * 2048, 8192, 1024), idesc_pv, ks > 0 ? 1u : 0u);
          tc_commit(&o_full[t]);
        }
        tc_commit(v_empty);
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;       // query tile
    const int wq = warp & 3;             // TMEM lane quarter
    const int r = wq * 32 + lane;        // row within the tile
    const int q = t * 128 + r;           // query index
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + t * 256;
    const float c2 = p.scale * 1.4426950408889634f;
    uint8_t* prow = sP[t] + r * 128;
    const int sw = r & 7;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int b = item / p.H, h = item - b * p.H;
      mbar_wait(&s_full[t], it & 1);
      tc_fence_after();
      // Both passes stream the 13 sixteen-column chunks of the row with the TMEM load of chunk j+1 in flight while
      // chunk j is processed (fully unrolled: the two register buffers alternate at compile time).
      // pass 1: row max over the valid keys
      float m = -INFINITY;
      {
        uint32_t v[2][16];
        tmem_ld16(trow, v[0]);
#pragma unroll
        for (int j = 0; j < AT_KSTEPS; ++j) {
          tmem_ld_wait();
          if (j + 1 < AT_KSTEPS) tmem_ld16(trow + (j + 1) * 16, v[(j + 1) & 1]);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float s = __uint_as_float(v[j & 1][e]);
            if (j * 16 + e < p.N) m = fmaxf(m, s);
          }
        }
      }
      // pass 2: p = exp2((s - m) c2), row sum, bf16 P tile (K-major, 128-byte swizzle)
      float l = 0.f;
      const float mc = m * c2;
      {
        uint32_t v[2][16];
        tmem_ld16(trow, v[0]);
#pragma unroll
        for (int j = 0; j < AT_KSTEPS; ++j) {
          tmem_ld_wait();
          if (j + 1 < AT_KSTEPS) tmem_ld16(trow + (j + 1) * 16, v[(j + 1) & 1]);
          float pv[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float s = __uint_as_float(v[j & 1][e]);
            pv[e] = (j * 16 + e < p.N) ? ex2_approx_ftz(fmaf(s, c2, -mc)) : 0.f;
            l += pv[e];
          }
          uint4 lo, hi;
          lo.x = pack_bf16x2(pv[0], pv[1]), lo.y = pack_bf16x2(pv[2], pv[3]);
          lo.z = pack_bf16x2(pv[4], pv[5]), lo.w = pack_bf16x2(pv[6], pv[7]);
          hi.x = pack_bf16x2(pv[8], pv[9]), hi.y = pack_bf16x2(pv[10], pv[11]);
          hi.z = pack_bf16x2(pv[12], pv[13]), hi.w = pack_bf16x2(pv[14], pv[15]);
          uint8_t* blk = prow + (j >> 2) * 16384;
          const int ck = (j & 3) * 2;
          *reinterpret_cast<uint4*>(blk + (((ck) ^ sw) << 4)) = lo;
          *reinterpret_cast<uint4*>(blk + (((ck + 1) ^ sw) << 4)) = hi;
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      // O_t = P_t V (aliases the first 64 columns of S_t)
      mbar_wait(&o_full[t], it & 1);
      tc_fence_after();
      uint32_t o[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld16(trow + j * 16, o[j]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&t_free[t]);
      if (q < p.N) {
        const float inv = 1.f / l;
        bf16* dst = p.out + (static_cast<long long>(b) * p.N + q) * D + h * AT_HD;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
        if (p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.N + q] = m * p.scale + __logf(l);
      }
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
// backward:  dV = P^T dO ; dP = dO V^T ; dS = P o (dP - delta) * scale ; dQ = dS K ; dK = dS^T Q
//   stage A_t (query tile t): S_t = Q_t K^T, dP_t = dO_t V^T -> TMEM; dS_t (bf16, smem) ; dQ_t = dS_t K
//   stage B_u (key tile u):   S^T_u = K_u Q^T, dP^T_u = V_u dO^T -> TMEM; P^T_u (smem) -> dV_u = P^T_u dO;
//                             dS^T_u (kept packed in registers, then smem) -> dK_u = dS^T_u Q
// P and dS are elementwise given the saved row log-sum-exp and delta = rowsum(dO o O), so the 208 columns
// of a row are split across 4 warps (16 compute warps); no atomics, S and dP are recomputed once (7 GEMM
// units instead of 5) so that nothing but the 64-KB operand buffer leaves TMEM.
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 2-3 delta / lse of the NEXT item (two-deep smem ring,
// so the global-load latency of that prologue is off the critical path: 0.486 -> 0.42 ms per layer), 4-19 compute.
// (Issuing the next stage's score MMAs before the outputs of the previous one are drained -- outputs in separate
// TMEM columns -- was measured SLOWER, 0.54 ms, and is not done.)
// --------------------------------------------------------------------------------------------
constexpr int ATB_THREADS = 640;
constexpr int ATB_COMPUTE = 512;
constexpr int ATB_SMEM = 4 * AT_TILE + AT_PBUF + 4 * AT_ROWS * 4 + 256 + 1024;  // lse / delta double buffered

struct AttnBwdParams {
  const bf16* out;
  const bf16* dout;
  const float* lse;
  bf16* dqkv;
  int B, N, H, D;
  int items;
  float scale;
};

__device__ __forceinline__ void pack16_store(uint8_t* prow, int j, int sw, const float (&v)[16]) {
  uint4 lo, hi;
  lo.x = pack_bf16x2(v[0], v[1]), lo.y = pack_bf16x2(v[2], v[3]);
  lo.z = pack_bf16x2(v[4], v[5]), lo.w = pack_bf16x2(v[6], v[7]);
  hi.x = pack_bf16x2(v[8], v[9]), hi.y = pack_bf16x2(v[10], v[11]);
  hi.z = pack_bf16x2(v[12], v[13]), hi.w = pack_bf16x2(v[14], v[15]);
  uint8_t* blk = prow + (j >> 2) * 16384;
  const int ck = (j & 3) * 2;
  *reinterpret_cast<uint4*>(blk + ((ck ^ sw) << 4)) = lo;
  *reinterpret_cast<uint4*>(blk + (((ck + 1) ^ sw) << 4)) = hi;
}

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

__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw_at[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_at) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + AT_TILE;
  uint8_t* sV = smem + 2 * AT_TILE;
  uint8_t* sdO = smem + 3 * AT_TILE;
  uint8_t* sPB = smem + 4 * AT_TILE;
  float* sLseAll = reinterpret_cast<float*>(smem + 4 * AT_TILE + AT_PBUF);  // [2][AT_ROWS] lse * log2(e); +inf beyond N
  float* sDelAll = sLseAll + 2 * AT_ROWS;                                   // [2][AT_ROWS] rowsum(dO o O)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelAll + 2 * AT_ROWS);
  uint64_t* in_full = bars + 0;
  uint64_t* in_empty = bars + 1;
  uint64_t* sd_full = bars + 2;
  uint64_t* pb_full = bars + 3;
  uint64_t* acc_full = bars + 4;
  uint64_t* tm_free = bars + 5;
  uint64_t* del_full = bars + 6;   // [2]
  uint64_t* del_empty = bars + 8;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    mbar_init(sd_full, 1);
    mbar_init(pb_full, ATB_COMPUTE);
    mbar_init(acc_full, 1);
    mbar_init(tm_free, ATB_COMPUTE);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&del_full[i], 64);
      mbar_init(&del_empty[i], ATB_COMPUTE);
    }
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
  // TMEM columns: S 0-207 | dP 256-463; once a stage's S / dP have been consumed, dQ_t / dV_u reuse 0-63 (R2)
  // and dK_u reuses 256-319 (RK)
  const uint32_t R0 = tmem_base, R1 = tmem_base + 256, R2 = R0, RK = R1;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int b = item / p.H, h = item - b * p.H;
        const int row0 = b * p.N, col0 = h * AT_HD;
        mbar_wait(in_empty, ph ^ 1);
        ph ^= 1;
        mbar_expect_tx(in_full, 4 * AT_TILE);
        tma_load_2d(&tmQKV, sQ, in_full, col0, row0);
        tma_load_2d(&tmQKV, sK, in_full, D + col0, row0);
        tma_load_2d(&tmQKV, sV, in_full, 2 * D + col0, row0);
        tma_load_2d(&tmDO, sdO, in_full, col0, row0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, AT_ROWS, 0, 0);   // [128 x 208] = A(K-major) B(K-major)^T
      const uint32_t idesc_o = make_idesc_bf16(128, AT_HD, 0, 1);     // [128 x 64]  = A(K-major) B(MN-major)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aO = smem_u32(sdO), aP = smem_u32(sPB);
      uint32_t ph_in = 0, ph_tm = 0, ph_pb = 0;
      auto mma_scores = [&](uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(R0, make_smem_desc(a0 + k * 32, 16, 1024), make_smem_desc(b0 + k * 32, 16, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(R1, make_smem_desc(a1 + k * 32, 16, 1024), make_smem_desc(b1 + k * 32, 16, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
        tc_commit(sd_full);
      };
      auto mma_out = [&](uint32_t dst, uint32_t bmn) {  // dst[128 x 64] = PB[128 x 208] * B[208 x 64]
        mbar_wait(pb_full, ph_pb);
        ph_pb ^= 1;
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < AT_KSTEPS; ++ks)
          tc_mma_bf16(dst, make_smem_desc(aP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                      make_smem_desc(bmn + ks * 2048, 8192, 1024), idesc_o, ks > 0 ? 1u : 0u);
        tc_commit(acc_full);
      };
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        mbar_wait(in_full, ph_in);
        ph_in ^= 1;
        for (int t = 0; t < 2; ++t) {  // stage A_t
          mbar_wait(tm_free, ph_tm ^ 1);
          ph_tm ^= 1;
          tc_fence_after();
          mma_scores(aQ + t * 16384, aK, aO + t * 16384, aV);
          mma_out(R2, aK);  // dQ_t = dS_t K
        }
        for (int u = 0; u < 2; ++u) {  // stage B_u
          mbar_wait(tm_free, ph_tm ^ 1);
          ph_tm ^= 1;
          tc_fence_after();
          mma_scores(aK + u * 16384, aQ, aV + u * 16384, aO);
          mma_out(R2, aO);  // dV_u = P^T_u dO
          mma_out(RK, aQ);  // dK_u = dS^T_u Q
        }
        tc_commit(in_empty);
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
          sDel[row] = acc;
          sLse[row] = row < p.N ? p.lse[(static_cast<long long>(b) * p.H + h) * p.N + row] * 1.4426950408889634f
                                : INFINITY;
        }
      }
      mbar_arrive(&del_full[buf]);
      buf ^= 1;
      if (buf == 0) ph ^= 1;
    }
  } else if (warp >= 4) {
    const int cw = warp - 4;
    const int wq = cw & 3;   // TMEM lane quarter (== warp % 4)
    const int cg = cw >> 2;  // column group: 16-column chunks j with (j & 3) == cg
    const int r = wq * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
    const float c2 = p.scale * 1.4426950408889634f;
    uint8_t* prow = sPB + r * 128;
    const int sw = r & 7;
    uint32_t ph_sd = 0, ph_acc = 0, dph = 0;
    int dbuf = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      // lse / delta of this item were produced by warps 2-3 while the previous item ran
      mbar_wait(&del_full[dbuf], dph);
      const float* sLse = sLseAll + dbuf * AT_ROWS;
      const float* sDel = sDelAll + dbuf * AT_ROWS;

      // ---------------- stage A_t : rows = queries, columns = keys ----------------
      for (int t = 0; t < 2; ++t) {
        const int q = t * 128 + r;
        const bool qok = q < p.N;
        const float lse_r = qok ? sLse[q] : INFINITY;
        const float del_r = qok ? sDel[q] : 0.f;
        mbar_wait(sd_full, ph_sd);
        ph_sd ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int j = cg; j < AT_KSTEPS; j += 4) {
          uint32_t s[16], dp[16];
          tmem_ld16(R0 + lane_sel + j * 16, s);
          tmem_ld16(R1 + lane_sel + j * 16, dp);
          tmem_ld_wait();
          float ds[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float pe = ex2_approx_ftz(fmaf(__uint_as_float(s[e]), c2, -lse_r));
            const float v = pe * (__uint_as_float(dp[e]) - del_r) * p.scale;
            ds[e] = (qok && (j * 16 + e < p.N)) ? v : 0.f;
          }
          pack16_store(prow, j, sw, ds);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pb_full);
        // dQ_t
        mbar_wait(acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(R2 + lane_sel + cg * 16, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tm_free);
        if (qok) store16_bf16(p.dqkv + (static_cast<long long>(b) * p.N + q) * 3 * D + h * AT_HD + cg * 16, o);
      }
      // ---------------- stage B_u : rows = keys, columns = queries ----------------
      for (int u = 0; u < 2; ++u) {
        const int key = u * 128 + r;
        mbar_wait(sd_full, ph_sd);
        ph_sd ^= 1;
        tc_fence_after();
        uint32_t dsp[4][8];  // dS^T of this thread's chunks, packed bf16
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = cg + 4 * jj;
          if (j < AT_KSTEPS) {
            uint32_t s[16], dp[16];
            tmem_ld16(R0 + lane_sel + j * 16, s);
            tmem_ld16(R1 + lane_sel + j * 16, dp);
            tmem_ld_wait();
            float pt[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 l4 = *reinterpret_cast<const float4*>(sLse + j * 16 + e4 * 4);
              const float4 d4 = *reinterpret_cast<const float4*>(sDel + j * 16 + e4 * 4);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
              const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
              float dsv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = e4 * 4 + e;
                const bool cok = (j * 16 + col) < p.N;
                const float pe = cok ? ex2_approx_ftz(fmaf(__uint_as_float(s[col]), c2, -ls[e])) : 0.f;
                pt[col] = pe;
                dsv[e] = cok ? pe * (__uint_as_float(dp[col]) - dl[e]) * p.scale : 0.f;
              }
              dsp[jj][e4 * 2] = pack_bf16x2(dsv[0], dsv[1]);
              dsp[jj][e4 * 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
            }
            pack16_store(prow, j, sw, pt);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pb_full);  // P^T_u ready -> dV_u
        mbar_wait(acc_full, ph_acc);
        ph_acc ^= 1;  // dV_u done: the operand buffer may be overwritten
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = cg + 4 * jj;
          if (j < AT_KSTEPS) {
            uint8_t* blk = prow + (j >> 2) * 16384;
            const int ck = (j & 3) * 2;
            *reinterpret_cast<uint4*>(blk + ((ck ^ sw) << 4)) = make_uint4(dsp[jj][0], dsp[jj][1], dsp[jj][2], dsp[jj][3]);
            *reinterpret_cast<uint4*>(blk + (((ck + 1) ^ sw) << 4)) = make_uint4(dsp[jj][4], dsp[jj][5], dsp[jj][6], dsp[jj][7]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(pb_full);  // dS^T_u ready -> dK_u
        mbar_wait(acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        uint32_t ov[16], ok_[16];
        tmem_ld16(R2 + lane_sel + cg * 16, ov);
        tmem_ld16(RK + lane_sel + cg * 16, ok_);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(tm_free);
        if (key < p.N) {
          bf16* base = p.dqkv + (static_cast<long long>(b) * p.N + key) * 3 * D + h * AT_HD + cg * 16;
          store16_bf16(base + D, ok_);
          store16_bf16(base + 2 * D, ov);
        }
      }
      mbar_arrive(&del_empty[dbuf]);  // this thread no longer reads the item's lse / delta
      dbuf ^= 1;
      if (dbuf == 0) dph ^= 1;
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

static int encode_qkv_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols) {
  uint64_t dims[2] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows)};
  uint64_t strides[1] = {static_cast<uint64_t>(cols) * 2};
  uint32_t box[2] = {64, AT_ROWS};
  return encode_tensor_map(tm, ptr, 2, dims, strides, box);
}

extern "C" int theia_attention_tc_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream) {
  if (N > AT_ROWS - 0 || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention: sequence length %d > %d", N, AT_ROWS);
  const int D = H * AT_HD;
  CUtensorMap tm;
  int rc = encode_qkv_map(&tm, qkv, static_cast<long long>(B) * N, 3LL * D);
  if (rc) return rc;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATF_SMEM);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn_tc fwd attr: %s", cudaGetErrorString(e));
    done = true;
  }
  AttnFwdParams p;
  p.out = static_cast<bf16*>(out);
  p.lse = lse;
  p.B = B, p.N = N, p.H = H, p.D = D;
  p.items = B * H;
  p.scale = 0.125f;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_tc_fwd_kernel<<<grid, ATF_THREADS, ATF_SMEM, static_cast<cudaStream_t>(stream)>>>(tm, p);
  THEIA_CHECK_LAUNCH("attention_tc_fwd");
  return THEIA_OK;
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
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATB_SMEM);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn_tc bwd attr: %s", cudaGetErrorString(e));
    done = true;
  }
  AttnBwdParams p;
  p.out = static_cast<const bf16*>(out);
  p.dout = static_cast<const bf16*>(dout);
  p.lse = lse;
  p.dqkv = static_cast<bf16*>(dqkv);
  p.B = B, p.N = N, p.H = H, p.D = D;
  p.items = B * H;
  p.scale = 0.125f;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  attn_tc_bwd_kernel<<<grid, ATB_THREADS, ATB_SMEM, static_cast<cudaStream_t>(stream)>>>(tmq, tmd, p);
  THEIA_CHECK_LAUNCH("attention_tc_bwd");
  return THEIA_OK;
}
