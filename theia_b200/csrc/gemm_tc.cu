// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// One persistent, warp-specialised kernel:
//   warp 0   TMA producer   (cp.async.bulk.tensor 2-D / 4-D, 128-byte swizzle, OOB zero fill
//                            = the convolution padding)
//   warp 1   MMA issuer     (one thread, tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16)
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     (tcgen05.ld -> registers -> swizzled smem transpose -> coalesced
//                            global stores with the fused epilogue; two warps per TMEM lane quarter,
//                            alternating 32-column chunks; epilogue flags are compile-time)
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand modes (include/theia_b200.h): K-major 2-D, MN-major 2-D (wgrad), K-major NHWC
// gather with the taps folded into K (forward / dgrad convolutions), MN-major NHWC gather of
// one tap (convolution wgrad, tap = z slice).
#include <cuda.h>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

struct GemmK {
  int M, N;
  int num_kb, kb_per_split, splits, batch_z;
  int m_tiles, n_tiles, total_items;
  int a_mode, b_mode;
  int cchunks;  // CONV_K: 64-channel chunks per tap
  int es;       // gather stride of the NHWC operand
  int b_tap_rows;
  int wtap[9];
  int dh[9], dw[9];
  int tile_w, tile_h, tiles_per_img;
  int kb_per_img, rows_per_kb;  // CONV_MN
  int conv_out;                 // output rows follow the conv geometry
  int out_h, out_w, out_img_rows, out_row_off, out_wpitch, sy, sx, py, px;
  int epi;
  void* out;
  long long ldo;
  void* out2;
  const float* bias;
  const bf16* aux;
  const float* pos;
  const float* cls;
  int tokens, tok_p0, tok_p1;
  float* stats;
  float* colsum;
  long long out_z_stride;
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;
constexpr int STAGING_BYTES = EPI_WARPS * 32 * 32 * 4;
constexpr int SMEM_LIMIT = 232448;  // 227 KB

constexpr int AUX_SLOTS = 3;                                  // per-warp ring depth (32x32 bf16 chunks)
constexpr int AUX_RING_BYTES = EPI_WARPS * AUX_SLOTS * 2048;  // 48 KB
constexpr int AUXF = THEIA_EPI_RESID | THEIA_EPI_MUL_AUX | THEIA_EPI_MUL_RELUMASK;

template <int BN, bool RING>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EXTRA = STAGING_BYTES + (RING ? AUX_RING_BYTES : 0);
  static constexpr int STAGES = (SMEM_LIMIT - EXTRA - 1024 - 256) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EXTRA + 256 + 1024;
};

struct Item {
  int m_blk, n_blk, z, kb0, kb1;
};

__device__ __forceinline__ Item decode_item(const GemmK& p, int item) {
  Item it;
  it.n_blk = item % p.n_tiles;
  int t = item / p.n_tiles;
  it.m_blk = t % p.m_tiles;
  t /= p.m_tiles;
  it.z = t % p.batch_z;
  const int s = t / p.batch_z;
  it.kb0 = s * p.kb_per_split;
  it.kb1 = min(p.num_kb, it.kb0 + p.kb_per_split);
  return it;
}

// Exact-erf GELU pieces from Abramowitz-Stegun 7.1.26 (|erf err| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2
// (shared with the Gaussian pdf) and five FMAs -- ~15 instructions per element instead of erff()+expf().
// SoA helpers over small register arrays so the compiler interleaves the independent dependency chains
// (the epilogue runs with only two warps per scheduler: ILP has to hide the fixed-latency stalls).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int NV, bool WANT_PDF>
__device__ __forceinline__ void normal_cdf_pdf(const float (&x)[NV], float (&cdf)[NV], float (&pdf)[NV]) {
  float t[NV], e[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float ax = fabsf(x[k]);
    t[k] = rcp_approx(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.0f));
    const float u = ax * 0.84932180028801907f;  // sqrt(log2(e) / 2)
    e[k] = ex2_approx(-(u * u));                // exp(-x^2 / 2)
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float pl = fmaf(t[k], 0.5f * 1.061405429f, 0.5f * -1.453152027f);
    pl = fmaf(pl, t[k], 0.5f * 1.421413741f);
    pl = fmaf(pl, t[k], 0.5f * -0.284496736f);
    pl = fmaf(pl, t[k], 0.5f * 0.254829592f);
    const float hq = pl * t[k] * e[k];  // 0.5 * erfc(|x| / sqrt 2)
    cdf[k] = 0.5f + copysignf(0.5f - hq, x[k]);
    if (WANT_PDF) pdf[k] = 0.39894228040143268f * e[k];
  }
}

// EPI_CT >= 0: epilogue flags are a compile-time constant (hot combinations); -1: read p.epi.
template <int BN, int EPI_CT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmK p) {
  // compile-time specialised kernels that read an aux operand stage it through a cp.async ring in shared
  // memory (48 KB in flight per SM: registers alone cannot keep enough HBM reads outstanding)
  constexpr bool RING = (EPI_CT >= 0) && ((EPI_CT & AUXF) != 0);
  using C = Cfg<BN, RING>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint8_t* aux_ring_all = smem + C::STAGES * C::STAGE_BYTES + STAGING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EXTRA);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        const int m0 = it.m_blk * BM, n0 = it.n_blk * BN;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          // ---- A ----
          if (p.a_mode == THEIA_OP_K2D) {
            tma_load_2d(&tmA, sa, &full[stage], kb * BK, m0);
          } else if (p.a_mode == THEIA_OP_MN2D) {
            tma_load_2d(&tmA, sa, &full[stage], m0, kb * BK);
            tma_load_2d(&tmA, sa + 8192, &full[stage], m0 + 64, kb * BK);
          } else {  // CONV_K
            const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
            const int b = it.m_blk / p.tiles_per_img, ht = it.m_blk - b * p.tiles_per_img;
            tma_load_4d(&tmA, sa, &full[stage], cc * 64, p.dw[tap], p.es * ht * p.tile_h + p.dh[tap], b);
          }
          // ---- B ----
          if (p.b_mode == THEIA_OP_K2D) {
            if (p.b_tap_rows > 0) {  // tap-major weight pack [tap][rows][C]
              const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
              tma_load_2d(&tmB, sb, &full[stage], cc * 64, p.wtap[tap] * p.b_tap_rows + n0);
            } else {
              tma_load_2d(&tmB, sb, &full[stage], kb * BK, n0);
            }
          } else if (p.b_mode == THEIA_OP_MN2D) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(&tmB, sb + i * 8192, &full[stage], n0 + 64 * i, kb * BK);
          } else {  // CONV_MN: tap = z
            const int b = kb / p.kb_per_img, hb = kb - b * p.kb_per_img;
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(&tmB, sb + i * 8192, &full[stage], n0 + 64 * i, p.dw[it.z],
                          p.es * hb * p.rows_per_kb + p.dh[it.z], b);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t db = make_smem_desc(b_addr + k * p.b_kstep, p.b_lbo, p.b_sbo);
            tc_mma_bf16(d_tmem, da, db, p.idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[acc]);  // accumulator complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int ew = warp - 4;      // 0..7
    const int wq = ew & 3;        // TMEM lane quarter == warp % 4
    const int hsel = ew >> 2;     // this warp takes the 32-column chunks with (chunk & 1) == hsel
    float* stg = staging_all + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int epi = (EPI_CT >= 0) ? EPI_CT : p.epi;
    const int c = lane & 7;
    const int rsub = lane >> 3;
    constexpr int NCHUNK = BN / 32;
    // ---- aux ring (RING kernels): chunk q of this warp lives in slot q % AUX_SLOTS; the prefetch cursor
    // (pitem, pch) runs AUX_SLOTS chunks ahead of the processing position, across tile boundaries ----
    uint8_t* ring = aux_ring_all + ew * (AUX_SLOTS * 2048);
    int pitem = blockIdx.x, pch = hsel, slot = 0;
    auto ring_issue = [&](int sl) {
      if (pitem < p.total_items) {
        const Item pit = decode_item(p, pitem);
        const int nn = pit.n_blk * BN + pch * 32 + c * 4;
        const int pimg = p.conv_out ? pit.m_blk / p.tiles_per_img : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = wq * 32 + i * 4 + rsub;
          long long orow;
          bool okr;
          if (p.conv_out) {
            const int ht = pit.m_blk - pimg * p.tiles_per_img;
            const int hh = r / p.tile_w, ww = r - hh * p.tile_w;
            const int h = ht * p.tile_h + hh;
            okr = (h < p.out_h) && (ww < p.out_w);
            orow = static_cast<long long>(pimg) * p.out_img_rows + p.out_row_off +
                   static_cast<long long>(h * p.sy + p.py) * p.out_wpitch + (ww * p.sx + p.px);
          } else {
            orow = pit.m_blk * BM + r;
            okr = orow < p.M;
          }
          const uint32_t dst = smem_u32(ring + sl * 2048 + (i * 32 + lane) * 8);
          if (okr && nn < p.N) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(p.aux + orow * p.ldo + nn)
                         : "memory");
          } else {
            asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(dst), "r"(0u) : "memory");
          }
        }
        pch += 2;
        if (pch >= NCHUNK) {
          pch = hsel;
          pitem += gridDim.x;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (RING) {
#pragma unroll
      for (int sidx = 0; sidx < AUX_SLOTS; ++sidx) ring_issue(sidx);
    }
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      // ---- per-tile row bookkeeping: this lane touches rows rr = 4*i + rsub of its warp's slab ----
      long long rowoff[8];
      uint32_t okmask = 0;
      int img = 0;
      if (p.conv_out) img = it.m_blk / p.tiles_per_img;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = wq * 32 + i * 4 + rsub;
        const int m = it.m_blk * BM + r;
        long long orow;
        bool ok;
        if (p.conv_out) {
          const int ht = it.m_blk - img * p.tiles_per_img;
          const int hh = r / p.tile_w, ww = r - hh * p.tile_w;
          const int h = ht * p.tile_h + hh;
          ok = (h < p.out_h) && (ww < p.out_w);
          orow = static_cast<long long>(img) * p.out_img_rows + p.out_row_off +
                 static_cast<long long>(h * p.sy + p.py) * p.out_wpitch + (ww * p.sx + p.px);
        } else {
          ok = m < p.M;
          orow = m;
        }
        rowoff[i] = orow * p.ldo;
        okmask |= (ok ? 1u : 0u) << i;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      float st_s = 0.f, st_ss = 0.f;
#pragma unroll 1
      for (int ch = hsel; ch < NCHUNK; ch += 2) {
        const bool last = (ch + 2 >= NCHUNK);
        const int nbase = it.n_blk * BN + ch * 32;
        if (RING) asm volatile("cp.async.wait_group %0;" ::"n"(AUX_SLOTS - 1) : "memory");  // this chunk's aux landed
        if (nbase >= p.N) {  // nothing to store; still release the accumulator
          if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          if (RING) {
            ring_issue(slot);
            slot = (slot + 1 == AUX_SLOTS) ? 0 : slot + 1;
          }
          continue;
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN + ch * 32, v);
        tmem_ld_wait();
        if (last) {  // all TMEM reads of this warp for this accumulator are done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        // phase 1: row-per-thread -> swizzled staging (conflict-free float4 stores)
        {
          float4* row = reinterpret_cast<float4*>(stg + lane * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 f;
            f.x = __uint_as_float(v[4 * j + 0]);
            f.y = __uint_as_float(v[4 * j + 1]);
            f.z = __uint_as_float(v[4 * j + 2]);
            f.w = __uint_as_float(v[4 * j + 3]);
            row[j ^ (lane & 7)] = f;
          }
        }
        __syncwarp();
        // phase 2: 4 rows x 128 B per warp instruction, coalesced global access; two passes of 4 rows
        // (16 values per lane) so the per-element math runs as 16 interleaved chains
        const int n = nbase + c * 4;
        const uint32_t ok = (n < p.N) ? okmask : 0u;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && ok) bias4 = *reinterpret_cast<const float4*>(p.bias + n);
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
        uint2 aux[8];
        if (RING) {
#pragma unroll
          for (int i = 0; i < 8; ++i) aux[i] = *reinterpret_cast<const uint2*>(ring + slot * 2048 + (i * 32 + lane) * 8);
        } else if (epi & AUXF) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            aux[i] = ((ok >> i) & 1u) ? *reinterpret_cast<const uint2*>(p.aux + rowoff[i] + n) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float xv[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = (half * 4 + j) * 4 + rsub;
            const float4 f = reinterpret_cast<const float4*>(stg + rr * 32)[c ^ (rr & 7)];
            xv[4 * j + 0] = f.x + bias4.x, xv[4 * j + 1] = f.y + bias4.y;
            xv[4 * j + 2] = f.z + bias4.z, xv[4 * j + 3] = f.w + bias4.w;
          }
          if (epi & THEIA_EPI_POSCLS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = half * 4 + j;
              if (!((ok >> i) & 1u)) continue;
              const int m = it.m_blk * BM + wq * 32 + i * 4 + rsub;
              const int t = m % p.tokens;
              const float4 ps = *reinterpret_cast<const float4*>(p.pos + static_cast<long long>(t) * p.N + n);
              if (t < p.tok_p0 || t >= p.tok_p1) {  // CLS / register token: the table holds the whole value
                xv[4 * j + 0] = ps.x, xv[4 * j + 1] = ps.y, xv[4 * j + 2] = ps.z, xv[4 * j + 3] = ps.w;
              } else {
                xv[4 * j + 0] += ps.x, xv[4 * j + 1] += ps.y, xv[4 * j + 2] += ps.z, xv[4 * j + 3] += ps.w;
              }
            }
          }
          if (epi & THEIA_EPI_GELU) {
            // gelu(x) = x Phi(x); its derivative Phi(x) + x phi(x) is stored (bf16) for the backward pass, so
            // the dgrad epilogue is a plain multiply instead of a second erf/exp evaluation
            float cdf[16], pdf[16];
            normal_cdf_pdf<16, true>(xv, cdf, pdf);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = half * 4 + j;
              uint2 dg;
              dg.x = pack_bf16x2(fmaf(xv[4 * j + 0], pdf[4 * j + 0], cdf[4 * j + 0]),
                                 fmaf(xv[4 * j + 1], pdf[4 * j + 1], cdf[4 * j + 1]));
              dg.y = pack_bf16x2(fmaf(xv[4 * j + 2], pdf[4 * j + 2], cdf[4 * j + 2]),
                                 fmaf(xv[4 * j + 3], pdf[4 * j + 3], cdf[4 * j + 3]));
              if ((ok >> i) & 1u) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out2) + rowoff[i] + n) = dg;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) xv[k] *= cdf[k];
          }
          if (epi & THEIA_EPI_RELU) {
#pragma unroll
            for (int k = 0; k < 16; ++k) xv[k] = fmaxf(xv[k], 0.f);
          }
          if (epi & AUXF) {
            float av[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 a01 = unpack_bf16x2(aux[half * 4 + j].x), a23 = unpack_bf16x2(aux[half * 4 + j].y);
              av[4 * j + 0] = a01.x, av[4 * j + 1] = a01.y, av[4 * j + 2] = a23.x, av[4 * j + 3] = a23.y;
            }
            if (epi & THEIA_EPI_MUL_AUX) {
#pragma unroll
              for (int k = 0; k < 16; ++k) xv[k] *= av[k];
            } else if (epi & THEIA_EPI_MUL_RELUMASK) {
#pragma unroll
              for (int k = 0; k < 16; ++k) xv[k] = av[k] > 0.f ? xv[k] : 0.f;
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) xv[k] += av[k];
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = half * 4 + j;
            const bool rok = (ok >> i) & 1u;
            const long long off = rowoff[i] + n;
            const float x0 = xv[4 * j + 0], x1 = xv[4 * j + 1], x2 = xv[4 * j + 2], x3 = xv[4 * j + 3];
            if (epi & THEIA_EPI_ATOMIC) {
              float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(it.z) * p.out_z_stride + off;
              if (rok)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(x0), "f"(x1), "f"(x2),
                             "f"(x3)
                             : "memory");
            } else if (epi & THEIA_EPI_OUT_F32) {
              if (rok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(x0, x1, x2, x3);
            } else {
              uint2 o;
              o.x = pack_bf16x2(x0, x1);
              o.y = pack_bf16x2(x2, x3);
              if (rok) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + off) = o;
              if (epi & (THEIA_EPI_STATS | THEIA_EPI_COLSUM)) {
                const float msk = rok ? 1.f : 0.f;
                const float2 q01 = unpack_bf16x2(o.x), q23 = unpack_bf16x2(o.y);
                if (epi & THEIA_EPI_STATS) {
                  st_s += msk * ((q01.x + q01.y) + (q23.x + q23.y));
                  st_ss += msk * ((q01.x * q01.x + q01.y * q01.y) + (q23.x * q23.x + q23.y * q23.y));
                }
                if (epi & THEIA_EPI_COLSUM)
                  cs0 += msk * q01.x, cs1 += msk * q01.y, cs2 += msk * q23.x, cs3 += msk * q23.y;
              }
            }
          }
        }
        if (RING) {  // aux of this chunk is consumed: refill the slot with the chunk AUX_SLOTS ahead
          ring_issue(slot);
          slot = (slot + 1 == AUX_SLOTS) ? 0 : slot + 1;
        }
        if (epi & THEIA_EPI_COLSUM) {  // column sums of the stored tile rows (bias gradient of the consumer)
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, 8), cs1 += __shfl_xor_sync(0xffffffffu, cs1, 8);
          cs2 += __shfl_xor_sync(0xffffffffu, cs2, 8), cs3 += __shfl_xor_sync(0xffffffffu, cs3, 8);
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16), cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
          cs2 += __shfl_xor_sync(0xffffffffu, cs2, 16), cs3 += __shfl_xor_sync(0xffffffffu, cs3, 16);
          if (rsub == 0 && n < p.N)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + n), "f"(cs0), "f"(cs1),
                         "f"(cs2), "f"(cs3)
                         : "memory");
        }
        __syncwarp();
      }
      if (epi & THEIA_EPI_STATS) {
        st_s = warp_sum(st_s);
        st_ss = warp_sum(st_ss);
        if (lane == 0) {
          atomicAdd(p.stats + 2 * img, st_s);
          atomicAdd(p.stats + 2 * img + 1, st_ss);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (RING) asm volatile("cp.async.wait_all;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static long long g_dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};

// ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream ----
constexpr int PROF_RING = 8192;
static bool g_prof_on = false;
static cudaEvent_t g_ev0[PROF_RING], g_ev1[PROF_RING];
static double g_prof_flops[PROF_RING];
static int g_prof_meta[PROF_RING][8];
static int g_prof_n = 0;
static bool g_prof_init = false;

static int encode_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                     uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {pitch_elems * 2};
  uint32_t box[2] = {box_inner, box_outer};
  return encode_tensor_map(tm, ptr, 2, dims, strides, box);
}

template <int BN, int EPI_CT>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmK& k, cudaStream_t stream) {
  using C = Cfg<BN, (EPI_CT >= 0) && ((EPI_CT & AUXF) != 0)>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI_CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  int grid = k.total_items < num_sms() ? k.total_items : num_sms();
  const bool prof = g_prof_on && g_prof_n < PROF_RING;
  if (prof) cudaEventRecord(g_ev0[g_prof_n], stream);
  gemm_tc_kernel<BN, EPI_CT><<<grid, NTHREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, k);
  if (prof) {
    cudaEventRecord(g_ev1[g_prof_n], stream);
    g_prof_flops[g_prof_n] = 2.0 * k.M * (double)k.N * (double)k.num_kb * BK * k.batch_z;
    int* mt = g_prof_meta[g_prof_n];
    mt[0] = k.M, mt[1] = k.N, mt[2] = k.num_kb * BK, mt[3] = k.a_mode, mt[4] = k.b_mode, mt[5] = k.epi;
    mt[6] = k.splits * 1000 + k.batch_z, mt[7] = BN;
    ++g_prof_n;
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  return THEIA_OK;
}

// hot epilogue combinations get a compile-time specialisation; anything else runs the generic kernel
template <int BN>
static int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmK& k, cudaStream_t stream) {
  switch (k.epi) {
#define THEIA_EPI_CASE(E) \
  case (E):               \
    return launch<BN, (E)>(tmA, tmB, k, stream);
    THEIA_EPI_CASE(0)
    THEIA_EPI_CASE(THEIA_EPI_GELU)
    THEIA_EPI_CASE(THEIA_EPI_RESID)
    THEIA_EPI_CASE(THEIA_EPI_OUT_F32)
    THEIA_EPI_CASE(THEIA_EPI_ATOMIC)
    THEIA_EPI_CASE(THEIA_EPI_MUL_AUX)
    THEIA_EPI_CASE(THEIA_EPI_MUL_AUX | THEIA_EPI_COLSUM)
    THEIA_EPI_CASE(THEIA_EPI_POSCLS)
    THEIA_EPI_CASE(THEIA_EPI_STATS)
    THEIA_EPI_CASE(THEIA_EPI_RELU | THEIA_EPI_STATS)
#undef THEIA_EPI_CASE
    default:
      return launch<BN, -1>(tmA, tmB, k, stream);
  }
}

}  // namespace theia

using namespace theia;

extern "C" int theia_debug_set(int key, long long value) {
  if (key == 0) {
    for (auto& v : g_dbg) v = 0;
    return 0;
  }
  if (key < 0 || key >= 8) return THEIA_ERR_ARG;
  g_dbg[key] = value;
  return 0;
}

extern "C" int theia_prof_enable(int on) {
  if (on && !g_prof_init) {
    for (int i = 0; i < PROF_RING; ++i) {
      if (cudaEventCreate(&g_ev0[i]) != cudaSuccess || cudaEventCreate(&g_ev1[i]) != cudaSuccess)
        return set_error(THEIA_ERR_CUDA, "cudaEventCreate failed");
    }
    g_prof_init = true;
  }
  g_prof_on = on != 0;
  if (on) g_prof_n = 0;
  return THEIA_OK;
}

// Per-launch record i of the current ring: ms, M, N, K, a_mode, b_mode, epi, splits*1000+batch_z, BN
extern "C" int theia_prof_record(int i, double* ms, int* meta8) {
  if (i < 0 || i >= g_prof_n) return THEIA_ERR_ARG;
  cudaEventSynchronize(g_ev1[i]);
  float t = 0.f;
  if (cudaEventElapsedTime(&t, g_ev0[i], g_ev1[i]) != cudaSuccess) return set_error(THEIA_ERR_CUDA, "elapsed");
  *ms = t;
  for (int k = 0; k < 8; ++k) meta8[k] = g_prof_meta[i][k];
  return THEIA_OK;
}

// Sums the recorded GEMM launches (device time in ms, executed flops), then resets the ring.
extern "C" int theia_prof_collect(double* total_ms, double* total_flops, long long* launches) {
  double ms = 0.0, fl = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    cudaEventSynchronize(g_ev1[i]);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_ev0[i], g_ev1[i]) != cudaSuccess)
      return set_error(THEIA_ERR_CUDA, "cudaEventElapsedTime failed");
    ms += t;
    fl += g_prof_flops[i];
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = g_prof_n;
  g_prof_n = 0;
  return THEIA_OK;
}

extern "C" int theia_gemm(const theia_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !d->A || !d->B || !d->out) return set_error(THEIA_ERR_ARG, "theia_gemm: null pointer");
  if (d->N % 8 != 0) return set_error(THEIA_ERR_ARG, "theia_gemm: N must be a multiple of 8 (got %d)", d->N);
  if (d->ldo % 4 != 0) return set_error(THEIA_ERR_ARG, "theia_gemm: ldo must be a multiple of 4");
  int bn = d->bn;
  if (bn == 0) {
    if (d->N % 256 == 0) bn = 256;
    else if (d->N % 192 == 0) bn = 192;
    else if (d->N % 128 == 0) bn = 128;
    else bn = d->N > 192 ? 256 : (d->N > 128 ? 192 : 128);
  }
  if (bn != 128 && bn != 192 && bn != 256) return set_error(THEIA_ERR_ARG, "theia_gemm: bn must be 128/192/256");

  GemmK k;
  memset(&k, 0, sizeof(k));
  k.M = d->M;
  k.N = d->N;
  k.a_mode = d->a_mode;
  k.b_mode = d->b_mode;
  k.epi = d->epi;
  k.out = d->out;
  k.ldo = d->ldo;
  k.out2 = d->out2;
  k.bias = d->bias;
  k.aux = static_cast<const bf16*>(d->aux);
  k.pos = d->pos;
  k.cls = d->cls;
  k.tokens = d->tokens > 0 ? d->tokens : 1;
  k.tok_p0 = d->tok_p0, k.tok_p1 = d->tok_p1;
  k.stats = d->stats;
  k.colsum = d->colsum;
  k.out_z_stride = d->out_z_stride;
  k.batch_z = d->batch_z > 0 ? d->batch_z : 1;
  k.splits = d->splits > 0 ? d->splits : 1;
  if (k.splits > 1 && !(d->epi & THEIA_EPI_ATOMIC)) return set_error(THEIA_ERR_ARG, "split-K needs EPI_ATOMIC");
  if ((d->epi & THEIA_EPI_GELU) && !d->out2) return set_error(THEIA_ERR_ARG, "EPI_GELU needs out2");
  if ((d->epi & (THEIA_EPI_RESID | THEIA_EPI_MUL_AUX | THEIA_EPI_MUL_RELUMASK)) && !d->aux)
    return set_error(THEIA_ERR_ARG, "epilogue needs aux");
  if ((d->epi & THEIA_EPI_POSCLS) && !d->pos) return set_error(THEIA_ERR_ARG, "POSCLS needs the token table (pos)");
  if ((d->epi & THEIA_EPI_COLSUM) && !d->colsum) return set_error(THEIA_ERR_ARG, "COLSUM needs colsum");
  if ((d->epi & THEIA_EPI_STATS) && (!d->stats || d->a_mode != THEIA_OP_CONV_K))
    return set_error(THEIA_ERR_ARG, "STATS needs stats and a CONV_K A operand");

  const theia_conv_geom& g = d->conv;
  CUtensorMap tmA, tmB;
  int rc;
  const int a_mn = (d->a_mode == THEIA_OP_MN2D);
  const int b_mn = (d->b_mode == THEIA_OP_MN2D || d->b_mode == THEIA_OP_CONV_MN);

  // ---- A ----
  if (d->a_mode == THEIA_OP_K2D) {
    rc = encode_2d(&tmA, d->A, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, 64, BM);
  } else if (d->a_mode == THEIA_OP_MN2D) {
    rc = encode_2d(&tmA, d->A, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, 64, 64);
  } else if (d->a_mode == THEIA_OP_CONV_K) {
    if (g.tile_w * g.tile_h != BM || g.C % 64 != 0 || g.ntaps < 1 || g.ntaps > 9)
      return set_error(THEIA_ERR_ARG, "CONV_K: tile must be 128 pixels, C %% 64 == 0");
    uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    uint64_t strides[3] = {(uint64_t)g.stride_w * 2, (uint64_t)g.stride_h * 2, (uint64_t)g.stride_b * 2};
    const uint32_t es = g.in_stride > 1 ? (uint32_t)g.in_stride : 1u;
    uint32_t box[4] = {64, es * (uint32_t)g.tile_w, es * (uint32_t)g.tile_h, 1};
    uint32_t estr[4] = {1, es, es, 1};
    rc = encode_tensor_map(&tmA, d->A, 4, dims, strides, box, estr);
    k.cchunks = g.C / 64;
    k.tile_w = g.tile_w;
    k.tile_h = g.tile_h;
    k.tiles_per_img = (g.out_h + g.tile_h - 1) / g.tile_h;
    k.conv_out = 1;
    k.out_h = g.out_h, k.out_w = g.out_w, k.out_img_rows = g.out_img_rows, k.out_row_off = g.out_row_off;
    k.out_wpitch = g.out_wpitch, k.sy = g.sy > 0 ? g.sy : 1, k.sx = g.sx > 0 ? g.sx : 1, k.py = g.py, k.px = g.px;
    if (d->M != g.B * k.tiles_per_img * BM) return set_error(THEIA_ERR_ARG, "CONV_K: M != B*tiles*128");
    if (d->K != g.ntaps * g.C) return set_error(THEIA_ERR_ARG, "CONV_K: K != ntaps*C");
  } else {
    return set_error(THEIA_ERR_ARG, "bad a_mode");
  }
  if (rc) return rc;
  // ---- B ----
  if (d->b_mode == THEIA_OP_K2D && d->a_mode == THEIA_OP_CONV_K && g.b_tap_rows > 0) {
    // tap-major weight pack [9][b_tap_rows][C]: one 2-D map over all taps
    rc = encode_2d(&tmB, d->B, (uint64_t)g.C, (uint64_t)9 * g.b_tap_rows, (uint64_t)g.C, 64, bn);
    k.b_tap_rows = g.b_tap_rows;
    for (int i = 0; i < 9; ++i) k.wtap[i] = g.wtap[i];
  } else if (d->b_mode == THEIA_OP_K2D) {
    rc = encode_2d(&tmB, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, 64, bn);
  } else if (d->b_mode == THEIA_OP_MN2D) {
    rc = encode_2d(&tmB, d->B, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->ldb, 64, 64);
  } else if (d->b_mode == THEIA_OP_CONV_MN) {
    // pixel grid = out_h x tile_w (the dY grid); the gathered tensor (conv.H x conv.W) may be
    // smaller (14x14 tokens under the 16x16 grid of the pad convolution)
    if (g.tile_w < 1 || 64 % g.tile_w != 0 || (g.out_h * g.tile_w) % 64 != 0)
      return set_error(THEIA_ERR_ARG, "CONV_MN: needs tile_w dividing 64 and out_h*tile_w %% 64 == 0");
    if (d->K != g.B * g.out_h * g.tile_w) return set_error(THEIA_ERR_ARG, "CONV_MN: K != B*out_h*tile_w");
    uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.B};
    uint64_t strides[3] = {(uint64_t)g.stride_w * 2, (uint64_t)g.stride_h * 2, (uint64_t)g.stride_b * 2};
    const uint32_t es = g.in_stride > 1 ? (uint32_t)g.in_stride : 1u;
    uint32_t box[4] = {64, es * (uint32_t)g.tile_w, es * (uint32_t)(64 / g.tile_w), 1};
    uint32_t estr[4] = {1, es, es, 1};
    rc = encode_tensor_map(&tmB, d->B, 4, dims, strides, box, estr);
    k.rows_per_kb = 64 / g.tile_w;
    k.kb_per_img = (g.out_h * g.tile_w) / 64;
    if (k.batch_z != g.ntaps) return set_error(THEIA_ERR_ARG, "CONV_MN: batch_z must equal ntaps");
  } else {
    return set_error(THEIA_ERR_ARG, "bad b_mode");
  }
  if (rc) return rc;
  for (int i = 0; i < 9; ++i) k.dh[i] = g.dh[i], k.dw[i] = g.dw[i];
  k.es = g.in_stride > 1 ? g.in_stride : 1;

  k.num_kb = (d->K + BK - 1) / BK;
  if (k.splits > k.num_kb) k.splits = k.num_kb;
  k.kb_per_split = (k.num_kb + k.splits - 1) / k.splits;
  k.splits = (k.num_kb + k.kb_per_split - 1) / k.kb_per_split;  // no empty split
  k.m_tiles = (d->M + BM - 1) / BM;
  k.n_tiles = (d->N + bn - 1) / bn;
  k.total_items = k.m_tiles * k.n_tiles * k.batch_z * k.splits;
  k.idesc = make_idesc_bf16(BM, bn, a_mn, b_mn);
  k.a_lbo = a_mn ? 8192 : 16;
  k.a_sbo = 1024;
  k.a_kstep = a_mn ? 2048 : 32;
  k.b_lbo = b_mn ? 8192 : 16;
  k.b_sbo = 1024;
  k.b_kstep = b_mn ? 2048 : 32;
  // bring-up overrides
  if (g_dbg[1]) k.a_lbo = (uint32_t)g_dbg[1];
  if (g_dbg[2]) k.a_sbo = (uint32_t)g_dbg[2];
  if (g_dbg[3]) k.b_lbo = (uint32_t)g_dbg[3];
  if (g_dbg[4]) k.b_sbo = (uint32_t)g_dbg[4];
  if (g_dbg[5]) k.a_kstep = (uint32_t)g_dbg[5];
  if (g_dbg[6]) k.b_kstep = (uint32_t)g_dbg[6];

  if (bn == 128) return dispatch<128>(tmA, tmB, k, stream);
  if (bn == 192) return dispatch<192>(tmA, tmB, k, stream);
  return dispatch<256>(tmA, tmB, k, stream);
}
