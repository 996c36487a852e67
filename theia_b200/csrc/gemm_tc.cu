// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// One persistent, warp-specialised kernel:
//   warp 0   TMA producer   (cp.async.bulk.tensor 2-D / 4-D, 128-byte swizzle, OOB zero fill
//                            = the convolution padding)
//   warp 1   MMA issuer     (one thread, tcgen05.mma.kind::f16: cta_group::1, 128 x BN x 16 per CTA -- or, PAIR = 2,
//                            cta_group::2, 256 x BN x 16 per cluster of two CTAs, issued by the leader CTA only)
//   warp 2   TMEM allocator
//   warps 4-11 epilogue     (tcgen05.ld 32x32b: every thread owns one output row and 32 consecutive columns
//                            of a chunk; the fused epilogue runs on that layout and stores with 256-bit
//                            st.global (one full 32-byte sector per lane); two warps per TMEM lane quarter,
//                            alternating chunks; epilogue flags are compile-time)
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand modes (include/theia_b200.h): K-major 2-D, MN-major 2-D (wgrad), K-major NHWC
// gather with the taps folded into K (forward / dgrad convolutions), MN-major NHWC gather of
// one tap (convolution wgrad, tap = z slice).
#include <stdlib.h>
#include <cuda.h>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

struct GemmK {
  int M, N;
  int num_kb, kb_per_split, splits, batch_z;
  int m_tiles, n_tiles, total_items;
  int m_units;  // m_tiles / CTAs per MMA (rounded up): the M extent of the tile scheduler
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
  int dbg;  // diagnostics (theia_debug_set key 7): 1 = no TMA loads after the first ring fill, 2 = skip the epilogue
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 128 + EPI_WARPS * 32;
constexpr int BIAS_BYTES = EPI_WARPS * 4 * 32 * 4;  // per-warp bias slice of the current tile (<= 4 chunks x 32 floats)
constexpr int SMEM_LIMIT = 232448;  // 227 KB

constexpr int AUX_SLOTS = 3;                                  // per-warp ring depth (32x32 bf16 chunks)
constexpr int AUX_RING_BYTES = EPI_WARPS * AUX_SLOTS * 2048;  // 48 KB
constexpr int AUXF = THEIA_EPI_RESID | THEIA_EPI_MUL_AUX | THEIA_EPI_MUL_RELUMASK;
// 8-bit code of gelu'(x) in [-0.1289, 1.1289] (THEIA_EPI_AUX_U8): q = rint((g + 0.129) * 255 / 1.258), step 4.9e-3.
// The fc1 GEMM writes two [M, 4D] tensors and is bound by HBM WRITE bandwidth (K / 2 = 384 flop per byte written);
// one byte instead of two for the derivative lifts that bound by a third and halves the dgrad's aux read.
constexpr float GD_SCALE = 255.0f / 1.258f, GD_BIAS = 0.129f * GD_SCALE, GD_ISCALE = 1.258f / 255.0f, GD_IBIAS = -0.129f;
// aux operands that go through the per-warp cp.async ring (the u8 codes are 32 bytes per row chunk: direct loads)
__host__ __device__ constexpr bool epi_uses_ring(int e) { return e >= 0 && (e & AUXF) != 0 && (e & THEIA_EPI_AUX_U8) == 0; }

// PAIR = CTAs per MMA: 1 = cta_group::1 (128 x BN tile per CTA), 2 = cta_group::2 (a cluster of two CTAs on
// one TPC computes a 256 x BN tile; each CTA stages its own 128 A rows and HALF of the B tile, so the operand
// traffic per CTA -- L2 -> SM and shared-memory writes -- drops from 48 to 32 KB per 64-deep K block).
template <int BN, bool RING, int PAIR>
struct Cfg {
  static constexpr int B_ROWS = BN / PAIR;  // B rows (K-major) / columns (MN-major) staged by one CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EXTRA = BIAS_BYTES + (RING ? AUX_RING_BYTES : 0);
  static constexpr int STAGES = (SMEM_LIMIT - EXTRA - 1024 - 256) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EXTRA + 256 + 1024;
};

struct Item {
  int m_blk, n_blk, z, kb0, kb1;
};

// m_blk is returned in scheduling units (CTA pairs under cta_group::2): the caller adds its cluster rank
__device__ __forceinline__ Item decode_item(const GemmK& p, int item) {
  Item it;
  it.n_blk = item % p.n_tiles;
  int t = item / p.n_tiles;
  it.m_blk = t % p.m_units;
  t /= p.m_units;
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
#ifdef THEIA_GELU_AS7125
  // Abramowitz-Stegun 7.1.25 (three terms, |erf err| <= 2.5e-5): two FMAs fewer per element
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float ax = fabsf(x[k]);
    t[k] = rcp_approx(fmaf(ax, 0.47047f * 0.70710678118654752f, 1.0f));
    const float u = ax * 0.84932180028801907f;
    e[k] = ex2_approx(-(u * u));
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float pl = fmaf(t[k], 0.5f * 0.7478556f, 0.5f * -0.0958798f);
    pl = fmaf(pl, t[k], 0.5f * 0.3480242f);
    const float hq = pl * t[k] * e[k];  // 0.5 * erfc(|x| / sqrt 2)
    cdf[k] = x[k] > 0.f ? 1.0f - hq : hq;
    if (WANT_PDF) pdf[k] = 0.39894228040143268f * e[k];
  }
#else
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
    cdf[k] = x[k] > 0.f ? 1.0f - hq : hq;  // one FADD + FSEL (was 0.5 + copysign(0.5 - hq, x): three)
    if (WANT_PDF) pdf[k] = 0.39894228040143268f * e[k];
  }
#endif
}

// EPI_CT >= 0: epilogue flags are a compile-time constant (hot combinations); -1: read p.epi.
template <int BN, int EPI_CT, int PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmK p) {
  // compile-time specialised kernels that read an aux operand stage it through a cp.async ring in shared
  // memory (48 KB in flight per SM: registers alone cannot keep enough HBM reads outstanding)
  constexpr bool RING = epi_uses_ring(EPI_CT);
  using C = Cfg<BN, RING, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* bias_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint8_t* aux_ring_all = smem + C::STAGES * C::STAGE_BYTES + BIAS_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::EXTRA);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // scheduling unit = the CTA (PAIR 1) or the two-CTA cluster (PAIR 2); rank 0 of a pair is the MMA leader:
  // it owns the full / tempty barriers both CTAs signal, and its commits are multicast to both CTAs
  const int cta_rank = (PAIR == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit0 = (PAIR == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int nunits = (PAIR == 2) ? (gridDim.x >> 1) : gridDim.x;

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
      mbar_init(&tempty[s], EPI_WARPS * PAIR);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_g<PAIR>(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_g<PAIR>();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();  // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    {  // whole warp, elected lane issues (uniform operands stay in uniform registers; see the MMA issuer)
      int stage = 0;
      uint32_t phase = 0;
      int filled = 0;
      for (int item = unit0; item < p.total_items; item += nunits) {
        const Item it = decode_item(p, item);
        const int m_blk = it.m_blk * PAIR + cta_rank;
        const int m0 = m_blk * BM, n0 = it.n_blk * BN + cta_rank * C::B_ROWS;  // this CTA's share of the operands
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if ((p.dbg & 1) && filled >= C::STAGES) {  // diagnostic: MMA rate without operand traffic
            if (cta_rank == 0 && lane == 0) mbar_arrive(&full[stage]);
            __syncwarp();
            if (++stage == C::STAGES) stage = 0, phase ^= 1;
            continue;
          }
          ++filled;
          if (elect_one_sync()) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          // both CTAs of a pair complete their bytes on the LEADER's barrier, which expects the sum
          if (cta_rank == 0) mbar_expect_tx(&full[stage], C::STAGE_BYTES * PAIR);
          const uint32_t fb = (PAIR == 2) ? mapa_u32(smem_u32(&full[stage]), 0) : smem_u32(&full[stage]);
          // ---- A ----
          if (p.a_mode == THEIA_OP_K2D) {
            tma_load_2d_g<PAIR>(&tmA, sa, fb, kb * BK, m0);
          } else if (p.a_mode == THEIA_OP_MN2D) {
            tma_load_2d_g<PAIR>(&tmA, sa, fb, m0, kb * BK);
            tma_load_2d_g<PAIR>(&tmA, sa + 8192, fb, m0 + 64, kb * BK);
          } else {  // CONV_K
            const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
            const int b = m_blk / p.tiles_per_img, ht = m_blk - b * p.tiles_per_img;
            tma_load_4d_g<PAIR>(&tmA, sa, fb, cc * 64, p.dw[tap], p.es * ht * p.tile_h + p.dh[tap], b);
          }
          // ---- B ----
          if (p.b_mode == THEIA_OP_K2D) {
            if (p.b_tap_rows > 0) {  // tap-major weight pack [tap][rows][C]
              const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
              tma_load_2d_g<PAIR>(&tmB, sb, fb, cc * 64, p.wtap[tap] * p.b_tap_rows + n0);
            } else {
              tma_load_2d_g<PAIR>(&tmB, sb, fb, kb * BK, n0);
            }
          } else if (p.b_mode == THEIA_OP_MN2D) {
#pragma unroll
            for (int i = 0; i < C::B_ROWS / 64; ++i) tma_load_2d_g<PAIR>(&tmB, sb + i * 8192, fb, n0 + 64 * i, kb * BK);
          } else {  // CONV_MN: tap = z
            const int b = kb / p.kb_per_img, hb = kb - b * p.kb_per_img;
#pragma unroll
            for (int i = 0; i < C::B_ROWS / 64; ++i)
              tma_load_4d_g<PAIR>(&tmB, sb + i * 8192, fb, n0 + 64 * i, p.dw[it.z],
                                  p.es * hb * p.rows_per_kb + p.dh[it.z], b);
          }
          }  // elected lane
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The WHOLE warp runs this loop (warp-uniform control flow, every value below is uniform), one elected lane
    // issues: the compiler then keeps the descriptors in uniform registers.  With `if (lane == 0)` around the loop
    // it could not prove uniformity and wrapped every tcgen05.mma in an ELECT / R2UR broadcast loop plus the full
    // 64-bit descriptor arithmetic -- ~35 SASS instructions per MMA on ONE thread, as long as the 64-cycle
    // execution of a 128 x 256 x 16 MMA (r2 ncu source view).  The descriptor's high word is constant and the low
    // word (address >> 4 | LBO << 16) only advances by constants: two 32-bit adds per MMA.
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t a_hi = ((p.a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t b_hi = ((p.b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem0 >> 4) & 0x3FFFu) | (((p.a_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = (((smem0 + A_BYTES) >> 4) & 0x3FFFu) | (((p.b_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t a_step = p.a_kstep >> 4, b_step = p.b_kstep >> 4;
      constexpr uint32_t STAGE16 = C::STAGE_BYTES >> 4;
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      for (int item = unit0; item < p.total_items; item += nunits) {
        const Item it = decode_item(p, item);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = it.kb0; kb < it.kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + k * a_step);
              const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + k * b_step);
              tc_mma_bf16_g<PAIR>(d_tmem, da, db, p.idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
            }
            tc_commit_g<PAIR>(&empty[stage]);  // smem slot (of both CTAs) reusable once these MMAs have read it
          }
          __syncwarp();
          a_lo += STAGE16, b_lo += STAGE16;
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
            a_lo = a_lo0, b_lo = b_lo0;
          }
        }
        if (elect_one_sync()) tc_commit_g<PAIR>(&tfull[acc]);  // accumulator complete (each CTA drains its own 128 TMEM lanes)
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    // Row-per-thread layout straight out of TMEM: lane l of the warp owns tile row 32*wq + l and the 32
    // consecutive columns of the current chunk.  No shared-memory transpose: a lane's 32 outputs are 64
    // contiguous bytes (bf16) written as two 256-bit stores (full sectors); per-row addressing is computed
    // once per tile.  Aux operands arrive through a per-warp cp.async ring in the same layout.
    const int ew = warp - 4;      // 0..7
    const int wq = ew & 3;        // TMEM lane quarter == warp % 4
    const int hsel = ew >> 2;     // this warp takes the 32-column chunks with (chunk & 1) == hsel
    const int r = wq * 32 + lane; // tile row of this lane
    int acc = 0;
    uint32_t acc_phase = 0;
    const int epi = (EPI_CT >= 0) ? EPI_CT : p.epi;
    constexpr int NCHUNK = BN / 32;
    const bool wide = (p.ldo & 15) == 0;  // rows start 32-byte aligned: 256-bit accesses allowed
    auto map_row = [&](int m_blk, long long& off, bool& ok, int& img) {
      if (p.conv_out) {
        img = m_blk / p.tiles_per_img;
        const int ht = m_blk - img * p.tiles_per_img;
        const int hh = r / p.tile_w, ww = r - hh * p.tile_w;
        const int h = ht * p.tile_h + hh;
        ok = (h < p.out_h) && (ww < p.out_w) && (m_blk < p.m_tiles);
        off = (static_cast<long long>(img) * p.out_img_rows + p.out_row_off +
               static_cast<long long>(h * p.sy + p.py) * p.out_wpitch + (ww * p.sx + p.px)) * p.ldo;
      } else {
        img = 0;
        const long long m = static_cast<long long>(m_blk) * BM + r;
        ok = m < p.M;
        off = m * p.ldo;
      }
    };
    // ---- aux ring (RING kernels): chunk q of this warp lives in slot q % AUX_SLOTS ([lane][4 x 16 B], the
    // 16-byte index XOR-swizzled against bank conflicts); the prefetch cursor (pitem, pch) runs AUX_SLOTS chunks
    // ahead of the processing position, across tile boundaries ----
    uint8_t* ring = aux_ring_all + ew * (AUX_SLOTS * 2048);
    const int rsw = (lane >> 1) & 3;
    int pitem = unit0, pch = hsel, slot = 0;
    auto ring_issue = [&](int sl) {
      if (pitem < p.total_items) {
        const Item pit = decode_item(p, pitem);
        long long poff;
        bool pok;
        int pimg;
        map_row(pit.m_blk * PAIR + cta_rank, poff, pok, pimg);
        const int nn = pit.n_blk * BN + pch * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t dst = smem_u32(ring + sl * 2048 + lane * 64 + ((j ^ rsw) << 4));
          if (pok && nn + 8 * j < p.N) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p.aux + poff + nn + 8 * j) : "memory");
          } else {
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
          }
        }
        pch += 2;
        if (pch >= NCHUNK) {
          pch = hsel;
          pitem += nunits;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (RING) {
#pragma unroll
      for (int sidx = 0; sidx < AUX_SLOTS; ++sidx) ring_issue(sidx);
    }
    const uint32_t te0 = (PAIR == 2) ? mapa_u32(smem_u32(&tempty[0]), 0) : smem_u32(&tempty[0]);
    const uint32_t te1 = (PAIR == 2) ? mapa_u32(smem_u32(&tempty[1]), 0) : smem_u32(&tempty[1]);
    // per-warp bias slice of the current tile ([chunk][32] floats)
    const uint32_t sbias = smem_u32(bias_all + ew * 128);
    if (unit0 < p.total_items) {
      const int n_blk0 = unit0 % p.n_tiles;
#pragma unroll
      for (int ci = 0; ci < NCHUNK / 2; ++ci) {
        const int n = n_blk0 * BN + (hsel + 2 * ci) * 32 + lane;
        sts32(sbias + (ci * 32 + lane) * 4, __float_as_uint((p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f));
      }
      __syncwarp();
    }
    for (int item = unit0; item < p.total_items; item += nunits) {
      Item it = decode_item(p, item);
      it.m_blk = it.m_blk * PAIR + cta_rank;
      long long rowoff;
      bool rok;
      int img;
      map_row(it.m_blk, rowoff, rok, img);
      // the NEXT tile's bias slice is fetched now (one value per lane and chunk) and parked in shared memory
      // after this tile's chunks are done: its global-load latency hides behind a whole tile of work
      float nbias[NCHUNK / 2];
      {
        const int nitem = item + nunits;
        const int nn_blk = (nitem < p.total_items) ? nitem % p.n_tiles : 0;
#pragma unroll
        for (int ci = 0; ci < NCHUNK / 2; ++ci) {
          const int n = nn_blk * BN + (hsel + 2 * ci) * 32 + lane;
          nbias[ci] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (p.dbg & 2) {  // diagnostic: mainloop rate without an epilogue
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc ? te1 : te0);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      float st_s = 0.f, st_ss = 0.f;
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BN;
      uint32_t vbuf[2][32];  // TMEM loads are software pipelined: chunk ci+1 is in flight while ci is processed
      tmem_ld32(tacc + hsel * 32, vbuf[0]);
      // 8-bit aux codes (32 bytes per lane and chunk) are fetched one chunk ahead as well
      uint4 a8[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)}, a8n[2] = {a8[0], a8[0]};
      auto load_a8 = [&](int ch, uint4 (&dst)[2]) {
        const int nb = it.n_blk * BN + ch * 32;
        if (rok && nb + 32 <= p.N) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.aux) + rowoff + nb);
          dst[0] = __ldg(src), dst[1] = __ldg(src + 1);
        }
      };
      if ((epi & THEIA_EPI_MUL_AUX) && (epi & THEIA_EPI_AUX_U8)) load_a8(hsel, a8n);
#pragma unroll
      for (int ci = 0; ci < NCHUNK / 2; ++ci) {
        const int ch = hsel + 2 * ci;
        const bool last = (ci + 1 == NCHUNK / 2);
        const int nbase = it.n_blk * BN + ch * 32;
        uint32_t(&v)[32] = vbuf[ci & 1];
        tmem_ld_wait();
        if (!last) {
          tmem_ld32(tacc + (ch + 2) * 32, vbuf[(ci + 1) & 1]);
        } else {  // all TMEM reads of this warp for this accumulator are done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc ? te1 : te0);
        }
        if (RING) asm volatile("cp.async.wait_group %0;" ::"n"(AUX_SLOTS - 1) : "memory");  // this chunk's aux landed
        if ((epi & THEIA_EPI_MUL_AUX) && (epi & THEIA_EPI_AUX_U8)) {
          a8[0] = a8n[0], a8[1] = a8n[1];
          if (!last) load_a8(ch + 2, a8n);
        }
        const int ncols = max(0, min(32, p.N - nbase));  // multiple of 8; 0 = nothing to store in this chunk
        float xv[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) xv[k] = __uint_as_float(v[k]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 b4 = lds128(sbias + (ci * 32 + 4 * j) * 4);  // smem broadcast
            xv[4 * j + 0] += __uint_as_float(b4.x), xv[4 * j + 1] += __uint_as_float(b4.y);
            xv[4 * j + 2] += __uint_as_float(b4.z), xv[4 * j + 3] += __uint_as_float(b4.w);
          }
        }
        if (epi & THEIA_EPI_POSCLS) {
          if (rok) {
            const int t = static_cast<int>((static_cast<long long>(it.m_blk) * BM + r) % p.tokens);
            const bool patch = (t >= p.tok_p0) && (t < p.tok_p1);
            const float* tab = p.pos + static_cast<long long>(t) * p.N + nbase;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (4 * j < ncols) {
                const float4 ps = *reinterpret_cast<const float4*>(tab + 4 * j);
                if (patch) {
                  xv[4 * j + 0] += ps.x, xv[4 * j + 1] += ps.y, xv[4 * j + 2] += ps.z, xv[4 * j + 3] += ps.w;
                } else {  // CLS / register token: the table holds the whole value
                  xv[4 * j + 0] = ps.x, xv[4 * j + 1] = ps.y, xv[4 * j + 2] = ps.z, xv[4 * j + 3] = ps.w;
                }
              }
            }
          }
        }
        uint32_t pk[16];
        auto store_bf16_row = [&](bf16* base) {  // pk[16] = this lane's 32 bf16 outputs
          if (!rok) return;
          bf16* dst = base + rowoff + nbase;
          if (wide && ncols == 32) {
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                         : "memory");
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16), "r"(pk[8]),
                         "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
                         : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < ncols) *reinterpret_cast<uint2*>(dst + 4 * j) = make_uint2(pk[2 * j], pk[2 * j + 1]);
          }
        };
        if (epi & THEIA_EPI_GELU) {
          // gelu(x) = x Phi(x); its derivative Phi(x) + x phi(x) is stored (bf16) for the backward pass, so
          // the dgrad epilogue is a plain multiply instead of a second erf/exp evaluation
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float xs[16], cdf[16], pdf[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) xs[k] = xv[16 * hh + k];
            normal_cdf_pdf<16, true>(xs, cdf, pdf);
            if (epi & THEIA_EPI_AUX_U8) {  // four 8-bit codes per word: pk[4 hh .. 4 hh + 3]
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                uint32_t c[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float g = fmaf(xs[4 * k4 + e], pdf[4 * k4 + e], cdf[4 * k4 + e]);
                  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(c[e]) : "f"(fmaf(g, GD_SCALE, GD_BIAS)));
                }
                pk[4 * hh + k4] = __byte_perm(__byte_perm(c[0], c[1], 0x0040), __byte_perm(c[2], c[3], 0x0040), 0x5410);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                pk[8 * hh + k] = pack_bf16x2(fmaf(xs[2 * k], pdf[2 * k], cdf[2 * k]),
                                             fmaf(xs[2 * k + 1], pdf[2 * k + 1], cdf[2 * k + 1]));
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) xv[16 * hh + k] = xs[k] * cdf[k];
          }
          if (epi & THEIA_EPI_AUX_U8) {  // 32 bytes per lane: one full sector (N % 32 == 0 is checked by the launcher)
            if (rok && ncols == 32) {
              uint8_t* d8 = reinterpret_cast<uint8_t*>(p.out2) + rowoff + nbase;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d8), "r"(pk[0]), "r"(pk[1]),
                           "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                           : "memory");
            }
          } else {
            store_bf16_row(reinterpret_cast<bf16*>(p.out2));
          }
        }
        if (epi & THEIA_EPI_GELU_FWD) {  // inference: gelu(x) = x Phi(x), no derivative output
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float xs[16], cdf[16], pdf[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) xs[k] = xv[16 * hh + k];
            normal_cdf_pdf<16, false>(xs, cdf, pdf);
#pragma unroll
            for (int k = 0; k < 16; ++k) xv[16 * hh + k] = xs[k] * cdf[k];
          }
        }
        if (epi & THEIA_EPI_QUICK_GELU) {  // x sigmoid(1.702 x)  (hf:activations.py QuickGELUActivation; CLIP)
#pragma unroll
          for (int k = 0; k < 32; ++k)
            xv[k] = xv[k] * rcp_approx(1.0f + ex2_approx(xv[k] * (-1.702f * 1.4426950408889634f)));
        }
        if (epi & THEIA_EPI_RELU) {
#pragma unroll
          for (int k = 0; k < 32; ++k) xv[k] = fmaxf(xv[k], 0.f);
        }
        if ((epi & THEIA_EPI_MUL_AUX) && (epi & THEIA_EPI_AUX_U8)) {  // v *= gelu' decoded from the 8-bit codes
          const uint32_t w8[8] = {a8[0].x, a8[0].y, a8[0].z, a8[0].w, a8[1].x, a8[1].y, a8[1].z, a8[1].w};
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float code = static_cast<float>((w8[k >> 2] >> (8 * (k & 3))) & 0xffu);
            xv[k] *= fmaf(code, GD_ISCALE, GD_IBIAS);
          }
        }
        if ((epi & AUXF) && !(epi & THEIA_EPI_AUX_U8)) {
          uint32_t au[16];
          if (RING) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 q = lds128(smem_u32(ring + slot * 2048 + lane * 64 + ((j ^ rsw) << 4)));
              au[4 * j + 0] = q.x, au[4 * j + 1] = q.y, au[4 * j + 2] = q.z, au[4 * j + 3] = q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint2 q = make_uint2(0u, 0u);
              if (rok && 4 * j < ncols) q = *reinterpret_cast<const uint2*>(p.aux + rowoff + nbase + 4 * j);
              au[2 * j] = q.x, au[2 * j + 1] = q.y;
            }
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float2 a2 = unpack_bf16x2(au[k]);
            if (epi & THEIA_EPI_MUL_AUX) {
              xv[2 * k] *= a2.x, xv[2 * k + 1] *= a2.y;
            } else if (epi & THEIA_EPI_MUL_RELUMASK) {
              xv[2 * k] = a2.x > 0.f ? xv[2 * k] : 0.f, xv[2 * k + 1] = a2.y > 0.f ? xv[2 * k + 1] : 0.f;
            } else {
              xv[2 * k] += a2.x, xv[2 * k + 1] += a2.y;
            }
          }
        }
        if (epi & THEIA_EPI_RESID_F32) {  // fp32 residual stream (teacher inference): aux is fp32, indexed like out
          if (rok) {
            const float* a32 = reinterpret_cast<const float*>(p.aux) + rowoff + nbase;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (4 * j < ncols) {
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(a32 + 4 * j));
                xv[4 * j + 0] += q4.x, xv[4 * j + 1] += q4.y, xv[4 * j + 2] += q4.z, xv[4 * j + 3] += q4.w;
              }
            }
          }
        }
        if (RING) {  // aux of this chunk is consumed: refill the slot with the chunk AUX_SLOTS ahead
          ring_issue(slot);
          slot = (slot + 1 == AUX_SLOTS) ? 0 : slot + 1;
        }
        if (epi & THEIA_EPI_ATOMIC) {
          if (rok) {
            float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(it.z) * p.out_z_stride + rowoff + nbase;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (4 * j < ncols)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(xv[4 * j]),
                             "f"(xv[4 * j + 1]), "f"(xv[4 * j + 2]), "f"(xv[4 * j + 3])
                             : "memory");
          }
        } else if (epi & THEIA_EPI_OUT_F32) {
          if (rok) {
            float* o = reinterpret_cast<float*>(p.out) + rowoff + nbase;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (8 * j < ncols) {
                if (wide) {
                  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 8 * j), "f"(xv[8 * j]),
                               "f"(xv[8 * j + 1]), "f"(xv[8 * j + 2]), "f"(xv[8 * j + 3]), "f"(xv[8 * j + 4]),
                               "f"(xv[8 * j + 5]), "f"(xv[8 * j + 6]), "f"(xv[8 * j + 7])
                               : "memory");
                } else {
                  *reinterpret_cast<float4*>(o + 8 * j) = make_float4(xv[8 * j], xv[8 * j + 1], xv[8 * j + 2], xv[8 * j + 3]);
                  *reinterpret_cast<float4*>(o + 8 * j + 4) =
                      make_float4(xv[8 * j + 4], xv[8 * j + 5], xv[8 * j + 6], xv[8 * j + 7]);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) pk[k] = pack_bf16x2(xv[2 * k], xv[2 * k + 1]);
          store_bf16_row(reinterpret_cast<bf16*>(p.out));
          if (epi & (THEIA_EPI_STATS | THEIA_EPI_COLSUM)) {
            float q[32];  // the stored (rounded) values; rows / columns that are not stored count as zero
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float2 q2 = unpack_bf16x2(pk[k]);
              const bool cok = rok && (2 * k < ncols);
              q[2 * k] = cok ? q2.x : 0.f, q[2 * k + 1] = cok ? q2.y : 0.f;
            }
            if (epi & THEIA_EPI_STATS) {
#pragma unroll
              for (int k = 0; k < 32; ++k) st_s += q[k], st_ss += q[k] * q[k];
            }
            if (epi & THEIA_EPI_COLSUM) {
              // warp reduce-scatter: after the butterfly lane l holds the sum over the warp's 32 rows of
              // column l of the chunk (bias gradient of the consumer), one coalesced red per warp and chunk
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < off; ++i) {
                  const float send = upper ? q[i] : q[i + off];
                  const float keep = upper ? q[i + off] : q[i];
                  q[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              if (lane < ncols) atomicAdd(p.colsum + nbase + lane, q[0]);
            }
          }
        }
      }
      __syncwarp();  // every lane has read this tile's bias slice
#pragma unroll
      for (int ci = 0; ci < NCHUNK / 2; ++ci) sts32(sbias + (ci * 32 + lane) * 4, __float_as_uint(nbias[ci]));
      __syncwarp();
      if (epi & THEIA_EPI_STATS) {
        st_s = warp_sum(st_s);
        st_ss = warp_sum(st_ss);
        if (lane == 0 && it.m_blk < p.m_tiles) {
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
  if (PAIR == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_g<PAIR>(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static long long g_dbg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

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

template <int BN, int EPI_CT, int PAIR>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmK& k, cudaStream_t stream) {
  using C = Cfg<BN, epi_uses_ring(EPI_CT), PAIR>;
  static PerDeviceOnce attr_once;
  if (attr_once.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI_CT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
  }
  const int max_units = num_sms() / PAIR;
  const int grid = (k.total_items < max_units ? k.total_items : max_units) * PAIR;
  const bool prof = g_prof_on && g_prof_n < PROF_RING;
  if (prof) cudaEventRecord(g_ev0[g_prof_n], stream);
  if (PAIR == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(NTHREADS), cfg.dynamicSmemBytes = C::SMEM_BYTES, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI_CT, PAIR>, tmA, tmB, k);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "gemm cluster launch: %s", cudaGetErrorString(e));
  } else {
    gemm_tc_kernel<BN, EPI_CT, PAIR><<<grid, NTHREADS, C::SMEM_BYTES, stream>>>(tmA, tmB, k);
  }
  if (prof) {
    cudaEventRecord(g_ev1[g_prof_n], stream);
    g_prof_flops[g_prof_n] = 2.0 * k.M * (double)k.N * (double)k.num_kb * BK * k.batch_z;
    int* mt = g_prof_meta[g_prof_n];
    mt[0] = k.M, mt[1] = k.N, mt[2] = k.num_kb * BK, mt[3] = k.a_mode, mt[4] = k.b_mode, mt[5] = k.epi;
    mt[6] = k.splits * 1000 + k.batch_z, mt[7] = BN + (PAIR == 2 ? 1 : 0);  // odd BN = CTA-pair kernel
    ++g_prof_n;
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
  return THEIA_OK;
}

// hot epilogue combinations get a compile-time specialisation; anything else runs the generic kernel
template <int BN, int PAIR>
static int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmK& k, cudaStream_t stream) {
  switch (k.epi) {
#define THEIA_EPI_CASE(E) \
  case (E):               \
    return launch<BN, (E), PAIR>(tmA, tmB, k, stream);
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
    THEIA_EPI_CASE(THEIA_EPI_GELU_FWD)
    THEIA_EPI_CASE(THEIA_EPI_QUICK_GELU)
    THEIA_EPI_CASE(THEIA_EPI_GELU | THEIA_EPI_AUX_U8)
    THEIA_EPI_CASE(THEIA_EPI_MUL_AUX | THEIA_EPI_COLSUM | THEIA_EPI_AUX_U8)
    THEIA_EPI_CASE(THEIA_EPI_RESID_F32 | THEIA_EPI_OUT_F32)
    THEIA_EPI_CASE(THEIA_EPI_POSCLS | THEIA_EPI_OUT_F32)
#undef THEIA_EPI_CASE
    default:
      return launch<BN, -1, PAIR>(tmA, tmB, k, stream);
  }
}

}  // namespace theia

using namespace theia;

// CTAs per MMA for a GEMM with this N tile and M extent: the 256-wide tile runs as a CTA pair (cta_group::2)
// whenever there are two M tiles to pair up.  theia_debug_set(8, 1) or the
// environment variable THEIA_GEMM_SINGLE_CTA forces single-CTA kernels, theia_debug_set(8, 2) pairs everywhere (A/B timing).
int theia::gemm_pair_mode(int bn, int m_tiles, int a_mode) {
  static const bool env_single = getenv("THEIA_GEMM_SINGLE_CTA") != nullptr;
  // measured (r01, B200): the pair wins 2-14 % on every 2-D operand shape of the step, but loses ~8 % on the
  // implicit-GEMM convolutions whose A operand is the 4-D TMA gather -- the two gathers of a pair have to
  // finish in lockstep -- so those keep the single-CTA kernel
  static const bool env_pair_conv = getenv("THEIA_GEMM_PAIR_CONV") != nullptr;  // A/B switch for tuning runs
  if (a_mode == THEIA_OP_CONV_K && g_dbg[8] != 2 && !env_pair_conv) return 1;
  return (bn == 256 && m_tiles >= 2 && g_dbg[8] != 1 && !env_single) ? 2 : 1;
}

extern "C" int theia_debug_set(int key, long long value) {
  if (key == 0) {
    for (auto& v : g_dbg) v = 0;
    return 0;
  }
  if (key < 0 || key >= 16) return THEIA_ERR_ARG;
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

  const int pair = gemm_pair_mode(bn, (d->M + BM - 1) / BM, d->a_mode);
  const int bbox = bn / pair;  // B rows / columns one CTA stages

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
  if ((d->epi & THEIA_EPI_AUX_U8) &&
      (!(d->epi & (THEIA_EPI_GELU | THEIA_EPI_MUL_AUX)) || (d->epi & (THEIA_EPI_RESID | THEIA_EPI_MUL_RELUMASK)) || d->N % 32 != 0 ||
       d->ldo % 32 != 0))
    return set_error(THEIA_ERR_ARG, "AUX_U8 goes with GELU (out2) or MUL_AUX (aux) only, N and ldo multiples of 32");
  if ((d->epi & THEIA_EPI_RESID_F32) && (!(d->epi & THEIA_EPI_OUT_F32) || (d->epi & AUXF) || (d->ldo & 3)))
    return set_error(THEIA_ERR_ARG, "RESID_F32 goes with OUT_F32 only (fp32 aux, no bf16 aux flag)");
  if ((d->epi & (THEIA_EPI_RESID | THEIA_EPI_MUL_AUX | THEIA_EPI_MUL_RELUMASK | THEIA_EPI_RESID_F32)) && !d->aux)
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
    rc = encode_2d(&tmB, d->B, (uint64_t)g.C, (uint64_t)9 * g.b_tap_rows, (uint64_t)g.C, 64, bbox);
    k.b_tap_rows = g.b_tap_rows;
    for (int i = 0; i < 9; ++i) k.wtap[i] = g.wtap[i];
  } else if (d->b_mode == THEIA_OP_K2D) {
    rc = encode_2d(&tmB, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, 64, bbox);
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
  k.m_units = (k.m_tiles + pair - 1) / pair;
  k.total_items = k.m_units * k.n_tiles * k.batch_z * k.splits;
  k.idesc = make_idesc_bf16(BM * pair, bn, a_mn, b_mn);
  k.a_lbo = a_mn ? 8192 : 16;
  k.a_sbo = 1024;
  k.a_kstep = a_mn ? 2048 : 32;
  k.b_lbo = b_mn ? 8192 : 16;
  k.b_sbo = 1024;
  k.b_kstep = b_mn ? 2048 : 32;
  // bring-up overrides
  k.dbg = (int)g_dbg[7];
  if (g_dbg[1]) k.a_lbo = (uint32_t)g_dbg[1];
  if (g_dbg[2]) k.a_sbo = (uint32_t)g_dbg[2];
  if (g_dbg[3]) k.b_lbo = (uint32_t)g_dbg[3];
  if (g_dbg[4]) k.b_sbo = (uint32_t)g_dbg[4];
  if (g_dbg[5]) k.a_kstep = (uint32_t)g_dbg[5];
  if (g_dbg[6]) k.b_kstep = (uint32_t)g_dbg[6];

  if (bn == 128) return dispatch<128, 1>(tmA, tmB, k, stream);
  if (bn == 192) return dispatch<192, 1>(tmA, tmB, k, stream);
  if (pair == 2) return dispatch<256, 2>(tmA, tmB, k, stream);
  return dispatch<256, 1>(tmA, tmB, k, stream);
}
