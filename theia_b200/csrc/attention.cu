// Non-causal multi-head self-attention for the 197-token ViT sequence (head_dim 64), forward
// and backward, one CTA per (image, head).  The whole K/V (and for backward Q/dO) of a head
// lives in shared memory (197 x 64 bf16 = 25 KB each), so softmax is single-pass: no online
// rescaling.  Replaces hf:models/vit/modeling_vit.py:171-196,232-246 (SDPA / eager attention).
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16 -> fp32) with ldmatrix from
// XOR-swizzled shared memory; attention is 4% (base) / 15% (tiny) of the step FLOPs.  The
// tcgen05 version (S and dP accumulators in TMEM) is the planned upgrade.
//
// Layout: qkv [B*N, 3*D] bf16 row-major (q | k | v, head h at columns h*64..h*64+63 of each),
// out / dout [B*N, D], lse [B, H, N] fp32 (natural-log-sum-exp of the scaled logits).
#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

constexpr int HD = 64;        // head dim
constexpr int NPAD = 208;     // 197 padded to 13 tiles of 16
constexpr int NTILE = NPAD / 16;
constexpr int ROWB = HD * 2;  // 128 bytes per smem row

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * ROWB + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Load rows [0,N) x 64 bf16 of one head (global row stride ld elements) into a swizzled smem
// tile of NPAD rows; padding rows are zero.
__device__ __forceinline__ void load_head_tile(uint8_t* sm, const bf16* g, long long ld, int N, int tid, int nthr) {
  for (int i = tid; i < NPAD * 8; i += nthr) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < N) v = *reinterpret_cast<const uint4*>(g + static_cast<long long>(r) * ld + c * 8);
    *reinterpret_cast<uint4*>(sm + swz(r, c)) = v;
  }
}

// A fragments (16 rows x 64 k) of a row tile: 4 k16 steps x 4 regs.
__device__ __forceinline__ void load_a_frags(uint32_t smbase, int row0, int lane, uint32_t (&a)[4][4]) {
  const int r = row0 + (lane & 15);
  const int chalf = lane >> 4;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(smbase + swz(r, ks * 2 + chalf), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}
// B fragments for C[m][n] = sum_k A[m][k] * T[n][k] (T stored [n][k] row-major, "non-trans"):
// 16 n-rows starting at n0, one k16 step ks -> (b0,b1) of n-tile n0 and (b0,b1) of n-tile n0+8.
__device__ __forceinline__ void load_b_nt(uint32_t smbase, int n0, int ks, int lane, uint32_t (&b)[4]) {
  const int r = n0 + (lane & 7) + ((lane >> 4) << 3);
  const int ch = ks * 2 + ((lane >> 3) & 1);
  ldsm_x4(smbase + swz(r, ch), b[0], b[1], b[2], b[3]);
}
// B fragments for C[m][n] = sum_k A[m][k] * T[k][n] (T stored [k][n] row-major, needs .trans):
// k16 rows starting at k0, 16 n-columns starting at n0 -> (b0,b1) of n-tile n0, (b0,b1) of n0+8.
__device__ __forceinline__ void load_b_t(uint32_t smbase, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int r = k0 + (lane & 7) + (((lane >> 3) & 1) << 3);
  const int ch = (n0 >> 3) + (lane >> 4);
  ldsm_x4_t(smbase + swz(r, ch), b[0], b[1], b[2], b[3]);
}

// --------------------------------------------------------------------------------------------
// forward
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                       float* __restrict__ lse, int N, int H, int D, float scale) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* sQ = sm;
  uint8_t* sK = sm + NPAD * ROWB;
  uint8_t* sV = sm + 2 * NPAD * ROWB;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ld = 3LL * D;
  const bf16* base = qkv + static_cast<long long>(b) * N * ld + h * HD;
  load_head_tile(sQ, base, ld, N, tid, 128);
  load_head_tile(sK, base + D, ld, N, tid, 128);
  load_head_tile(sV, base + 2 * D, ld, N, tid, 128);
  __syncthreads();
  const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV);
  const int g = lane >> 2, tig = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;

  for (int mt = warp; mt < NTILE; mt += 4) {
    const int row0 = mt * 16;
    if (row0 >= N) break;
    uint32_t aq[4][4];
    load_a_frags(uQ, row0, lane, aq);
    float s[NTILE * 2][4];
#pragma unroll
    for (int j = 0; j < NTILE * 2; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NTILE; ++kt) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4];
        load_b_nt(uK, kt * 16, ks, lane, bk);
        mma16816(s[2 * kt], aq[ks], bk[0], bk[1]);
        mma16816(s[2 * kt + 1], aq[ks], bk[2], bk[3]);
      }
    }
    // mask padded keys, row max
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NTILE * 2; ++j) {
      const int col = j * 8 + 2 * tig;
      if (col >= N) s[j][0] = -INFINITY, s[j][2] = -INFINITY;
      if (col + 1 >= N) s[j][1] = -INFINITY, s[j][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NTILE * 2; ++j) {
      s[j][0] = exp2f((s[j][0] - m0) * sl2);
      s[j][1] = exp2f((s[j][1] - m0) * sl2);
      s[j][2] = exp2f((s[j][2] - m1) * sl2);
      s[j][3] = exp2f((s[j][3] - m1) * sl2);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P V
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NTILE; ++kt) {
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
      ap[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
      ap[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      ap[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        uint32_t bv[4];
        load_b_t(uV, kt * 16, nd * 16, lane, bv);
        mma16816(o[2 * nd], ap, bv[0], bv[1]);
        mma16816(o[2 * nd + 1], ap, bv[2], bv[3]);
      }
    }
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = row0 + g, r1 = row0 + g + 8;
    bf16* ob = out + static_cast<long long>(b) * N * D + h * HD;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + 2 * tig;
      if (r0 < N) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * D + col) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
      if (r1 < N) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * D + col) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
    if (lse != nullptr && tig == 0) {
      float* lp = lse + (static_cast<long long>(b) * H + h) * N;
      if (r0 < N) lp[r0] = m0 * scale + logf(l0);
      if (r1 < N) lp[r1] = m1 * scale + logf(l1);
    }
  }
}

// --------------------------------------------------------------------------------------------
// backward:  dV = P^T dO ; dP = dO V^T ; dS = P o (dP - delta) ; dQ = scale dS K ; dK = scale dS^T Q
// Phase A (per 16-query tile): recompute S,P chunk by chunk, accumulate dQ.
// Phase B (per 16-key tile):   recompute S^T,P^T chunk by chunk, accumulate dK, dV.
// No atomics, no P/dS round trip through memory; S and dP are recomputed once (7 GEMM units vs 5).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                                       const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                       bf16* __restrict__ dqkv, int N, int H, int D, float scale) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* sQ = sm;
  uint8_t* sK = sm + NPAD * ROWB;
  uint8_t* sV = sm + 2 * NPAD * ROWB;
  uint8_t* sdO = sm + 3 * NPAD * ROWB;
  float* sLse = reinterpret_cast<float*>(sm + 4 * NPAD * ROWB);  // lse * log2e ; +inf for padding rows
  float* sDel = sLse + NPAD;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ld = 3LL * D;
  const bf16* base = qkv + static_cast<long long>(b) * N * ld + h * HD;
  const bf16* ob = out + static_cast<long long>(b) * N * D + h * HD;
  const bf16* dob = dout + static_cast<long long>(b) * N * D + h * HD;
  load_head_tile(sQ, base, ld, N, tid, 128);
  load_head_tile(sK, base + D, ld, N, tid, 128);
  load_head_tile(sV, base + 2 * D, ld, N, tid, 128);
  load_head_tile(sdO, dob, D, N, tid, 128);
  // delta_i = sum_d dO[i,d] * O[i,d]: 8 threads per row (one 16-byte chunk each)
  for (int i = tid; i < NPAD * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    float acc = 0.f;
    if (r < N) {
      const uint4 a = *reinterpret_cast<const uint4*>(ob + static_cast<long long>(r) * D + c * 8);
      const uint4 d = *reinterpret_cast<const uint4*>(dob + static_cast<long long>(r) * D + c * 8);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* pd = reinterpret_cast<const uint32_t*>(&d);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = unpack_bf16x2(pa[e]), y = unpack_bf16x2(pd[e]);
        acc += x.x * y.x + x.y * y.y;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (c == 0) {
      sDel[r] = acc;
      sLse[r] = r < N ? lse[(static_cast<long long>(b) * H + h) * N + r] * 1.4426950408889634f : INFINITY;
    }
  }
  __syncthreads();
  const uint32_t uQ = smem_u32(sQ), uK = smem_u32(sK), uV = smem_u32(sV), udO = smem_u32(sdO);
  const int g = lane >> 2, tig = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  bf16* dq_base = dqkv + static_cast<long long>(b) * N * ld + h * HD;

  // ---------------- phase A: dQ ----------------
  for (int mt = warp; mt < NTILE; mt += 4) {
    const int row0 = mt * 16;
    if (row0 >= N) break;
    uint32_t aq[4][4], ado[4][4];
    load_a_frags(uQ, row0, lane, aq);
    load_a_frags(udO, row0, lane, ado);
    const float lse0 = sLse[row0 + g], lse1 = sLse[row0 + g + 8];
    const float del0 = sDel[row0 + g], del1 = sDel[row0 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll 1
    for (int kt = 0; kt < NTILE; ++kt) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4], bv[4];
        load_b_nt(uK, kt * 16, ks, lane, bk);
        load_b_nt(uV, kt * 16, ks, lane, bv);
        mma16816(s[0], aq[ks], bk[0], bk[1]);
        mma16816(s[1], aq[ks], bk[2], bk[3]);
        mma16816(dp[0], ado[ks], bv[0], bv[1]);
        mma16816(dp[1], ado[ks], bv[2], bv[3]);
      }
      uint32_t ads[4];
      float ds[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int col = kt * 16 + t * 8 + 2 * tig;
        const bool ok0 = col < N, ok1 = col + 1 < N;
        const float p0 = ok0 ? exp2f(s[t][0] * sl2 - lse0) : 0.f;
        const float p1 = ok1 ? exp2f(s[t][1] * sl2 - lse0) : 0.f;
        const float p2 = ok0 ? exp2f(s[t][2] * sl2 - lse1) : 0.f;
        const float p3 = ok1 ? exp2f(s[t][3] * sl2 - lse1) : 0.f;
        ds[t][0] = p0 * (dp[t][0] - del0) * scale;
        ds[t][1] = p1 * (dp[t][1] - del0) * scale;
        ds[t][2] = p2 * (dp[t][2] - del1) * scale;
        ds[t][3] = p3 * (dp[t][3] - del1) * scale;
      }
      ads[0] = pack_bf16x2(ds[0][0], ds[0][1]);
      ads[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      ads[2] = pack_bf16x2(ds[1][0], ds[1][1]);
      ads[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        uint32_t bk[4];
        load_b_t(uK, kt * 16, nd * 16, lane, bk);
        mma16816(dq[2 * nd], ads, bk[0], bk[1]);
        mma16816(dq[2 * nd + 1], ads, bk[2], bk[3]);
      }
    }
    const int r0 = row0 + g, r1 = row0 + g + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + 2 * tig;
      if (r0 < N) *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r0) * ld + col) = pack_bf16x2(dq[j][0], dq[j][1]);
      if (r1 < N) *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r1) * ld + col) = pack_bf16x2(dq[j][2], dq[j][3]);
    }
  }

  // ---------------- phase B: dK, dV ----------------
  for (int nt = warp; nt < NTILE; nt += 4) {
    const int key0 = nt * 16;
    if (key0 >= N) break;
    uint32_t ak[4][4], av[4][4];
    load_a_frags(uK, key0, lane, ak);
    load_a_frags(uV, key0, lane, av);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
#pragma unroll 1
    for (int qt = 0; qt < NTILE; ++qt) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bq[4], bd[4];
        load_b_nt(uQ, qt * 16, ks, lane, bq);
        load_b_nt(udO, qt * 16, ks, lane, bd);
        mma16816(st[0], ak[ks], bq[0], bq[1]);
        mma16816(st[1], ak[ks], bq[2], bq[3]);
        mma16816(dpt[0], av[ks], bd[0], bd[1]);
        mma16816(dpt[1], av[ks], bd[2], bd[3]);
      }
      uint32_t apt[4], adst[4];
      float pt[2][4], dst[2][4];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = qt * 16 + t * 8 + 2 * tig;  // query index = column of S^T
        const float lq0 = sLse[q], lq1 = sLse[q + 1];
        const float dq0 = sDel[q], dq1 = sDel[q + 1];
        pt[t][0] = exp2f(st[t][0] * sl2 - lq0);
        pt[t][1] = exp2f(st[t][1] * sl2 - lq1);
        pt[t][2] = exp2f(st[t][2] * sl2 - lq0);
        pt[t][3] = exp2f(st[t][3] * sl2 - lq1);
        dst[t][0] = pt[t][0] * (dpt[t][0] - dq0) * scale;
        dst[t][1] = pt[t][1] * (dpt[t][1] - dq1) * scale;
        dst[t][2] = pt[t][2] * (dpt[t][2] - dq0) * scale;
        dst[t][3] = pt[t][3] * (dpt[t][3] - dq1) * scale;
      }
      apt[0] = pack_bf16x2(pt[0][0], pt[0][1]);
      apt[1] = pack_bf16x2(pt[0][2], pt[0][3]);
      apt[2] = pack_bf16x2(pt[1][0], pt[1][1]);
      apt[3] = pack_bf16x2(pt[1][2], pt[1][3]);
      adst[0] = pack_bf16x2(dst[0][0], dst[0][1]);
      adst[1] = pack_bf16x2(dst[0][2], dst[0][3]);
      adst[2] = pack_bf16x2(dst[1][0], dst[1][1]);
      adst[3] = pack_bf16x2(dst[1][2], dst[1][3]);
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        uint32_t bd[4], bq[4];
        load_b_t(udO, qt * 16, nd * 16, lane, bd);
        load_b_t(uQ, qt * 16, nd * 16, lane, bq);
        mma16816(dv[2 * nd], apt, bd[0], bd[1]);
        mma16816(dv[2 * nd + 1], apt, bd[2], bd[3]);
        mma16816(dk[2 * nd], adst, bq[0], bq[1]);
        mma16816(dk[2 * nd + 1], adst, bq[2], bq[3]);
      }
    }
    const int r0 = key0 + g, r1 = key0 + g + 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + 2 * tig;
      if (r0 < N) {
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r0) * ld + D + col) = pack_bf16x2(dk[j][0], dk[j][1]);
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r0) * ld + 2 * D + col) = pack_bf16x2(dv[j][0], dv[j][1]);
      }
      if (r1 < N) {
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r1) * ld + D + col) = pack_bf16x2(dk[j][2], dk[j][3]);
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<long long>(r1) * ld + 2 * D + col) = pack_bf16x2(dv[j][2], dv[j][3]);
      }
    }
  }
}

}  // namespace theia

using namespace theia;

extern "C" int theia_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, void* stream) {
  if (N > NPAD || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention: sequence length %d > %d", N, NPAD);
  const int D = H * HD;
  const int smem = 3 * NPAD * ROWB;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn fwd attr: %s", cudaGetErrorString(e));
    done = true;
  }
  attn_fwd_kernel<<<B * H, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, N, H, D, 0.125f);
  THEIA_CHECK_LAUNCH("attention_fwd");
  return THEIA_OK;
}

extern "C" int theia_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                   int B, int N, int H, void* stream) {
  if (N > NPAD || N < 1) return set_error(THEIA_ERR_UNSUPPORTED, "attention: sequence length %d > %d", N, NPAD);
  const int D = H * HD;
  const int smem = 4 * NPAD * ROWB + 2 * NPAD * sizeof(float);
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "attn bwd attr: %s", cudaGetErrorString(e));
    done = true;
  }
  attn_bwd_kernel<<<B * H, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(dout), lse,
      static_cast<bf16*>(dqkv), N, H, D, 0.125f);
  THEIA_CHECK_LAUNCH("attention_bwd");
  return THEIA_OK;
}
