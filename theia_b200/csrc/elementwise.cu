// HBM-bound kernels of the distillation step: row LayerNorm (fwd/bwd), per-image LayerNorm over
// [C,H,W] used by the lconv heads (apply / bwd), fused loss reduction + gradient, image
// pre-processing into patch rows, parameter packing, column reductions.
// All are warp-shuffle reductions with 16-byte vector accesses; fp32 statistics, bf16 I/O.
#include <math.h>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "theia_b200.h"

namespace theia {

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

// ============================================================================================
// Row LayerNorm.  One warp per row, D <= 1024, D % 8 == 0.  Two-pass statistics in registers.
// ============================================================================================
// NCH = 16-byte chunks per lane (D <= 256 NCH), R = consecutive rows per warp: a warp keeps R rows in flight so that
// narrow rows (D = 192: 384 B) still put enough bytes on the wire per SM to cover HBM latency
template <int NCH, int R>
__global__ void __launch_bounds__(256, (NCH * R >= 6) ? 3 : 4)
layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
                     float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * R;
  if (row0 >= M) return;
  const int nch = D >> 3;
  uint4 raw[R][NCH];  // the rows stay packed (bf16) in registers; unpacked on the fly by each pass
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bf16* xr = x + static_cast<long long>(row0 + r) * D;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      raw[r][c] = make_uint4(0u, 0u, 0u, 0u);
      if (ch < nch && row0 + r < M) raw[r][c] = *reinterpret_cast<const uint4*>(xr + ch * 8);
    }
  }
  auto unpack8 = [](const uint4& u, float (&f)[8]) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= M) break;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {  // chunks beyond the row are zero: no effect on the sum
      float f[8];
      unpack8(raw[r][c], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += f[e];
    }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (lane + 32 * c < nch) {
        float f[8];
        unpack8(raw[r][c], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = f[e] - mean;
          ss += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    bf16* yr = y + static_cast<long long>(row0 + r) * D;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        float f[8], g[8], b[8], o[8];
        unpack8(raw[r][c], f);
        load8f(gamma + ch * 8, g);
        load8f(beta + ch * 8, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (f[e] - mean) * rstd * g[e] + b[e];
        store8(yr + ch * 8, o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row0 + r] = mean;
      if (rstd_out) rstd_out[row0 + r] = rstd;
    }
  }
}

// fp32 rows in (the fp32 residual stream of the teacher path), bf16 or fp32 rows out.  NV = float4 per lane.
template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(256, 4)
layernorm_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         void* __restrict__ y, int M, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nv = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int i = lane + 32 * c;
    v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nv) v[c] = xr[i];
    s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
  }
  const float mean = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    if (lane + 32 * c < nv) {
      const float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, d = v[c].w - mean;
      ss += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int i = lane + 32 * c;
    if (i < nv) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[i], b = reinterpret_cast<const float4*>(beta)[i];
      const float o0 = (v[c].x - mean) * rstd * g.x + b.x, o1 = (v[c].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[c].z - mean) * rstd * g.z + b.z, o3 = (v[c].w - mean) * rstd * g.w + b.w;
      if (OUT_F32)
        reinterpret_cast<float4*>(static_cast<float*>(y) + static_cast<long long>(row) * D)[i] = make_float4(o0, o1, o2, o3);
      else
        reinterpret_cast<uint2*>(static_cast<bf16*>(y) + static_cast<long long>(row) * D)[i] =
            make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma;  optional + dadd (residual grad).
// dgamma/dbeta (and optionally the column sums of dx = bias gradient of the producer Linear): per-lane
// register partials over the rows a warp visits, block-reduced in shared memory, one atomic per column
// per block.
// HBM-bound (3 reads + 1 write of [M, D] bf16), so the rows are staged through shared memory by the bulk
// async-copy engine: every warp owns a ring of STAGES row slots (dy | x | dadd, contiguous 2*D bytes each)
// filled by cp.async.bulk with mbarrier completion -- the bytes in flight per SM (~100 KB) no longer depend on
// registers, which are left to the 9*NCH*8 fp32 column accumulators.  One warp per row, NCH = 16-byte chunks
// per lane (D <= 256*NCH).
constexpr int LNB_WARPS = 12;

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// resident blocks per SM: narrow rows need more warps in flight (a warp works on one row at a time)
constexpr int lnb_blocks_per_sm(int nch) { return nch == 1 ? 2 : 1; }

template <int NCH, int LNB_STAGES>
__global__ void __launch_bounds__(LNB_WARPS * 32, lnb_blocks_per_sm(NCH))
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const bf16* __restrict__ dadd, bf16* __restrict__ dx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dxsum, int M, int D) {
  extern __shared__ __align__(128) uint8_t lnb_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nch = D >> 3;
  const uint32_t rowb = static_cast<uint32_t>(D) * 2;  // bytes of one bf16 row
  const uint32_t slotb = 3 * rowb;
  float* red = reinterpret_cast<float*>(lnb_smem);  // [3][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(lnb_smem + 3 * D * sizeof(float)) + wib * LNB_STAGES;
  const uint32_t ring = smem_u32(lnb_smem) + 3 * D * sizeof(float) + LNB_WARPS * LNB_STAGES * 8 +
                        wib * LNB_STAGES * slotb;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < LNB_STAGES; ++st) mbar_init(&bars[st], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int stride = gridDim.x * LNB_WARPS;
  const int row0 = blockIdx.x * LNB_WARPS + wib;
  auto issue = [&](int row, int st) {  // lane 0: stage one row
    if (row < M) {
      const long long off = static_cast<long long>(row) * D;
      mbar_expect_tx(&bars[st], dadd ? 3 * rowb : 2 * rowb);
      bulk_g2s(ring + st * slotb, dy + off, rowb, &bars[st]);
      bulk_g2s(ring + st * slotb + rowb, x + off, rowb, &bars[st]);
      if (dadd) bulk_g2s(ring + st * slotb + 2 * rowb, dadd + off, rowb, &bars[st]);
    }
  };
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < LNB_STAGES; ++st) issue(row0 + st * stride, st);
  }

  float ag[NCH][8], ab[NCH][8], ax[NCH][8], gm[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[c][e] = 0.f, ab[c][e] = 0.f, ax[c][e] = 0.f, gm[c][e] = 0.f;
    if (ch < nch) load8f(gamma + ch * 8, gm[c]);
  }
  float mean_n = 0.f, rstd_n = 0.f;  // statistics of the row about to be processed (fetched one row ahead)
  if (row0 < M) mean_n = mean_in[row0], rstd_n = rstd_in[row0];
  int st = 0;
  uint32_t parity = 0;
  const float invD = 1.0f / D;
  for (int row = row0; row < M; row += stride) {
    const float rstd = rstd_n, nmr = -mean_n * rstd_n;
    if (row + stride < M) mean_n = mean_in[row + stride], rstd_n = rstd_in[row + stride];
    mbar_wait(&bars[st], parity);
    const uint32_t slot = ring + st * slotb;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        const uint4 pd = lds128(slot + ch * 16), px = lds128(slot + rowb + ch * 16);
        const uint32_t* ud = reinterpret_cast<const uint32_t*>(&pd);
        const uint32_t* ux = reinterpret_cast<const uint32_t*>(&px);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 d = unpack_bf16x2(ud[e2]), xv = unpack_bf16x2(ux[e2]);
          const float xh0 = fmaf(xv.x, rstd, nmr), xh1 = fmaf(xv.y, rstd, nmr);
          const float g0 = d.x * gm[c][2 * e2], g1 = d.y * gm[c][2 * e2 + 1];
          s1 += g0 + g1;
          s2 = fmaf(g0, xh0, fmaf(g1, xh1, s2));
          ag[c][2 * e2] = fmaf(d.x, xh0, ag[c][2 * e2]), ag[c][2 * e2 + 1] = fmaf(d.y, xh1, ag[c][2 * e2 + 1]);
          ab[c][2 * e2] += d.x, ab[c][2 * e2 + 1] += d.y;
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    const float ns2 = -s2;
    const long long base = static_cast<long long>(row) * D;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        const uint4 pd = lds128(slot + ch * 16), px = lds128(slot + rowb + ch * 16);
        uint4 pa = make_uint4(0u, 0u, 0u, 0u);
        if (dadd) pa = lds128(slot + 2 * rowb + ch * 16);
        const uint32_t* ud = reinterpret_cast<const uint32_t*>(&pd);
        const uint32_t* ux = reinterpret_cast<const uint32_t*>(&px);
        const uint32_t* ua = reinterpret_cast<const uint32_t*>(&pa);
        uint4 outp;
        uint32_t* uo = reinterpret_cast<uint32_t*>(&outp);
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 d = unpack_bf16x2(ud[e2]), xv = unpack_bf16x2(ux[e2]), a = unpack_bf16x2(ua[e2]);
          const float xh0 = fmaf(xv.x, rstd, nmr), xh1 = fmaf(xv.y, rstd, nmr);
          const float o0 = fmaf(rstd, fmaf(xh0, ns2, fmaf(d.x, gm[c][2 * e2], -s1)), a.x);
          const float o1 = fmaf(rstd, fmaf(xh1, ns2, fmaf(d.y, gm[c][2 * e2 + 1], -s1)), a.y);
          uo[e2] = pack_bf16x2(o0, o1);
          if (dxsum) ax[c][2 * e2] += o0, ax[c][2 * e2 + 1] += o1;
        }
        *reinterpret_cast<uint4*>(dx + base + ch * 8) = outp;
      }
    }
    __syncwarp();  // every lane is done reading the slot
    if (lane == 0) issue(row + LNB_STAGES * stride, st);
    if (++st == LNB_STAGES) st = 0, parity ^= 1;
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&red[ch * 8 + e], ag[c][e]);
        atomicAdd(&red[D + ch * 8 + e], ab[c][e]);
        if (dxsum) atomicAdd(&red[2 * D + ch * 8 + e], ax[c][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, red[i]);
    atomicAdd(dbeta + i, red[D + i]);
    if (dxsum) atomicAdd(dxsum + i, red[2 * D + i]);
  }
}

// ============================================================================================
// LayerNorm over [C,H,W] per image (lconv heads), NHWC storage: n = H*W*C elements per image.
// Statistics (sum, sumsq) come from the producing GEMM's epilogue.
// ============================================================================================
// One thread = 8 consecutive elements j of the [H,W,C] map, for LN3D_APPLY_G images: the fp32 affine (64 B per
// thread) is fetched once and reused, so the L2 traffic of gamma/beta does not exceed the HBM traffic of x.
constexpr int LN3D_APPLY_G = 8;
__global__ void __launch_bounds__(256) ln3d_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         bf16* __restrict__ y, int n, int B, float eps, int C, int Wp,
                                                         int Wv) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= n) return;
  const int b0 = blockIdx.y * LN3D_APPLY_G;
  float ns = static_cast<float>(n);
  bool pad = false;
  if (Wp > 0) {  // padded storage [Wp x Wp] of a [Wv x Wv] map: padding stays zero and is not counted
    const int pos = j / C, w = pos % Wp, h = pos / Wp;
    ns = static_cast<float>(Wv) * Wv * C;
    pad = (w >= Wv || h >= Wv);
  }
  float g[8], bb[8];
  if (!pad) {
    load8f(gamma + j, g);
    load8f(beta + j, bb);
  }
#pragma unroll
  for (int u0 = 0; u0 < LN3D_APPLY_G; u0 += 4) {
    uint4 px[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = b0 + u0 + u;
      px[u] = make_uint4(0u, 0u, 0u, 0u);
      if (!pad && b < B) px[u] = *reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * n + j);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = b0 + u0 + u;
      if (b < B) {
        uint4 outp = make_uint4(0u, 0u, 0u, 0u);
        if (!pad) {
          const float mean = stats[2 * b] / ns;
          const float var = fmaxf(stats[2 * b + 1] / ns - mean * mean, 0.f);
          const float rstd = rsqrtf(var + eps);
          const uint32_t* ux = reinterpret_cast<const uint32_t*>(&px[u]);
          uint32_t* uo = reinterpret_cast<uint32_t*>(&outp);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const float2 v = unpack_bf16x2(ux[e2]);
            uo[e2] = pack_bf16x2((v.x - mean) * rstd * g[2 * e2] + bb[2 * e2],
                                 (v.y - mean) * rstd * g[2 * e2 + 1] + bb[2 * e2 + 1]);
          }
        }
        *reinterpret_cast<uint4*>(y + static_cast<long long>(b) * n + j) = outp;
      }
    }
  }
}

// pass 1: per-image sums of g = dy*gamma and g*xhat (-> red[b][2]); dgamma/dbeta over the batch.
// grid (chunks, image groups); block 256 threads x 8 elements.
constexpr int LN3D_GROUP = 16;
__global__ void __launch_bounds__(256) ln3d_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, float* __restrict__ red,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              int n, int B, float eps, int C, int Wp, int Wv) {
  __shared__ float sred[LN3D_GROUP][2];
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  bool act = j < n;
  float ns = static_cast<float>(n);
  if (Wp > 0) {
    ns = static_cast<float>(Wv) * Wv * C;
    if (act) {
      const int pos = j / C, w = pos % Wp, h = pos / Wp;
      act = (w < Wv) && (h < Wv);
    }
  }
  const int b0 = blockIdx.y * LN3D_GROUP;
  if (threadIdx.x < LN3D_GROUP * 2) (&sred[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  float g8[8], ag[8], ab[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ag[e] = 0.f, ab[e] = 0.f, g8[e] = 0.f;
  if (act) load8f(gamma + j, g8);
  const int lane = threadIdx.x & 31;
  for (int bi = 0; bi < LN3D_GROUP; ++bi) {
    const int b = b0 + bi;
    if (b >= B) break;
    float s1 = 0.f, s2 = 0.f;
    if (act) {
      const float mean = stats[2 * b] / ns;
      const float var = fmaxf(stats[2 * b + 1] / ns - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      float d[8], xv[8];
      const long long off = static_cast<long long>(b) * n + j;
      load8(dy + off, d);
      load8(x + off, xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xv[e] - mean) * rstd;
        const float g = d[e] * g8[e];
        s1 += g;
        s2 += g * xh;
        ag[e] += d[e] * xh;
        ab[e] += d[e];
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(&sred[bi][0], s1);
      atomicAdd(&sred[bi][1], s2);
    }
  }
  if (act) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(dgamma + j + e, ag[e]);
      atomicAdd(dbeta + j + e, ab[e]);
    }
  }
  __syncthreads();
  if (threadIdx.x < LN3D_GROUP * 2) {
    const int bi = threadIdx.x >> 1;
    if (b0 + bi < B) atomicAdd(red + 2 * (b0 + bi) + (threadIdx.x & 1), sred[bi][threadIdx.x & 1]);
  }
}

// pass 2: dx = rstd * (g - mean(g) - xhat*mean(g*xhat)); optional ReLU mask of the producer
// (x is the post-ReLU conv output, so x > 0 <=> pre-activation > 0).
__global__ void __launch_bounds__(256) ln3d_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                             const float* __restrict__ stats,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ red, bf16* __restrict__ dx,
                                                             int n, long long total, float eps, int relu_mask, int C,
                                                             int Wp, int Wv) {
  const long long i8 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i8 >= total) return;
  const int b = static_cast<int>(i8 / n);
  const int j = static_cast<int>(i8 - static_cast<long long>(b) * n);
  float ns = static_cast<float>(n);
  if (Wp > 0) {
    const int pos = j / C, w = pos % Wp, h = pos / Wp;
    ns = static_cast<float>(Wv) * Wv * C;
    if (w >= Wv || h >= Wv) {
      *reinterpret_cast<uint4*>(dx + i8) = make_uint4(0u, 0u, 0u, 0u);
      return;
    }
  }
  const float mean = stats[2 * b] / ns;
  const float var = fmaxf(stats[2 * b + 1] / ns - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float c1 = red[2 * b] / ns, c2 = red[2 * b + 1] / ns;
  float d[8], xv[8], g8[8], o[8];
  load8(dy + i8, d);
  load8(x + i8, xv);
  load8f(gamma + j, g8);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float xh = (xv[e] - mean) * rstd;
    float r = rstd * (d[e] * g8[e] - c1 - xh * c2);
    if (relu_mask && !(xv[e] > 0.f)) r = 0.f;
    o[e] = r;
  }
  store8(dx + i8, o);
}

// ============================================================================================
// Loss (rvfm.py:153-176).  pred fp32 [B,n], target fp32 or bf16 [B,n].
//   acc[b][5] = { sum d^2, sum smoothl1(d), sum p^2, sum t^2, sum p*t }
// ============================================================================================
template <typename TT>
__device__ __forceinline__ void load_t4(const TT* p, float (&f)[4]);
template <>
__device__ __forceinline__ void load_t4<float>(const float* p, float (&f)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w;
}
template <>
__device__ __forceinline__ void load_t4<bf16>(const bf16* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y;
}

template <typename TT>
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ pred, const TT* __restrict__ tgt,
                                                          float* __restrict__ acc, int n, int chunk) {
  __shared__ float sh[5][8];
  const int b = blockIdx.y;
  const long long base = static_cast<long long>(b) * n;
  const int j0 = blockIdx.x * chunk;
  const int j1 = min(n, j0 + chunk);
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = j0 + threadIdx.x * 4; j < j1; j += blockDim.x * 4) {
    float p[4], t[4];
    load_t4<float>(pred + base + j, p);
    load_t4<TT>(tgt + base + j, t);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float d = p[e] - t[e];
      const float ad = fabsf(d);
      s[0] += d * d;
      s[1] += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
      s[2] += p[e] * p[e];
      s[3] += t[e] * t[e];
      s[4] += p[e] * t[e];
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    s[k] = warp_sum(s[k]);
    if (lane == 0) sh[k][w] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[threadIdx.x][i];
    atomicAdd(acc + b * 5 + threadIdx.x, t);
  }
}

// out[3] = {mse, cos, l1} of one teacher (unweighted means, as nn.MSELoss / CosineEmbeddingLoss /
// SmoothL1Loss return them).  One block.
__global__ void loss_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out, int B, int n) {
  __shared__ float sh[3][32];
  float m = 0.f, c = 0.f, l = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* a = acc + b * 5;
    m += a[0];
    l += a[1];
    const float pn = fmaxf(sqrtf(a[2]), 1e-12f), tn = fmaxf(sqrtf(a[3]), 1e-12f);
    // F.normalize then CosineEmbeddingLoss (eps 1e-8 on the unit vectors)
    const float pp = a[2] / (pn * pn), tt = a[3] / (tn * tn);
    const float cosv = (a[4] / (pn * tn)) / sqrtf((pp + 1e-8f) * (tt + 1e-8f));
    c += 1.f - cosv;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  m = warp_sum(m), c = warp_sum(c), l = warp_sum(l);
  if (lane == 0) sh[0][w] = m, sh[1][w] = c, sh[2][w] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float M_ = 0.f, C_ = 0.f, L_ = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) M_ += sh[0][i], C_ += sh[1][i], L_ += sh[2][i];
    const double cnt = static_cast<double>(B) * n;
    out[0] = static_cast<float>(M_ / cnt);
    out[1] = C_ / B;
    out[2] = static_cast<float>(L_ / cnt);
  }
}

// dpred = gm*2d/(B n) + gl*clamp(d,-1,1)/(B n) + gc/B * d(1-cos_b)/dp ; g* = upstream grad * weight,
// read from device memory (coef[3] = {g_mse*w, g_cos*w, g_l1*w}) so no host sync is needed.
template <typename TT, typename TO>
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ pred, const TT* __restrict__ tgt,
                                                        const float* __restrict__ acc, const float* __restrict__ coef,
                                                        TO* __restrict__ dpred, int n, int B, bf16* __restrict__ dpred_bf16) {
  const int b = blockIdx.y;
  const long long base = static_cast<long long>(b) * n;
  const float* a = acc + b * 5;
  const float pn = fmaxf(sqrtf(a[2]), 1e-12f), tn = fmaxf(sqrtf(a[3]), 1e-12f);
  const float inv_cnt = 1.f / (static_cast<float>(B) * static_cast<float>(n));
  const float km = coef[0] * 2.f * inv_cnt;
  const float kl = coef[2] * inv_cnt;
  const float kc = coef[1] / B;
  // cos = pt/(pn tn):  dcos/dp = t/(pn tn) - pt p/(pn^3 tn);  loss term is (1 - cos)
  const float ct = -kc / (pn * tn);
  const float cp = kc * a[4] / (pn * pn * pn * tn);
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= n) return;
  float p[4], t[4], o[4];
  load_t4<float>(pred + base + j, p);
  load_t4<TT>(tgt + base + j, t);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float d = p[e] - t[e];
    o[e] = km * d + kl * fminf(fmaxf(d, -1.f), 1.f) + ct * t[e] + cp * p[e];
  }
  if (dpred_bf16 != nullptr) {  // bf16 copy for the head-Linear backward GEMMs, written in the same pass
    uint2 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(dpred_bf16 + base + j) = u;
  }
  if constexpr (sizeof(TO) == 4) {
    *reinterpret_cast<float4*>(dpred + base + j) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    uint2 u;
    u.x = pack_bf16x2(o[0], o[1]);
    u.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(dpred + base + j) = u;
  }
}

// ============================================================================================
// Image pre-processing (DeiT processor without resize): uint8 -> normalised bf16 patch rows.
// out[b*197 + 0][:] = 0 (CLS slot), out[b*197 + 1 + p][c*256 + i*16 + j] = norm(img[b, py*16+i, px*16+j, c])
// ============================================================================================
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int NT, int P0, int B,
                                                         int chw, float s0, float s1, float s2, float o0, float o1,
                                                         float o2) {
  // one thread per 8 consecutive k of one patch row
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * NT * 96;
  if (t >= total) return;
  const int k8 = static_cast<int>(t % 96);
  const long long row = t / 96;
  const int tok = static_cast<int>(row % NT);
  const int b = static_cast<int>(row / NT);
  float o[8];
  if (tok < P0 || tok >= P0 + 196) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
  } else {
    const int p = tok - P0, py = p / 14, px = p % 14;
    const int k = k8 * 8;
    const int c = k >> 8, i = (k >> 4) & 15, j = k & 15;
    const int yy = py * 16 + i, xx = px * 16 + j;
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float of = c == 0 ? o0 : (c == 1 ? o1 : o2);
    if (chw) {
      const uint8_t* src = img + ((static_cast<long long>(b) * 3 + c) * 224 + yy) * 224 + xx;
      const uint2 u = *reinterpret_cast<const uint2*>(src);  // xx % 8 == 0
      const uint8_t* q = reinterpret_cast<const uint8_t*>(&u);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (static_cast<float>(q[e]) - of) * sc;
    } else {
      const uint8_t* src = img + ((static_cast<long long>(b) * 224 + yy) * 224 + xx) * 3 + c;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (static_cast<float>(src[e * 3]) - of) * sc;
    }
  }
  store8(out + row * 768 + k8 * 8, o);
}

// --------------------------------------------------------------------------------------------
// DeiT processor WITH resize (the reference's default): bicubic antialias 224 -> 256 (torch
// `_upsample_bicubic2d_aa`, a = -0.5, float accumulation: horizontal taps first, then vertical),
// clamp, round-half-even to uint8 levels, centre crop 224 (= resized pixels 16..239), normalise.
// The per-output tap tables are computed on the host exactly as ATen does and kept in constant memory.
// --------------------------------------------------------------------------------------------
struct ResizeTable {
  int xmin[256];
  int xsize[256];
  float w[256][5];
  // torchvision's CPU uint8 path (ATen upsample_bicubic2d_aa on uint8, the Pillow-SIMD scheme): the same taps as
  // int16 fixed-point weights with `prec` fractional bits; horizontal pass first, its result rounded and clamped to
  // uint8, then the vertical pass
  int wi[256][5];
  int prec;
};
__constant__ ResizeTable c_rt;

__global__ void __launch_bounds__(256) preprocess_resize_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int NT, int P0,
                                                                int B, int chw, float s0, float s1, float s2, float o0,
                                                                float o1, float o2, uint8_t* __restrict__ dbg_u8, int fixed) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * NT * 96;
  if (t >= total) return;
  const int k8 = static_cast<int>(t % 96);
  const long long row = t / 96;
  const int tok = static_cast<int>(row % NT);
  const int b = static_cast<int>(row / NT);
  float o[8];
  if (tok < P0 || tok >= P0 + 196) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
  } else {
    const int p = tok - P0, py = p / 14, px = p % 14;
    const int k = k8 * 8;
    const int c = k >> 8, i = (k >> 4) & 15, j = k & 15;
    const int oy = py * 16 + i + 16;
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float of = c == 0 ? o0 : (c == 1 ? o1 : o2);
    const int ymin = c_rt.xmin[oy], ysize = c_rt.xsize[oy];
    const long long pix_stride = chw ? 1 : 3;
    const long long row_stride = chw ? 224 : 224 * 3;
    const uint8_t* base = chw ? img + (static_cast<long long>(b) * 3 + c) * 224 * 224
                              : img + static_cast<long long>(b) * 224 * 224 * 3 + c;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ox = px * 16 + j + e + 16;
      const int xmin = c_rt.xmin[ox], xsize = c_rt.xsize[ox];
      float acc = 0.f;
      if (fixed) {  // CPU-tensor semantics of the reference's processor: integer arithmetic, bit-exact
        const int half = 1 << (c_rt.prec - 1);
        int sv = half;
        for (int y = 0; y < ysize; ++y) {
          const uint8_t* src = base + (ymin + y) * row_stride + xmin * pix_stride;
          int sh = half;
          for (int x = 0; x < xsize; ++x) sh += static_cast<int>(src[x * pix_stride]) * c_rt.wi[ox][x];
          sh >>= c_rt.prec;
          sh = sh < 0 ? 0 : (sh > 255 ? 255 : sh);
          sv += sh * c_rt.wi[oy][y];
        }
        sv >>= c_rt.prec;
        acc = static_cast<float>(sv < 0 ? 0 : (sv > 255 ? 255 : sv));
      } else {
      for (int y = 0; y < ysize; ++y) {
        const uint8_t* src = base + (ymin + y) * row_stride + xmin * pix_stride;
        float r = static_cast<float>(src[0]) * c_rt.w[ox][0];
        for (int x = 1; x < xsize; ++x) r += static_cast<float>(src[x * pix_stride]) * c_rt.w[ox][x];
        if (y == 0) acc = r * c_rt.w[oy][0];
        else acc += r * c_rt.w[oy][y];
      }
      acc = rintf(fminf(fmaxf(acc, 0.f), 255.f));
      }
      o[e] = (acc - of) * sc;
      // test hook: the resized + centre-cropped uint8 image [B,224,224,3] the reference's processor produces
      if (dbg_u8 != nullptr)
        dbg_u8[((static_cast<long long>(b) * 224 + (py * 16 + i)) * 224 + (px * 16 + j + e)) * 3 + c] = static_cast<uint8_t>(acc);
    }
  }
  store8(out + row * 768 + k8 * 8, o);
}

// ============================================================================================
// Generic strided gather (parameter packing / gradient unpacking).
//   out[((a*n1 + b)*n2 + c)*n3 + d] = in[base + a*s0 + b*s1 + c*s2 + d*s3]
// ============================================================================================
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gather4_kernel(const TI* __restrict__ in, TO* __restrict__ out, int n0, int n1,
                                                      int n2, int n3, long long s0, long long s1, long long s2,
                                                      long long s3, long long base) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n0) * n1 * n2 * n3;
  if (t >= total) return;
  long long r = t;
  const int d = static_cast<int>(r % n3);
  r /= n3;
  const int c = static_cast<int>(r % n2);
  r /= n2;
  const int b = static_cast<int>(r % n1);
  const int a = static_cast<int>(r / n1);
  const float v = static_cast<float>(in[base + a * s0 + b * s1 + c * s2 + d * s3]);
  out[t] = static_cast<TO>(v);
}

// Fused optimizer tail (train_rvfm.py:126-133, optimizers/utils.py:8-35): AdamW over the flat fp32
// buffers in ONE pass (torch.optim.AdamW arithmetic: decoupled weight decay, bias correction), two weight-decay
// groups through a per-64-element flag table, optional global-norm clipping through a device-side coefficient.
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long k = i; k < n; ++k) acc += g[k] * g[k];
    }
  }
  acc = warp_sum(acc);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}
// coef[0] = min(1, max_norm / (sqrt(sumsq) + 1e-6))  (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float c = max_norm / (sqrtf(sumsq[0]) + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
}
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         const uint8_t* __restrict__ flag64, long long n, float lr,
                                                         float beta1, float beta2, float eps, float wd, float bc1,
                                                         float sqrt_bc2, const float* __restrict__ gscale,
                                                         const int* __restrict__ pack_table, bf16* __restrict__ packbf) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const uint8_t fl = flag64[i >> 6];
  if (fl & 2) return;  // parameter without a gradient (frozen / unused head): left bit-identical, like torch
  const float gs = gscale ? gscale[0] : 1.f;
  const float decay = (fl & 1) ? 1.f - lr * wd : 1.f;
  const float step = lr / bc1;
  float4 P = *reinterpret_cast<float4*>(p + i), G = *reinterpret_cast<const float4*>(g + i);
  float4 M = *reinterpret_cast<float4*>(m + i), V = *reinterpret_cast<float4*>(v + i);
  float* pp = &P.x;
  float* mm = &M.x;
  float* vv = &V.x;
  const float* gg = &G.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float gk = gg[k] * gs;
    float pk = pp[k] * decay;
    const float mk = mm[k] + (gk - mm[k]) * (1.f - beta1);  // lerp_
    const float vk = vv[k] * beta2 + gk * gk * (1.f - beta2);
    const float denom = sqrtf(vk) / sqrt_bc2 + eps;
    pk -= step * (mk / denom);
    pp[k] = pk, mm[k] = mk, vv[k] = vk;
  }
  *reinterpret_cast<float4*>(p + i) = P;
  *reinterpret_cast<float4*>(m + i) = M;
  *reinterpret_cast<float4*>(v + i) = V;
  if (pack_table != nullptr) {  // refresh the bf16 GEMM-operand copy of this weight in the same pass
    const int dst = pack_table[i >> 6];
    if (dst >= 0) {
      uint2 u;
      u.x = pack_bf16x2(P.x, P.y);
      u.y = pack_bf16x2(P.z, P.w);
      *reinterpret_cast<uint2*>(packbf + (static_cast<long long>(dst) << 6) + (i & 63)) = u;
    }
  }
}

// bf16 operand copies of every parameter block the pack table names (Linear / patch-embed weights: the copy is a
// plain cast, 64-element blocks): ONE launch instead of a cast kernel per weight.
__global__ void __launch_bounds__(256) pack_cast_kernel(const float* __restrict__ p, const int* __restrict__ pack_table,
                                                        bf16* __restrict__ packbf, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const int dst = pack_table[i >> 6];
  if (dst < 0) return;
  const float4 P = *reinterpret_cast<const float4*>(p + i);
  uint2 u;
  u.x = pack_bf16x2(P.x, P.y);
  u.y = pack_bf16x2(P.z, P.w);
  *reinterpret_cast<uint2*>(packbf + (static_cast<long long>(dst) << 6) + (i & 63)) = u;
}

// Segmented strided gather: ONE launch runs a table of layout conversions (conv-weight packs, LayerNorm[C,H,W]
// affine <-> NHWC, conv weight-gradient scratch -> reference layout).  Segment s:
//   out[((a*n1 + b)*n2 + c)*n3 + d] = (b < lim1 && c < lim2) ? in[base + a*s0 + b*s1 + c*s2 + d*s3] : 0
__global__ void __launch_bounds__(256) perm_seg_kernel(const theia_perm_seg* __restrict__ segs, int nseg,
                                                       uintptr_t out_rebase) {
  int lo = 0, hi = nseg - 1;
  const long long blk = blockIdx.x;
  while (lo < hi) {  // last segment whose first_block <= blk
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].first_block <= blk) lo = mid;
    else hi = mid - 1;
  }
  const theia_perm_seg sg = segs[lo];
  const long long t = (blk - sg.first_block) * 256 + threadIdx.x;
  const long long total = static_cast<long long>(sg.n0) * sg.n1 * sg.n2 * sg.n3;
  if (t >= total) return;
  long long r = t;
  const int d = static_cast<int>(r % sg.n3);
  r /= sg.n3;
  const int c = static_cast<int>(r % sg.n2);
  r /= sg.n2;
  const int b = static_cast<int>(r % sg.n1);
  const int a = static_cast<int>(r / sg.n1);
  float v = 0.f;
  if ((sg.lim1 <= 0 || b < sg.lim1) && (sg.lim2 <= 0 || c < sg.lim2))
    v = static_cast<const float*>(sg.in)[sg.base + a * sg.s0 + b * sg.s1 + c * sg.s2 + d * sg.s3];
  void* outp = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(sg.out) + out_rebase);
  if (sg.out_f32) static_cast<float*>(outp)[t] = v;
  else static_cast<bf16*>(outp)[t] = __float2bfloat16_rn(v);
}

// Target ingest (src/theia/dataset/data_utils.py:152-153,342-355): teacher embedding stored
// [C,H,W] bf16 -> "(h w) c" and z-scored with the per-channel bf16 mean / std, on the GPU instead of in the
// dataloader workers.  Arithmetic reproduces torch's bf16 ops bit for bit: r = bf16(x - mean); y = bf16(r / std).
__global__ void target_ingest_kernel(const bf16* __restrict__ in, const bf16* __restrict__ mean,
                                     const bf16* __restrict__ stdv, bf16* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bf16* src = in + static_cast<long long>(b) * C * HW;
  bf16* dst = out + static_cast<long long>(b) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(src[static_cast<long long>(c) * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      float v = tile[threadIdx.x][i];
      if (mean != nullptr && stdv != nullptr) {
        const float r = __bfloat162float(__float2bfloat16_rn(v - __bfloat162float(mean[c])));
        v = r / __bfloat162float(stdv[c]);
      }
      dst[static_cast<long long>(p) * C + c] = __float2bfloat16_rn(v);
    }
  }
}

// LayerNorm[C,H,W] affine: [C][Hv][Wv] (reference layout) <-> [Hp][Wp][C] (NHWC, zero padded)
__global__ void __launch_bounds__(256) chw_to_hwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                         int Hv, int Wv, int Hp, int Wp) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(Hp) * Wp * C) return;
  const int c = static_cast<int>(t % C);
  const int pos = static_cast<int>(t / C), w = pos % Wp, h = pos / Wp;
  out[t] = (h < Hv && w < Wv) ? in[(static_cast<long long>(c) * Hv + h) * Wv + w] : 0.f;
}
__global__ void __launch_bounds__(256) hwc_to_chw_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                                         int Hv, int Wv, int Hp, int Wp) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(C) * Hv * Wv) return;
  const int w = static_cast<int>(t % Wv);
  const int h = static_cast<int>((t / Wv) % Hv);
  const int c = static_cast<int>(t / (static_cast<long long>(Wv) * Hv));
  out[t] = in[(static_cast<long long>(h) * Wp + w) * C + c];
}

// tiled transpose-cast: out[c][r] = (bf16) in[r][c]   (fp32 [R,Cc] -> bf16 [Cc,R])
__global__ void transpose_cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? in[static_cast<long long>(r) * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[static_cast<long long>(c) * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 f = *reinterpret_cast<const float4*>(in + i);
    uint2 u;
    u.x = pack_bf16x2(f.x, f.y);
    u.y = pack_bf16x2(f.z, f.w);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (long long k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
  }
}

// ============================================================================================
// Column sum of a bf16 [M,N] matrix into fp32 out[N] (+=): bias gradients.
// rows with (m % skip_mod == 0) are skipped when skip_mod > 0 (CLS rows for the patch bias).
// period > 0: out has `period*N` entries and row m adds into out[(m % period)*N + n] (pos-emb grad).
// ============================================================================================
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int M, int N,
                                                     long long ld, int skip_mod, int rows_per_block, int t0, int t1) {
  __shared__ float sh[8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + tx) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 0.f;
  if (n < N) {
    for (int r = r0 + ty; r < r1; r += 8) {
      if (skip_mod > 0) {
        const int tk = r % skip_mod;
        if (tk < t0 || tk >= t1) continue;
      }
      float v[8];
      load8(x + static_cast<long long>(r) * ld + n, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[ty][tx * 8 + e] = a[e];
  __syncthreads();
  const int col = threadIdx.x;  // 256 columns per block
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += sh[i][col];
  const int gn = blockIdx.x * 256 + col;
  if (gn < N) atomicAdd(out + gn, s);
}

// out[j] (=) sum_b x[b*n + j] : batch reduction (pos-emb / cls gradients), x bf16, out fp32
__global__ void __launch_bounds__(256) batchsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int B, int n) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (j >= n) return;
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 0.f;
  for (int b = 0; b < B; ++b) {
    float v[8];
    load8(x + static_cast<long long>(b) * n + j, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) out[j + e] = a[e];
}

// --------------------------------------------------------------------------------------------
// Conv-weight layout conversions through shared-memory tiles (coalesced on both sides).
//   reference layout  W[a][b][9]  (a, b < C; Conv2d: a = Cout, b = Cin; ConvTranspose2d: a = Cin, b = Cout)
//   GEMM pack         P[X][T][Y] ("tap in the middle") or P[T][X][Y] (tap-major), (X, Y) = (a, b) or (b, a),
//                     T = t or 8 - t (flipped kernel)
// pack:   W (fp32 master) -> up to two bf16 packs per weight (forward and dgrad operand)
// unpack: tap-major fp32 weight-gradient scratch ws[T][X][Y] -> reference layout (fp32 gradient buffer)
// One launch per direction for ALL conv weights of the model (segment table in device memory).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ long long conv_perm_index(int a, int b, int t, int C, int flags) {
  const int X = (flags & THEIA_CP_SWAP) ? b : a, Y = (flags & THEIA_CP_SWAP) ? a : b;
  const int T = (flags & THEIA_CP_FLIP) ? 8 - t : t;
  return (flags & THEIA_CP_TAPMAJOR) ? (static_cast<long long>(T) * C + X) * C + Y
                                     : (static_cast<long long>(X) * 9 + T) * C + Y;
}

__global__ void __launch_bounds__(256) conv_pack_kernel(const theia_conv_perm* __restrict__ segs, int tiles_per_conv,
                                                        int tiles_b) {
  __shared__ float tile[32][32 * 9 + 1];
  const theia_conv_perm sg = segs[blockIdx.x / tiles_per_conv];
  const int tix = blockIdx.x % tiles_per_conv;
  const int a0 = (tix / tiles_b) * 32, b0 = (tix % tiles_b) * 32;
  const int C = sg.C;
  const float* in = static_cast<const float*>(sg.ref);
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {  // row a0 + i / 288: 288 contiguous floats (32 b x 9 taps)
    const int ar = i / 288, k = i - ar * 288;
    tile[ar][k] = in[(static_cast<long long>(a0 + ar) * C + b0) * 9 + k];
  }
  __syncthreads();
  for (int which = 0; which < 2; ++which) {
    bf16* out = static_cast<bf16*>(which ? sg.pack1 : sg.pack0);
    const int flags = which ? sg.flags1 : sg.flags0;
    if (out == nullptr) continue;
    // 32 consecutive Y per (X, T): Y = b without SWAP, a with SWAP
    for (int i = threadIdx.x; i < 32 * 9 * 32; i += 256) {
      const int y = i & 31, xt = i >> 5, t = xt % 9, x = xt / 9;
      const int a = (flags & THEIA_CP_SWAP) ? y : x, b = (flags & THEIA_CP_SWAP) ? x : y;
      out[conv_perm_index(a0 + a, b0 + b, t, C, flags)] = __float2bfloat16_rn(tile[a][b * 9 + t]);
    }
  }
}

__global__ void __launch_bounds__(256) conv_unpack_kernel(const theia_conv_perm* __restrict__ segs, int tiles_per_conv,
                                                          int tiles_b, uintptr_t out_rebase) {
  __shared__ float tile[32][32 * 9 + 1];
  const theia_conv_perm sg = segs[blockIdx.x / tiles_per_conv];
  const int tix = blockIdx.x % tiles_per_conv;
  const int a0 = (tix / tiles_b) * 32, b0 = (tix % tiles_b) * 32;
  const int C = sg.C;
  const float* ws = static_cast<const float*>(sg.pack0);
  const int flags = sg.flags0 | THEIA_CP_TAPMAJOR;
  for (int i = threadIdx.x; i < 32 * 9 * 32; i += 256) {
    const int y = i & 31, xt = i >> 5, t = xt % 9, x = xt / 9;
    const int a = (flags & THEIA_CP_SWAP) ? y : x, b = (flags & THEIA_CP_SWAP) ? x : y;
    tile[a][b * 9 + t] = ws[conv_perm_index(a0 + a, b0 + b, t, C, flags)];
  }
  __syncthreads();
  float* out = reinterpret_cast<float*>(reinterpret_cast<uintptr_t>(sg.ref) + out_rebase);
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int ar = i / 288, k = i - ar * 288;
    out[(static_cast<long long>(a0 + ar) * C + b0) * 9 + k] = tile[ar][k];
  }
}

}  // namespace theia

using namespace theia;

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" int theia_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                   float* rstd, int M, int D, float eps, void* stream) {
  if (D % 8 != 0 || D > 1280) return set_error(THEIA_ERR_ARG, "layernorm: D %% 8 == 0 and D <= 1280 required");
  if (M <= 0) return THEIA_OK;
  const int wpb = 8;
#define LNF(NCH, R)                                                                                             \
  layernorm_fwd_kernel<NCH, R><<<((M + R - 1) / R + wpb - 1) / wpb, wpb * 32, 0, S(stream)>>>(                  \
      static_cast<const bf16*>(x), gamma, beta, static_cast<bf16*>(y), mean, rstd, M, D, eps)
  if (D <= 256) LNF(1, 4);
  else if (D <= 512) LNF(2, 2);
  else if (D <= 768) LNF(3, 1);
  else if (D <= 1024) LNF(4, 1);
  else LNF(5, 1);  // ViT-H teacher (D = 1280): forward only
#undef LNF
  THEIA_CHECK_LAUNCH("layernorm_fwd");
  return THEIA_OK;
}

extern "C" int theia_layernorm_fwd_f32(const float* x, const float* gamma, const float* beta, void* y, int y_is_f32, int M,
                                       int D, float eps, void* stream) {
  if (D % 4 != 0 || D > 1280) return set_error(THEIA_ERR_ARG, "layernorm (fp32 rows): D %% 4 == 0 and D <= 1280 required");
  if (M <= 0) return THEIA_OK;
  const unsigned grid = static_cast<unsigned>((M + 7) / 8);
#define LNF32(NV)                                                                                     \
  do {                                                                                                \
    if (y_is_f32) layernorm_fwd_f32_kernel<NV, true><<<grid, 256, 0, S(stream)>>>(x, gamma, beta, y, M, D, eps);  \
    else layernorm_fwd_f32_kernel<NV, false><<<grid, 256, 0, S(stream)>>>(x, gamma, beta, y, M, D, eps);          \
  } while (0)
  if (D <= 512) LNF32(4);
  else if (D <= 1024) LNF32(8);
  else LNF32(10);
#undef LNF32
  THEIA_CHECK_LAUNCH("layernorm_fwd_f32");
  return THEIA_OK;
}

extern "C" int theia_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                   const float* rstd, const void* dadd, void* dx, float* dgamma, float* dbeta,
                                   float* dxsum, int M, int D, void* stream) {
  if (D % 8 != 0 || D > 1024) return set_error(THEIA_ERR_ARG, "layernorm: D %% 8 == 0 and D <= 1024 required");
  if (M <= 0) return THEIA_OK;
  int grid = num_sms() * lnb_blocks_per_sm(D <= 256 ? 1 : (D <= 512 ? 2 : 3));  // persistent blocks
  if (grid > (M + LNB_WARPS - 1) / LNB_WARPS) grid = (M + LNB_WARPS - 1) / LNB_WARPS;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dadd)) & 15)
    return set_error(THEIA_ERR_ARG, "layernorm_bwd: dy / x / dadd must be 16-byte aligned");
#define LNB(N, ST)                                                                                                   \
  do {                                                                                                               \
    const size_t sm = 3 * D * sizeof(float) + LNB_WARPS * ST * 8 + static_cast<size_t>(LNB_WARPS) * ST * 6 * D;      \
    static PerDeviceOnce attr_once;                                                                                  \
    if (attr_once.first_use())                                                                                       \
      cudaFuncSetAttribute(layernorm_bwd_kernel<N, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);    \
    layernorm_bwd_kernel<N, ST><<<grid, LNB_WARPS * 32, sm, S(stream)>>>(                                            \
        static_cast<const bf16*>(dy), static_cast<const bf16*>(x), gamma, mean, rstd, static_cast<const bf16*>(dadd), \
        static_cast<bf16*>(dx), dgamma, dbeta, dxsum, M, D);                                                         \
  } while (0)
  if (D <= 256) LNB(1, 4);
  else if (D <= 512) LNB(2, 4);
  else if (D <= 768) LNB(3, 3);
  else LNB(4, 2);
#undef LNB
  THEIA_CHECK_LAUNCH("layernorm_bwd");
  return THEIA_OK;
}

extern "C" int theia_ln3d_apply(const void* x, const float* stats, const float* gamma_hwc, const float* beta_hwc,
                                void* y, int B, int n, float eps, int C, int Wp, int Wv, void* stream) {
  if (n % 8 != 0 || C % 8 != 0) return set_error(THEIA_ERR_ARG, "ln3d: n, C %% 8 != 0");
  dim3 ga((n / 8 + 255) / 256, (B + LN3D_APPLY_G - 1) / LN3D_APPLY_G);
  ln3d_apply_kernel<<<ga, 256, 0, S(stream)>>>(static_cast<const bf16*>(x), stats, gamma_hwc, beta_hwc,
                                               static_cast<bf16*>(y), n, B, eps, C, Wp, Wv);
  THEIA_CHECK_LAUNCH("ln3d_apply");
  return THEIA_OK;
}

extern "C" int theia_ln3d_bwd(const void* dy, const void* x, const float* stats, const float* gamma_hwc, float* red,
                              void* dx, float* dgamma_hwc, float* dbeta_hwc, int B, int n, float eps, int relu_mask,
                              int C, int Wp, int Wv, void* stream) {
  if (n % 8 != 0 || C % 8 != 0) return set_error(THEIA_ERR_ARG, "ln3d: n, C %% 8 != 0");
  cudaError_t e = cudaMemsetAsync(red, 0, sizeof(float) * 2 * B, S(stream));
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  dim3 g1((n / 8 + 255) / 256, (B + LN3D_GROUP - 1) / LN3D_GROUP);
  ln3d_bwd_reduce_kernel<<<g1, 256, 0, S(stream)>>>(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), stats,
                                                    gamma_hwc, red, dgamma_hwc, dbeta_hwc, n, B, eps, C, Wp, Wv);
  THEIA_CHECK_LAUNCH("ln3d_bwd_reduce");
  const long long total = static_cast<long long>(B) * n;
  ln3d_bwd_apply_kernel<<<static_cast<unsigned>((total / 8 + 255) / 256), 256, 0, S(stream)>>>(
      static_cast<const bf16*>(dy), static_cast<const bf16*>(x), stats, gamma_hwc, red, static_cast<bf16*>(dx), n,
      total, eps, relu_mask, C, Wp, Wv);
  THEIA_CHECK_LAUNCH("ln3d_bwd_apply");
  return THEIA_OK;
}

extern "C" int theia_loss_fwd(const float* pred, const void* target, int target_is_bf16, float* acc, float* out3,
                              int B, int n, void* stream) {
  if (n % 4 != 0) return set_error(THEIA_ERR_ARG, "loss: n %% 4 != 0");
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float) * 5 * B, S(stream));
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  const int chunk = 16384;
  dim3 g((n + chunk - 1) / chunk, B);
  if (target_is_bf16)
    loss_reduce_kernel<bf16><<<g, 256, 0, S(stream)>>>(pred, static_cast<const bf16*>(target), acc, n, chunk);
  else
    loss_reduce_kernel<float><<<g, 256, 0, S(stream)>>>(pred, static_cast<const float*>(target), acc, n, chunk);
  THEIA_CHECK_LAUNCH("loss_reduce");
  loss_finalize_kernel<<<1, 256, 0, S(stream)>>>(acc, out3, B, n);
  THEIA_CHECK_LAUNCH("loss_finalize");
  return THEIA_OK;
}

extern "C" int theia_loss_bwd(const float* pred, const void* target, int target_is_bf16, const float* acc,
                              const float* coef3, void* dpred, int dpred_is_f32, void* dpred_bf16_copy, int B, int n,
                              void* stream) {
  if (n % 4 != 0) return set_error(THEIA_ERR_ARG, "loss: n %% 4 != 0");
  dim3 g((n / 4 + 255) / 256, B);
  const bf16* tb = static_cast<const bf16*>(target);
  const float* tf = static_cast<const float*>(target);
  bf16* cp = dpred_is_f32 ? static_cast<bf16*>(dpred_bf16_copy) : nullptr;
  if (target_is_bf16 && dpred_is_f32)
    loss_grad_kernel<bf16, float><<<g, 256, 0, S(stream)>>>(pred, tb, acc, coef3, static_cast<float*>(dpred), n, B, cp);
  else if (target_is_bf16)
    loss_grad_kernel<bf16, bf16><<<g, 256, 0, S(stream)>>>(pred, tb, acc, coef3, static_cast<bf16*>(dpred), n, B, cp);
  else if (dpred_is_f32)
    loss_grad_kernel<float, float><<<g, 256, 0, S(stream)>>>(pred, tf, acc, coef3, static_cast<float*>(dpred), n, B, cp);
  else
    loss_grad_kernel<float, bf16><<<g, 256, 0, S(stream)>>>(pred, tf, acc, coef3, static_cast<bf16*>(dpred), n, B, cp);
  THEIA_CHECK_LAUNCH("loss_grad");
  return THEIA_OK;
}

static double bicubic_aa_filter(float xf) {
  // ATen upsample aa bicubic filter (a = -0.5); its literals are doubles, so it evaluates in double
  double x = xf;
  const float a = -0.5f;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}

static int upload_resize_table() {
  static bool done[64] = {false};  // __constant__ memory is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (done[dev]) return 0;
  static ResizeTable rt;
  const float scale = 224.0f / 256.0f;
  const float support = 2.0f;  // interp_size 4 * 0.5 (scale < 1: no widening)
  const float invscale = 1.0f;
  for (int o = 0; o < 256; ++o) {
    const float center = scale * (o + 0.5f);
    int xmin = static_cast<int>(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5f);
    if (xmax > 224) xmax = 224;
    const int xsize = xmax - xmin;
    const float xmin_m_center = xmin - center;
    float total = 0.f;
    float w[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < xsize && j < 5; ++j) {
      const float wj = static_cast<float>(bicubic_aa_filter((j + xmin_m_center + 0.5f) * invscale));
      w[j] = wj;
      total += wj;
    }
    for (int j = 0; j < xsize && j < 5; ++j)
      if (total != 0.f) w[j] /= total;
    rt.xmin[o] = xmin;
    rt.xsize[o] = xsize > 5 ? 5 : xsize;
    for (int j = 0; j < 5; ++j) rt.w[o][j] = w[j];
  }
  {  // fixed-point weights (double arithmetic, as ATen's _compute_indices_int16_weights_aa)
    static double wd[256][5];
    double wt_max = 0.0;
    const double dscale = 224.0 / 256.0;
    for (int o = 0; o < 256; ++o) {
      const double center = dscale * (o + 0.5);
      double total = 0.0;
      for (int j = 0; j < 5; ++j) wd[o][j] = 0.0;
      for (int j = 0; j < rt.xsize[o]; ++j) {
        double x = (j + rt.xmin[o] - center + 0.5);
        if (x < 0.0) x = -x;
        const double a = -0.5;
        double wv = 0.0;
        if (x < 1.0) wv = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
        else if (x < 2.0) wv = (((x - 5.0) * x + 8.0) * x - 4.0) * a;
        wd[o][j] = wv;
        total += wv;
      }
      for (int j = 0; j < rt.xsize[o]; ++j) {
        wd[o][j] /= total;
        if (wd[o][j] > wt_max) wt_max = wd[o][j];
      }
    }
    int prec = 0;
    for (; prec < 22; ++prec) {
      const int next = static_cast<int>(0.5 + wt_max * (1 << (prec + 1)));
      if (next >= (1 << 15)) break;
    }
    rt.prec = prec;
    for (int o = 0; o < 256; ++o)
      for (int j = 0; j < 5; ++j) {
        const double v = wd[o][j];
        rt.wi[o][j] = v < 0 ? static_cast<int>(-0.5 + v * (1 << prec)) : static_cast<int>(0.5 + v * (1 << prec));
      }
  }
  cudaError_t e = cudaMemcpyToSymbol(c_rt, &rt, sizeof(rt));
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "resize table upload: %s", cudaGetErrorString(e));
  done[dev] = true;
  return 0;
}

static uint8_t* g_resize_dbg_u8 = nullptr;
// Test hook: when set, the next theia_preprocess(do_resize=1) calls also write the resized + centre-cropped uint8
// image [B,224,224,3] (the byte stage of the reference's processor) to this device buffer.  NULL switches it off.
extern "C" int theia_preprocess_debug_u8(void* resized_u8_out) {
  g_resize_dbg_u8 = static_cast<uint8_t*>(resized_u8_out);
  return THEIA_OK;
}

extern "C" int theia_preprocess(const uint8_t* images, void* patches, int B, int channels_first, int do_resize,
                                int do_rescale, int do_normalize, const float* mean3, const float* std3, int tokens,
                                int patch_off, void* stream) {
  if (tokens < patch_off + 196 || patch_off < 0) return set_error(THEIA_ERR_ARG, "preprocess: bad token layout");
  // out = (x - off) * scale  per channel
  float sc[3], of[3];
  for (int c = 0; c < 3; ++c) {
    if (do_normalize) {
      const float m = do_rescale ? mean3[c] * 255.f : mean3[c];
      const float s = do_rescale ? std3[c] * 255.f : std3[c];
      of[c] = m;
      sc[c] = 1.f / s;
    } else {
      of[c] = 0.f;
      sc[c] = do_rescale ? (1.f / 255.f) : 1.f;
    }
  }
  const long long total = static_cast<long long>(B) * tokens * 96;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (do_resize) {
    int rc = upload_resize_table();
    if (rc) return rc;
    preprocess_resize_kernel<<<grid, 256, 0, S(stream)>>>(images, static_cast<bf16*>(patches), tokens, patch_off, B,
                                                          channels_first,
                                                          sc[0], sc[1], sc[2], of[0], of[1], of[2], g_resize_dbg_u8,
                                                          do_resize == 2 ? 1 : 0);
  } else {
    preprocess_kernel<<<grid, 256, 0, S(stream)>>>(images, static_cast<bf16*>(patches), tokens, patch_off, B,
                                                   channels_first, sc[0],
                                                   sc[1], sc[2], of[0], of[1], of[2]);
  }
  THEIA_CHECK_LAUNCH("preprocess");
  return THEIA_OK;
}

// --------------------------------------------------------------------------------------------
// Any input size (the reference's processor takes whatever the caller has: backbones.py:337-339).  do_resize: bicubic
// antialias H x W -> 256 x 256 (per-axis tap tables as ATen builds them, support widened by the scale when
// down-sampling), centre crop 224.  do_resize = 0: centre crop, zero padding when the image is smaller than 224
// (hf:image_processing_backends.py center_crop).  The 224 x 224 kernels above stay the fast path.
// --------------------------------------------------------------------------------------------
namespace theia {
struct AxisDev {
  const int* xmin;   // float path (CUDA-tensor semantics): first tap, tap count, normalised weights [256][T]
  const int* xsize;
  const float* w;
  int T;
  const int* xmin_i;  // fixed-point path (CPU uint8 semantics): int16 weights with `prec` fractional bits [256][Ti]
  const int* xsize_i;
  const int* wi;
  int Ti, prec;
};

__global__ void __launch_bounds__(256) preprocess_any_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ out, int NT, int P0,
                                                             int B, int chw, int in_h, int in_w, int mode, AxisDev ax, AxisDev ay,
                                                             float s0, float s1, float s2, float o0, float o1, float o2,
                                                             uint8_t* __restrict__ dbg_u8) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * NT * 96;
  if (t >= total) return;
  const int k8 = static_cast<int>(t % 96);
  const long long row = t / 96;
  const int tok = static_cast<int>(row % NT);
  const int b = static_cast<int>(row / NT);
  float o[8];
  if (tok < P0 || tok >= P0 + 196) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
  } else {
    const int p = tok - P0, py = p / 14, px = p % 14;
    const int k = k8 * 8;
    const int c = k >> 8, i = (k >> 4) & 15, j = k & 15;
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float of = c == 0 ? o0 : (c == 1 ? o1 : o2);
    const long long pix_stride = chw ? 1 : 3;
    const long long row_stride = chw ? in_w : static_cast<long long>(in_w) * 3;
    const uint8_t* base = chw ? img + (static_cast<long long>(b) * 3 + c) * in_h * in_w
                              : img + static_cast<long long>(b) * in_h * in_w * 3 + c;
    const int yy = py * 16 + i;
#pragma unroll 1
    for (int e = 0; e < 8; ++e) {
      const int xx = px * 16 + j + e;
      float acc;
      if (mode == 0) {  // centre crop / zero pad
        const int sy = yy + (in_h >= 224 ? (in_h - 224) / 2 : -((224 - in_h) / 2));
        const int sx = xx + (in_w >= 224 ? (in_w - 224) / 2 : -((224 - in_w) / 2));
        acc = (sy >= 0 && sy < in_h && sx >= 0 && sx < in_w) ? static_cast<float>(base[sy * row_stride + sx * pix_stride]) : 0.f;
      } else if (mode == 2) {  // integer arithmetic: horizontal pass rounded and clamped to uint8, then the vertical one
        const int oy = yy + 16, ox = xx + 16;
        const int ymin = ay.xmin_i[oy], ysize = ay.xsize_i[oy], xmin = ax.xmin_i[ox], xsize = ax.xsize_i[ox];
        const int* wx = ax.wi + ox * ax.Ti;
        const int* wy = ay.wi + oy * ay.Ti;
        int sv = 1 << (ay.prec - 1);
        for (int y = 0; y < ysize; ++y) {
          const uint8_t* src = base + (ymin + y) * row_stride + xmin * pix_stride;
          int sh = 1 << (ax.prec - 1);
          for (int x = 0; x < xsize; ++x) sh += static_cast<int>(src[x * pix_stride]) * wx[x];
          sh >>= ax.prec;
          sh = sh < 0 ? 0 : (sh > 255 ? 255 : sh);
          sv += sh * wy[y];
        }
        sv >>= ay.prec;
        acc = static_cast<float>(sv < 0 ? 0 : (sv > 255 ? 255 : sv));
      } else {  // float arithmetic: horizontal taps first, then vertical, round-half-even to uint8 levels
        const int oy = yy + 16, ox = xx + 16;
        const int ymin = ay.xmin[oy], ysize = ay.xsize[oy], xmin = ax.xmin[ox], xsize = ax.xsize[ox];
        const float* wx = ax.w + ox * ax.T;
        const float* wy = ay.w + oy * ay.T;
        acc = 0.f;
        for (int y = 0; y < ysize; ++y) {
          const uint8_t* src = base + (ymin + y) * row_stride + xmin * pix_stride;
          float r = static_cast<float>(src[0]) * wx[0];
          for (int x = 1; x < xsize; ++x) r += static_cast<float>(src[x * pix_stride]) * wx[x];
          if (y == 0) acc = r * wy[0];
          else acc += r * wy[y];
        }
        acc = rintf(fminf(fmaxf(acc, 0.f), 255.f));
      }
      o[e] = (acc - of) * sc;
      if (dbg_u8 != nullptr) dbg_u8[((static_cast<long long>(b) * 224 + yy) * 224 + xx) * 3 + c] = static_cast<uint8_t>(acc);
    }
  }
  store8(out + row * 768 + k8 * 8, o);
}

// float images (the processor also takes float tensors: values in [0, 255], or in [0, 1] with do_rescale = 0): centre crop /
// zero pad to 224 x 224 and the same fused (x - offset) * scale; no resize on this path
__global__ void __launch_bounds__(256) preprocess_f32_kernel(const float* __restrict__ img, bf16* __restrict__ out, int NT, int P0, int B,
                                                             int chw, int in_h, int in_w, float s0, float s1, float s2, float o0,
                                                             float o1, float o2) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * NT * 96;
  if (t >= total) return;
  const int k8 = static_cast<int>(t % 96);
  const long long row = t / 96;
  const int tok = static_cast<int>(row % NT);
  const int b = static_cast<int>(row / NT);
  float o[8];
  if (tok < P0 || tok >= P0 + 196) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
  } else {
    const int p = tok - P0, py = p / 14, px = p % 14;
    const int k = k8 * 8;
    const int c = k >> 8, i = (k >> 4) & 15, j = k & 15;
    const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const float of = c == 0 ? o0 : (c == 1 ? o1 : o2);
    const long long pix_stride = chw ? 1 : 3;
    const long long row_stride = chw ? in_w : static_cast<long long>(in_w) * 3;
    const float* base = chw ? img + (static_cast<long long>(b) * 3 + c) * in_h * in_w
                            : img + static_cast<long long>(b) * in_h * in_w * 3 + c;
    const int sy = py * 16 + i + (in_h >= 224 ? (in_h - 224) / 2 : -((224 - in_h) / 2));
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int sx = px * 16 + j + e + (in_w >= 224 ? (in_w - 224) / 2 : -((224 - in_w) / 2));
      const float v = (sy >= 0 && sy < in_h && sx >= 0 && sx < in_w) ? base[sy * row_stride + sx * pix_stride] : 0.f;
      o[e] = (v - of) * sc;
    }
  }
  store8(out + row * 768 + k8 * 8, o);
}

// Float tap table of one axis (in -> 256), built ON THE DEVICE with the expressions of ATen's CUDA kernel
// (upsample_antialias::_compute_weights_span / _compute_weights / BicubicFilterFunctor, ATen/native/cuda/UpSample.cuh):
// its cubic is evaluated in float with nvcc's FMA contraction, which a host restatement does not reproduce bit for
// bit once the taps fall between pixel centres (down-sampling).
__device__ __forceinline__ float aa_cubic_f32(float x) {
  const float a = -0.5f;
  if (x < 0) x = -x;
  if (x < 1) return ((a + 2) * x - (a + 3)) * x * x + 1;
  if (x < 2) return (((x - 5) * x + 8) * x - 4) * a;
  return 0;
}
__global__ void axis_float_table_kernel(int in, int T, int* __restrict__ xmin_out, int* __restrict__ xsize_out,
                                        float* __restrict__ w_out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= 256) return;
  const float scale = static_cast<float>(in) / 256;
  const float support = (scale >= 1.0) ? (4 * 0.5f) * scale : 4 * 0.5f;
  const float center = scale * (o + 0.5f);
  const int xmin = max(static_cast<int>(center - support + 0.5f), 0);
  int xsize = min(static_cast<int>(center + support + 0.5f), in) - xmin;
  if (xsize > T) xsize = T;
  const float xmin_m_center = xmin - center;
  const float invscale = (scale >= 1.0) ? 1.0 / scale : 1.0;
  float total_w = 0.0;
  float* wt = w_out + o * T;
  int j = 0;
  for (j = 0; j < xsize; j++) {
    const float w = aa_cubic_f32((j + xmin_m_center + 0.5f) * invscale);
    wt[j] = w;
    total_w += w;
  }
  for (j = 0; j < xsize; j++)
    if (total_w != 0.0) wt[j] /= total_w;
  for (; j < T; j++) wt[j] = 0.f;
  xmin_out[o] = xmin, xsize_out[o] = xsize;
}

}  // namespace theia
namespace {

struct AxisHost {
  int in = 0;
  uint8_t* blob = nullptr;  // device
  AxisDev dev{};
};

// per-output tap tables of one axis (in -> 256), as ATen computes them: float (upsample_gen2d_aa_out_frame, CUDA) and
// double -> int16 (_compute_indices_int16_weights_aa, CPU uint8)
int build_axis(int in, AxisHost* ah) {
  const int OUT = 256;
  const float scale = static_cast<float>(in) / OUT;
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const int T = static_cast<int>(ceilf(support)) * 2 + 1;
  const double dscale = static_cast<double>(in) / OUT;
  const double dsupport = dscale >= 1.0 ? 2.0 * dscale : 2.0;
  const double dinv = dscale >= 1.0 ? 1.0 / dscale : 1.0;
  const int Ti = static_cast<int>(ceil(dsupport)) * 2 + 1;
  std::vector<int> ints(4 * OUT + static_cast<size_t>(OUT) * Ti, 0);
  std::vector<float> wf(static_cast<size_t>(OUT) * T, 0.f);
  std::vector<double> wd(static_cast<size_t>(OUT) * Ti, 0.0);
  int* xmin = ints.data();
  int* xsize = xmin + OUT;
  int* xmin_i = xsize + OUT;
  int* xsize_i = xmin_i + OUT;
  int* wi = xsize_i + OUT;
  double wt_max = 0.0;
  for (int o = 0; o < OUT; ++o) {
    {
      const float center = scale * (o + 0.5f);
      int lo = static_cast<int>(center - support + 0.5f);
      if (lo < 0) lo = 0;
      int hi = static_cast<int>(center + support + 0.5f);
      if (hi > in) hi = in;
      int n = hi - lo;
      if (n > T) n = T;
      if (n < 0) n = 0;
      const float lo_m_center = lo - center;
      float total = 0.f;
      for (int j = 0; j < n; ++j) {
        const float wj = static_cast<float>(bicubic_aa_filter((j + lo_m_center + 0.5f) * invscale));
        wf[static_cast<size_t>(o) * T + j] = wj;
        total += wj;
      }
      if (total != 0.f)
        for (int j = 0; j < n; ++j) wf[static_cast<size_t>(o) * T + j] /= total;
      xmin[o] = lo, xsize[o] = n;
    }
    {
      const double center = dscale * (o + 0.5);
      long long lo = static_cast<long long>(center - dsupport + 0.5);
      if (lo < 0) lo = 0;
      long long hi = static_cast<long long>(center + dsupport + 0.5);
      if (hi > in) hi = in;
      long long n = hi - lo;
      if (n > Ti) n = Ti;
      if (n < 0) n = 0;
      double total = 0.0;
      for (int j = 0; j < n; ++j) {
        double x = (j + lo - center + 0.5) * dinv;
        if (x < 0.0) x = -x;
        const double a = -0.5;
        double wv = 0.0;
        if (x < 1.0) wv = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
        else if (x < 2.0) wv = (((x - 5.0) * x + 8.0) * x - 4.0) * a;
        wd[static_cast<size_t>(o) * Ti + j] = wv;
        total += wv;
      }
      if (total != 0.0)
        for (int j = 0; j < n; ++j) {
          double& v = wd[static_cast<size_t>(o) * Ti + j];
          v /= total;
          if (v > wt_max) wt_max = v;
        }
      xmin_i[o] = static_cast<int>(lo), xsize_i[o] = static_cast<int>(n);
    }
  }
  int prec = 0;
  for (; prec < 22; ++prec) {
    const int next = static_cast<int>(0.5 + wt_max * (1 << (prec + 1)));
    if (next >= (1 << 15)) break;
  }
  for (size_t q = 0; q < wd.size(); ++q) {
    const double v = wd[q];
    wi[q] = v < 0 ? static_cast<int>(-0.5 + v * (1 << prec)) : static_cast<int>(0.5 + v * (1 << prec));
  }
  const size_t ibytes = ints.size() * sizeof(int), fbytes = wf.size() * sizeof(float);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ah->blob), ibytes + fbytes);  // first use of this size only
  if (e == cudaSuccess) e = cudaMemcpy(ah->blob, ints.data(), ibytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(ah->blob + ibytes, wf.data(), fbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {  // float path: xmin / xsize / weights as the CUDA kernel of ATen computes them
    int* dix = reinterpret_cast<int*>(ah->blob);
    axis_float_table_kernel<<<1, 256>>>(in, T, dix, dix + OUT, reinterpret_cast<float*>(ah->blob + ibytes));
    e = cudaDeviceSynchronize();
  }
  if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "resize table (%d -> 256): %s", in, cudaGetErrorString(e));
  const int* di = reinterpret_cast<const int*>(ah->blob);
  ah->in = in;
  ah->dev.xmin = di, ah->dev.xsize = di + OUT, ah->dev.xmin_i = di + 2 * OUT, ah->dev.xsize_i = di + 3 * OUT, ah->dev.wi = di + 4 * OUT;
  ah->dev.w = reinterpret_cast<const float*>(ah->blob + ibytes);
  ah->dev.T = T, ah->dev.Ti = Ti, ah->dev.prec = prec;
  return THEIA_OK;
}

// tap tables are cached per (device, input extent); a handful of extents per process in practice
int axis_table(int in, AxisDev* out) {
  static std::vector<AxisHost> cache[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (const AxisHost& a : cache[dev])
    if (a.in == in) {
      *out = a.dev;
      return THEIA_OK;
    }
  AxisHost ah;
  const int rc = build_axis(in, &ah);
  if (rc) return rc;
  cache[dev].push_back(ah);
  *out = ah.dev;
  return THEIA_OK;
}
}  // namespace

extern "C" int theia_preprocess_hw(const uint8_t* images, int in_h, int in_w, void* patches, int B, int channels_first,
                                   int do_resize, int do_rescale, int do_normalize, const float* mean3, const float* std3,
                                   int tokens, int patch_off, void* stream) {
  if (in_h == 224 && in_w == 224)
    return theia_preprocess(images, patches, B, channels_first, do_resize, do_rescale, do_normalize, mean3, std3, tokens,
                            patch_off, stream);
  if (tokens < patch_off + 196 || patch_off < 0) return set_error(THEIA_ERR_ARG, "preprocess: bad token layout");
  if (in_h < 1 || in_w < 1 || in_h > 8192 || in_w > 8192) return set_error(THEIA_ERR_ARG, "preprocess: image %d x %d", in_h, in_w);
  float sc[3], of[3];
  for (int c = 0; c < 3; ++c) {
    if (do_normalize) {
      of[c] = do_rescale ? mean3[c] * 255.f : mean3[c];
      sc[c] = 1.f / (do_rescale ? std3[c] * 255.f : std3[c]);
    } else {
      of[c] = 0.f;
      sc[c] = do_rescale ? (1.f / 255.f) : 1.f;
    }
  }
  AxisDev ax{}, ay{};
  if (do_resize) {
    int rc = axis_table(in_w, &ax);
    if (rc) return rc;
    rc = axis_table(in_h, &ay);
    if (rc) return rc;
  }
  const long long total = static_cast<long long>(B) * tokens * 96;
  preprocess_any_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, S(stream)>>>(
      images, static_cast<bf16*>(patches), tokens, patch_off, B, channels_first, in_h, in_w, do_resize, ax, ay, sc[0], sc[1],
      sc[2], of[0], of[1], of[2], g_resize_dbg_u8);
  THEIA_CHECK_LAUNCH("preprocess_any");
  return THEIA_OK;
}

extern "C" int theia_preprocess_f32(const float* images, int in_h, int in_w, void* patches, int B, int channels_first,
                                    int do_rescale, int do_normalize, const float* mean3, const float* std3, int tokens,
                                    int patch_off, void* stream) {
  if (tokens < patch_off + 196 || patch_off < 0) return set_error(THEIA_ERR_ARG, "preprocess: bad token layout");
  if (in_h < 1 || in_w < 1 || in_h > 8192 || in_w > 8192) return set_error(THEIA_ERR_ARG, "preprocess: image %d x %d", in_h, in_w);
  float sc[3], of[3];
  for (int c = 0; c < 3; ++c) {
    if (do_normalize) {
      of[c] = do_rescale ? mean3[c] * 255.f : mean3[c];
      sc[c] = 1.f / (do_rescale ? std3[c] * 255.f : std3[c]);
    } else {
      of[c] = 0.f;
      sc[c] = do_rescale ? (1.f / 255.f) : 1.f;
    }
  }
  const long long total = static_cast<long long>(B) * tokens * 96;
  preprocess_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, S(stream)>>>(
      images, static_cast<bf16*>(patches), tokens, patch_off, B, channels_first, in_h, in_w, sc[0], sc[1], sc[2], of[0], of[1],
      of[2]);
  THEIA_CHECK_LAUNCH("preprocess_f32");
  return THEIA_OK;
}

extern "C" int theia_gather4(const void* in, void* out, int in_is_f32, int out_is_f32, int n0, int n1, int n2, int n3,
                             long long s0, long long s1, long long s2, long long s3, long long base, void* stream) {
  const long long total = static_cast<long long>(n0) * n1 * n2 * n3;
  if (total == 0) return THEIA_OK;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (in_is_f32 && !out_is_f32)
    gather4_kernel<float, bf16><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(in), static_cast<bf16*>(out), n0,
                                                             n1, n2, n3, s0, s1, s2, s3, base);
  else if (in_is_f32 && out_is_f32)
    gather4_kernel<float, float><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(in), static_cast<float*>(out),
                                                              n0, n1, n2, n3, s0, s1, s2, s3, base);
  else
    return set_error(THEIA_ERR_UNSUPPORTED, "gather4: unsupported dtype combination");
  THEIA_CHECK_LAUNCH("gather4");
  return THEIA_OK;
}

extern "C" int theia_adamw_flat(float* p, const float* g, float* m, float* v, const uint8_t* flag64, long long n,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                float max_grad_norm, float* scratch2, const int* pack_table, void* packbf, void* stream) {
  if (n <= 0) return THEIA_OK;
  if (n % 4 != 0) return set_error(THEIA_ERR_ARG, "adamw: n %% 4 != 0");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const float* gscale = nullptr;
  if (max_grad_norm > 0.f) {
    if (!scratch2) return set_error(THEIA_ERR_ARG, "adamw: clipping needs scratch2");
    cudaError_t e = cudaMemsetAsync(scratch2, 0, 2 * sizeof(float), S(stream));
    if (e != cudaSuccess) return set_error(THEIA_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    sumsq_kernel<<<num_sms() * 4, 256, 0, S(stream)>>>(g, n, scratch2);
    THEIA_CHECK_LAUNCH("sumsq");
    clip_coef_kernel<<<1, 1, 0, S(stream)>>>(scratch2, max_grad_norm, scratch2 + 1);
    THEIA_CHECK_LAUNCH("clip_coef");
    gscale = scratch2 + 1;
  }
  adamw_flat_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, S(stream)>>>(
      p, g, m, v, flag64, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), gscale,
      packbf ? pack_table : nullptr, static_cast<bf16*>(packbf));
  THEIA_CHECK_LAUNCH("adamw_flat");
  return THEIA_OK;
}

extern "C" int theia_conv_pack(const theia_conv_perm* segs_dev, int nconv, int C, void* stream) {
  if (nconv <= 0) return THEIA_OK;
  if (C % 32 != 0) return set_error(THEIA_ERR_ARG, "conv_pack: C %% 32 != 0");
  const int tb = C / 32;
  conv_pack_kernel<<<nconv * tb * tb, 256, 0, S(stream)>>>(segs_dev, tb * tb, tb);
  THEIA_CHECK_LAUNCH("conv_pack");
  return THEIA_OK;
}

extern "C" int theia_conv_unpack(const theia_conv_perm* segs_dev, int nconv, int C, const void* out_rebase, void* stream) {
  if (nconv <= 0) return THEIA_OK;
  if (C % 32 != 0) return set_error(THEIA_ERR_ARG, "conv_unpack: C %% 32 != 0");
  const int tb = C / 32;
  conv_unpack_kernel<<<nconv * tb * tb, 256, 0, S(stream)>>>(segs_dev, tb * tb, tb, reinterpret_cast<uintptr_t>(out_rebase));
  THEIA_CHECK_LAUNCH("conv_unpack");
  return THEIA_OK;
}

extern "C" int theia_pack_cast(const float* p, const int* pack_table, void* packbf, long long n, void* stream) {
  if (n <= 0) return THEIA_OK;
  if (n % 4 != 0) return set_error(THEIA_ERR_ARG, "pack_cast: n %% 4 != 0");
  pack_cast_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, S(stream)>>>(p, pack_table,
                                                                                     static_cast<bf16*>(packbf), n);
  THEIA_CHECK_LAUNCH("pack_cast");
  return THEIA_OK;
}

extern "C" int theia_perm_segments(const theia_perm_seg* segs_dev, int nseg, long long total_blocks,
                                   const void* out_rebase, void* stream) {
  if (nseg <= 0 || total_blocks <= 0) return THEIA_OK;
  perm_seg_kernel<<<static_cast<unsigned>(total_blocks), 256, 0, S(stream)>>>(segs_dev, nseg,
                                                                             reinterpret_cast<uintptr_t>(out_rebase));
  THEIA_CHECK_LAUNCH("perm_segments");
  return THEIA_OK;
}

extern "C" int theia_target_ingest(const void* emb_chw, const void* mean_c, const void* std_c, void* out_hwc, int B,
                                   int C, int HW, void* stream) {
  if (B <= 0) return THEIA_OK;
  dim3 g((HW + 31) / 32, (C + 31) / 32, B), blk(32, 8);
  target_ingest_kernel<<<g, blk, 0, S(stream)>>>(static_cast<const bf16*>(emb_chw), static_cast<const bf16*>(mean_c),
                                                 static_cast<const bf16*>(std_c), static_cast<bf16*>(out_hwc), C, HW);
  THEIA_CHECK_LAUNCH("target_ingest");
  return THEIA_OK;
}

extern "C" int theia_chw_to_hwc(const float* in, float* out, int C, int Hv, int Wv, int Hp, int Wp, void* stream) {
  const long long total = static_cast<long long>(Hp) * Wp * C;
  chw_to_hwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, S(stream)>>>(in, out, C, Hv, Wv, Hp, Wp);
  THEIA_CHECK_LAUNCH("chw_to_hwc");
  return THEIA_OK;
}
extern "C" int theia_hwc_to_chw(const float* in, float* out, int C, int Hv, int Wv, int Hp, int Wp, void* stream) {
  const long long total = static_cast<long long>(C) * Hv * Wv;
  hwc_to_chw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, S(stream)>>>(in, out, C, Hv, Wv, Hp, Wp);
  THEIA_CHECK_LAUNCH("hwc_to_chw");
  return THEIA_OK;
}

extern "C" int theia_cast_bf16(const float* in, void* out, long long n, void* stream) {
  if (n <= 0) return THEIA_OK;
  cast_kernel<<<static_cast<unsigned>((n / 4 + 256) / 256), 256, 0, S(stream)>>>(in, static_cast<bf16*>(out), n);
  THEIA_CHECK_LAUNCH("cast");
  return THEIA_OK;
}

extern "C" int theia_transpose_cast_bf16(const float* in, void* out, int R, int C, void* stream) {
  dim3 g((C + 31) / 32, (R + 31) / 32), b(32, 8);
  transpose_cast_kernel<<<g, b, 0, S(stream)>>>(in, static_cast<bf16*>(out), R, C);
  THEIA_CHECK_LAUNCH("transpose_cast");
  return THEIA_OK;
}

extern "C" int theia_colsum(const void* x, float* out, int M, int N, long long ld, int skip_mod, void* stream) {
  if (N % 8 != 0) return set_error(THEIA_ERR_ARG, "colsum: N %% 8 != 0");
  const int rows_per_block = 512;
  dim3 g((N + 255) / 256, (M + rows_per_block - 1) / rows_per_block);
  // skip_mod > 0: rows with (m % skip_mod) == 0 are skipped (legacy form: CLS rows)
  colsum_kernel<<<g, 256, 0, S(stream)>>>(static_cast<const bf16*>(x), out, M, N, ld, skip_mod, rows_per_block, 1,
                                          skip_mod > 0 ? skip_mod : 1);
  THEIA_CHECK_LAUNCH("colsum");
  return THEIA_OK;
}

extern "C" int theia_colsum_tokens(const void* x, float* out, int M, int N, long long ld, int period, int t0, int t1,
                                   void* stream) {
  if (N % 8 != 0 || period < 1) return set_error(THEIA_ERR_ARG, "colsum_tokens: bad arguments");
  const int rows_per_block = 512;
  dim3 g((N + 255) / 256, (M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<g, 256, 0, S(stream)>>>(static_cast<const bf16*>(x), out, M, N, ld, period, rows_per_block, t0, t1);
  THEIA_CHECK_LAUNCH("colsum_tokens");
  return THEIA_OK;
}

extern "C" int theia_batchsum(const void* x, float* out, int B, int n, void* stream) {
  if (n % 8 != 0) return set_error(THEIA_ERR_ARG, "batchsum: n %% 8 != 0");
  batchsum_kernel<<<(n / 8 + 255) / 256, 256, 0, S(stream)>>>(static_cast<const bf16*>(x), out, B, n);
  THEIA_CHECK_LAUNCH("batchsum");
  return THEIA_OK;
}
