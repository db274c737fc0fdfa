// fused_ln_kernels.cuh -- y = LayerNorm(a + dropout(b)) in one pass each way, for sm_100a.
//
// The post-norm residual blocks of the hot path -- DefAttnLayer (transoar/models/backbones/decoder_blocks.py:163-177:
// `src = norm1(src + dropout1(attn))`, `src = norm2(src + dropout3(ffn))`) and FocusedDecoderLayer
// (transoar/models/necks/focused_decoder.py:166-189) -- run in the reference as three ATen kernels forward (dropout, add,
// layer_norm: 8 passes over a [234000, 384] fp32 tensor = 360 MB each at VISCERAL) and four backward (layer-norm input gradient,
// the gamma/beta reduction, the dropout mask scale, gradient accumulation).  Fused: forward reads a, b and writes z = a + dropout(b)
// (kept for the backward) and y; backward reads dy, z and writes da (= dz) and db (= dz * mask / (1 - p)); the dropout mask is
// never stored -- it is a counter-based hash of (seed, element index) evaluated again in the backward.
//
// One warp per row, lane l owns the float4 column groups l, l + 32, ... (C % 4 == 0, C <= 1024); row statistics by warp
// shuffles (two-pass: mean, then centred variance, both from registers); persistent grid, gamma/beta gradients accumulated in
// registers over all rows of a warp, reduced across the CTA in shared memory, one partial per CTA, summed by a second kernel.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hash_rng.cuh"

namespace fusedln {

using hashrng::keep4;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxNV = 8;                     // float4 groups per lane: C <= 4 * 32 * 8 = 1024

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// four consecutive elements of the branch tensor b / its gradient db, stored as fp32 or bf16 (the bf16 route: the branch is the output
// of a bf16 GEMM, the residual stream a / z / y stays fp32 -- what torch.autocast does, whose layer_norm runs and returns fp32)
__device__ __forceinline__ float4 ldb4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ldb4(const __nv_bfloat16 *p)
{
  const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&t.x), hi = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
  return make_float4(__bfloat162float(lo.x), __bfloat162float(lo.y), __bfloat162float(hi.x), __bfloat162float(hi.y));
}
__device__ __forceinline__ void stb4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void stb4(__nv_bfloat16 *p, float4 v)
{
  uint2 t;
  *reinterpret_cast<__nv_bfloat162 *>(&t.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162 *>(&t.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2 *>(p) = t;
}

template <int NV, typename TB = float>
__global__ void __launch_bounds__(kThreads)
fwd_kernel(const float *__restrict__ a, const TB *__restrict__ b, const float *__restrict__ gamma, const float *__restrict__ beta,
           long long rows, int C, float eps, uint32_t thresh, float scale, uint64_t seed, const unsigned long long *__restrict__ epoch,
           float *__restrict__ z, float *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd)
{
  seed = hashrng::with_epoch(seed, epoch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c4 = C >> 2;
  float g[NV][4], be[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * 32 + lane;
    const float4 gv = i < c4 ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bv = i < c4 ? __ldg(reinterpret_cast<const float4 *>(beta) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[k][0] = gv.x; g[k][1] = gv.y; g[k][2] = gv.z; g[k][3] = gv.w;
    be[k][0] = bv.x; be[k][1] = bv.y; be[k][2] = bv.z; be[k][3] = bv.w;
  }
  const float inv_c = 1.f / (float)C;
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < rows; row += (long long)gridDim.x * kWarps) {
    const float4 *ar = reinterpret_cast<const float4 *>(a + row * C);
    const TB *br = b != nullptr ? b + row * C : nullptr;
    float v[NV][4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * 32 + lane;
      if (i < c4) {
        const float4 av = __ldg(ar + i);
        v[k][0] = av.x; v[k][1] = av.y; v[k][2] = av.z; v[k][3] = av.w;
        if (br != nullptr) {
          const float4 bv = ldb4(br + 4 * i);
          float m[4] = {1.f, 1.f, 1.f, 1.f};
          if (thresh != 0u) keep4(seed, (uint64_t)row * c4 + i, thresh, scale, m);
          v[k][0] = fmaf(bv.x, m[0], v[k][0]); v[k][1] = fmaf(bv.y, m[1], v[k][1]);
          v[k][2] = fmaf(bv.z, m[2], v[k][2]); v[k][3] = fmaf(bv.w, m[3], v[k][3]);
        }
        s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
      } else {
        v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.f;
      }
    }
    const float mu = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (k * 32 + lane < c4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = v[k][j] - mu; q = fmaf(d, d, q); }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_c + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    float4 *zr = z != nullptr ? reinterpret_cast<float4 *>(z + row * C) : nullptr;
    float4 *yr = reinterpret_cast<float4 *>(y + row * C);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * 32 + lane;
      if (i < c4) {
        if (zr != nullptr) zr[i] = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
        yr[i] = make_float4(fmaf((v[k][0] - mu) * rs, g[k][0], be[k][0]), fmaf((v[k][1] - mu) * rs, g[k][1], be[k][1]),
                            fmaf((v[k][2] - mu) * rs, g[k][2], be[k][2]), fmaf((v[k][3] - mu) * rs, g[k][3], be[k][3]));
      }
    }
  }
}

// part [gridDim.x][2][C]: per-CTA sums of dy * xhat (dgamma) and dy (dbeta)
template <int NV, typename TB = float>
__global__ void __launch_bounds__(kThreads)
bwd_kernel(const float *__restrict__ dy, const float *__restrict__ z, const float *__restrict__ gamma, const float *__restrict__ mean,
           const float *__restrict__ rstd, long long rows, int C, uint32_t thresh, float scale, uint64_t seed,
           const unsigned long long *__restrict__ epoch, float *__restrict__ da, TB *__restrict__ db, float *__restrict__ part)
{
  seed = hashrng::with_epoch(seed, epoch);
  extern __shared__ float red[];                                   // [kWarps][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c4 = C >> 2;
  float g[NV][4], dg[NV][4], dbt[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * 32 + lane;
    const float4 gv = i < c4 ? __ldg(reinterpret_cast<const float4 *>(gamma) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[k][0] = gv.x; g[k][1] = gv.y; g[k][2] = gv.z; g[k][3] = gv.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) { dg[k][j] = 0.f; dbt[k][j] = 0.f; }
  }
  const float inv_c = 1.f / (float)C;
  for (long long row = (long long)blockIdx.x * kWarps + warp; row < rows; row += (long long)gridDim.x * kWarps) {
    const float4 *dr = reinterpret_cast<const float4 *>(dy + row * C);
    const float4 *zr = reinterpret_cast<const float4 *>(z + row * C);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float gy[NV][4], xh[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * 32 + lane;
      if (i < c4) {
        const float4 dv = __ldg(dr + i), zv = __ldg(zr + i);
        const float d4[4] = {dv.x, dv.y, dv.z, dv.w}, z4[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[k][j] = (z4[j] - mu) * rs;
          dg[k][j] = fmaf(d4[j], xh[k][j], dg[k][j]);
          dbt[k][j] += d4[j];
          gy[k][j] = d4[j] * g[k][j];
          s1 += gy[k][j];
          s2 = fmaf(gy[k][j], xh[k][j], s2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { gy[k][j] = 0.f; xh[k][j] = 0.f; }
      }
    }
    const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
    float4 *ar = reinterpret_cast<float4 *>(da + row * C);
    TB *br = db != nullptr ? db + row * C : nullptr;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = k * 32 + lane;
      if (i < c4) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rs * (gy[k][j] - m1 - xh[k][j] * m2);
        ar[i] = make_float4(o[0], o[1], o[2], o[3]);
        if (br != nullptr) {
          float m[4] = {1.f, 1.f, 1.f, 1.f};
          if (thresh != 0u) keep4(seed, (uint64_t)row * c4 + i, thresh, scale, m);
          stb4(br + 4 * i, make_float4(o[0] * m[0], o[1] * m[1], o[2] * m[2], o[3] * m[3]));
        }
      }
    }
  }
  // CTA reduction of the parameter gradients, then one partial per CTA
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = k * 32 + lane;
    if (i < c4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[(warp * 2 + 0) * C + i * 4 + j] = dg[k][j];
        red[(warp * 2 + 1) * C + i * 4 + j] = dbt[k][j];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w * 2 * C + i];
    part[(long long)blockIdx.x * 2 * C + i] = s;
  }
}

// 256 threads per 32 entries of [dgamma | dbeta]: warp w sums partials w, w + 8, ..., combined in shared memory
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const float *__restrict__ part, int ctas, int C, float *__restrict__ dgamma,
                                                          float *__restrict__ dbeta)
{
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (i < 2 * C)
    for (int b = w; b < ctas; b += 8) s += part[(long long)b * 2 * C + i];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < 2 * C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][lane];
    if (i < C) dgamma[i] = (float)s;
    else dbeta[i - C] = (float)s;
  }
}

}  // namespace fusedln
