// stem_conv_kernels.cuh -- the first convolution of the AttnFPN encoder (1 input channel -> CO feature maps, 3x3x3, stride 1,
// padding 1, no bias: EncoderCnnBlock._block[0] of stage 0, transoar/models/backbones/encoder_blocks.py:28-33 with
// in_channels = 1, attn_fpn.py:170-182) as a direct stencil, forward and weight gradient, for sm_100a.
//
// Why not a GEMM: the reduction length is 27 (one channel x 27 taps), the output is 629 MB per sample at 160x160x256 and the
// whole layer is 8.5 GFLOP per sample -- an HBM-bound stencil (1.26 GB of output per batch of 2 = 0.19 ms at the measured peak).
// cuDNN runs it as an implicit GEMM on a generic "indexed, no shared memory" kernel: 7.8 ms forward + 4.0 ms weight gradient
// per step on B200.  Here:
//   forward : one CTA per (sample, d, h) row segment of 256 voxels, two voxels per thread, the 27 x CO weights broadcast from
//             shared memory (one LDS.128 feeds 8 FMAs), fp32 FMA (no TF32 rounding), output written channels-last
//             (NDHWC) so the next layer's tensor-core convolution and the fused InstanceNorm read it without a transpose.
//   wgrad   : dW[co][kd][kh][kw] = sum_voxels dy[voxel][co] * x[voxel + tap].  A group of 9 lanes walks one row of voxels;
//             lane j owns the row tap (kd, kh) = j with a 3-wide sliding window of x along w and keeps CO x 3 accumulators, the
//             CO gradients of the voxel arrive as warp-broadcast 16-byte loads.  Per-CTA partial sums go to a workspace and a
//             second kernel adds them (deterministic, no atomics).
// The input gradient is not provided: the layer's input is the CT volume.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace stemconv {

constexpr int kFwdThreads = 128;
constexpr int kFwdVox = 2 * kFwdThreads;      // voxels of one row per CTA
constexpr int kWgThreads = 128;               // 4 warps x 3 groups of 9 lanes
constexpr int kWgGroups = (kWgThreads / 32) * 3;

template <int CO>
__global__ void __launch_bounds__(kFwdThreads)
fwd_kernel(const float *__restrict__ x, const float *__restrict__ weight, int N, int D, int H, int W, float *__restrict__ y)
{
  __shared__ float4 ws[27][CO / 4];
  for (int i = threadIdx.x; i < 27 * CO; i += kFwdThreads) {
    const int tap = i / CO, co = i % CO;
    reinterpret_cast<float *>(&ws[tap][0])[co] = weight[co * 27 + tap];            // weight [CO][1][3][3][3]
  }
  __syncthreads();
  const int wtiles = (W + kFwdVox - 1) / kFwdVox;
  const long long rows = (long long)N * D * H;
  for (long long job = blockIdx.x; job < rows * wtiles; job += gridDim.x) {
    const int wt = (int)(job % wtiles);
    const long long row = job / wtiles;
    const int h = (int)(row % H), d = (int)((row / H) % D);
    const long long n = row / ((long long)H * D);
    const int w0 = wt * kFwdVox + threadIdx.x, w1 = w0 + kFwdThreads;
    float a0[CO], a1[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      if (dd < 0 || dd >= D) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hh = h + kh - 1;
        if (hh < 0 || hh >= H) continue;
        const float *xr = x + ((n * D + dd) * H + hh) * (long long)W;
        float p[3], q[3];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int u0 = w0 + kw - 1, u1 = w1 + kw - 1;
          p[kw] = (u0 >= 0 && u0 < W) ? __ldg(xr + u0) : 0.f;
          q[kw] = (u1 >= 0 && u1 < W) ? __ldg(xr + u1) : 0.f;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 wv = ws[tap][c4];
            a0[c4 * 4 + 0] = fmaf(p[kw], wv.x, a0[c4 * 4 + 0]); a0[c4 * 4 + 1] = fmaf(p[kw], wv.y, a0[c4 * 4 + 1]);
            a0[c4 * 4 + 2] = fmaf(p[kw], wv.z, a0[c4 * 4 + 2]); a0[c4 * 4 + 3] = fmaf(p[kw], wv.w, a0[c4 * 4 + 3]);
            a1[c4 * 4 + 0] = fmaf(q[kw], wv.x, a1[c4 * 4 + 0]); a1[c4 * 4 + 1] = fmaf(q[kw], wv.y, a1[c4 * 4 + 1]);
            a1[c4 * 4 + 2] = fmaf(q[kw], wv.z, a1[c4 * 4 + 2]); a1[c4 * 4 + 3] = fmaf(q[kw], wv.w, a1[c4 * 4 + 3]);
          }
        }
      }
    }
    float *yr = y + row * (long long)W * CO;
    if (w0 < W) {
#pragma unroll
      for (int c4 = 0; c4 < CO / 4; ++c4)
        __stcs(reinterpret_cast<float4 *>(yr + (long long)w0 * CO) + c4, make_float4(a0[c4 * 4], a0[c4 * 4 + 1], a0[c4 * 4 + 2], a0[c4 * 4 + 3]));
    }
    if (w1 < W) {
#pragma unroll
      for (int c4 = 0; c4 < CO / 4; ++c4)
        __stcs(reinterpret_cast<float4 *>(yr + (long long)w1 * CO) + c4, make_float4(a1[c4 * 4], a1[c4 * 4 + 1], a1[c4 * 4 + 2], a1[c4 * 4 + 3]));
    }
  }
}

// part [gridDim.x][27][CO]
template <int CO>
__global__ void __launch_bounds__(kWgThreads)
wgrad_partial_kernel(const float *__restrict__ dy, const float *__restrict__ x, int N, int D, int H, int W, float *__restrict__ part)
{
  __shared__ float acc_s[27 * CO];
  for (int i = threadIdx.x; i < 27 * CO; i += kWgThreads) acc_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / 9, j = lane % 9;                         // lanes 27..31: grp == 3, idle
  const bool active = grp < 3;
  const int kd = j / 3, kh = j % 3;
  float acc[3][CO];
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[kw][c] = 0.f;
  const long long rows = (long long)N * D * H;
  const long long g0 = (long long)blockIdx.x * kWgGroups + warp * 3 + (active ? grp : 0);
  const long long gstride = (long long)gridDim.x * kWgGroups;
  if (active) {
    for (long long row = g0; row < rows; row += gstride) {
      const int h = (int)(row % H), d = (int)((row / H) % D);
      const long long n = row / ((long long)H * D);
      const int dd = d + kd - 1, hh = h + kh - 1;
      const bool row_ok = dd >= 0 && dd < D && hh >= 0 && hh < H;
      const float *xr = x + ((n * D + (row_ok ? dd : 0)) * H + (row_ok ? hh : 0)) * (long long)W;
      const float4 *gy = reinterpret_cast<const float4 *>(dy + row * (long long)W * CO);
      float xm = 0.f, xc = row_ok ? __ldg(xr) : 0.f;            // x[w-1], x[w]
      for (int w = 0; w < W; ++w) {
        const float xp = (row_ok && w + 1 < W) ? __ldg(xr + w + 1) : 0.f;
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) {
          const float4 g = __ldg(gy + (long long)w * (CO / 4) + c4);
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[0][c4 * 4 + i] = fmaf(gv[i], xm, acc[0][c4 * 4 + i]);
            acc[1][c4 * 4 + i] = fmaf(gv[i], xc, acc[1][c4 * 4 + i]);
            acc[2][c4 * 4 + i] = fmaf(gv[i], xp, acc[2][c4 * 4 + i]);
          }
        }
        xm = xc; xc = xp;
      }
    }
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int c = 0; c < CO; ++c) atomicAdd(&acc_s[((kd * 3 + kh) * 3 + kw) * CO + c], acc[kw][c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * CO; i += kWgThreads) part[(long long)blockIdx.x * (27 * CO) + i] = acc_s[i];
}

// dweight [CO][27] = sum over CTAs of part [ctas][27][CO]
__global__ void wgrad_finalize_kernel(const float *__restrict__ part, int ctas, int CO, float *__restrict__ dweight)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;             // i = tap * CO + co
  if (i >= 27 * CO) return;
  double s = 0.0;
  for (int b = 0; b < ctas; ++b) s += part[(long long)b * (27 * CO) + i];
  dweight[(i % CO) * 27 + i / CO] = (float)s;
}

}  // namespace stemconv
