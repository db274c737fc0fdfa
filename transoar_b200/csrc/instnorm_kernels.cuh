// instnorm_kernels.cuh -- fused InstanceNorm3d(affine) + ReLU, forward and gradient, for sm_100a.
//
// Replaces the `InstanceNorm3d -> ReLU(inplace)` pairs of the reference's EncoderCnnBlock
// (transoar/models/backbones/encoder_blocks.py:28-46), which ATen runs as cuDNN batch-norm kernels with one "batch" per
// (sample, channel): at the VISCERAL shape the first encoder stage normalises 24 instances of 6.55 M voxels (629 MB per
// activation) and the cuDNN kernels take 8.4 ms forward / 12.8 ms backward per layer pair on B200 -- ~25 % of the
// whole-model step -- for what is 1.9 GB / 3.1 GB of compulsory HBM traffic (0.3 / 0.5 ms at the measured peak).
//
// Layout: x, y [I][V] with I = N*C instances of V = D*H*W contiguous voxels (NCDHW), gamma/beta [C], c = i % C.
//   forward : y = relu((x - mean_i) * rstd_i * gamma_c + beta_c), mean / biased var over V, rstd = 1/sqrt(var + eps)
//   backward: dz = dy * (y > 0);  dbeta_c = sum dz;  dgamma_c = sum dz * xhat;
//             dx = gamma_c * rstd_i * (dz - mean_V(dz) - xhat * mean_V(dz * xhat))
// Three streaming passes each way (partial statistics -> finalize -> apply); statistics are combined per block with
// Chan's parallel-variance formula in fp32 and merged across blocks in fp64.  16-byte vector loads, grid sized so
// that every SM holds several CTAs, storage fp32 or bf16 with fp32 arithmetic.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace instnorm {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 8;                 // 16-byte vectors per thread per chunk
template <typename T> struct Pack;               // 16-byte vector of T <-> floats
template <> struct Pack<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float *p, float (&v)[4])
  {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Pack<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8])
  {
    const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __bfloat162float(h[i].x); v[2 * i + 1] = __bfloat162float(h[i].y); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8])
  {
    uint4 t;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i].x = __float2bfloat16_rn(v[2 * i]); h[i].y = __float2bfloat16_rn(v[2 * i + 1]); }
    *reinterpret_cast<uint4 *>(p) = t;
  }
};

template <typename T> __host__ __device__ constexpr int chunk_elems() { return kThreads * kVecPerThread * Pack<T>::N; }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// block-wide sum of two values; result valid in thread 0
__device__ __forceinline__ void block_sum2(float &a, float &b)
{
  __shared__ float sa[kThreads / 32], sb[kThreads / 32];
  a = warp_sum(a); b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = l < kThreads / 32 ? sa[l] : 0.f;
    b = l < kThreads / 32 ? sb[l] : 0.f;
    a = warp_sum(a); b = warp_sum(b);
  }
  __syncthreads();
}

// ---- forward pass 1: per (instance, chunk) count / mean / M2 ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
stats_partial_kernel(const T *__restrict__ x, long long V, int chunks, int vec_ok, float *__restrict__ part)
{
  constexpr int N = Pack<T>::N, CH = chunk_elems<T>();
  const int inst = blockIdx.y, ch = blockIdx.x;
  const long long beg = (long long)ch * CH, end = min(V, beg + CH);
  const T *base = x + (long long)inst * V;
  float s = 0.f;
  float vals[kVecPerThread][N];
  bool have[kVecPerThread];
#pragma unroll
  for (int k = 0; k < kVecPerThread; ++k) {
    const long long e = beg + ((long long)k * kThreads + threadIdx.x) * N;
    have[k] = vec_ok && e + N <= end;
    if (have[k]) {
      Pack<T>::load(base + e, vals[k]);
#pragma unroll
      for (int j = 0; j < N; ++j) s += vals[k][j];
    } else {
      for (long long t = e; t < end && t < e + N; ++t) s += (float)base[t];      // ragged tail / unaligned instance
    }
  }
  float unused = 0.f;
  block_sum2(s, unused);
  __shared__ float s_mean;
  if (threadIdx.x == 0) s_mean = s / (float)(end - beg);
  __syncthreads();
  const float mean = s_mean;
  float m2 = 0.f;
#pragma unroll
  for (int k = 0; k < kVecPerThread; ++k) {
    const long long e = beg + ((long long)k * kThreads + threadIdx.x) * N;
    if (have[k]) {
#pragma unroll
      for (int j = 0; j < N; ++j) { const float d = vals[k][j] - mean; m2 = fmaf(d, d, m2); }
    } else {
      for (long long t = e; t < end && t < e + N; ++t) { const float d = (float)base[t] - mean; m2 = fmaf(d, d, m2); }
    }
  }
  block_sum2(m2, unused);
  if (threadIdx.x == 0) {
    float *p = part + ((long long)inst * chunks + ch) * 3;
    p[0] = (float)(end - beg); p[1] = mean; p[2] = m2;
  }
}

// ---- forward pass 2: merge the chunks of an instance (Chan et al.), one warp per instance --------------------------
__global__ void stats_finalize_kernel(const float *__restrict__ part, int chunks, int instances, float eps, float *__restrict__ mean,
                                      float *__restrict__ rstd)
{
  const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (inst >= instances) return;
  double n = 0.0, mu = 0.0, m2 = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    const float *p = part + ((long long)inst * chunks + c) * 3;
    const double nb = p[0], mb = p[1], m2b = p[2];
    const double nn = n + nb, delta = mb - mu;
    mu += delta * nb / nn;
    m2 += m2b + delta * delta * n * nb / nn;
    n = nn;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double nb = __shfl_xor_sync(0xffffffffu, n, d), mb = __shfl_xor_sync(0xffffffffu, mu, d), m2b = __shfl_xor_sync(0xffffffffu, m2, d);
    const double nn = n + nb;
    if (nn > 0.0) {
      const double delta = mb - mu;
      mu += delta * nb / nn;
      m2 += m2b + delta * delta * n * nb / nn;
    }
    n = nn;
  }
  if (lane == 0) {
    mean[inst] = (float)mu;
    rstd[inst] = (float)(1.0 / sqrt(m2 / n + (double)eps));
  }
}

// ---- forward pass 3: normalise + affine + ReLU ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ mean,
             const float *__restrict__ rstd, long long V, int C, int vec_ok, T *__restrict__ y)
{
  constexpr int N = Pack<T>::N, CH = chunk_elems<T>();
  const int inst = blockIdx.y;
  const long long beg = (long long)blockIdx.x * CH, end = min(V, beg + CH);
  const float a = rstd[inst] * gamma[inst % C], b = beta[inst % C] - mean[inst] * a;
  const T *xi = x + (long long)inst * V;
  T *yi = y + (long long)inst * V;
#pragma unroll
  for (int k = 0; k < kVecPerThread; ++k) {
    const long long e = beg + ((long long)k * kThreads + threadIdx.x) * N;
    if (vec_ok && e + N <= end) {
      float v[N];
      Pack<T>::load(xi + e, v);
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaxf(fmaf(v[j], a, b), 0.f);
      Pack<T>::store(yi + e, v);
    } else {
      for (long long t = e; t < end && t < e + N; ++t) yi[t] = (T)fmaxf(fmaf((float)xi[t], a, b), 0.f);
    }
  }
}

// ---- backward pass 1: per (instance, chunk) sums of dz and dz * xhat -----------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
bwd_partial_kernel(const T *__restrict__ dy, const T *__restrict__ x, const T *__restrict__ y, const float *__restrict__ mean,
                   const float *__restrict__ rstd, long long V, int chunks, int vec_ok, float *__restrict__ part)
{
  constexpr int N = Pack<T>::N, CH = chunk_elems<T>();
  const int inst = blockIdx.y, ch = blockIdx.x;
  const long long beg = (long long)ch * CH, end = min(V, beg + CH);
  const long long off = (long long)inst * V;
  const float mu = mean[inst], rs = rstd[inst];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < kVecPerThread; ++k) {
    const long long e = beg + ((long long)k * kThreads + threadIdx.x) * N;
    if (vec_ok && e + N <= end) {
      float g[N], xv[N], yv[N];
      Pack<T>::load(dy + off + e, g); Pack<T>::load(x + off + e, xv); Pack<T>::load(y + off + e, yv);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float dz = yv[j] > 0.f ? g[j] : 0.f;
        s1 += dz;
        s2 = fmaf(dz, (xv[j] - mu) * rs, s2);
      }
    } else {
      for (long long t = e; t < end && t < e + N; ++t) {
        const float dz = (float)y[off + t] > 0.f ? (float)dy[off + t] : 0.f;
        s1 += dz;
        s2 = fmaf(dz, ((float)x[off + t] - mu) * rs, s2);
      }
    }
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    float *p = part + ((long long)inst * chunks + ch) * 2;
    p[0] = s1; p[1] = s2;
  }
}

// ---- backward pass 2: per-instance sums (fp64 merge), then dgamma / dbeta per channel: one warp per instance -----------
__global__ void bwd_finalize_kernel(const float *__restrict__ part, int chunks, int instances, float *__restrict__ sums)
{
  const int inst = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (inst >= instances) return;
  double s1 = 0.0, s2 = 0.0;
  for (int c = lane; c < chunks; c += 32) {
    const float *p = part + ((long long)inst * chunks + c) * 2;
    s1 += p[0]; s2 += p[1];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, d); s2 += __shfl_xor_sync(0xffffffffu, s2, d); }
  if (lane == 0) { sums[2 * inst] = (float)s1; sums[2 * inst + 1] = (float)s2; }
}

__global__ void bwd_param_kernel(const float *__restrict__ sums, int batch, int C, float *__restrict__ dgamma, float *__restrict__ dbeta)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, b = 0.0;
  for (int n = 0; n < batch; ++n) { b += sums[2 * (n * C + c)]; g += sums[2 * (n * C + c) + 1]; }
  dgamma[c] = (float)g; dbeta[c] = (float)b;
}

// ---- backward pass 3: dx ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
bwd_apply_kernel(const T *__restrict__ dy, const T *__restrict__ x, const T *__restrict__ y, const float *__restrict__ gamma,
                 const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ sums, long long V, int C,
                 int vec_ok, T *__restrict__ dx)
{
  constexpr int N = Pack<T>::N, CH = chunk_elems<T>();
  const int inst = blockIdx.y;
  const long long beg = (long long)blockIdx.x * CH, end = min(V, beg + CH);
  const long long off = (long long)inst * V;
  const float mu = mean[inst], rs = rstd[inst], a = gamma[inst % C] * rs;
  const float m1 = sums[2 * inst] / (float)V, m2 = sums[2 * inst + 1] / (float)V;
#pragma unroll
  for (int k = 0; k < kVecPerThread; ++k) {
    const long long e = beg + ((long long)k * kThreads + threadIdx.x) * N;
    if (vec_ok && e + N <= end) {
      float g[N], xv[N], yv[N];
      Pack<T>::load(dy + off + e, g); Pack<T>::load(x + off + e, xv); Pack<T>::load(y + off + e, yv);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float dz = yv[j] > 0.f ? g[j] : 0.f;
        g[j] = a * (dz - m1 - (xv[j] - mu) * rs * m2);
      }
      Pack<T>::store(dx + off + e, g);
    } else {
      for (long long t = e; t < end && t < e + N; ++t) {
        const float dz = (float)y[off + t] > 0.f ? (float)dy[off + t] : 0.f;
        dx[off + t] = (T)(a * (dz - m1 - ((float)x[off + t] - mu) * rs * m2));
      }
    }
  }
}

// =====================================================================================================================
// Channels-last (NDHWC) variants, fp32.  x, y [B][V][C]: what cuDNN's tensor-core convolutions produce and consume, so a
// channels-last encoder needs no NCDHW <-> NDHWC transposes around every convolution (14 ms of the 126 ms VISCERAL step).
// A CTA streams a run of voxels of one sample; thread t owns the float4 channel group t % (C/4) of voxel rows t / (C/4),
// so a warp still reads whole 128-byte lines.  Per-channel sums are reduced over the voxel rows of the CTA in shared memory.
// Statistics use the shifted single-pass form (shift = the chunk's first voxel) and the same fp64 merge as above.
// Requires C % 4 == 0 and (C/4) | 192 -- every channel count of the reference's encoders (24 ... 768).
// =====================================================================================================================
constexpr int kClThreads = 192;
constexpr int kClUnroll = 8;

__host__ __device__ inline bool cl_supported(int C) { return C > 0 && C % 4 == 0 && kClThreads % (C / 4) == 0; }

// four channels of storage type T (fp32: 16 bytes, bf16: 8 bytes) <-> fp32 registers; the arithmetic is fp32 for both
__device__ __forceinline__ float4 cl_ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 cl_ld4(const __nv_bfloat16 *p)
{
  const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
  return make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(b.x), __bfloat162float(b.y));
}
__device__ __forceinline__ void cl_st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void cl_st4(__nv_bfloat16 *p, float4 v)
{
  uint2 t;
  *reinterpret_cast<__nv_bfloat162 *>(&t.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162 *>(&t.y) = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2 *>(p) = t;
}

// sum NV float4-wide values over the voxel rows of the CTA; result valid in the threads of row 0 (t < cg)
template <int NV>
__device__ __forceinline__ void cl_reduce_rows(float (&v)[NV][4], int cg, int rows, float *red /* [kClThreads][NV * 4] */)
{
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[t * (NV * 4) + i * 4 + j] = v[i][j];
  for (int s = rows >> 1; s > 0; s >>= 1) {
    __syncthreads();
    if (t < s * cg) {
#pragma unroll
      for (int k = 0; k < NV * 4; ++k) red[t * (NV * 4) + k] += red[(t + s * cg) * (NV * 4) + k];
    }
  }
  __syncthreads();
  if (t < cg) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = red[t * (NV * 4) + i * 4 + j];
  }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads)
cl_stats_partial_kernel(const T *__restrict__ x, long long V, int C, int chunks, long long chunk_vox, float *__restrict__ part)
{
  __shared__ float red[kClThreads * 8];
  const int cg = C >> 2, rows = kClThreads / cg, g = threadIdx.x % cg, r = threadIdx.x / cg;
  const int b = blockIdx.y, ch = blockIdx.x;
  const long long v0 = (long long)ch * chunk_vox, v1 = min(V, v0 + chunk_vox);
  const T *base = x + (long long)b * V * C + g * 4;
  const float4 K = cl_ld4(base + v0 * C);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (long long v = v0 + r; v < v1; v += (long long)rows * kClUnroll) {
    float4 q[kClUnroll];
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      q[u] = vv < v1 ? cl_ld4(base + vv * C) : K;
    }
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const float d0 = q[u].x - K.x, d1 = q[u].y - K.y, d2 = q[u].z - K.z, d3 = q[u].w - K.w;
      acc[0][0] += d0; acc[0][1] += d1; acc[0][2] += d2; acc[0][3] += d3;
      acc[1][0] = fmaf(d0, d0, acc[1][0]); acc[1][1] = fmaf(d1, d1, acc[1][1]);
      acc[1][2] = fmaf(d2, d2, acc[1][2]); acc[1][3] = fmaf(d3, d3, acc[1][3]);
    }
  }
  cl_reduce_rows<2>(acc, cg, rows, red);
  if (threadIdx.x < cg) {
    const float n = (float)(v1 - v0);
    const float k4[4] = {K.x, K.y, K.z, K.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float *p = part + (((long long)b * C + g * 4 + j) * chunks + ch) * 3;
      p[0] = n; p[1] = k4[j] + acc[0][j] / n; p[2] = fmaxf(acc[1][j] - acc[0][j] * acc[0][j] / n, 0.f);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads)
cl_apply_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ mean,
                const float *__restrict__ rstd, long long V, int C, long long chunk_vox, T *__restrict__ y)
{
  const int cg = C >> 2, rows = kClThreads / cg, g = threadIdx.x % cg, r = threadIdx.x / cg;
  const int b = blockIdx.y;
  const long long v0 = (long long)blockIdx.x * chunk_vox, v1 = min(V, v0 + chunk_vox);
  float a[4], o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = g * 4 + j;
    a[j] = rstd[b * C + c] * gamma[c];
    o[j] = beta[c] - mean[b * C + c] * a[j];
  }
  const long long off = (long long)b * V * C + g * 4;
  for (long long v = v0 + r; v < v1; v += (long long)rows * kClUnroll) {
    float4 q[kClUnroll];
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1) q[u] = cl_ld4(x + off + vv * C);
    }
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1)
        cl_st4(y + off + vv * C, make_float4(fmaxf(fmaf(q[u].x, a[0], o[0]), 0.f), fmaxf(fmaf(q[u].y, a[1], o[1]), 0.f),
                                             fmaxf(fmaf(q[u].z, a[2], o[2]), 0.f), fmaxf(fmaf(q[u].w, a[3], o[3]), 0.f)));
    }
  }
}

// The ReLU mask is recomputed from x with the forward's own arithmetic (fmaf(x, rstd * gamma, beta - mean * rstd * gamma) > 0), so the
// backward reads two tensors per pass (dy, x) instead of three (dy, x, y): 5 instead of 7 passes over the activation in total.
template <typename T>
__global__ void __launch_bounds__(kClThreads)
cl_bwd_partial_kernel(const T *__restrict__ dy, const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                      const float *__restrict__ mean, const float *__restrict__ rstd, long long V, int C, int chunks, long long chunk_vox,
                      float *__restrict__ part)
{
  __shared__ float red[kClThreads * 8];
  const int cg = C >> 2, rows = kClThreads / cg, g = threadIdx.x % cg, r = threadIdx.x / cg;
  const int b = blockIdx.y, ch = blockIdx.x;
  const long long v0 = (long long)ch * chunk_vox, v1 = min(V, v0 + chunk_vox);
  float mu[4], rs[4], fa[4], fo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = g * 4 + j;
    mu[j] = mean[b * C + c]; rs[j] = rstd[b * C + c];
    fa[j] = rs[j] * gamma[c]; fo[j] = beta[c] - mu[j] * fa[j];                 // exactly cl_apply_kernel's a / o
  }
  const long long off = (long long)b * V * C + g * 4;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (long long v = v0 + r; v < v1; v += (long long)rows * kClUnroll) {
    float4 gq[kClUnroll], xq[kClUnroll];
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1) {
        gq[u] = cl_ld4(dy + off + vv * C);
        xq[u] = cl_ld4(x + off + vv * C);
      }
    }
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1) {
        const float gg[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w}, xx[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dz = fmaf(xx[j], fa[j], fo[j]) > 0.f ? gg[j] : 0.f;
          acc[0][j] += dz;
          acc[1][j] = fmaf(dz, (xx[j] - mu[j]) * rs[j], acc[1][j]);
        }
      }
    }
  }
  cl_reduce_rows<2>(acc, cg, rows, red);
  if (threadIdx.x < cg) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float *p = part + (((long long)b * C + g * 4 + j) * chunks + ch) * 2;
      p[0] = acc[0][j]; p[1] = acc[1][j];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kClThreads)
cl_bwd_apply_kernel(const T *__restrict__ dy, const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ sums, long long V, int C,
                    long long chunk_vox, T *__restrict__ dx)
{
  const int cg = C >> 2, rows = kClThreads / cg, g = threadIdx.x % cg, r = threadIdx.x / cg;
  const int b = blockIdx.y;
  const long long v0 = (long long)blockIdx.x * chunk_vox, v1 = min(V, v0 + chunk_vox);
  float mu[4], rs[4], a[4], fo[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = g * 4 + j, i = b * C + c;
    mu[j] = mean[i]; rs[j] = rstd[i]; a[j] = rs[j] * gamma[c]; fo[j] = beta[c] - mu[j] * a[j];
    m1[j] = sums[2 * i] / (float)V; m2[j] = sums[2 * i + 1] / (float)V;
  }
  const long long off = (long long)b * V * C + g * 4;
  for (long long v = v0 + r; v < v1; v += (long long)rows * kClUnroll) {
    float4 gq[kClUnroll], xq[kClUnroll];
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1) {
        gq[u] = cl_ld4(dy + off + vv * C);
        xq[u] = cl_ld4(x + off + vv * C);
      }
    }
#pragma unroll
    for (int u = 0; u < kClUnroll; ++u) {
      const long long vv = v + (long long)u * rows;
      if (vv < v1) {
        const float gg[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w}, xx[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dz = fmaf(xx[j], a[j], fo[j]) > 0.f ? gg[j] : 0.f;
          o[j] = a[j] * (dz - m1[j] - (xx[j] - mu[j]) * rs[j] * m2[j]);
        }
        cl_st4(dx + off + vv * C, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }
}

}  // namespace instnorm
