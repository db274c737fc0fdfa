// msda3d_kernels.cuh -- sm_100a kernels for 3D multi-scale deformable attention (forward + gradient).
//
// Replaces the reference's kernels in transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh ("cuh" below):
//   forward  ms_deformable_im2col_gpu_kernel                      cuh:370-439   (+ trilinear device fn cuh:31-114)
//   backward ms_deformable_col2im_gpu_kernel_* (seven variants)   cuh:441-1092  (+ col2im device fns cuh:116-367)
//
// Design (DESIGN.md has the long form):
//   * one "unit" = one (batch, query, head) triple = C output channels.  The reference spends one THREAD per output
//     element, so all C threads of a unit redo the coordinate arithmetic and re-read loc / attn_weight; here a group
//     of G lanes owns a unit, each lane keeps 16 bytes (4 fp32 / 8 bf16 channels) x NV in registers, the L*P samples
//     of the unit are set up ONCE (lane j prepares sample j) and broadcast with warp shuffles.
//   * every corner fetch is one 16-byte read-only load per lane: a unit's corner is a single contiguous C*4-byte
//     segment, fully coalesced.  All 8 corner loads of a sample are issued before the first FMA.
//   * backward: grad_value goes out as 128-bit vector reductions (red.global.add.v4.f32 -> REDG.E.ADD.F32x4),
//     4x fewer L2 atomic transactions than the reference's scalar atomicAdd; grad_loc / grad_attn_weight are reduced
//     over channels with log2(G) shuffle steps instead of a shared-memory tree with 2+log2(C) block barriers per
//     sample (cuh:632-643), and written once per sample, coalesced.
//   * arithmetic that decides WHICH voxels are read is kept bit-identical to the compiled reference:
//     pixel = fma(loc, size, -0.5) (single rounding -- what nvcc -O3 emits for cuh:424-426, evidence in
//     profiles/r01_reference_fwd_sass.txt), range test on the pixel coordinate, floor, frac = pixel - floor.
//     The blend keeps the reference's operation order too, so fp32 forward results are bit-identical.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda3d {

constexpr int kMaxLevels = 16;     // levels cached in shared memory by the vector kernels
constexpr int kThreads = 256;      // 8 warps per CTA

// ---------------------------------------------------------------------------------------------------------------
// Sampling-index arithmetic (shared by every kernel and by the debug hook)
// ---------------------------------------------------------------------------------------------------------------
template <typename CT> struct Arith;
template <> struct Arith<float> {
  static __device__ __forceinline__ float pix(float loc, int size) { return __fmaf_rn(loc, __int2float_rn(size), -0.5f); }
  static __device__ __forceinline__ int floor_i(float x) { return __float2int_rd(x); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};
template <> struct Arith<double> {
  static __device__ __forceinline__ double pix(double loc, int size) { return __fma_rn(loc, (double)size, -0.5); }
  static __device__ __forceinline__ int floor_i(double x) { return __double2int_rd(x); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};

template <typename CT> struct Sample {
  unsigned mask;          // bit k set <=> corner k (reference order v1..v8: bit2=d high, bit1=h high, bit0=w high) is read
  int d_low, h_low, w_low;
  CT ld, lh, lw;
};

// cuh:424-428 (pixel coords + range test) and cuh:36-46 (floor / fractions) and the corner guards cuh:60-107.
template <typename CT>
__device__ __forceinline__ Sample<CT> locate(CT loc_w, CT loc_h, CT loc_d, int D, int H, int W)
{
  Sample<CT> s;
  const CT d_im = Arith<CT>::pix(loc_d, D);
  const CT h_im = Arith<CT>::pix(loc_h, H);
  const CT w_im = Arith<CT>::pix(loc_w, W);
  const bool in_range = d_im > (CT)-1 && h_im > (CT)-1 && w_im > (CT)-1 && d_im < (CT)D && h_im < (CT)H && w_im < (CT)W;
  s.d_low = Arith<CT>::floor_i(d_im);
  s.h_low = Arith<CT>::floor_i(h_im);
  s.w_low = Arith<CT>::floor_i(w_im);
  s.ld = d_im - (CT)s.d_low;
  s.lh = h_im - (CT)s.h_low;
  s.lw = w_im - (CT)s.w_low;
  unsigned mask = 0;
  if (in_range) {
    const unsigned dl = s.d_low >= 0, dh = s.d_low + 1 <= D - 1;
    const unsigned hl = s.h_low >= 0, hh = s.h_low + 1 <= H - 1;
    const unsigned wl = s.w_low >= 0, wh = s.w_low + 1 <= W - 1;
    const unsigned wm = wl | (wh << 1);                 // bits for (w low, w high)
    const unsigned hm = (hl ? wm : 0u) | ((hh ? wm : 0u) << 2);
    mask = (dl ? hm : 0u) | ((dh ? hm : 0u) << 4);
  }
  s.mask = mask;
  return s;
}

// cuh:109-110, products evaluated left to right.
template <typename CT>
__device__ __forceinline__ void corner_weights(CT ld, CT lh, CT lw, CT (&w)[8])
{
  using A = Arith<CT>;
  const CT hd = (CT)1 - ld, hh = (CT)1 - lh, hw = (CT)1 - lw;
  const CT a = A::mul(hd, hh), b = A::mul(hd, lh), c = A::mul(ld, hh), d = A::mul(ld, lh);
  w[0] = A::mul(a, hw); w[1] = A::mul(a, lw); w[2] = A::mul(b, hw); w[3] = A::mul(b, lw);
  w[4] = A::mul(c, hw); w[5] = A::mul(c, lw); w[6] = A::mul(d, hw); w[7] = A::mul(d, lw);
}

// cuh:112 in the operation order of the compiled reference: w2*v2, then fma(w1,v1,.), fma(w3,v3,.) ... fma(w8,v8,.).
template <typename CT>
__device__ __forceinline__ CT blend(const CT (&w)[8], CT v0, CT v1, CT v2, CT v3, CT v4, CT v5, CT v6, CT v7)
{
  using A = Arith<CT>;
  CT acc = A::mul(w[1], v1);
  acc = A::fma(w[0], v0, acc);
  acc = A::fma(w[2], v2, acc);
  acc = A::fma(w[3], v3, acc);
  acc = A::fma(w[4], v4, acc);
  acc = A::fma(w[5], v5, acc);
  acc = A::fma(w[6], v6, acc);
  acc = A::fma(w[7], v7, acc);
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Storage-type helpers
// ---------------------------------------------------------------------------------------------------------------
template <typename VT> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float up(float v) { return v; }
  static __device__ __forceinline__ float down(float v) { return v; }
};
template <> struct Cvt<double> {
  static __device__ __forceinline__ double up(double v) { return v; }
  static __device__ __forceinline__ double down(double v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float up(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 down(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float up(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half down(float v) { return __float2half_rn(v); }
};

// 16-byte vectors of the storage type, widened to fp32 in registers.
template <typename VT> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float *p, float (&v)[4])
  {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void load_stream(const float *p, float (&v)[4])
  {
    float4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&v)[4])
  {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <typename H2, typename H> struct Vec16Half {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4 t, float (&v)[8])
  {
    const H2 *h = reinterpret_cast<const H2 *>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = Cvt<H>::up(h[i].x);
      v[2 * i + 1] = Cvt<H>::up(h[i].y);
    }
  }
  static __device__ __forceinline__ void load(const H *p, float (&v)[8]) { unpack(__ldg(reinterpret_cast<const uint4 *>(p)), v); }
  static __device__ __forceinline__ void load_stream(const H *p, float (&v)[8])
  {
    uint4 t;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    unpack(t, v);
  }
  static __device__ __forceinline__ void store(H *p, const float (&v)[8])
  {
    uint4 t;
    H2 *h = reinterpret_cast<H2 *>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i].x = Cvt<H>::down(v[2 * i]);
      h[i].y = Cvt<H>::down(v[2 * i + 1]);
    }
    __stcs(reinterpret_cast<uint4 *>(p), t);
  }
};
template <> struct Vec16<__nv_bfloat16> : Vec16Half<__nv_bfloat162, __nv_bfloat16> {};
template <> struct Vec16<__half> : Vec16Half<__half2, __half> {};

__device__ __forceinline__ float ldg_stream(const float *p)
{
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// 128-bit vector reduction into global memory (sm_90+): one L2 atomic transaction per 4 floats.
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Vector kernels: C = G * NV * Vec16<VT>::N channels per unit, G lanes per unit (G | 32)
// ---------------------------------------------------------------------------------------------------------------
struct GroupSample {       // what lane j of a group prepares for sample (s0 + j) and the group then broadcasts
  unsigned mask;
  int off;                 // element offset of the (d_low,h_low,w_low) corner inside this batch element's value slab
  float ld, lh, lw, aw;
};

__device__ __forceinline__ GroupSample prepare_sample(const int4 *lv, const float *__restrict__ loc_u,
                                                      const float *__restrict__ aw_u, int s, int LP, int P, int MC)
{
  GroupSample g;
  g.mask = 0; g.off = 0; g.ld = g.lh = g.lw = g.aw = 0.f;
  if (s < LP) {
    const int4 li = lv[s / P];
    const float x = ldg_stream(loc_u + 3 * s), y = ldg_stream(loc_u + 3 * s + 1), z = ldg_stream(loc_u + 3 * s + 2);
    g.aw = ldg_stream(aw_u + s);
    const Sample<float> sm = locate<float>(x, y, z, li.x, li.y, li.z);
    g.mask = sm.mask; g.ld = sm.ld; g.lh = sm.lh; g.lw = sm.lw;
    g.off = (li.w + (sm.d_low * li.y + sm.h_low) * li.z + sm.w_low) * MC;
  }
  return g;
}

template <typename VT, int G, int NV>
__global__ void __launch_bounds__(kThreads)
fwd_vec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
               const float *__restrict__ loc, const float *__restrict__ aw, int N, int S, int M, int L, int Lq, int P,
               VT *__restrict__ out)
{
  using V = Vec16<VT>;
  constexpr int VEC = V::N, CPL = VEC * NV, C = G * CPL;
  __shared__ int4 lv[kMaxLevels];
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();

  const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  const long long stride = (long long)gridDim.x * (kThreads / 32) * (32 / G);
  for (long long u = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * (32 / G) + grp; u < total; u += stride) {
    const int m = (int)(u % M);
    const long long b = u / ((long long)M * Lq);
    const VT *vbase = value + b * (long long)S * MC + m * C + gl * CPL;
    const float *loc_u = loc + u * LP * 3;
    const float *aw_u = aw + u * LP;
    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;

    int l_cur = 0, p_cur = 0;
    int4 li = lv[0];
    for (int s0 = 0; s0 < LP; s0 += G) {
      const GroupSample mine = prepare_sample(lv, loc_u, aw_u, s0 + gl, LP, P, MC);
      const int cnt = min(G, LP - s0);
      for (int j = 0; j < cnt; ++j) {
        const int sH = li.z * MC, sD = li.y * sH;
        if (++p_cur == P) { p_cur = 0; ++l_cur; li = lv[min(l_cur, L - 1)]; }
        const unsigned mask = __shfl_sync(gmask, mine.mask, j, G);
        if (mask == 0) continue;                                  // cuh:428 -- uniform inside the group
        const int off = __shfl_sync(gmask, mine.off, j, G);
        const float ld = __shfl_sync(gmask, mine.ld, j, G);
        const float lh = __shfl_sync(gmask, mine.lh, j, G);
        const float lw = __shfl_sync(gmask, mine.lw, j, G);
        const float wa = __shfl_sync(gmask, mine.aw, j, G);
        float w[8];
        corner_weights<float>(ld, lh, lw, w);
        const VT *p0 = vbase + off;
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
          float v[8][VEC];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (mask & (1u << k)) {
              V::load(p0 + ((k & 4) ? sD : 0) + ((k & 2) ? sH : 0) + ((k & 1) ? MC : 0) + nv * VEC, v[k]);
            } else {
#pragma unroll
              for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
            }
          }
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            const float val = blend<float>(w, v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
            acc[nv * VEC + c] = __fmaf_rn(wa, val, acc[nv * VEC + c]);   // cuh:430
          }
        }
      }
    }
    VT *o = out + u * C + gl * CPL;
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
      float t[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) t[c] = acc[nv * VEC + c];
      V::store(o + nv * VEC, t);
    }
  }
}

template <int G> __device__ __forceinline__ float group_sum(unsigned gmask, float v)
{
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(gmask, v, d, G);
  return v;
}

template <typename VT, int G, int NV>
__global__ void __launch_bounds__(kThreads)
bwd_vec_kernel(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ starts, const float *__restrict__ loc, const float *__restrict__ aw, int N, int S,
               int M, int L, int Lq, int P, float *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_aw)
{
  using V = Vec16<VT>;
  constexpr int VEC = V::N, CPL = VEC * NV, C = G * CPL;
  __shared__ int4 lv[kMaxLevels];
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();

  const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  const long long stride = (long long)gridDim.x * (kThreads / 32) * (32 / G);
  for (long long u = ((long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * (32 / G) + grp; u < total; u += stride) {
    const int m = (int)(u % M);
    const long long b = u / ((long long)M * Lq);
    const long long slab = b * (long long)S * MC + m * C + gl * CPL;
    const VT *vbase = value + slab;
    float *gbase = grad_value + slab;
    const float *loc_u = loc + u * LP * 3;
    const float *aw_u = aw + u * LP;
    float top[CPL];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
      float t[VEC];
      V::load_stream(grad_out + u * C + gl * CPL + nv * VEC, t);
#pragma unroll
      for (int c = 0; c < VEC; ++c) top[nv * VEC + c] = t[c];
    }

    int l_cur = 0, p_cur = 0;
    int4 li = lv[0];
    for (int s0 = 0; s0 < LP; s0 += G) {
      const GroupSample mine = prepare_sample(lv, loc_u, aw_u, s0 + gl, LP, P, MC);
      float r_a = 0.f, r_w = 0.f, r_h = 0.f, r_d = 0.f;          // results of "my" sample (lane j keeps sample s0+j)
      const int cnt = min(G, LP - s0);
      for (int j = 0; j < cnt; ++j) {
        const int sH = li.z * MC, sD = li.y * sH;
        const float fD = __int2float_rn(li.x), fH = __int2float_rn(li.y), fW = __int2float_rn(li.z);
        if (++p_cur == P) { p_cur = 0; ++l_cur; li = lv[min(l_cur, L - 1)]; }
        const unsigned mask = __shfl_sync(gmask, mine.mask, j, G);
        if (mask == 0) continue;                                  // all four gradients of this sample stay 0 (cuh:618-621)
        const int off = __shfl_sync(gmask, mine.off, j, G);
        const float ld = __shfl_sync(gmask, mine.ld, j, G);
        const float lh = __shfl_sync(gmask, mine.lh, j, G);
        const float lw = __shfl_sync(gmask, mine.lw, j, G);
        const float wa = __shfl_sync(gmask, mine.aw, j, G);
        float w[8];
        corner_weights<float>(ld, lh, lw, w);
        const float hd = 1.f - ld, hh = 1.f - lh, hw = 1.f - lw;
        // d(weight)/d(frac) factors, cuh:159-231
        const float d0 = hh * hw, d1 = hh * lw, d2 = lh * hw, d3 = lh * lw;
        const float h0 = hd * hw, h1 = hd * lw, h2 = ld * hw, h3 = ld * lw;
        const float w0 = hd * hh, w1 = hd * lh, w2 = ld * hh, w3 = ld * lh;
        float pa = 0.f, pw = 0.f, ph = 0.f, pd = 0.f;
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
          float v[8][VEC];
          int koff[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            koff[k] = off + ((k & 4) ? sD : 0) + ((k & 2) ? sH : 0) + ((k & 1) ? MC : 0) + nv * VEC;
            if (mask & (1u << k)) {
              V::load(vbase + koff[k], v[k]);
            } else {
#pragma unroll
              for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
            }
          }
          float tgv[VEC];
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            const float t = top[nv * VEC + c];
            tgv[c] = t * wa;                                                           // cuh:151
            const float val = blend<float>(w, v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
            float gd = -d0 * v[0][c], gh = -h0 * v[0][c], gw = -w0 * v[0][c];          // cuh:159-231
            gd -= d1 * v[1][c]; gh -= h1 * v[1][c]; gw += w0 * v[1][c];
            gd -= d2 * v[2][c]; gh += h0 * v[2][c]; gw -= w1 * v[2][c];
            gd -= d3 * v[3][c]; gh += h1 * v[3][c]; gw += w1 * v[3][c];
            gd += d0 * v[4][c]; gh -= h2 * v[4][c]; gw -= w2 * v[4][c];
            gd += d1 * v[5][c]; gh -= h3 * v[5][c]; gw += w2 * v[5][c];
            gd += d2 * v[6][c]; gh += h2 * v[6][c]; gw -= w3 * v[6][c];
            gd += d3 * v[7][c]; gh += h3 * v[7][c]; gw += w3 * v[7][c];
            pa = fmaf(t, val, pa);                                                     // cuh:237
            pw = fmaf(fW * gw, tgv[c], pw);                                            // cuh:238-240
            ph = fmaf(fH * gh, tgv[c], ph);
            pd = fmaf(fD * gd, tgv[c], pd);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (mask & (1u << k)) {
#pragma unroll
              for (int c4 = 0; c4 < VEC; c4 += 4)
                red_add_v4(gbase + koff[k] + c4, w[k] * tgv[c4], w[k] * tgv[c4 + 1], w[k] * tgv[c4 + 2], w[k] * tgv[c4 + 3]);
            }
          }
        }
        pa = group_sum<G>(gmask, pa);
        pw = group_sum<G>(gmask, pw);
        ph = group_sum<G>(gmask, ph);
        pd = group_sum<G>(gmask, pd);
        if (gl == j) { r_a = pa; r_w = pw; r_h = ph; r_d = pd; }
      }
      const int s = s0 + gl;
      if (s < LP) {
        float *gl_ = grad_loc + (u * LP + s) * 3;
        gl_[0] = r_w; gl_[1] = r_h; gl_[2] = r_d;
        grad_aw[u * LP + s] = r_a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Generic kernels: any channel count, any level count, fp32 / fp64 / 16-bit storage.  One warp per unit, lanes
// stride over channels.  Used when no vector instantiation fits (odd C, fp64 gradcheck shapes, >16 levels, ...).
// ---------------------------------------------------------------------------------------------------------------
template <typename VT, typename CT>
__global__ void __launch_bounds__(kThreads)
fwd_generic_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                   const CT *__restrict__ loc, const CT *__restrict__ aw, int N, int S, int M, int C, int L, int Lq, int P,
                   VT *__restrict__ out)
{
  const int lane = threadIdx.x & 31;
  const long long total = (long long)N * Lq * M;
  const long long nwarps = (long long)gridDim.x * (kThreads / 32);
  const long long MC = (long long)M * C;
  for (long long u = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); u < total; u += nwarps) {
    const int m = (int)(u % M);
    const long long b = u / ((long long)M * Lq);
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      CT col = 0;
      for (int l = 0; l < L; ++l) {
        const int D = (int)shapes[3 * l], H = (int)shapes[3 * l + 1], W = (int)shapes[3 * l + 2];
        const VT *base = value + (b * S + starts[l]) * MC + (long long)m * C;
        for (int p = 0; p < P; ++p) {
          const long long si = (u * L + l) * P + p;
          const Sample<CT> s = locate<CT>(loc[si * 3], loc[si * 3 + 1], loc[si * 3 + 2], D, H, W);
          if (s.mask == 0 || c >= C) continue;
          CT w[8], v[8];
          corner_weights<CT>(s.ld, s.lh, s.lw, w);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const long long vox = ((long long)(s.d_low + ((k >> 2) & 1)) * H + (s.h_low + ((k >> 1) & 1))) * W + (s.w_low + (k & 1));
            v[k] = (s.mask & (1u << k)) ? (CT)Cvt<VT>::up(base[vox * MC + c]) : (CT)0;
          }
          col = Arith<CT>::fma(aw[si], blend<CT>(w, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]), col);
        }
      }
      if (c < C) out[u * C + c] = Cvt<VT>::down(col);
    }
  }
}

template <typename VT, typename CT>
__global__ void __launch_bounds__(kThreads)
bwd_generic_kernel(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                   const int64_t *__restrict__ starts, const CT *__restrict__ loc, const CT *__restrict__ aw, int N, int S,
                   int M, int C, int L, int Lq, int P, CT *__restrict__ grad_value, CT *__restrict__ grad_loc,
                   CT *__restrict__ grad_aw)
{
  const int lane = threadIdx.x & 31;
  const long long total = (long long)N * Lq * M;
  const long long nwarps = (long long)gridDim.x * (kThreads / 32);
  const long long MC = (long long)M * C;
  for (long long u = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); u < total; u += nwarps) {
    const int m = (int)(u % M);
    const long long b = u / ((long long)M * Lq);
    for (int l = 0; l < L; ++l) {
      const int D = (int)shapes[3 * l], H = (int)shapes[3 * l + 1], W = (int)shapes[3 * l + 2];
      const long long lvl = (b * S + starts[l]) * MC + (long long)m * C;
      for (int p = 0; p < P; ++p) {
        const long long si = (u * L + l) * P + p;
        const Sample<CT> s = locate<CT>(loc[si * 3], loc[si * 3 + 1], loc[si * 3 + 2], D, H, W);
        CT pa = 0, pw = 0, ph = 0, pd = 0;
        if (s.mask != 0) {
          const CT wa = aw[si];
          CT w[8];
          corner_weights<CT>(s.ld, s.lh, s.lw, w);
          const CT ld = s.ld, lh = s.lh, lw = s.lw, hd = 1 - ld, hh = 1 - lh, hw = 1 - lw;
          const CT dd[8] = {-(hh * hw), -(hh * lw), -(lh * hw), -(lh * lw), hh * hw, hh * lw, lh * hw, lh * lw};
          const CT dh[8] = {-(hd * hw), -(hd * lw), hd * hw, hd * lw, -(ld * hw), -(ld * lw), ld * hw, ld * lw};
          const CT dw[8] = {-(hd * hh), hd * hh, -(hd * lh), hd * lh, -(ld * hh), ld * hh, -(ld * lh), ld * lh};
          for (int c = lane; c < C; c += 32) {
            const CT t = (CT)Cvt<VT>::up(grad_out[u * C + c]);
            const CT tgv = t * wa;
            CT v[8], gd = 0, gh = 0, gw = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              v[k] = 0;
              if (s.mask & (1u << k)) {
                const long long vox = ((long long)(s.d_low + ((k >> 2) & 1)) * H + (s.h_low + ((k >> 1) & 1))) * W + (s.w_low + (k & 1));
                const long long a = lvl + vox * MC + c;
                v[k] = (CT)Cvt<VT>::up(value[a]);
                gd += dd[k] * v[k]; gh += dh[k] * v[k]; gw += dw[k] * v[k];
                atomicAdd(grad_value + a, w[k] * tgv);
              }
            }
            pa += t * blend<CT>(w, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            pw += (CT)W * gw * tgv; ph += (CT)H * gh * tgv; pd += (CT)D * gd * tgv;
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            pa += __shfl_xor_sync(0xffffffffu, pa, d);
            pw += __shfl_xor_sync(0xffffffffu, pw, d);
            ph += __shfl_xor_sync(0xffffffffu, ph, d);
            pd += __shfl_xor_sync(0xffffffffu, pd, d);
          }
        }
        if (lane == 0) {
          grad_loc[si * 3] = pw; grad_loc[si * 3 + 1] = ph; grad_loc[si * 3 + 2] = pd;
          grad_aw[si] = pa;
        }
      }
    }
  }
}

// Test hook: index arithmetic of one sample per thread.
template <typename CT>
__global__ void indices_kernel(const int64_t *__restrict__ shapes, const CT *__restrict__ loc, long long T, int L, int P,
                               int32_t *__restrict__ idx, CT *__restrict__ frac)
{
  for (long long si = (long long)blockIdx.x * blockDim.x + threadIdx.x; si < T; si += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((si / P) % L);
    const Sample<CT> s = locate<CT>(loc[si * 3], loc[si * 3 + 1], loc[si * 3 + 2], (int)shapes[3 * l], (int)shapes[3 * l + 1],
                                    (int)shapes[3 * l + 2]);
    idx[si * 4] = s.mask != 0; idx[si * 4 + 1] = s.d_low; idx[si * 4 + 2] = s.h_low; idx[si * 4 + 3] = s.w_low;
    frac[si * 3] = s.ld; frac[si * 3 + 1] = s.lh; frac[si * 3 + 2] = s.lw;
  }
}

}  // namespace msda3d
