// win_attn_kernels.cuh -- Swin3D window attention (forward + gradient) for sm_100a.
//
// Replaces the attention core of the reference's WindowAttention3D.forward (transoar/models/backbones/encoder_blocks.py:259-285):
//     attn = (q * scale) @ k^T  + relative_position_bias[h]  (+ shift mask[w])  -> softmax -> @ v
// which the reference runs as two batched GEMMs, two elementwise adds over a materialised [B*nW, heads, n, n] score tensor and a
// softmax (n = 125 tokens per 5x5x5 window, head dim 16 for every stage: 48/3, 96/6, 192/12, 384/24).  At 160x160x256 stage 2 has
// 6656 windows x 3 heads per sample: the score tensor alone is 1.25 GB there and is written / read five times.
//
// Here one CTA owns one (window, head): K and V of the window (n x 16 floats each) sit in shared memory, thread i owns query row i
// and never materialises more than one score at a time (two passes over the 125 keys: row maximum, then exp / sum / P V).  The
// kernel reads q, k, v straight out of the qkv Linear's output [B*nW, n, 3, heads, 16] and writes [B*nW, n, heads*16] -- the layout
// the output projection consumes -- so the permute / reshape copies of the reference disappear as well.
//
// Backward, per (window, head), two phases that each recompute the scores from q, k, bias (nothing n x n is ever stored):
//   row phase    (thread i = query row):  dS_ij = P_ij (dO_i . V_j - D_i);  dq_i = scale * sum_j dS_ij k_j;  dbias[i][j] += dS_ij
//   column phase (thread j = key row):    dk_j = sum_i dS_ij (scale q_i);   dv_j = sum_i P_ij dO_i
// The bias gradient is accumulated over all windows a CTA processes in a shared-memory tile (row i is only touched by thread i:
// no atomics; the row stride n is padded to an odd number so the accesses are bank-conflict free) and flushed once per CTA with
// atomicAdd -- a few hundred flushes instead of one atomic per (window, i, j).
//
// bias is passed TRANSPOSED for the row phases (biasT[h][j][i]: for a fixed key j the 125 threads read consecutive floats) and in its
// natural layout for the column phase; the shift mask is symmetric (mask[w][i][j] = -100 iff tokens i and j carry different region
// labels, encoder_blocks.py:387-400), so one copy serves both.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace winattn {

constexpr int kThreads = 128;     // one thread per token of the window (n <= 128)
constexpr int HD = 16;            // head dim of every Swin stage of the reference

__device__ __forceinline__ float dot16(const float (&a)[HD], const float *__restrict__ b)
{
  const float4 b0 = *reinterpret_cast<const float4 *>(b), b1 = *reinterpret_cast<const float4 *>(b + 4);
  const float4 b2 = *reinterpret_cast<const float4 *>(b + 8), b3 = *reinterpret_cast<const float4 *>(b + 12);
  float s = a[0] * b0.x;
  s = fmaf(a[1], b0.y, s); s = fmaf(a[2], b0.z, s); s = fmaf(a[3], b0.w, s);
  s = fmaf(a[4], b1.x, s); s = fmaf(a[5], b1.y, s); s = fmaf(a[6], b1.z, s); s = fmaf(a[7], b1.w, s);
  s = fmaf(a[8], b2.x, s); s = fmaf(a[9], b2.y, s); s = fmaf(a[10], b2.z, s); s = fmaf(a[11], b2.w, s);
  s = fmaf(a[12], b3.x, s); s = fmaf(a[13], b3.y, s); s = fmaf(a[14], b3.z, s); s = fmaf(a[15], b3.w, s);
  return s;
}

__device__ __forceinline__ void load16(const float *__restrict__ p, float (&v)[HD])
{
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p + c));
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
  }
}

__device__ __forceinline__ void axpy16(float a, const float *__restrict__ x, float (&y)[HD])
{
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 t = *reinterpret_cast<const float4 *>(x + c);
    y[c] = fmaf(a, t.x, y[c]); y[c + 1] = fmaf(a, t.y, y[c + 1]); y[c + 2] = fmaf(a, t.z, y[c + 2]); y[c + 3] = fmaf(a, t.w, y[c + 3]);
  }
}

// qkv [Bw][n][3][H][HD]; biasT [H][n(j)][n(i)]; mask [nW][n][n] or NULL; out [Bw][n][H*HD]; lse [Bw][H][n]
__global__ void __launch_bounds__(kThreads)
fwd_kernel(const float *__restrict__ qkv, const float *__restrict__ biasT, const float *__restrict__ mask, int n, int H, int nW, float scale,
           float *__restrict__ out, float *__restrict__ lse)
{
  __shared__ __align__(16) float sK[kThreads][HD];
  __shared__ __align__(16) float sV[kThreads][HD];
  const int bw = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
  const long long tok = 3LL * H * HD;                                 // floats per token in qkv
  const float *base = qkv + (long long)bw * n * tok + h * HD;
  float q[HD];
  if (i < n) {
    load16(base + i * tok, q);
    float kv[HD];
    load16(base + i * tok + (long long)H * HD, kv);
#pragma unroll
    for (int c = 0; c < HD; ++c) { sK[i][c] = kv[c]; q[c] *= scale; }
    load16(base + i * tok + 2LL * H * HD, kv);
#pragma unroll
    for (int c = 0; c < HD; ++c) sV[i][c] = kv[c];
  }
  __syncthreads();
  if (i >= n) return;
  const float *bT = biasT + (long long)h * n * n + i;                 // + j * n
  const float *mk = mask != nullptr ? mask + ((long long)(bw % nW) * n) * n + i : nullptr;   // symmetric: mask[w][j][i] == mask[w][i][j]
  // both passes fetch the bias / mask values of four keys before using them: the loops are bound by the latency of those loads otherwise
  float m = -CUDART_INF_F;
  for (int j0 = 0; j0 < n; j0 += 4) {
    float add[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = min(j0 + u, n - 1);
      add[u] = __ldg(bT + (long long)j * n) + (mk != nullptr ? __ldg(mk + (long long)j * n) : 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + u < n) m = fmaxf(m, dot16(q, sK[j0 + u]) + add[u]);
  }
  float l = 0.f, o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = 0.f;
  for (int j0 = 0; j0 < n; j0 += 4) {
    float add[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = min(j0 + u, n - 1);
      add[u] = __ldg(bT + (long long)j * n) + (mk != nullptr ? __ldg(mk + (long long)j * n) : 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + u < n) {
        const float p = __expf(dot16(q, sK[j0 + u]) + add[u] - m);
        l += p;
        axpy16(p, sV[j0 + u], o);
      }
    }
  }
  const float inv = 1.f / l;
  float *dst = out + ((long long)bw * n + i) * H * HD + h * HD;
#pragma unroll
  for (int c = 0; c < HD; c += 4)
    *reinterpret_cast<float4 *>(dst + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
  lse[((long long)bw * H + h) * n + i] = m + __logf(l);
}

// Backward.  grid = (chunks, H): CTA (c, h) walks windows c, c + chunks, ... of head h.  256 threads: threads 0..127 run the row phase
// (thread = query row) while threads 128..255 run the column phase (thread = key row) of the same window at the same time -- both only
// read the window's tiles -- so an SM holds twice the warps for the same shared memory.  The bias / mask values of four keys are loaded
// before they are used (the loops are latency-bound on those loads otherwise).  dynamic smem: tiles + dbias tile [n][ld] (ld odd).
// dqkv [Bw][n][3][H][HD]; dbias [H][n][n] (natural layout, accumulated with atomicAdd: zero it first).
constexpr int kBwdThreads = 2 * kThreads;

__global__ void __launch_bounds__(kBwdThreads, 2)
bwd_kernel(const float *__restrict__ qkv, const float *__restrict__ bias, const float *__restrict__ biasT, const float *__restrict__ mask,
           const float *__restrict__ out, const float *__restrict__ dout, const float *__restrict__ lse, int Bw, int n, int H, int nW, float scale,
           float *__restrict__ dqkv, float *__restrict__ dbias)
{
  extern __shared__ __align__(16) float smem[];
  float (*sQ)[HD] = reinterpret_cast<float (*)[HD]>(smem);                          // scale * q
  float (*sK)[HD] = reinterpret_cast<float (*)[HD]>(smem + kThreads * HD);
  float (*sV)[HD] = reinterpret_cast<float (*)[HD]>(smem + 2 * kThreads * HD);
  float (*sdO)[HD] = reinterpret_cast<float (*)[HD]>(smem + 3 * kThreads * HD);
  float *sL = smem + 4 * kThreads * HD;                                              // lse per row
  float *sD = sL + kThreads;                                                         // D_i = dO_i . O_i
  float *sB = sD + kThreads;                                                         // dbias tile [n][ld]
  const int ld = n | 1;
  const int h = blockIdx.y, role = threadIdx.x / kThreads, t = threadIdx.x % kThreads;
  const long long tok = 3LL * H * HD;
  for (int e = threadIdx.x; e < n * ld; e += kBwdThreads) sB[e] = 0.f;

  for (int bw = blockIdx.x; bw < Bw; bw += gridDim.x) {
    __syncthreads();                                                                 // previous window's phases are done with the tiles
    const float *base = qkv + (long long)bw * n * tok + h * HD;
    if (t < n) {
      float v[HD];
      if (role == 0) {
        load16(base + t * tok, v);
#pragma unroll
        for (int c = 0; c < HD; ++c) sQ[t][c] = v[c] * scale;
        load16(base + t * tok + (long long)H * HD, v);
#pragma unroll
        for (int c = 0; c < HD; ++c) sK[t][c] = v[c];
      } else {
        float dO[HD];
        load16(base + t * tok + 2LL * H * HD, v);
#pragma unroll
        for (int c = 0; c < HD; ++c) sV[t][c] = v[c];
        const long long orow = ((long long)bw * n + t) * H * HD + h * HD;
        load16(dout + orow, dO);
        load16(out + orow, v);
        float Di = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) { sdO[t][c] = dO[c]; Di = fmaf(dO[c], v[c], Di); }
        sL[t] = lse[((long long)bw * H + h) * n + t];
        sD[t] = Di;
      }
    }
    __syncthreads();
    if (t >= n) continue;                                                            // (no barrier below this point inside the iteration)
    const float *mk = mask != nullptr ? mask + ((long long)(bw % nW) * n) * n + t : nullptr;   // mask[w][x][t], x = the loop index
    if (role == 0) {
      // ---- row phase: thread t = query row i
      float qs[HD], dO[HD], dq[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) { qs[c] = sQ[t][c]; dO[c] = sdO[t][c]; dq[c] = 0.f; }
      const float li = sL[t], Di = sD[t];
      const float *bT = biasT + (long long)h * n * n + t;
      for (int j0 = 0; j0 < n; j0 += 4) {
        float add[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, n - 1);
          add[u] = __ldg(bT + (long long)j * n) + (mk != nullptr ? __ldg(mk + (long long)j * n) : 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < n) {
            const float p = __expf(dot16(qs, sK[j]) + add[u] - li);
            const float ds = p * (dot16(dO, sV[j]) - Di);
            axpy16(ds, sK[j], dq);
            sB[t * ld + j] += ds;
          }
        }
      }
      float *dst = dqkv + ((long long)bw * n + t) * tok + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4)
        *reinterpret_cast<float4 *>(dst + c) = make_float4(dq[c] * scale, dq[c + 1] * scale, dq[c + 2] * scale, dq[c + 3] * scale);
    } else {
      // ---- column phase: thread t = key row j
      float kj[HD], vj[HD], dk[HD], dv[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) { kj[c] = sK[t][c]; vj[c] = sV[t][c]; dk[c] = 0.f; dv[c] = 0.f; }
      const float *bN = bias + (long long)h * n * n + t;                                // bias[h][i][j = t]
      for (int i0 = 0; i0 < n; i0 += 4) {
        float add[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = min(i0 + u, n - 1);
          add[u] = __ldg(bN + (long long)i * n) + (mk != nullptr ? __ldg(mk + (long long)i * n) : 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if (i < n) {
            const float p = __expf(dot16(kj, sQ[i]) + add[u] - sL[i]);
            const float ds = p * (dot16(vj, sdO[i]) - sD[i]);
            axpy16(ds, sQ[i], dk);
            axpy16(p, sdO[i], dv);
          }
        }
      }
      float *dkp = dqkv + ((long long)bw * n + t) * tok + (long long)H * HD + h * HD;
      float *dvp = dkp + (long long)H * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        *reinterpret_cast<float4 *>(dkp + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
        *reinterpret_cast<float4 *>(dvp + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
      }
    }
  }
  __syncthreads();
  float *db = dbias + (long long)h * n * n;
  for (int e = threadIdx.x; e < n * n; e += kBwdThreads) atomicAdd(db + e, sB[(e / n) * ld + e % n]);
}

}  // namespace winattn
