// roi_attn_kernels.cuh -- fused RoI-restricted cross-attention (forward + gradient) for sm_100a.
//
// Replaces the attention core of the reference's FocusedAttn.forward, transoar/models/necks/focused_decoder.py:238-254:
//     attn = q @ k^T ; attn += mask(-inf outside the query's RoI box) ; softmax(-1) ; x = attn @ v
// The reference materialises attn as a dense fp32 [B, heads, Nq, Nkv] tensor (1.77 GB per sample and layer at the
// VISCERAL shape) of which ~94 % is -inf.  generate_attn_masks (focused_decoder.py:138-159) only ever produces
// axis-aligned boxes, shared by `num_queries_per_organ` consecutive queries, so a CTA owns one (batch, query group,
// head), walks ONLY the key/value voxels inside the group's box in chunks of TK tokens and keeps the running softmax
// statistics in registers (flash-attention style): no score tensor, no mask tensor.
//
// Shapes: q [B, Nq, H, HD] (already scaled), k / v [B, Nkv, H, HD], Nkv = X*Y*Z tokens in row-major (x, y, z) order
// (= src.flatten(2) of a [B, C, X, Y, Z] map), groups int32 [Gn, 8] = {q0, nq (<= 32), x1, y1, z1, x2, y2, z2},
// out [B, Nq, H*HD], lse [B, H, Nq] (log-sum-exp of the in-box scores, kept for the backward).
// An empty box gives NaN rows, exactly as softmax over an all -inf row does in the reference.
//
// fp32 on the CUDA cores: inside the boxes the whole layer is ~10 GFLOP; the tensor-core work of this block is the
// K/V projection over all tokens, which stays a cuBLAS GEMM in the module.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace roiattn {

constexpr int TQ = 32;        // query rows per CTA (one RoI group, padded)
constexpr int TK = 64;        // key/value tokens per chunk
constexpr int kThreads = 128;

struct Group {
  int q0, nq, x1, y1, z1, x2, y2, z2;
};

// A group's box clamped to the key/value grid (X = Nkv / (Y*Z)): a box built for another feature-map shape can then never address a
// token outside [0, Nkv) -- neither the forward's reads nor the backward's dk / dv reductions.  (The python module raises on such a
// mismatch like the reference's `attn += mask` does; this is the memory-safety net underneath the C ABI.)
__device__ __forceinline__ Group load_group(const int *__restrict__ groups, int gi, int X, int Y, int Z)
{
  const int *gp = groups + gi * 8;
  Group g;
  g.q0 = gp[0]; g.nq = gp[1];
  g.x1 = max(gp[2], 0); g.y1 = max(gp[3], 0); g.z1 = max(gp[4], 0);
  g.x2 = min(gp[5], X); g.y2 = min(gp[6], Y); g.z2 = min(gp[7], Z);
  return g;
}

// token id of the j-th voxel of the box (z fastest, then y, then x), or -1 past the end
__device__ __forceinline__ int box_token(const Group &g, int j, int Y, int Z)
{
  const int bz = g.z2 - g.z1, by = g.y2 - g.y1, bx = g.x2 - g.x1;
  if (bz <= 0 || by <= 0 || bx <= 0 || j >= bx * by * bz) return -1;
  const int iz = j % bz, r = j / bz, iy = r % by, ix = r / by;
  return ((g.x1 + ix) * Y + (g.y1 + iy)) * Z + (g.z1 + iz);
}

// Forward.  grid = (Gn * S, H, B): S CTAs split the tokens of a box (flash-decoding style) so that 20 organs x 8 heads
// still fill 148 SMs; S > 1 writes partial softmax states that combine_kernel merges.  Thread t owns score rows 4*(t/16)..+3 and columns (t%16) + 16*j (j < 4) of the TQ x TK tile
// (consecutive lanes -> consecutive K rows: with a row stride of HD+4 floats the float4 reads are bank-conflict free), and
// output rows 4*(t/16)..+3, columns (t%16) + 16*i (i < HD/16).
template <int HD>
__global__ void __launch_bounds__(kThreads)
fwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, const int *__restrict__ groups,
           int Nq, int Nkv, int H, int Y, int Z, float *__restrict__ out, float *__restrict__ lse, int S, float *__restrict__ part)
{
  static_assert(HD % 16 == 0 && HD <= 128, "head dim");
  constexpr int OC = HD / 16;
  extern __shared__ __align__(16) float smem_f[];
  float (*sQ)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f);
  float (*sK)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f + TQ * (HD + 4));
  float (*sV)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f + (TQ + TK) * (HD + 4));
  float (*sP)[TK + 4] = reinterpret_cast<float (*)[TK + 4]>(smem_f + (TQ + 2 * TK) * (HD + 4));
  __shared__ int sTok[TK];

  const int gi = blockIdx.x / S, split = blockIdx.x % S, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const Group g = load_group(groups, gi, Nkv / (Y * Z), Y, Z);
  const int ntok = max(0, g.x2 - g.x1) * max(0, g.y2 - g.y1) * max(0, g.z2 - g.z1);
  const long long HHD = (long long)H * HD;
  // this CTA's share of the box: chunks [c_beg, c_end) of TK tokens (S > 1: partial softmax state goes to `part`)
  const int nchunk = (ntok + TK - 1) / TK, cps = (nchunk + S - 1) / S;
  const int tok_beg = min(split * cps, nchunk) * TK, tok_end = min(min((split + 1) * cps, nchunk) * TK, ntok);

  for (int i = t; i < TQ * (HD / 4); i += kThreads) {
    const int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < g.nq) val = *reinterpret_cast<const float4 *>(q + ((long long)b * Nq + g.q0 + r) * HHD + h * HD + c4);
    *reinterpret_cast<float4 *>(&sQ[r][c4]) = val;
  }

  const int rg = t / 16, cg = t % 16;           // row group (4 rows), column group
  float m_run[4], l_run[4], o[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F; l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < OC; ++c) o[i][c] = 0.f;
  }

  for (int base = tok_beg; base < tok_end; base += TK) {
    __syncthreads();
    if (t < TK) sTok[t] = box_token(g, base + t, Y, Z);
    __syncthreads();
    for (int i = t; i < TK * (HD / 4); i += kThreads) {
      const int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      const int tok = sTok[r];
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (tok >= 0) {
        const long long off = ((long long)b * Nkv + tok) * HHD + h * HD + c4;
        kv = __ldg(reinterpret_cast<const float4 *>(k + off));
        vv = __ldg(reinterpret_cast<const float4 *>(v + off));
      }
      *reinterpret_cast<float4 *>(&sK[r][c4]) = kv;
      *reinterpret_cast<float4 *>(&sV[r][c4]) = vv;
    }
    __syncthreads();

    // S = Q K^T for my 4x4 tile
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < HD; d += 4) {
      float4 qa[4], kb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4 *>(&sQ[rg * 4 + i][d]);
#pragma unroll
      for (int j = 0; j < 4; ++j) kb[j] = *reinterpret_cast<const float4 *>(&sK[cg + 16 * j][d]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          s[i][j] += qa[i].x * kb[j].x + qa[i].y * kb[j].y + qa[i].z * kb[j].z + qa[i].w * kb[j].w;
    }
    // mask the tail of the last chunk, running max / sum per row (16 lanes share a row group)
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (sTok[cg + 16 * j] < 0) s[i][j] = -CUDART_INF_F;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int dlt = 8; dlt > 0; dlt >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, dlt));
      const float m_new = fmaxf(m_run[i], mx);
      alpha[i] = (m_run[i] == -CUDART_INF_F) ? 0.f : __expf(m_run[i] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (s[i][j] == -CUDART_INF_F) ? 0.f : __expf(s[i][j] - m_new);
        s[i][j] = p;
        sum += p;
      }
#pragma unroll
      for (int dlt = 8; dlt > 0; dlt >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dlt);
      l_run[i] = l_run[i] * alpha[i] + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) sP[rg * 4 + i][cg + 16 * j] = s[i][j];
    }
    __syncthreads();
    // O = O * alpha + P V
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < OC; ++c) o[i][c] *= alpha[i];
#pragma unroll 2
    for (int kk = 0; kk < TK; kk += 4) {                       // four tokens per step: P rows come as float4
      float4 p4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p4[i] = *reinterpret_cast<const float4 *>(&sP[rg * 4 + i][kk]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float vv[OC];
#pragma unroll
        for (int c = 0; c < OC; ++c) vv[c] = sV[kk + u][cg + 16 * c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = u == 0 ? p4[i].x : (u == 1 ? p4[i].y : (u == 2 ? p4[i].z : p4[i].w));
#pragma unroll
          for (int c = 0; c < OC; ++c) o[i][c] = fmaf(p, vv[c], o[i][c]);
        }
      }
    }
  }

  if (S > 1) {
    // partial state of rows q0..q0+nq-1: part[b][h][q][split][HD + 2] = {o (unnormalised), m, l}
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg * 4 + i;
      if (r < g.nq) {
        float *dst = part + ((((long long)b * H + h) * Nq + g.q0 + r) * S + split) * (HD + 2);
#pragma unroll
        for (int c = 0; c < OC; ++c) dst[cg + 16 * c] = o[i][c];
        if (cg == 0) { dst[HD] = m_run[i]; dst[HD + 1] = l_run[i]; }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rg * 4 + i;
    if (r < g.nq) {
      const float inv = 1.f / l_run[i];          // empty box: 0 * inf = NaN, like softmax of an all -inf row
      float *dst = out + ((long long)b * Nq + g.q0 + r) * HHD + h * HD;
#pragma unroll
      for (int c = 0; c < OC; ++c) dst[cg + 16 * c] = o[i][c] * inv;
      if (cg == 0) lse[((long long)b * H + h) * Nq + g.q0 + r] = m_run[i] + __logf(l_run[i]);
    }
  }
}

// Merge the S partial softmax states of every (b, h, q) row: one warp per row.
template <int HD>
__global__ void __launch_bounds__(128)
combine_kernel(const float *__restrict__ part, int rows, int Nq, int H, int S, float *__restrict__ out, float *__restrict__ lse)
{
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;     // row = (b*H + h)*Nq + q
  if (row >= rows) return;
  const float *p = part + (long long)row * S * (HD + 2);
  float M = -CUDART_INF_F;
  for (int s_ = 0; s_ < S; ++s_) M = fmaxf(M, p[s_ * (HD + 2) + HD]);
  float L = 0.f;
  for (int s_ = 0; s_ < S; ++s_) {
    const float m = p[s_ * (HD + 2) + HD];
    L += (m == -CUDART_INF_F) ? 0.f : p[s_ * (HD + 2) + HD + 1] * __expf(m - M);
  }
  const int q_ = row % Nq, bh = row / Nq, h = bh % H, b = bh / H;
  float *dst = out + ((long long)b * Nq + q_) * H * HD + h * HD;
  const float inv = 1.f / L;
  for (int c = lane; c < HD; c += 32) {
    float acc = 0.f;
    for (int s_ = 0; s_ < S; ++s_) {
      const float m = p[s_ * (HD + 2) + HD];
      if (m != -CUDART_INF_F) acc += p[s_ * (HD + 2) + c] * __expf(m - M);
    }
    dst[c] = acc * inv;
  }
  if (lane == 0) lse[row] = M + __logf(L);
}

// Backward.  grid = (Gn * S, H, B).  P = exp(S - lse), dV += P^T dO, dP = dO V^T, dS = P * (dP - D), D = rowsum(dO * O),
// dQ += dS K, dK += dS^T Q.  dk / dv are accumulated with atomics (boxes of different organs overlap); dq is exclusive.
template <int HD>
__global__ void __launch_bounds__(kThreads)
bwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v, const int *__restrict__ groups,
           const float *__restrict__ out, const float *__restrict__ dout, const float *__restrict__ lse, int Nq, int Nkv, int H,
           int Y, int Z, float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv, int S)
{
  constexpr int OC = HD / 16;
  extern __shared__ __align__(16) float smem_f[];
  float (*sQ)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f);
  float (*sdO)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f + TQ * (HD + 4));
  float (*sK)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f + 2 * TQ * (HD + 4));
  float (*sV)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(smem_f + (2 * TQ + TK) * (HD + 4));
  float (*sP)[TK + 4] = reinterpret_cast<float (*)[TK + 4]>(smem_f + (2 * TQ + 2 * TK) * (HD + 4));   // P, then dS
  __shared__ float sLse[TQ], sD[TQ];
  __shared__ int sTok[TK];

  const int gi = blockIdx.x / S, split = blockIdx.x % S, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const Group g = load_group(groups, gi, Nkv / (Y * Z), Y, Z);
  const int ntok = max(0, g.x2 - g.x1) * max(0, g.y2 - g.y1) * max(0, g.z2 - g.z1);
  const long long HHD = (long long)H * HD;
  const int nchunk = (ntok + TK - 1) / TK, cps = (nchunk + S - 1) / S;
  const int tok_beg = min(split * cps, nchunk) * TK, tok_end = min(min((split + 1) * cps, nchunk) * TK, ntok);

  for (int i = t; i < TQ * (HD / 4); i += kThreads) {
    const int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
    float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), da = qa;
    if (r < g.nq) {
      const long long off = ((long long)b * Nq + g.q0 + r) * HHD + h * HD + c4;
      qa = *reinterpret_cast<const float4 *>(q + off);
      da = *reinterpret_cast<const float4 *>(dout + off);
    }
    *reinterpret_cast<float4 *>(&sQ[r][c4]) = qa;
    *reinterpret_cast<float4 *>(&sdO[r][c4]) = da;
  }
  if (t < TQ) {
    float dsum = 0.f, l = 0.f;
    if (t < g.nq) {
      const long long off = ((long long)b * Nq + g.q0 + t) * HHD + h * HD;
      for (int d = 0; d < HD; ++d) dsum += dout[off + d] * out[off + d];
      l = lse[((long long)b * H + h) * Nq + g.q0 + t];
    }
    sD[t] = dsum; sLse[t] = l;
  }

  const int rg = t / 16, cg = t % 16;
  float dqa[4][OC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < OC; ++c) dqa[i][c] = 0.f;

  for (int base = tok_beg; base < tok_end; base += TK) {
    __syncthreads();
    if (t < TK) sTok[t] = box_token(g, base + t, Y, Z);
    __syncthreads();
    for (int i = t; i < TK * (HD / 4); i += kThreads) {
      const int r = i / (HD / 4), c4 = (i % (HD / 4)) * 4;
      const int tok = sTok[r];
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (tok >= 0) {
        const long long off = ((long long)b * Nkv + tok) * HHD + h * HD + c4;
        kv = __ldg(reinterpret_cast<const float4 *>(k + off));
        vv = __ldg(reinterpret_cast<const float4 *>(v + off));
      }
      *reinterpret_cast<float4 *>(&sK[r][c4]) = kv;
      *reinterpret_cast<float4 *>(&sV[r][c4]) = vv;
    }
    __syncthreads();

    // S = Q K^T and dP = dO V^T for my 4x4 tile
    float s[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
#pragma unroll 2
    for (int d = 0; d < HD; d += 4) {
      float4 qa[4], da[4], kb[4], vb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qa[i] = *reinterpret_cast<const float4 *>(&sQ[rg * 4 + i][d]);
        da[i] = *reinterpret_cast<const float4 *>(&sdO[rg * 4 + i][d]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        kb[j] = *reinterpret_cast<const float4 *>(&sK[cg + 16 * j][d]);
        vb[j] = *reinterpret_cast<const float4 *>(&sV[cg + 16 * j][d]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] += qa[i].x * kb[j].x + qa[i].y * kb[j].y + qa[i].z * kb[j].z + qa[i].w * kb[j].w;
          dp[i][j] += da[i].x * vb[j].x + da[i].y * vb[j].y + da[i].z * vb[j].z + da[i].w * vb[j].w;
        }
    }
    // P and dS; P goes to smem first (dV needs it), dS replaces it afterwards
    float ds[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rg * 4 + i;
      float pr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool valid = sTok[cg + 16 * j] >= 0 && r < g.nq;
        pr[j] = valid ? __expf(s[i][j] - sLse[r]) : 0.f;
        ds[i][j] = pr[j] * (dp[i][j] - sD[r]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sP[r][cg + 16 * j] = pr[j];
    }
    __syncthreads();
    // dV[tok][c] += sum_r P[r][tok] * dO[r][c]
    {
      // 64 tokens x HD channels; thread handles tokens tg*4..+3 (tg = t/8 in 0..15) and channels (t%8) + 8*i (i < HD/8)
      const int tg = t / 8, cl = t % 8;
      float acc[4][HD / 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) acc[i][c] = 0.f;
      for (int r = 0; r < TQ; ++r) {
        float dv_[HD / 8];
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) dv_[c] = sdO[r][cl + 8 * c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = sP[r][tg * 4 + i];
#pragma unroll
          for (int c = 0; c < HD / 8; ++c) acc[i][c] = fmaf(p, dv_[c], acc[i][c]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = sTok[tg * 4 + i];
        if (tok >= 0) {
          float *dst = dv + ((long long)b * Nkv + tok) * HHD + h * HD;
#pragma unroll
          for (int c = 0; c < HD / 8; ++c) atomicAdd(dst + cl + 8 * c, acc[i][c]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sP[rg * 4 + i][cg + 16 * j] = ds[i][j];
    __syncthreads();
    // dQ += dS K
#pragma unroll 4
    for (int kk = 0; kk < TK; ++kk) {
      float kv[OC];
#pragma unroll
      for (int c = 0; c < OC; ++c) kv[c] = sK[kk][cg + 16 * c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d_ = sP[rg * 4 + i][kk];
#pragma unroll
        for (int c = 0; c < OC; ++c) dqa[i][c] = fmaf(d_, kv[c], dqa[i][c]);
      }
    }
    // dK[tok][c] += sum_r dS[r][tok] * Q[r][c]
    {
      const int tg = t / 8, cl = t % 8;
      float acc[4][HD / 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) acc[i][c] = 0.f;
      for (int r = 0; r < TQ; ++r) {
        float qv[HD / 8];
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) qv[c] = sQ[r][cl + 8 * c];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d_ = sP[r][tg * 4 + i];
#pragma unroll
          for (int c = 0; c < HD / 8; ++c) acc[i][c] = fmaf(d_, qv[c], acc[i][c]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = sTok[tg * 4 + i];
        if (tok >= 0) {
          float *dst = dk + ((long long)b * Nkv + tok) * HHD + h * HD;
#pragma unroll
          for (int c = 0; c < HD / 8; ++c) atomicAdd(dst + cl + 8 * c, acc[i][c]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = rg * 4 + i;
    if (r < g.nq) {
      float *dst = dq + ((long long)b * Nq + g.q0 + r) * HHD + h * HD;
#pragma unroll
      for (int c = 0; c < OC; ++c) {
        if (S > 1) atomicAdd(dst + cg + 16 * c, dqa[i][c]);     // dq zero-filled by the host wrapper when the box is split
        else dst[cg + 16 * c] = dqa[i][c];
      }
    }
  }
}

template <int HD> constexpr size_t fwd_smem_bytes() { return sizeof(float) * ((TQ + 2 * TK) * (HD + 4) + TQ * (TK + 4)); }
template <int HD> constexpr size_t bwd_smem_bytes() { return sizeof(float) * ((2 * TQ + 2 * TK) * (HD + 4) + TQ * (TK + 4)); }

}  // namespace roiattn
