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

// Four-channel vectors of the storage type (16 bytes fp32, 8 bytes bf16 / fp16), widened to fp32 in registers.
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
// 16-bit storage uses 8-byte vectors so that a lane owns the same four channels as in fp32: the fp32 grad_value
// reductions of neighbouring lanes then stay contiguous (one full sector per lane pair).  With 16-byte / 8-channel
// vectors each lane's two reductions straddle half sectors and the L2 atomic units see twice the transactions
// (measured: bf16 backward 11.1 ms vs 5.7 ms fp32, profiles/r01_sweep_before_bf16_fix.md).
template <typename H2, typename H> struct Vec16Half {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint2 t, float (&v)[4])
  {
    const H2 *h = reinterpret_cast<const H2 *>(&t);
    v[0] = Cvt<H>::up(h[0].x); v[1] = Cvt<H>::up(h[0].y); v[2] = Cvt<H>::up(h[1].x); v[3] = Cvt<H>::up(h[1].y);
  }
  static __device__ __forceinline__ void load(const H *p, float (&v)[4]) { unpack(__ldg(reinterpret_cast<const uint2 *>(p)), v); }
  static __device__ __forceinline__ void load_stream(const H *p, float (&v)[4])
  {
    uint2 t;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(t.x), "=r"(t.y) : "l"(p));
    unpack(t, v);
  }
  static __device__ __forceinline__ void store(H *p, const float (&v)[4])
  {
    uint2 t;
    H2 *h = reinterpret_cast<H2 *>(&t);
    h[0].x = Cvt<H>::down(v[0]); h[0].y = Cvt<H>::down(v[1]); h[1].x = Cvt<H>::down(v[2]); h[1].y = Cvt<H>::down(v[3]);
    __stcs(reinterpret_cast<uint2 *>(p), t);
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
// Vector kernels: C = G * NV * 4 channels per unit, G lanes per unit (G | 32), 32/G units per warp.
//
// Control flow is warp-uniform (every lane runs every loop iteration, shuffles use the full mask); only the
// "is this sample in range" test diverges, per group.  Per chunk of G samples, lane j of a group prepares sample j
// (coordinates -> clamped corner offset, zero-padded axis weights, strides) and parks it in a per-warp shared-memory
// slot; the group then walks the chunk reading each sample back with three broadcast LDS.128.
//
// Zero padding without predicated loads: the low corner is clamped into the volume and the stride along an axis is 0
// when one of its two sides is outside, so all eight corner loads are unconditional and in bounds; the axis weight of
// an outside side is forced to 0, which zeroes exactly the corner weights the reference zero-pads (cuh:60-107).  For
// finite inputs the blend is bit-identical: fma(w, 0, acc) == fma(0, v, acc) == acc.
// ---------------------------------------------------------------------------------------------------------------
struct PreparedSample {
  float4 a;   // {corner offset (uint bits), attention weight, hd_e, ld_e}
  float4 b;   // {hh_e, lh_e, hw_e, lw_e}
  int4 c;     // {stride d, stride h, stride w (elements; 0 when clamped), flags: bit0 in range, bits1..6 = dl,dh,hl,hh,wl,wh valid}
};

struct UnitCoords {
  bool active;
  int m;
  long long u, b;
  long long row;   // b * Lq + q = u / M: the unit's (batch, query) row (fused prologue: row of the reference points / merged projection)
  int q;           // query index inside the batch element
};

// Work scheduling.  A "block slot" is the set of UPB = 8 * 32/G units one CTA processes together.  Slots are handed
// out in contiguous runs (CTA i owns slots [i*per, (i+1)*per)), so what a CTA leaves in L1 is what it needs next.
//   linear order : slot t = units [t*UPB, (t+1)*UPB) in memory order (head fastest, then query).
//   brick order  : used when Lq == S, i.e. queries are the voxels of the level pyramid (the FPN refinement's
//                  self-attention, decoder_blocks.py:107-131).  A slot is one head x one BDxBHxBW brick of
//                  neighbouring query voxels: neighbouring queries sample neighbouring voxels, so the 8*32/G units of
//                  a CTA share corner lines in L1 instead of each pulling its own from L2.  Pure scheduling: any
//                  order gives the same results.
struct BrickPlan {
  int4 lvb[kMaxLevels];    // per level: bricks along d, h, w and index of the level's first brick
  int nb;                  // bricks per (batch, head)
};

template <int UPB> struct BrickDims;    // UPB = BD * BH * BW, w fastest
template <> struct BrickDims<8> { static constexpr int BD = 2, BH = 2, BW = 2; };
template <> struct BrickDims<16> { static constexpr int BD = 2, BH = 2, BW = 4; };
template <> struct BrickDims<32> { static constexpr int BD = 2, BH = 4, BW = 4; };
template <> struct BrickDims<64> { static constexpr int BD = 4, BH = 4, BW = 4; };
template <> struct BrickDims<128> { static constexpr int BD = 4, BH = 4, BW = 8; };
template <> struct BrickDims<256> { static constexpr int BD = 4, BH = 8, BW = 8; };

template <int UPB>
__device__ __forceinline__ void make_brick_plan(BrickPlan &bp, const int4 *lv, int L)
{
  using B = BrickDims<UPB>;
  int first = 0;
  for (int l = 0; l < L; ++l) {
    const int nd = (lv[l].x + B::BD - 1) / B::BD, nh = (lv[l].y + B::BH - 1) / B::BH, nw = (lv[l].z + B::BW - 1) / B::BW;
    bp.lvb[l] = make_int4(nd, nh, nw, first);
    first += nd * nh * nw;
  }
  bp.nb = first;
}

// unit of (slot t, unit-in-block ui)
template <int UPB>
__device__ __forceinline__ UnitCoords slot_unit(bool brick, long long t, int ui, long long total, const BrickPlan &bp,
                                                const int4 *lv, int L, int M, int Lq)
{
  UnitCoords c;
  if (!brick) {
    const long long u = t * UPB + ui;
    c.active = u < total;
    c.u = c.active ? u : 0;
    c.row = c.u / M;
    c.m = (int)(c.u - c.row * M);
    c.b = c.row / Lq;
    c.q = (int)(c.row - c.b * Lq);
    return c;
  }
  using B = BrickDims<UPB>;
  const long long per_batch = (long long)bp.nb * M;
  c.b = t / per_batch;
  const int r = (int)(t - c.b * per_batch);
  c.m = r / bp.nb;
  int br = r - c.m * bp.nb;
  int l = 0;
  while (l + 1 < L && br >= bp.lvb[l + 1].w) ++l;
  const int4 nb = bp.lvb[l], li = lv[l];
  br -= nb.w;
  const int bw = br % nb.z, bh = (br / nb.z) % nb.y, bd = br / (nb.z * nb.y);
  const int w = bw * B::BW + ui % B::BW, h = bh * B::BH + (ui / B::BW) % B::BH, d = bd * B::BD + ui / (B::BW * B::BH);
  c.active = d < li.x && h < li.y && w < li.z;
  const long long q = c.active ? (long long)li.w + ((long long)d * li.y + h) * li.z + w : 0;
  c.row = c.b * Lq + q;
  c.q = (int)q;
  c.u = c.row * M + c.m;
  if (!c.active) { c.u = 0; c.m = 0; c.b = 0; c.row = 0; c.q = 0; }
  return c;
}

// Lane-side preparation of one sample given its normalised location (x, y, z) and attention weight
// (cuh:424-428 + cuh:36-46 + corner guards cuh:60-107).
__device__ __forceinline__ PreparedSample prepare_located(const int4 li, float x, float y, float z, float w, const UnitCoords &uc, int S, int MC, int C)
{
  PreparedSample ps;
  ps.a = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.b = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.c = make_int4(0, 0, 0, 0);
  const Sample<float> sm = locate<float>(x, y, z, li.x, li.y, li.z);
  if (sm.mask != 0) {
    const bool vdl = sm.d_low >= 0, vdh = sm.d_low + 1 <= li.x - 1;
    const bool vhl = sm.h_low >= 0, vhh = sm.h_low + 1 <= li.y - 1;
    const bool vwl = sm.w_low >= 0, vwh = sm.w_low + 1 <= li.z - 1;
    const int sW = MC, sH = li.z * MC, sD = li.y * sH;
    const unsigned vox = (unsigned)((max(sm.d_low, 0) * li.y + max(sm.h_low, 0)) * li.z + max(sm.w_low, 0));
    const unsigned off = ((unsigned)(uc.b * S) + (unsigned)li.w + vox) * (unsigned)MC + (unsigned)(uc.m * C);
    ps.a = make_float4(__uint_as_float(off), w, vdl ? 1.f - sm.ld : 0.f, vdh ? sm.ld : 0.f);
    ps.b = make_float4(vhl ? 1.f - sm.lh : 0.f, vhh ? sm.lh : 0.f, vwl ? 1.f - sm.lw : 0.f, vwh ? sm.lw : 0.f);
    ps.c = make_int4((vdl && vdh) ? sD : 0, (vhl && vhh) ? sH : 0, (vwl && vwh) ? sW : 0,
                     1 | (vdl << 1) | (vdh << 2) | (vhl << 3) | (vhh << 4) | (vwl << 5) | (vwh << 6));
  }
  return ps;
}

// Sample s of unit uc from the op's own inputs: sampling_loc / attn_weight (cuh:405-423).
__device__ __forceinline__ PreparedSample prepare_sample(const int4 *lv, const float *__restrict__ loc, const float *__restrict__ aw,
                                                         const UnitCoords &uc, int s, int LP, int P, int S, int MC, int C)
{
  if (uc.active && s < LP) {
    const long long si = uc.u * LP + s;
    const float x = ldg_stream(loc + 3 * si), y = ldg_stream(loc + 3 * si + 1), z = ldg_stream(loc + 3 * si + 2);
    return prepare_located(lv[s / P], x, y, z, ldg_stream(aw + si), uc, S, MC, C);
  }
  PreparedSample ps;
  ps.a = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.b = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.c = make_int4(0, 0, 0, 0);
  return ps;
}

template <int G> __device__ __forceinline__ float group_max(float v)
{
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
template <int G> __device__ __forceinline__ float group_sum_f(float v)
{
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Fused prologue (MSDeformAttn.forward, transoar/models/ops/modules/ms_deform_attn.py:115-126): the op reads the RAW outputs of the
// sampling_offsets / attention_weights Linear layers and does, for the L*P <= G samples of its unit,
//     attn = softmax(logits)                   (over the unit's L*P samples; lanes of the group hold one sample each)
//     loc  = ref[q, l] + offset / (W, H, D)_l  (division then addition, each rounded once -- the same two fp32 operations ATen runs)
// so sampling_loc and attn_weight are never written to or read from HBM.  `weight` returns the lane's softmax value (0 for idle lanes).
// Where unit u keeps its L*P offsets / logits: dense arrays [units][L*P][3] / [units][L*P] (ld == 0), or both inside one row-major
// [N*Lq, ld] tensor -- the output of ONE Linear layer that computes offsets (columns [0, 3*M*L*P)) and logits (columns from
// logit_col) together -- so that the two small projections are a single GEMM and their gradients arrive in one tensor.
__device__ __forceinline__ long long fused_off_base(const UnitCoords &uc, int LP, int M, long long ld)
{
  return ld ? uc.row * ld + (long long)uc.m * LP * 3 : uc.u * LP * 3;
}
__device__ __forceinline__ long long fused_logit_base(const UnitCoords &uc, int LP, int M, long long ld, int logit_col)
{
  return ld ? uc.row * ld + logit_col + (long long)uc.m * LP : uc.u * LP;
}

template <int G>
__device__ __forceinline__ PreparedSample prepare_sample_fused(const int4 *lv, const float *__restrict__ off, const float *__restrict__ logit,
                                                               const float *__restrict__ ref, long long ref_bstride, const UnitCoords &uc, int s,
                                                               int LP, int P, int L, int M, int Lq, int S, int MC, int C, float &weight)
{
  // `off` / `logit` already point at this unit's first sample (fused_off_base / fused_logit_base)
  const bool mine = uc.active && s < LP;
  const long long si = s;
  const float lg = mine ? ldg_stream(logit + si) : -INFINITY;
  const float mx = group_max<G>(lg);
  const float e = mine ? expf(lg - mx) : 0.f;
  const float sum = group_sum_f<G>(e);
  weight = mine ? __fdiv_rn(e, sum) : 0.f;
  if (mine) {
    const int l = s / P;
    const int4 li = lv[l];
    const float *r = ref + uc.b * ref_bstride + ((long long)uc.q * L + l) * 3;
    const float x = __fadd_rn(__ldg(r), __fdiv_rn(ldg_stream(off + 3 * si), __int2float_rn(li.z)));
    const float y = __fadd_rn(__ldg(r + 1), __fdiv_rn(ldg_stream(off + 3 * si + 1), __int2float_rn(li.y)));
    const float z = __fadd_rn(__ldg(r + 2), __fdiv_rn(ldg_stream(off + 3 * si + 2), __int2float_rn(li.x)));
    return prepare_located(li, x, y, z, weight, uc, S, MC, C);
  }
  PreparedSample ps;
  ps.a = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.b = make_float4(0.f, 0.f, 0.f, 0.f);
  ps.c = make_int4(0, 0, 0, 0);
  return ps;
}

// cuh:109-110 from the (zero-padded) axis weights, products left to right: (d*h)*w.
__device__ __forceinline__ void corner_weights_axes(float hd, float ld, float hh, float lh, float hw, float lw, float (&w)[8])
{
  const float a = __fmul_rn(hd, hh), b = __fmul_rn(hd, lh), c = __fmul_rn(ld, hh), d = __fmul_rn(ld, lh);
  w[0] = __fmul_rn(a, hw); w[1] = __fmul_rn(a, lw); w[2] = __fmul_rn(b, hw); w[3] = __fmul_rn(b, lw);
  w[4] = __fmul_rn(c, hw); w[5] = __fmul_rn(c, lw); w[6] = __fmul_rn(d, hw); w[7] = __fmul_rn(d, lw);
}

// FUSED = 1: `loc` / `aw` are the raw sampling offsets / attention logits and `ref` [N|1, Lq, L, 3] the reference points
// (prepare_sample_fused); requires L * P <= G.
template <typename VT, int G, int NV, int MINB, int FUSED = 0>
__global__ void __launch_bounds__(kThreads, MINB)
fwd_vec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
               const float *__restrict__ loc, const float *__restrict__ aw, int N, int S, int M, int L, int Lq, int P,
               VT *__restrict__ out, int brick, const float *__restrict__ ref = nullptr, long long ref_bstride = 0, long long fused_ld = 0,
               int logit_col = 0)
{
  using V = Vec16<VT>;
  constexpr int VEC = V::N, CPL = VEC * NV, C = G * CPL, UPW = 32 / G, WARPS = kThreads / 32, UPB = WARPS * UPW;
  __shared__ int4 lv[kMaxLevels];
  __shared__ BrickPlan bp;
  __shared__ float4 sA[WARPS][32];
  __shared__ float4 sB[WARPS][32];
  __shared__ int4 sC[WARPS][32];
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, g0 = lane - gl;
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  if (brick && threadIdx.x == 0) make_brick_plan<UPB>(bp, lv, L);
  __syncthreads();
  const long long slots = brick ? (long long)N * M * bp.nb : (total + UPB - 1) / UPB;
  const long long per = (slots + gridDim.x - 1) / gridDim.x;
  const long long t_end = min(slots, (blockIdx.x + 1) * per);
  const unsigned lane_off = gl * VEC;                               // vector nv of this lane starts at nv*G*VEC + gl*VEC
  for (long long t = blockIdx.x * per; t < t_end; ++t) {
    const UnitCoords uc = slot_unit<UPB>(brick != 0, t, warp * UPW + lane / G, total, bp, lv, L, M, Lq);
    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;

    for (int s0 = 0; s0 < LP; s0 += G) {
      float w_unused;
      const PreparedSample mine = FUSED ? prepare_sample_fused<G>(lv, loc + fused_off_base(uc, LP, M, fused_ld), aw + fused_logit_base(uc, LP, M, fused_ld, logit_col),
                                                                  ref, ref_bstride, uc, s0 + gl, LP, P, L, M, Lq, S, MC, C, w_unused)
                                        : prepare_sample(lv, loc, aw, uc, s0 + gl, LP, P, S, MC, C);
      __syncwarp();
      sA[warp][lane] = mine.a; sB[warp][lane] = mine.b; sC[warp][lane] = mine.c;
      __syncwarp();
      const int cnt = min(G, LP - s0);
      for (int j = 0; j < cnt; ++j) {
        const int4 pc = sC[warp][g0 + j];
        if (pc.w == 0) continue;                                    // cuh:428 failed (or idle tail group)
        const float4 pa = sA[warp][g0 + j], pb = sB[warp][g0 + j];
        float w[8];
        corner_weights_axes(pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, w);
        unsigned o[8];
        o[0] = __float_as_uint(pa.x) + lane_off; o[1] = o[0] + pc.z; o[2] = o[0] + pc.y; o[3] = o[2] + pc.z;
        o[4] = o[0] + pc.x; o[5] = o[4] + pc.z; o[6] = o[4] + pc.y; o[7] = o[6] + pc.z;
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
          float v[8][VEC];
#pragma unroll
          for (int k = 0; k < 8; ++k) V::load(value + (o[k] + nv * (G * VEC)), v[k]);
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            const float val = blend<float>(w, v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
            acc[nv * VEC + c] = __fmaf_rn(pa.y, val, acc[nv * VEC + c]);   // cuh:430
          }
        }
      }
    }
    if (uc.active) {
      VT *o = out + uc.u * C + lane_off;
#pragma unroll
      for (int nv = 0; nv < NV; ++nv) {
        float tv[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) tv[c] = acc[nv * VEC + c];
        V::store(o + nv * (G * VEC), tv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENT (msda3d_set_tuning("stage", 1); profiles/r02_experiments.md): the forward with the COARSEST level's slab of the current
// (batch, head) staged in shared memory by the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP) and gathered from there with
// LDS.128; the other levels keep the LDG path.  This is the part of BASELINE.json's "stage per-level 3D feature tiles into shared
// memory via TMA" design that fits on chip at all: sampling offsets are measured in voxels of the SAMPLED level, so a query brick's
// footprint on any level is brick + a halo of +-(n_points + jitter) voxels -- at 256 B per (voxel, head) the halo alone is
// 15^3 x 256 B = 864 KB for every level, against 227 KB of shared memory.  Only a whole small level fits: 5x5x8 x 256 B = 51 KB.
// Brick order only (all units of a CTA share one (batch, head)), fp32, 16 lanes x one float4 per unit.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
fwd_stage_kernel(const float *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ starts,
                 const float *__restrict__ loc, const float *__restrict__ aw, int N, int S, int M, int L, int Lq, int P,
                 float *__restrict__ out)
{
  constexpr int G = 16, VEC = 4, C = 64, UPW = 2, WARPS = kThreads / 32, UPB = WARPS * UPW;
  extern __shared__ __align__(128) float slab[];                  // [V][C]: coarsest level of the staged (batch, head)
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int4 lv[kMaxLevels];
  __shared__ BrickPlan bp;
  __shared__ float4 sA[WARPS][32];
  __shared__ float4 sB[WARPS][32];
  __shared__ int4 sC[WARPS][32];
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, g0 = lane - gl;
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  if (threadIdx.x == 0) {
    make_brick_plan<UPB>(bp, lv, L);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lc = L - 1;
  const int4 lic = lv[lc];
  const int V = lic.x * lic.y * lic.z;
  const long long slots = (long long)N * M * bp.nb;
  const long long per = (slots + gridDim.x - 1) / gridDim.x;
  const long long t_end = min(slots, (blockIdx.x + 1) * per);
  const unsigned lane_off = gl * VEC;
  long long staged = -1;
  unsigned phase = 0;
  for (long long t = blockIdx.x * per; t < t_end; ++t) {
    const long long bm = t / bp.nb;                                // b * M + m of this slot (brick order)
    if (bm != staged) {
      __syncthreads();                                             // nobody still gathers from the old slab
      const long long b = bm / M;
      const int m = (int)(bm - b * M);
      if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"((unsigned)(V * C * 4)) : "memory");
      __syncthreads();
      for (int v = threadIdx.x; v < V; v += kThreads) {
        const float *src = value + ((b * S + lic.w + v) * (long long)M + m) * C;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(slab + v * C)), "l"(src), "r"(C * 4), "r"(smem_u32(&mbar)) : "memory");
      }
      unsigned done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
      phase ^= 1u;
      staged = bm;
    }
    const UnitCoords uc = slot_unit<UPB>(true, t, warp * UPW + lane / G, total, bp, lv, L, M, Lq);
    float acc[VEC] = {0.f, 0.f, 0.f, 0.f};
    for (int s0 = 0; s0 < LP; s0 += G) {
      const int s = s0 + gl;
      PreparedSample mine;
      if (uc.active && s < LP && s / P == lc) {                    // coarsest level: offsets relative to the compact slab [V][C]
        UnitCoords cu = uc;
        cu.b = 0; cu.m = 0;
        const long long si = uc.u * LP + s;
        mine = prepare_located(make_int4(lic.x, lic.y, lic.z, 0), ldg_stream(loc + 3 * si), ldg_stream(loc + 3 * si + 1),
                               ldg_stream(loc + 3 * si + 2), ldg_stream(aw + si), cu, 0, C, C);
      } else {
        mine = prepare_sample(lv, loc, aw, uc, s, LP, P, S, MC, C);
      }
      __syncwarp();
      sA[warp][lane] = mine.a; sB[warp][lane] = mine.b; sC[warp][lane] = mine.c;
      __syncwarp();
      const int cnt = min(G, LP - s0);
      for (int j = 0; j < cnt; ++j) {
        const int4 pc = sC[warp][g0 + j];
        if (pc.w == 0) continue;
        const float4 pa = sA[warp][g0 + j], pb = sB[warp][g0 + j];
        float w[8];
        corner_weights_axes(pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, w);
        unsigned o[8];
        o[0] = __float_as_uint(pa.x) + lane_off; o[1] = o[0] + pc.z; o[2] = o[0] + pc.y; o[3] = o[2] + pc.z;
        o[4] = o[0] + pc.x; o[5] = o[4] + pc.z; o[6] = o[4] + pc.y; o[7] = o[6] + pc.z;
        float v[8][VEC];
        if ((s0 + j) / P == lc) {                                   // warp-uniform: both units of the warp are at sample j
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 q = *reinterpret_cast<const float4 *>(slab + o[k]);
            v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) Vec16<float>::load(value + o[k], v[k]);
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float val = blend<float>(w, v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
          acc[c] = __fmaf_rn(pa.y, val, acc[c]);
        }
      }
    }
    if (uc.active) Vec16<float>::store(out + uc.u * C + lane_off, acc);
  }
}

template <int G> __device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
  for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Backward.  With t_c = grad_output[c] and dot_k = sum_c t_c * v_k[c] (k = corner), every gradient of a sample is a
// trilinear form in the eight dots:
//   grad_attn_weight      = sum_k w_k dot_k                                   (cuh:236-237)
//   grad_loc.{z,y,x}      = size * attn * sum_k d(w_k)/d{ld,lh,lw} dot_k      (cuh:159-231,238-240)
// so the per-channel work is the eight dots (one FMA per corner and channel); the rest is ~40 scalar operations per
// sample, evaluated separably (along w, then h, then d).  The reference spends 24 FMAs per channel on the same sums.
struct SampleGrads { float a, w, h, d; };    // per-lane partial sums over this lane's channels (unscaled)

// The four trilinear forms of a sample from its eight dots, evaluated separably (along w, then h, then d).
__device__ __forceinline__ SampleGrads grads_from_dots(const float (&dot)[8], const float4 pa, const float4 pb, const int flags)
{
  // separable evaluation of the four trilinear forms; s?l / s?h = -1 / +1 where that side is inside, else 0
  const float sdl = (flags & 2) ? -1.f : 0.f, sdh = (flags & 4) ? 1.f : 0.f;
  const float shl = (flags & 8) ? -1.f : 0.f, shh = (flags & 16) ? 1.f : 0.f;
  const float swl = (flags & 32) ? -1.f : 0.f, swh = (flags & 64) ? 1.f : 0.f;
  float A[4], Bw[4];                                       // index = 2*kd + kh
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    A[i] = pb.z * dot[2 * i] + pb.w * dot[2 * i + 1];
    Bw[i] = swl * dot[2 * i] + swh * dot[2 * i + 1];
  }
  float A2[2], Bh[2], Bw2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    A2[i] = pb.x * A[2 * i] + pb.y * A[2 * i + 1];
    Bh[i] = shl * A[2 * i] + shh * A[2 * i + 1];
    Bw2[i] = pb.x * Bw[2 * i] + pb.y * Bw[2 * i + 1];
  }
  SampleGrads g;
  g.a = pa.z * A2[0] + pa.w * A2[1];
  g.d = sdl * A2[0] + sdh * A2[1];
  g.h = pa.z * Bh[0] + pa.w * Bh[1];
  g.w = pa.z * Bw2[0] + pa.w * Bw2[1];
  return g;
}

template <typename VT, int G, int NV>
__device__ __forceinline__ SampleGrads sample_backward(const VT *__restrict__ value, const float (&top)[Vec16<VT>::N * NV],
                                                       const float4 pa, const float4 pb, const int4 pc, unsigned lane_off,
                                                       float (&w)[8], unsigned (&o)[8])
{
  using V = Vec16<VT>;
  constexpr int VEC = V::N;
  corner_weights_axes(pa.z, pa.w, pb.x, pb.y, pb.z, pb.w, w);
  o[0] = __float_as_uint(pa.x) + lane_off; o[1] = o[0] + pc.z; o[2] = o[0] + pc.y; o[3] = o[2] + pc.z;
  o[4] = o[0] + pc.x; o[5] = o[4] + pc.z; o[6] = o[4] + pc.y; o[7] = o[6] + pc.z;
  float dot[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) dot[k] = 0.f;
#pragma unroll
  for (int nv = 0; nv < NV; ++nv) {
    float v[8][VEC];
#pragma unroll
    for (int k = 0; k < 8; ++k) V::load(value + (o[k] + nv * (G * VEC)), v[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) dot[k] = fmaf(top[nv * VEC + c], v[k][c], dot[k]);
    }
  }
  return grads_from_dots(dot, pa, pb, pc.w);
}

// grad_value gets w_k * (t_c * attn) per corner as 128-bit reductions (skipped where the corner weight is zero).  On
// B200 these reductions are what bounds the kernel: the L2 atomic units take ~6.6 TB/s of fp32 payload
// (tools/micro/red_bench.cu), the launch sends 8 corners x C floats per in-range sample.  Combining contributions of
// neighbouring queries in shared memory first was tried and is slower (profiles/r01_experiments.md, section 4).
// FUSED = 1 (see fwd_vec_kernel): grad_loc / grad_aw receive the gradients with respect to the raw offsets / logits:
//   d/d offset = d/d loc / (W, H, D)      d/d logit_j = attn_j * (g_j - sum_k attn_k g_k)   (softmax backward inside the group)
// PAIR = 1 (G = 16, NV = 1, brick order): the two units of a warp are w-neighbouring query voxels of one head.  Their samples of a
// level land in the same cell or in w-adjacent cells whenever their offsets agree (always at initialisation, mostly in a trained
// model: the offsets come from one Linear layer applied to neighbouring voxels), so up to all eight grad_value reductions of the
// second unit hit addresses the first unit reduces into as well.  The half-warps compare corner offsets, the first unit adds the
// second unit's contributions with warp shuffles and issues ONE reduction per shared address: 25-47 % fewer L2 atomics at the
// price of ~35 shuffles per sample and 122 registers (2 CTAs per SM).  Only the order of the fp32 sums changes.  MEASURED SLOWER on
// B200 (5.16 -> 5.98 ms without jitter, 5.69 -> 6.96 ms with; profiles/r01_experiments.md section 8), so it is off unless
// msda3d_set_tuning("pair", 1) asks for it.
template <typename VT, int G, int NV, int MINB, int SKIP_RED = 0, int FUSED = 0, int PAIR = 0, int ROT = 0>
__global__ void __launch_bounds__(kThreads, MINB)
bwd_vec_kernel(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ starts, const float *__restrict__ loc, const float *__restrict__ aw, int N, int S,
               int M, int L, int Lq, int P, float *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_aw, int brick, const float *__restrict__ ref = nullptr, long long ref_bstride = 0,
               long long fused_ld = 0, int logit_col = 0)
{
  using V = Vec16<VT>;
  constexpr int VEC = V::N, CPL = VEC * NV, C = G * CPL, UPW = 32 / G, WARPS = kThreads / 32, UPB = WARPS * UPW;
  __shared__ int4 lv[kMaxLevels];
  __shared__ BrickPlan bp;
  __shared__ float4 sA[WARPS][32];
  __shared__ float4 sB[WARPS][32];
  __shared__ int4 sC[WARPS][32];
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, g0 = lane - gl;
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  // Sample-order rotation (ROT: 1 = per warp, 2 = per unit).  In brick order the units of a CTA are neighbouring query voxels of one
  // head; walking their samples in the same order makes all of them reduce into the same few coarse-level voxels at the same moment
  // (same-address reductions serialise in one L2 slice).  Starting unit i at sample i spreads the CTA over the (level, point) slots.
  // ROT & 4: consecutive CTAs (the ones resident together) take their runs of slots from DIFFERENT (batch, head) slabs instead of from
  // neighbouring bricks of one slab, so concurrently running CTAs do not meet in the same coarse-level voxels either.
  const int rot0 = (ROT & 3) == 0 ? 0 : ((ROT & 3) == 1 ? warp * UPW : warp * UPW + lane / G);
  static_assert(ROT == 0 || PAIR == 0, "rotation and pair combining exclude each other");
  if (brick && threadIdx.x == 0) make_brick_plan<UPB>(bp, lv, L);
  __syncthreads();
  const long long slots = brick ? (long long)N * M * bp.nb : (total + UPB - 1) / UPB;
  const long long per = (slots + gridDim.x - 1) / gridDim.x;
  long long t_end = min(slots, (blockIdx.x + 1) * per);
  const unsigned lane_off = gl * VEC;                               // vector nv of this lane starts at nv*G*VEC + gl*VEC
  long long t_begin = blockIdx.x * per;
  if ((ROT & 4) != 0) {
    const int K = N * M, Rk = (int)gridDim.x / K;
    long long run = blockIdx.x;
    if (run < (long long)Rk * K) run = (run % K) * Rk + run / K;
    t_begin = run * per;
    t_end = min(slots, (run + 1) * per);
  }
  for (long long t = t_begin; t < t_end; ++t) {
    const UnitCoords uc = slot_unit<UPB>(brick != 0, t, warp * UPW + lane / G, total, bp, lv, L, M, Lq);
    float top[CPL];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) {
      float tv[VEC];
      V::load_stream(grad_out + uc.u * C + lane_off + nv * (G * VEC), tv);
#pragma unroll
      for (int c = 0; c < VEC; ++c) top[nv * VEC + c] = tv[c];
    }

    for (int s0 = 0; s0 < LP; s0 += G) {
      float w_mine = 0.f;
      const PreparedSample mine = FUSED ? prepare_sample_fused<G>(lv, loc + fused_off_base(uc, LP, M, fused_ld), aw + fused_logit_base(uc, LP, M, fused_ld, logit_col),
                                                                  ref, ref_bstride, uc, s0 + gl, LP, P, L, M, Lq, S, MC, C, w_mine)
                                        : prepare_sample(lv, loc, aw, uc, s0 + gl, LP, P, S, MC, C);
      __syncwarp();
      sA[warp][lane] = mine.a; sB[warp][lane] = mine.b; sC[warp][lane] = mine.c;
      __syncwarp();
      float r_a = 0.f, r_w = 0.f, r_h = 0.f, r_d = 0.f;            // lane j keeps the sums of sample s0 + j
      const int cnt = min(G, LP - s0);
      const int jr = (ROT & 3) == 0 ? 0 : rot0 % cnt;
      for (int jj = 0; jj < cnt; ++jj) {
        int j = jj;
        if ((ROT & 3) != 0) { j += jr; if (j >= cnt) j -= cnt; }
        const int4 pc = sC[warp][g0 + j];
        SampleGrads q = {0.f, 0.f, 0.f, 0.f};
        if constexpr (PAIR != 0) {
          static_assert(PAIR == 0 || (G == 16 && NV == 1), "pair combining is written for 16 lanes x one 16-byte vector per unit");
          const bool act = pc.w != 0;
          float w[8], ay = 0.f;
          unsigned o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { w[k] = 0.f; o[k] = 0u; }
          if (act) {
            const float4 pa = sA[warp][g0 + j], pb = sB[warp][g0 + j];
            q = sample_backward<VT, G, NV>(value, top, pa, pb, pc, lane_off, w, o);
            ay = pa.y;
          }
          // what the partner unit (other half-warp, same lane offset) is doing with this sample
          const unsigned po0 = __shfl_xor_sync(0xffffffffu, o[0], 16), po1 = __shfl_xor_sync(0xffffffffu, o[1], 16);
          const int px = __shfl_xor_sync(0xffffffffu, pc.x, 16), py = __shfl_xor_sync(0xffffffffu, pc.y, 16), pz = __shfl_xor_sync(0xffffffffu, pc.z, 16);
          const int pact = __shfl_xor_sync(0xffffffffu, (int)act, 16);
          const bool second = lane >= 16;                                         // this lane belongs to the unit that hands its sums over
          const bool geom = act && pact != 0 && pc.x == px && pc.y == py && pc.z == pz;
          const bool same = geom && o[0] == po0;                                  // same cell: all eight corners coincide
          const bool shift = geom && !same && pc.z != 0 && (second ? o[0] == po1 : po0 == o[1]);   // second unit one cell further in w
          unsigned need = 0;
          float c[8][4];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float wk = w[k] * ay;
            if (w[k] != 0.f) need |= 1u << k;
#pragma unroll
            for (int i = 0; i < 4; ++i) c[k][i] = wk * top[i];
          }
          if (same || shift) {                                                    // uniform over the warp's two units
            const unsigned pneed = __shfl_xor_sync(0xffffffffu, need, 16);
            if (same) {                                                           // corner k of the second unit lands on corner k of the first
#pragma unroll
              for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float got = __shfl_xor_sync(0xffffffffu, c[k][i], 16);
                  if (!second) c[k][i] += got;
                }
              }
              need = second ? 0u : (need | pneed);
            } else {                                                              // its w-low corner k - 1 lands on the first unit's w-high corner k
#pragma unroll
              for (int k = 1; k < 8; k += 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float got = __shfl_xor_sync(0xffffffffu, c[k - 1][i], 16);
                  if (!second) c[k][i] += got;
                }
              }
              need = second ? (need & 0xAAu) : (need | ((pneed & 0x55u) << 1));
            }
          }
          if (SKIP_RED == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (need >> k & 1u) red_add_v4(grad_value + o[k], c[k][0], c[k][1], c[k][2], c[k][3]);
          }
        } else if (pc.w != 0) {
          const float4 pa = sA[warp][g0 + j], pb = sB[warp][g0 + j];
          float w[8];
          unsigned o[8];
          q = sample_backward<VT, G, NV>(value, top, pa, pb, pc, lane_off, w, o);
          if (SKIP_RED == 0 || (SKIP_RED == 2 && (j & 3) == 0) || (SKIP_RED == 3 && (s0 + j) / P < L - 1) || (SKIP_RED == 4 && (s0 + j) / P < L - 2)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (w[k] != 0.f) {
                const float wk = w[k] * pa.y;                       // cuh:151,166: w_k * (top_grad * attn_weight)
#pragma unroll
                for (int c4 = 0; c4 < CPL; c4 += 4)
                  red_add_v4(grad_value + (o[k] + (c4 / VEC) * (G * VEC) + (c4 % VEC)), wk * top[c4], wk * top[c4 + 1],
                             wk * top[c4 + 2], wk * top[c4 + 3]);
              }
            }
          }
        }
        q.a = group_sum<G>(q.a); q.w = group_sum<G>(q.w); q.h = group_sum<G>(q.h); q.d = group_sum<G>(q.d);
        if (gl == j) { r_a = q.a; r_w = q.w; r_h = q.h; r_d = q.d; }
      }
      const int s = s0 + gl;
      if (FUSED) {
        const float dot = group_sum_f<G>(w_mine * r_a);             // softmax backward needs the whole unit: every lane takes part
        if (uc.active && s < LP) {
          const int4 li = lv[s / P];
          const float fw = __int2float_rn(li.z), fh = __int2float_rn(li.y), fd = __int2float_rn(li.x);
          float *gl_ = grad_loc + fused_off_base(uc, LP, M, fused_ld) + 3 * s;
          gl_[0] = __fdiv_rn(fw * (r_w * mine.a.y), fw);            // d loc (cuh:238-240), then the gradient of offset / W
          gl_[1] = __fdiv_rn(fh * (r_h * mine.a.y), fh);
          gl_[2] = __fdiv_rn(fd * (r_d * mine.a.y), fd);
          grad_aw[fused_logit_base(uc, LP, M, fused_ld, logit_col) + s] = w_mine * (r_a - dot);
        }
      } else if (uc.active && s < LP) {
        // my own sample: scale by attn * size (cuh:238-240); out-of-range samples carry exact zeros (cuh:618-621)
        const int4 li = lv[s / P];
        float *gl_ = grad_loc + (uc.u * LP + s) * 3;
        gl_[0] = __int2float_rn(li.z) * (r_w * mine.a.y);
        gl_[1] = __int2float_rn(li.y) * (r_h * mine.a.y);
        gl_[2] = __int2float_rn(li.x) * (r_d * mine.a.y);
        grad_aw[uc.u * LP + s] = r_a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bwd_duo_kernel: the backward with TWO w-neighbouring query voxels per 16-lane group (fp32, C = 64, brick order).
//
// The refinement's queries are the voxels of the pyramid and their offsets come from one Linear layer, so two w-neighbouring queries
// of one head sample, level by level, the SAME cell (coarser levels: the pair is 1/2 ... 1/8 voxel apart) or w-ADJACENT cells (the
// pair's own level) unless their offsets differ by a large part of a voxel.  A group that owns both queries then needs 8 (same cell)
// or 12 (adjacent cells) corner rows instead of 16: that many gathers (the rows feed both units' dots) and that many grad_value
// reductions (one per row, carrying both units' contributions).  The L2 atomic units and the L1 data pipe -- the two resources that
// bound bwd_vec_kernel -- see up to half of the traffic.  Detection is exact (corner offsets and clamped strides of the two prepared
// samples are compared), everything else -- pairs at volume borders, offsets that disagree -- takes the one-unit path of
// bwd_vec_kernel, unit by unit.  No shuffles and no shared-memory accumulation are involved (the pair lives in ONE lane's registers),
// which is what made the earlier combining experiments slower (PAIR above; profiles/r01_experiments.md sections 4 and 8).
// Rows i = 2 * kd + kh; columns c = position along w: A's low / high corners are columns 0 / 1, B's are delta / delta + 1.
// ---------------------------------------------------------------------------------------------------------------
// Sum eight per-lane values over the 16 lanes of a group with 8 shuffles instead of 32: every butterfly step halves the number of
// values a lane carries (it keeps one half of them and hands the other half to its partner).  Returns, in lane gl, the group's
// total of x[gl >> 1].
__device__ __forceinline__ float group_sum8_packed(const float (&x)[8], int gl)
{
  const bool h1 = (gl & 8) != 0, h2 = (gl & 4) != 0, h3 = (gl & 2) != 0;
  float y[4], z[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = (h1 ? x[4 + i] : x[i]) + __shfl_xor_sync(0xffffffffu, h1 ? x[i] : x[4 + i], 8);
#pragma unroll
  for (int i = 0; i < 2; ++i) z[i] = (h2 ? y[2 + i] : y[i]) + __shfl_xor_sync(0xffffffffu, h2 ? y[i] : y[2 + i], 4);
  float t = (h3 ? z[1] : z[0]) + __shfl_xor_sync(0xffffffffu, h3 ? z[0] : z[1], 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

template <int SKIP_RED>
__device__ __forceinline__ SampleGrads duo_single(const float *__restrict__ value, const float (&top)[4], const float4 pa, const float4 pb,
                                                  const int4 pc, unsigned lane_off, float *__restrict__ grad_value)
{
  float w[8];
  unsigned o[8];
  const SampleGrads q = sample_backward<float, 16, 1>(value, top, pa, pb, pc, lane_off, w, o);
  if (SKIP_RED == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (w[k] != 0.f) {
        const float wk = w[k] * pa.y;
        red_add_v4(grad_value + o[k], wk * top[0], wk * top[1], wk * top[2], wk * top[3]);
      }
    }
  }
  return q;
}

template <int FUSED, int ROT, int SKIP_RED = 0, int THREADS = kThreads, int MINB = 2>
__global__ void __launch_bounds__(THREADS, MINB)
bwd_duo_kernel(const float *__restrict__ grad_out, const float *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ starts, const float *__restrict__ loc, const float *__restrict__ aw, int N, int S,
               int M, int L, int Lq, int P, float *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_aw, const float *__restrict__ ref, long long ref_bstride, long long fused_ld, int logit_col)
{
  using V = Vec16<float>;
  constexpr int G = 16, VEC = 4, C = 64, WARPS = THREADS / 32, UPB = WARPS * 4;      // a slot = one head x one 2x4x4 brick = 32 units (256 threads)
  __shared__ int4 lv[kMaxLevels];
  __shared__ BrickPlan bp;
  __shared__ float4 sA[2][WARPS][32];
  __shared__ float4 sB[2][WARPS][32];
  __shared__ int4 sC[2][WARPS][32];
  __shared__ float4 sR[WARPS * 2][G][2];                          // per group and sample: the eight sums {A: a, w, h, d; B: a, w, h, d}
  if (threadIdx.x < L)
    lv[threadIdx.x] = make_int4((int)shapes[3 * threadIdx.x], (int)shapes[3 * threadIdx.x + 1],
                                (int)shapes[3 * threadIdx.x + 2], (int)starts[threadIdx.x]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % G, g0 = lane - gl, grp = warp * 2 + lane / G;
  const int MC = M * C, LP = L * P;
  const long long total = (long long)N * Lq * M;
  if (threadIdx.x == 0) make_brick_plan<UPB>(bp, lv, L);
  __syncthreads();
  const long long slots = (long long)N * M * bp.nb;
  const long long per = (slots + gridDim.x - 1) / gridDim.x;
  long long run = blockIdx.x;
  if ((ROT & 4) != 0) {                                           // consecutive CTAs on different (batch, head) slabs (bwd_vec_kernel, ROT)
    const int K = N * M, Rk = (int)gridDim.x / K;
    if (run < (long long)Rk * K) run = (run % K) * Rk + run / K;
  }
  const long long t_begin = run * per, t_end = min(slots, (run + 1) * per);
  const unsigned lane_off = gl * VEC;
  const int rot0 = (ROT & 3) == 0 ? 0 : ((ROT & 3) == 1 ? warp * 2 : grp);
  for (long long t = t_begin; t < t_end; ++t) {
    const UnitCoords ucA = slot_unit<UPB>(true, t, 2 * grp, total, bp, lv, L, M, Lq);
    const UnitCoords ucB = slot_unit<UPB>(true, t, 2 * grp + 1, total, bp, lv, L, M, Lq);
    float topA[VEC], topB[VEC];
    V::load_stream(grad_out + ucA.u * C + lane_off, topA);
    V::load_stream(grad_out + ucB.u * C + lane_off, topB);

    for (int s0 = 0; s0 < LP; s0 += G) {
      float wmA = 0.f, wmB = 0.f;
      const PreparedSample mA = FUSED ? prepare_sample_fused<G>(lv, loc + fused_off_base(ucA, LP, M, fused_ld), aw + fused_logit_base(ucA, LP, M, fused_ld, logit_col),
                                                                ref, ref_bstride, ucA, s0 + gl, LP, P, L, M, Lq, S, MC, C, wmA)
                                      : prepare_sample(lv, loc, aw, ucA, s0 + gl, LP, P, S, MC, C);
      const PreparedSample mB = FUSED ? prepare_sample_fused<G>(lv, loc + fused_off_base(ucB, LP, M, fused_ld), aw + fused_logit_base(ucB, LP, M, fused_ld, logit_col),
                                                                ref, ref_bstride, ucB, s0 + gl, LP, P, L, M, Lq, S, MC, C, wmB)
                                      : prepare_sample(lv, loc, aw, ucB, s0 + gl, LP, P, S, MC, C);
      const float attA = mA.a.y, attB = mB.a.y;
      __syncwarp();
      sA[0][warp][lane] = mA.a; sB[0][warp][lane] = mA.b; sC[0][warp][lane] = mA.c;
      sA[1][warp][lane] = mB.a; sB[1][warp][lane] = mB.b; sC[1][warp][lane] = mB.c;
      __syncwarp();
      float rA_a = 0.f, rA_w = 0.f, rA_h = 0.f, rA_d = 0.f, rB_a = 0.f, rB_w = 0.f, rB_h = 0.f, rB_d = 0.f;   // lane j keeps the sums of sample s0 + j
      const int cnt = min(G, LP - s0);
      const int jr = (ROT & 3) == 0 ? 0 : rot0 % cnt;
      for (int jj = 0; jj < cnt; ++jj) {
        int j = jj;
        if ((ROT & 3) != 0) { j += jr; if (j >= cnt) j -= cnt; }
        const int4 pcA = sC[0][warp][g0 + j], pcB = sC[1][warp][g0 + j];
        SampleGrads qA = {0.f, 0.f, 0.f, 0.f}, qB = {0.f, 0.f, 0.f, 0.f};
        const bool actA = pcA.w != 0, actB = pcB.w != 0;
        if (actA || actB) {
          const float4 paA = sA[0][warp][g0 + j], pbA = sB[0][warp][g0 + j], paB = sA[1][warp][g0 + j], pbB = sB[1][warp][g0 + j];
          const unsigned offA = __float_as_uint(paA.x), offB = __float_as_uint(paB.x);
          const bool geom = actA && actB && pcA.x == pcB.x && pcA.y == pcB.y && pcA.z == pcB.z;
          const bool same = geom && offA == offB;
          const bool shift = geom && pcA.z != 0 && offB == offA + (unsigned)pcA.z;
          if (same || shift) {
            unsigned orow[4];
            orow[0] = offA + lane_off; orow[1] = orow[0] + pcA.y; orow[2] = orow[0] + pcA.x; orow[3] = orow[2] + pcA.y;
            const unsigned sw = (unsigned)pcA.z;
            float v[4][3][VEC];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              V::load(value + orow[i], v[i][0]);
              V::load(value + (orow[i] + sw), v[i][1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (shift) {
                V::load(value + (orow[i] + 2u * sw), v[i][2]);
              } else {
#pragma unroll
                for (int c = 0; c < VEC; ++c) v[i][2][c] = 0.f;
              }
            }
            float dotA[8], dotB[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float dA0 = 0.f, dA1 = 0.f, dB0 = 0.f, dB1 = 0.f, dB2 = 0.f;
#pragma unroll
              for (int c = 0; c < VEC; ++c) {
                dA0 = fmaf(topA[c], v[i][0][c], dA0); dA1 = fmaf(topA[c], v[i][1][c], dA1);
                dB0 = fmaf(topB[c], v[i][0][c], dB0); dB1 = fmaf(topB[c], v[i][1][c], dB1); dB2 = fmaf(topB[c], v[i][2][c], dB2);
              }
              dotA[2 * i] = dA0; dotA[2 * i + 1] = dA1;
              dotB[2 * i] = shift ? dB1 : dB0; dotB[2 * i + 1] = shift ? dB2 : dB1;
            }
            qA = grads_from_dots(dotA, paA, pbA, pcA.w);
            qB = grads_from_dots(dotB, paB, pbB, pcB.w);
            if (SKIP_RED == 0) {
              // row weights (d x h) in the product order of corner_weights_axes, times the unit's attention weight (cuh:151,166)
              const float rwA[4] = {__fmul_rn(paA.z, pbA.x), __fmul_rn(paA.z, pbA.y), __fmul_rn(paA.w, pbA.x), __fmul_rn(paA.w, pbA.y)};
              const float rwB[4] = {__fmul_rn(paB.z, pbB.x), __fmul_rn(paB.z, pbB.y), __fmul_rn(paB.w, pbB.x), __fmul_rn(paB.w, pbB.y)};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float wAl = __fmul_rn(rwA[i], pbA.z), wAh = __fmul_rn(rwA[i], pbA.w);
                const float wBl = __fmul_rn(rwB[i], pbB.z), wBh = __fmul_rn(rwB[i], pbB.w);
                const float a0 = wAl * paA.y, a1 = wAh * paA.y;
                const float b0 = (shift ? 0.f : wBl) * paB.y, b1 = (shift ? wBl : wBh) * paB.y, b2 = wBh * paB.y;
                if (wAl != 0.f || (!shift && wBl != 0.f))
                  red_add_v4(grad_value + orow[i], fmaf(b0, topB[0], a0 * topA[0]), fmaf(b0, topB[1], a0 * topA[1]),
                             fmaf(b0, topB[2], a0 * topA[2]), fmaf(b0, topB[3], a0 * topA[3]));
                if (wAh != 0.f || (shift ? wBl != 0.f : wBh != 0.f))
                  red_add_v4(grad_value + (orow[i] + sw), fmaf(b1, topB[0], a1 * topA[0]), fmaf(b1, topB[1], a1 * topA[1]),
                             fmaf(b1, topB[2], a1 * topA[2]), fmaf(b1, topB[3], a1 * topA[3]));
                if (shift && wBh != 0.f)
                  red_add_v4(grad_value + (orow[i] + 2u * sw), b2 * topB[0], b2 * topB[1], b2 * topB[2], b2 * topB[3]);
              }
            }
          } else {
#pragma unroll 1
            for (int u = 0; u < 2; ++u) {                           // one code copy for both units (instruction-cache footprint)
              const float4 pa = u ? paB : paA, pb = u ? pbB : pbA;
              const int4 pc = u ? pcB : pcA;
              if (pc.w == 0) continue;
              const float top[VEC] = {u ? topB[0] : topA[0], u ? topB[1] : topA[1], u ? topB[2] : topA[2], u ? topB[3] : topA[3]};
              const SampleGrads q = duo_single<SKIP_RED>(value, top, pa, pb, pc, lane_off, grad_value);
              if (u) qB = q; else qA = q;
            }
          }
        }
        const float x[8] = {qA.a, qA.w, qA.h, qA.d, qB.a, qB.w, qB.h, qB.d};
        const float tot = group_sum8_packed(x, gl);                // lane gl: total of x[gl >> 1]
        if ((gl & 1) == 0) reinterpret_cast<float *>(&sR[grp][j][0])[gl >> 1] = tot;
      }
      __syncwarp();
      {
        if (gl < cnt) {                                            // lane j takes the sums of sample s0 + j
          const float4 ra = sR[grp][gl][0], rb = sR[grp][gl][1];
          rA_a = ra.x; rA_w = ra.y; rA_h = ra.z; rA_d = ra.w; rB_a = rb.x; rB_w = rb.y; rB_h = rb.z; rB_d = rb.w;
        }
      }
      const int s = s0 + gl;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const UnitCoords &uc = u ? ucB : ucA;
        const float r_a = u ? rB_a : rA_a, r_w = u ? rB_w : rA_w, r_h = u ? rB_h : rA_h, r_d = u ? rB_d : rA_d;
        const float att = u ? attB : attA, w_mine = u ? wmB : wmA;
        if (FUSED) {
          const float dot = group_sum_f<G>(w_mine * r_a);          // softmax backward needs the whole unit: every lane takes part
          if (uc.active && s < LP) {
            const int4 li = lv[s / P];
            const float fw = __int2float_rn(li.z), fh = __int2float_rn(li.y), fd = __int2float_rn(li.x);
            float *gl_ = grad_loc + fused_off_base(uc, LP, M, fused_ld) + 3 * s;
            gl_[0] = __fdiv_rn(fw * (r_w * att), fw);
            gl_[1] = __fdiv_rn(fh * (r_h * att), fh);
            gl_[2] = __fdiv_rn(fd * (r_d * att), fd);
            grad_aw[fused_logit_base(uc, LP, M, fused_ld, logit_col) + s] = w_mine * (r_a - dot);
          }
        } else if (uc.active && s < LP) {
          const int4 li = lv[s / P];
          float *gl_ = grad_loc + (uc.u * LP + s) * 3;
          gl_[0] = __int2float_rn(li.z) * (r_w * att);
          gl_[1] = __int2float_rn(li.y) * (r_h * att);
          gl_[2] = __int2float_rn(li.x) * (r_d * att);
          grad_aw[uc.u * LP + s] = r_a;
        }
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
