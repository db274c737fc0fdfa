// conv3d_tc_kernels.cuh -- 3x3x3 / stride 1 / padding 1 convolution of a channels-last (NDHWC) fp32 volume on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands by TMA), for the narrow full-resolution stages of the AttnFPN
// encoder (EncoderCnnBlock stage 0, second convolution: 24 -> 24 channels at 160x160x256, transoar/models/backbones/encoder_blocks.py:
// 34-40 via attn_fpn.py:170-182), forward and -- called with the flipped / transposed weights -- the gradient with respect to the input.
// cuDNN's choice for that input gradient on B200 is a "strided dgrad" kernel that takes 4.1 ms per step; the forward takes 1.5 ms.
//
// Implicit GEMM without an im2col copy.  One CTA tile = 1 x 8 x 16 INPUT columns (d, h, w) = the 128 rows of the MMA.  A CTA takes a
// work item (n, 40 output planes, h tile, w tile) and marches through its depth planes: the 10 x 16 halo of ONE input plane arrives per
// output plane by one 5-D TMA box (rows of 32 channels: the copy engine zero-fills channels >= CI and the out-of-bounds voxels, which
// is the convolution's padding; 128-byte swizzle) into a ring of 5 planes -- planes d - 1 and d are still there from the previous
// output planes, so the input is read from L2 1.5 x instead of 4.3 x.  In that form the A operand of a tap pair (kd, kh) is the ring
// slot of plane d + kd - 1 plus a START-ADDRESS OFFSET of kh tile rows (16 rows = two swizzle atoms) in the shared-memory descriptor.
// The three kw taps are NOT three more offsets: they are folded into N.  Row r = (h, wi) of the accumulator holds, in column block kw,
// the partial sum P_kw[wi] = sum_{kd, kh, ci} x[d + kd - 1, h + kh - 1, w0 - 1 + wi, ci] * w[kd, kh, kw, ci, co]; the output is
// y[w0 + j] = P_0[j] + P_1[j + 1] + P_2[j + 2], which the epilogue forms with two warp shuffles per channel (a warp owns two complete
// tile rows; 14 of the 16 columns of a tile row are outputs).  So a tile costs 9 x CI/8 = 27 MMAs of 128 x 96 x 8 instead of 81 of
// 128 x 32 x 8.  The weights of all 27 taps live in shared memory for the whole (persistent) kernel, rounded to TF32 when they are
// staged.  Roles as in tc_gemm_kernels.cuh: TMA producer warp, single-thread MMA issuer, four epilogue warps (TMEM -> registers ->
// shuffles -> one contiguous CO * 4-byte row per voxel), four TMEM accumulators.
#pragma once

#include "tc_gemm_kernels.cuh"

namespace convtc {

using namespace tcgemm;

constexpr int TH = 16, TW = 8;                 // weight-gradient tile: 1 x TH x TW voxels of dy
constexpr int HH = TH + 2, HW = TW + 2;        // and its halo tile of x: 3 x HH x HW voxels
constexpr int kThreadsConv = 192;
constexpr int NPAD = 32;                       // MMA N (output channels padded to a multiple of 16)

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, int c4)
{
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// no-swizzle shared-memory descriptor: start, leading-dimension byte offset, stride byte offset (all in 16-byte units), sm_100 version 1
__device__ __forceinline__ uint64_t desc_noswizzle(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ float to_tf32(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// forward / input-gradient tile geometry
constexpr int FTH = 8;                          // tile rows (h)
constexpr int FWI = 16;                         // input columns per tile row = MMA rows per tile row (two 8-row groups)
constexpr int FWO = FWI - 2;                    // output columns per tile row
constexpr int FHH = FTH + 2;                    // halo rows
constexpr int kFPlaneBytes = FHH * FWI * 128;   // one depth plane of the halo, rows of 32 channels (CI real + zero fill) = 128 B: 20480 bytes
constexpr int NF = 3 * NPAD;                    // MMA N: (kw, co)
constexpr int NACC = 4;                         // TMEM accumulators (4 x 96 columns): tiles in flight between the MMA issuer and the epilogue
constexpr int DSEG = 40;                        // output planes a CTA marches through per work item (2 extra halo planes per item)
static_assert(kFPlaneBytes % 1024 == 0 && FTH * FWI == 128 && (FWI * 128) % 1024 == 0, "tile geometry / swizzle atom alignment");

// One lane of a converged warp.  The MMA issuer runs its loop with the WHOLE warp converged and only the tcgen05 instructions under
// this predicate: every operand (descriptors, TMEM addresses, barrier addresses) is then warp-uniform for the compiler and lives in
// uniform registers.  Under `if (lane == 0)` the same values count as divergent and every MMA pays a vector-to-uniform "waterfall"
// (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~12 instructions) -- which, not the tensor pipe, set the pace of this kernel's MMA chain.
// (tcgemm::elect_one, tc_gemm_kernels.cuh)
template <int CI> struct ConvCfg {
  static constexpr int STAGE_BYTES = kFPlaneBytes;                           // one depth plane of the halo
  static constexpr int W_PAIR_BYTES = NF * 128;                              // B of one (kd, kh): [n = kw * 32 + co (96)][32 ci] rows of 128 B
  static constexpr int W_BYTES = 9 * W_PAIR_BYTES;                           // 110592
  static constexpr int STAGES = 5;                                           // ring of depth planes: 3 in use, 2 in flight
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + W_BYTES + 256 + 1024;
};

// K-major operand in the 128-byte swizzle: rows of 128 B (32 floats of K), 8-row atoms of 1024 B, 16-byte chunk c of row r at chunk
// c ^ (r % 8); a k-step of 8 floats is +32 B on the start address.  (The no-swizzle canonical layout -- chunk planes
// [ci / 4][voxel][4 floats], six 4-channel TMA boxes per plane -- was the first version of this kernel and runs the MMAs at exactly
// the same rate, ~100 clocks per 128 x 96 x 8; the swizzled form needs one TMA per plane instead of six.)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr)
{
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// x [N, D, H, W, CI] through the tensor map; wg [27][CO][CI] (tap-major; for the input gradient: flipped taps, transposed channels);
// y [N, D, H, W, CO].  CI % 8 == 0, CO % 4 == 0, CO <= 32.
template <int CI>
__global__ void __launch_bounds__(kThreadsConv, 1)
conv3d_k3_kernel(const __grid_constant__ CUtensorMap tmX, const float *__restrict__ wg, float *__restrict__ y, int N, int D, int H, int W, int CO,
                 int diag)   // diag (conv3d_tc_debug_mode, timing experiments only): 1 = epilogue does not read / store, 2 = no MMAs, 4 = no stores
{
  using C = ConvCfg<CI>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bars = wbase + C::W_BYTES;
  constexpr int R = C::STAGES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (R + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * R + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * R + NACC + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * R + 2 * NACC);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for the compiler

  // weights -> shared memory, B operand of the tap pair a = kd * 3 + kh: row n = kw * 32 + co holds wg[a * 3 + kw][co][0 .. CI) and
  // zeros up to 32 channels, 16-byte chunks swizzled; rows co >= CO zero
  for (int i = threadIdx.x; i < 9 * NF * 8; i += kThreadsConv) {
    const int c = i % 8, n = (i / 8) % NF, a = i / (8 * NF);
    const int kw = n / NPAD, co = n % NPAD;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (co < CO && 4 * c < CI) {
      const float4 g = __ldg(reinterpret_cast<const float4 *>(wg + ((long long)(a * 3 + kw) * CO + co) * CI + 4 * c));
      v = make_float4(to_tf32(g.x), to_tf32(g.y), to_tf32(g.z), to_tf32(g.w));
    }
    const uint32_t dst = wbase + (uint32_t)(a * C::W_PAIR_BYTES + n * 128 + ((c ^ (n & 7)) << 4));
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < R; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the generic-proxy weight stores must be visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // work item = (n, depth segment, h tile, w tile): the CTA marches through the segment's output planes; input plane d + 1 is the only
  // new data per output plane (planes d - 1 and d are still in the ring), so the halo is read from L2 1.05 x 1.43 times instead of 4.3
  const int th = (H + FTH - 1) / FTH, tw = (W + FWO - 1) / FWO, segs = (D + DSEG - 1) / DSEG;
  const long long items = (long long)N * segs * th * tw;
  auto item_coords = [&](long long t, int &iw, int &ih, int &d0, int &seg, int &n) {
    iw = (int)(t % tw); ih = (int)((t / tw) % th);
    const int sg = (int)((t / ((long long)tw * th)) % segs);
    n = (int)(t / ((long long)tw * th * segs));
    d0 = sg * DSEG; seg = min(DSEG, D - d0);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;                                             // running plane index: slot g % R, parity (g / R) & 1
      for (long long t = blockIdx.x; t < items; t += gridDim.x) {
        int iw, ih, d0, seg, n;
        item_coords(t, iw, ih, d0, seg, n);
        for (int p = 0; p < seg + 2; ++p, ++g) {
          const uint32_t slot = g % R, par = (g / R) & 1u;
          mbar_wait(empty(slot), par ^ 1u);
          mbar_expect_tx(full(slot), kFPlaneBytes);
          tma_load_5d(base + slot * C::STAGE_BYTES, &tmX, full(slot), 0, iw * FWO - 1, ih * FTH - 1, d0 - 1 + p, n);
        }
      }
    }
  } else if (warp == 1) {
    {
      // M = 128, N = 96, both operands K-major, TF32 in, fp32 out.  The whole warp walks the loop (see elect_one).
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NF >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int as = 0;
      uint32_t aphase = 0, g0 = 0;
      const uint64_t db0 = desc_k_sw128(wbase);
      for (long long t = blockIdx.x; t < items; t += gridDim.x) {
        int iw, ih, d0, seg, n;
        item_coords(t, iw, ih, d0, seg, n);
        for (int j = 0; j < seg; ++j) {
          mbar_wait(tempty(as), aphase ^ 1u);
          if (j == 0) {
            mbar_wait(full(g0 % R), (g0 / R) & 1u);
            mbar_wait(full((g0 + 1) % R), ((g0 + 1) / R) & 1u);
          }
          const uint32_t gn = g0 + j + 2;
          mbar_wait(full(gn % R), (gn / R) & 1u);
          tc_fence_after();
          const uint32_t acc = tmem_u + (uint32_t)(as * NF);
          // One descriptor per depth plane of the ring and per tile, then every (kh, k-step) only adds a compile-time constant to the
          // 14-bit start-address field (all shared-memory addresses are < 256 KB, so the field never carries into the LBO field):
          // kh = 16 rows = two swizzle atoms, k-step = 32 B.
          uint64_t da[3];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) da[kd] = desc_k_sw128(base + ((g0 + j + kd) % R) * C::STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int a = 0; a < 9; ++a) {
              const uint32_t a16 = (uint32_t)((a % 3) * FWI * 128 / 16);
              const uint32_t b16 = (uint32_t)(a * (C::W_PAIR_BYTES / 16));
#pragma unroll
              for (int s = 0; s < CI / 8; ++s)
                if (!(diag & 2)) umma_tf32(acc, da[a / 3] + (a16 + (uint32_t)(2 * s)), db0 + (b16 + (uint32_t)(2 * s)), idesc, (a | s) != 0 ? 1u : 0u);
            }
            umma_commit(tfull(as));
            umma_commit(empty((g0 + j) % R));                        // plane d - 1 is not needed again
            if (j == seg - 1) {                                      // end of the segment: its last two planes as well
              umma_commit(empty((g0 + j + 1) % R));
              umma_commit(empty((g0 + j + 2) % R));
            }
          }
          __syncwarp();
          if (++as == NACC) { as = 0; aphase ^= 1u; }
        }
        g0 += (uint32_t)(seg + 2);
      }
    }
  } else {
    // warp q owns accumulator rows q * 32 .. + 31 = tile rows 2 q and 2 q + 1, lane = (row parity, wi)
    const int q = warp & 3, hh = q * 2 + (lane >> 4), wi = lane & 15;
    int as = 0;
    uint32_t aphase = 0;
    for (long long t = blockIdx.x; t < items; t += gridDim.x) {
      int iw, ih, d0, seg, n;
      item_coords(t, iw, ih, d0, seg, n);
      const int h = ih * FTH + hh, w = iw * FWO + wi;
      const bool store = wi < FWO && h < H && w < W;
      for (int j = 0; j < seg; ++j) {
        mbar_wait(tfull(as), aphase);
        tc_fence_after();
        if (diag & 1) {
          tc_fence_before();
          mbar_arrive(tempty(as));
          if (++as == NACC) { as = 0; aphase ^= 1u; }
          continue;
        }
        float v[32], v1[32], v2[32];
        const uint32_t tacc = tmem_base + (uint32_t)(as * NF) + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(tacc, v);
        tmem_ld_32x32(tacc + NPAD, v1);
        tmem_ld_32x32(tacc + 2 * NPAD, v2);
        tc_fence_before();
        mbar_arrive(tempty(as));                                   // the accumulator is in registers: the next tile may overwrite it
#pragma unroll
        for (int c = 0; c < 32; ++c)                               // y[j] = P_0[j] + P_1[j + 1] + P_2[j + 2] (lanes j + 1, j + 2 of the same tile row)
          v[c] += __shfl_down_sync(0xffffffffu, v1[c], 1) + __shfl_down_sync(0xffffffffu, v2[c], 2);
        if (store && !(diag & 4)) {
          float *dst = y + ((((long long)n * D + d0 + j) * H + h) * W + w) * CO;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c < CO) *reinterpret_cast<float4 *>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
        if (++as == NACC) { as = 0; aphase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Probe (tests only) for the weight-gradient formulation.  MN-major TF32 operands are only accepted in the "128-byte swizzle with
// 32-byte atoms" layout (rows of 32 mn-elements = 128 bytes, 32-byte chunks XORed with row % 4; the no-swizzle MN-major form gives
// wrong results -- tried).  Stored as [voxel row][32 channels], a slab of 32 channels needs a leading-dimension offset to reach
// the next 32 mn-elements: with LBO = 128 bytes = ONE ROW the four slabs of an M = 128 operand are the same rows shifted by 0..3
// voxels -- exactly the kw taps of a convolution -- and a (kd, kh) tap is a start-address offset of whole rows.  The probe checks
// that the hardware accepts overlapping slabs and start addresses that are not aligned to the 512-byte swizzle pattern:
//   D[j * 32 + c][n] = sum_{k < 8} X[r0 + j + k][c] * Y[k][n]      X [24 rows][32], Y [8][32], D [128][32]
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swz32(uint32_t byte_off) { return byte_off ^ (((byte_off >> 7) & 3u) << 5); }

static __global__ void __launch_bounds__(128) mn_sw32_probe_kernel(const float *__restrict__ X, const float *__restrict__ Y, float *__restrict__ Dout, int r0)
{
  __shared__ __align__(1024) uint8_t sX[24 * 128];
  __shared__ __align__(1024) uint8_t sY[8 * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 24 * 32; i += 128) *reinterpret_cast<float *>(sX + swz32((uint32_t)i * 4u)) = to_tf32(X[i]);
  for (int i = t; i < 8 * 32; i += 128) *reinterpret_cast<float *>(sY + swz32((uint32_t)i * 4u)) = to_tf32(Y[i]);
  if (t == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (t == 0) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [](uint32_t addr, uint32_t lbo) {   // layout type 1 = SWIZZLE_128B_BASE32B, 4-row groups 512 bytes apart
      return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
    };
    umma_tf32(tm, desc(smem_u32(sX) + (uint32_t)r0 * 128u, 128), desc(smem_u32(sY), 128), idesc, 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  float v[32];
  tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) Dout[(size_t)t * 32 + n] = v[n];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][kd][kh][kw] = sum over voxels of dy[v][co] * x[v + (kd-1, kh-1, kw-1)][ci].
//
// The reduction index is the voxel, so both operands are MN-major (channels contiguous) -- for TF32 the tensor core takes that only in
// the 128-byte swizzle / 32-byte atom layout, i.e. rows of 32 channels.  TMA pads for free: a box of 32 channels over a tensor with
// 24 zero-fills channels 24..31 while it writes 128-byte rows.  One tile = 1 x 16 x 8 voxels; ONE TMA box brings its 3 x 18 x 10 halo of
// x as 540 rows [voxel][32 ch], another the 128 rows of dy.  Per (kd, kh) and tile row h, ONE MMA (128 x 32 x 8) accumulates
//     D_{kd,kh}[kw * 32 + ci][co] += sum_{w < 8} x[halo(kd, h + kh, w + kw)][ci] * dy[(h, w)][co]
// because the "next 32 mn-elements" of the A operand are reached through a leading-dimension offset of ONE ROW (128 bytes): the four
// slabs of the M = 128 operand are the same eight voxel rows shifted by kw = 0, 1, 2 (, 3: unused) -- the three kw taps cost nothing.
// 9 accumulators of 32 columns stay in TMEM for the CTA's whole (persistent) run; at the end every CTA writes its partial
// [9][128][32] to a workspace and a second kernel sums the CTAs.  cuDNN's kernel for the 24 -> 24 full-resolution layer takes 9.8 ms.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWgXBytes = 3 * HH * HW * 128;                     // 69120: halo tile rows of 128 bytes
constexpr int kWgDyBytes = TH * TW * 128;                        // 16384
constexpr int kWgStageBytes = (kWgXBytes + kWgDyBytes + 1023) / 1024 * 1024;   // 86016
constexpr int kWgSmemBytes = 2 * kWgStageBytes + 1024 + 128;
static_assert(kWgXBytes % 1024 == 512 || kWgXBytes % 512 == 0, "dy tile must start on a 512-byte swizzle pattern boundary");

// MN-major, 128B swizzle / 32B atoms: K rows of 128 bytes (32 channels), SBO = 4 rows; LBO = distance between the 32-wide slabs of the
// M / N extent: one row (128 B) for the x operand -- slab kw starts one voxel further, the slabs overlap -- and one tile row of dy
// (TW rows = 1024 B) for the dy operand, whose slabs are consecutive tile rows.
__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr, uint32_t lbo_bytes = 128)
{
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

// part [gridDim.x][9][128][32]
static __global__ void __launch_bounds__(kThreadsConv, 1)
conv3d_k3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, float *__restrict__ part, int N, int D,
                       int H, int W)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + 2 * kWgStageBytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (2 + s); };
  const uint32_t done = bars + 32, tmem_slot = bars + 64;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for the compiler
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int th = (H + TH - 1) / TH, tw = (W + TW - 1) / TW;
  const long long tiles = (long long)N * D * th * tw;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int iw = (int)(t % tw), ih = (int)((t / tw) % th), d = (int)((t / ((long long)tw * th)) % D), n = (int)(t / ((long long)tw * th * D));
        mbar_wait(empty(stage), phase ^ 1u);
        mbar_expect_tx(full(stage), kWgXBytes + kWgDyBytes);
        const uint32_t sx = base + stage * kWgStageBytes;
        tma_load_5d(sx, &tmX, full(stage), 0, iw * TW - 1, ih * TH - 1, d - 1, n);
        tma_load_5d(sx + kWgXBytes, &tmDy, full(stage), 0, iw * TW, ih * TH, d, n);
        if (++stage == 2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    {
      // The whole warp walks the loop, one elected lane issues (tcgemm::elect_one: operands stay in uniform registers).
      // M = 128 (kw * 32 + ci), N = 32 * (number of dy rows paired with this x row), both operands MN-major (bits 15 / 16), TF32 in, fp32 out
      constexpr uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0, later = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t sx = base + stage * kWgStageBytes;
        const uint64_t dx0 = desc_mn_sw32(sx), dy0 = desc_mn_sw32(sx + kWgXBytes, TW * 128);
        if (elect_one()) {
        // x row r of plane kd meets dy rows hh = r - kh (kh = 0..2, 0 <= hh < TH) in ONE MMA: the dy rows are consecutive N slabs, their
        // products land in accumulator kd at column block 2 - kh.  Row 2 goes first: it is the first to touch all three blocks at once,
        // so on the first tile it alone carries accumulate = 0.
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
          for (int i = 0; i < HH; ++i) {
            const int r = i == 0 ? 2 : (i <= 2 ? i - 1 : i);
            const int hh_lo = r - 2 > 0 ? r - 2 : 0, hh_hi = r < TH - 1 ? r : TH - 1, cnt = hh_hi - hh_lo + 1, kh_max = r - hh_lo;
            const uint32_t xrow16 = (uint32_t)(((kd * HH + r) * HW) * (128 / 16));
            umma_tf32(tmem_u + (uint32_t)(kd * 96 + (2 - kh_max) * 32), dx0 + xrow16, dy0 + (uint32_t)(hh_lo * TW * (128 / 16)),
                      idesc0 | ((uint32_t)((32 * cnt) >> 3) << 17), (i | later) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty(stage));
        }
        __syncwarp();
        later = 1u;
        if (++stage == 2) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    // after the last MMA: this CTA's partial sums, accumulator a rows m = kw * 32 + ci, columns co
    const int q = warp & 3;
    mbar_wait(done, 0);
    tc_fence_after();
    float *dst = part + ((long long)blockIdx.x * 9 * 128 + q * 32 + lane) * 32;
    const bool any = (long long)blockIdx.x < tiles;
#pragma unroll 1
    for (int a = 0; a < 9; ++a) {
      float v[32];
      tmem_ld_32x32(tmem_base + (uint32_t)((a / 3) * 96 + (2 - a % 3) * 32) + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4 *>(dst + (long long)a * 128 * 32 + c) = any ? make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// dw [CO][CI][3][3][3] = sum over CTAs of part[cta][kd * 3 + kh][kw * 32 + ci][co]; 256 threads per 32 outputs (8 warps split the CTAs)
static __global__ void __launch_bounds__(256) conv3d_k3_wgrad_finalize_kernel(const float *__restrict__ part, int ctas, int CI, int CO, float *__restrict__ dw)
{
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int o = blockIdx.x * 32 + lane, total = CO * CI * 27;
  double s = 0.0;
  if (o < total) {
    const int tap = o % 27, ci = (o / 27) % CI, co = o / (27 * CI);
    const int a = tap / 3, kw = tap % 3;
    const long long idx = ((long long)a * 128 + kw * 32 + ci) * 32 + co;
    for (int b = w; b < ctas; b += 8) s += part[(long long)b * 9 * 128 * 32 + idx];
  }
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && o < total) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][lane];
    dw[o] = (float)s;
  }
}

}  // namespace convtc
