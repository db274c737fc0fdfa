// conv3d_gen_kernels.cuh -- 3x3x3 convolutions of channels-last (NDHWC) fp32 volumes with ANY channel counts, stride 1 or 2, on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands by TMA): forward, gradient with respect to the
// input, gradient with respect to the weights.  These are the convolutions of the AttnFPN backbone that conv3d_tc_kernels.cuh (24 -> 24
// only: all 27 taps' weights resident in shared memory) does not cover: EncoderCnnBlock stages 1-5 (24 -> 48 -> ... -> 768 channels, the
// first convolution of each stage with stride 2; transoar/models/backbones/encoder_blocks.py:28-46 via attn_fpn.py:170-182) and the FPN's
// 3x3x3 output convolutions with bias (attn_fpn.py:65-74: 96 / 192 / 384 / 384 -> 384).  The reference runs them through cuDNN.
//
// Implicit GEMM, no im2col copy anywhere:
//
//   forward / input gradient, per-tap kernel (conv_kmajor_kernel) -- the coarse pyramid levels and stride 2
//     D[v, n] = sum over steps s, reduction channels r of  A_s[v + delta_s, r] * B[tap_s][n][r]      (+ bias[n])
//     A tile of 128 output voxels is a BW x BH x BD box of the volume.  One K-step = one tap and one chunk of 32 reduction channels:
//     ONE 5-D TMA box (32 channels, BW, BH, BD, 1) at the tile origin shifted by the tap brings the 128 x 32 A operand (rows of 128 bytes,
//     128-byte swizzle: exactly the K-major operand form of tc_gemm_kernels.cuh); voxels outside the volume and channels past the tensor's
//     count are zero-filled by the copy engine -- that is the convolution's zero padding and the channel padding.  The B operand comes
//     from the tap-major weight tensor [27][CO][CI] through a 3-D tensor map: K-major rows for the forward (reduction = ci),
//     MN-major 32 x 32 slabs for the input gradient (reduction = co, n = ci), so no transposed / re-laid-out weight copy exists.
//     Stride 2 without element strides: the eight parity classes of a volume (even / odd index per axis) are eight plain 5-D tensors with
//     doubled strides.  Forward: tap k reads class (k + 1) % 2 at index o + (k == 0 ? -1 : 0) per axis.  Input gradient: the inputs of
//     parity class p receive dx[2 j + p] = sum over the taps with k = 1 (p = 0) or k in {0, 2} (p = 1) per axis of dy[j + (k == 0)] W[k]:
//     eight stride-1 problems on the dy grid with 1 / 2 / 4 / 8 taps, each stored through the tensor map of ITS class of dx -- or, for narrow
//     layers, ONE problem whose N extent is the eight classes side by side (fold_cip: 8 wide K-steps instead of 27 narrow ones).
//     Few-tile layers split the (tap, chunk) loop over the idle SMs; the copy engine adds the partial tiles (cp.reduce.async.bulk.tensor).
//     The epilogue writes with 5-D bulk tensor stores (one per 32-voxel x 32-channel chunk), which also clip ragged volumes / channels.
//
//   forward / input gradient, halo kernel (conv_halo_kernel) -- stride 1 on volumes with >= 16 rows: see the comment at the kernel
//   weight gradient (conv_wgrad_kernel): halo tiles, kw taps as overlapping MN-major slabs: see the comment at the kernel
//
// Roles, barriers, rings and TMEM double buffering are those of tc_gemm_kernels.cuh.  What bounds a NARROW implicit GEMM (N <= 128) is the
// single issuing thread (a dependent instruction chain, ~5 clocks per instruction), not the tensor pipe (128 x N x 8 takes max(32 + N/4, N/2)
// clocks in every operand layout) and not L2: profiles/r02_experiments.md section 8a.  Hence unrolled constant taps, tile pairs, class folding.
#pragma once

#include "tc_gemm_kernels.cuh"

namespace convgen {

using namespace tcgemm;

constexpr int kMaxClasses = 8, kMaxSteps = 27;

struct Step {
  signed char amap, dw, dh, dd;   // which A tensor map (parity class), shift of the box origin in (that map's) voxels
  int tap;                        // row block of the weight tensor: kd * 9 + kh * 3 + kw
};

struct Problem {
  int batch, tw, th, td;          // tile grid (per class)
  int BW, BH, BD;                 // tile box, BW * BH * BD == 128, BW <= 32
  int qh, qd;                     // the 32 rows of one TMEM lane quarter as a box: BW x qh x qd
  int N;                          // columns of D (output channels of this launch)
  int chunks;                     // ceil(reduction channels / 32)
  int nclass;
  int fold_cip;                   // > 0: class-folded stride-2 input gradient: column block c * fold_cip .. belongs to parity class c of dx (see conv3d_gen_dgrad_s2_folded)
  int ksplit;                     // > 1: the (tap, chunk) K-steps of a tile are split over ksplit work items that ADD into a zero-filled D (few-tile layers)
  int nsteps[kMaxClasses];
  Step steps[kMaxClasses][kMaxSteps];
  CUtensorMap tmA[kMaxClasses];
  CUtensorMap tmD[kMaxClasses];
  CUtensorMap tmB;
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, int c4)
{
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// work item w -> (class, sample, tile origin, column tile); column tiles fastest so that neighbouring CTAs share the A boxes in L2
struct Item { int cls, n, w0, h0, d0, nt, ks; };
__device__ __forceinline__ Item decode(const Problem &p, long long w, int n_tiles)
{
  Item it;
  it.ks = (int)(w % p.ksplit); w /= p.ksplit;
  it.nt = (int)(w % n_tiles); w /= n_tiles;
  it.w0 = (int)(w % p.tw) * p.BW; w /= p.tw;
  it.h0 = (int)(w % p.th) * p.BH; w /= p.th;
  it.d0 = (int)(w % p.td) * p.BD; w /= p.td;
  it.n = (int)(w % p.batch);
  it.cls = (int)(w / p.batch);
  return it;
}

template <int BN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
conv_kmajor_kernel(const __grid_constant__ Problem p, const float *__restrict__ bias)
{
  using C = Cfg<BN, 1>;
  static_assert(BN % 32 == 0 && BN <= 256, "column tile");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bars = epi_base + C::EPI_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int n_tiles = (p.N + BN - 1) / BN;
  const long long work = (long long)p.nclass * p.batch * p.td * p.th * p.tw * n_tiles * p.ksplit;
  // K-steps kb = s * chunks + c of a tile; split ks takes [ks * per, (ks + 1) * per)
  auto k_range = [&](const Item &it, int &kb0, int &kb1) {
    const int total = p.nsteps[it.cls] * p.chunks, per = (total + p.ksplit - 1) / p.ksplit;
    kb0 = it.ks * per; kb1 = min(total, kb0 + per);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = blockIdx.x; w < work; w += gridDim.x) {
        const Item it = decode(p, w, n_tiles);
        const int n0 = it.nt * BN;
        int kb0, kb1;
        k_range(it, kb0, kb1);
        {
          int sidx = kb0 / p.chunks, c = kb0 % p.chunks;            // one division per work item, not per K-step: this thread's loop sets the pace
          Step st = p.steps[it.cls][sidx];
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty(stage), phase ^ 1u);
            mbar_expect_tx(full(stage), C::STAGE_BYTES);
            const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
            tma_load_5d(sa, &p.tmA[st.amap], full(stage), c * 32, it.w0 + st.dw, it.h0 + st.dh, it.d0 + st.dd, it.n);
            if (!B_MN) {
              tma_load_3d(sb, &p.tmB, full(stage), c * 32, n0, st.tap);                           // BN rows (n) x 32 reduction channels
            } else {
#pragma unroll
              for (int i = 0; i < BN / 32; ++i) tma_load_3d(sb + i * kSlabBytes, &p.tmB, full(stage), n0 + 32 * i, c * 32, st.tap);   // 32 reduction rows x 32 n
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            if (++c == p.chunks) { c = 0; ++sidx; if (kb + 1 < kb1) st = p.steps[it.cls][sidx]; }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc<BN, false, B_MN, 1, float>();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      const Item it = decode(p, w, n_tiles);
      int kb0, kb1;
      k_range(it, kb0, kb1);
      mbar_wait(tempty(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_u + (uint32_t)(as * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = smem_desc<false>(sa + k * kstep_bytes<false>()), db = smem_desc<B_MN>(sb + k * kstep_bytes<B_MN>());
            umma_tf32(acc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty(stage));
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(tfull(as));
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    int as = 0, tma_buf = 0;
    uint32_t aphase = 0;
    // rows 32 q .. 32 q + 31 of the tile (w fastest, then h, then d) are the box BW x qh x qd starting here
    const int r0 = 32 * q;
    const int qh0 = (r0 / p.BW) % p.BH, qd0 = r0 / (p.BW * p.BH);
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      const Item it = decode(p, w, n_tiles);
      int kb0, kb1;
      k_range(it, kb0, kb1);
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 32 * ((warp - 2) >> 2); c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
        float v[32];
        tmem_ld_32x32(tmem_base + (uint32_t)(as * BN + c0) + ((uint32_t)(q * 32) << 16), v);
        const int n0 = it.nt * BN + c0;
        if (n0 >= p.N || kb0 >= kb1) continue;                    // warp-uniform; an empty K split (a class with fewer taps than splits) adds nothing
        if (bias != nullptr && it.ks == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (n0 + j < p.N) ? __ldg(bias + n0 + j) : 0.f;
        }
        const uint32_t buf = epi_base + (uint32_t)(warp - 2) * C::EPI_WARP_BYTES + (uint32_t)(tma_buf * 4096);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + (uint32_t)(lane * 128 + ((i ^ (lane & 7)) << 4))), "f"(v[4 * i]),
                       "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          const int dcls = p.fold_cip > 0 ? n0 / p.fold_cip : it.cls, dch = p.fold_cip > 0 ? n0 % p.fold_cip : n0;
          if (p.ksplit > 1)                                        // partial sums of the K splits are combined by the copy engine (fp32 add in L2)
            asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                         ::"l"(&p.tmD[dcls]), "r"(buf), "r"(dch), "r"(it.w0), "r"(it.h0 + qh0), "r"(it.d0 + qd0), "r"(it.n) : "memory");
          else
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                         ::"l"(&p.tmD[dcls]), "r"(buf), "r"(dch), "r"(it.w0), "r"(it.h0 + qh0), "r"(it.d0 + qd0), "r"(it.n) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        tma_buf ^= 1;
      }
      tc_fence_before();
      mbar_arrive(tempty(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Forward / input gradient with the A operand read out of a HALO tile (conv_halo_kernel) -- the fast path for volumes with >= ~16 rows.
//
// conv_kmajor_kernel above fetches one 128 x 32 box per tap: every input voxel crosses L2 -> SM 27 times, and the L2 fabric (~45 B/clk per
// SM measured here) is what bounds it.  A K-major, 128-byte-swizzled MMA operand may start at ANY row of a tile in shared memory and its
// 8-row groups may be any number of rows apart (the swizzle is a function of the absolute shared-memory address, for the copy engine and
// for the tensor core alike: tools/probe_kshift.py, profiles/r02_experiments.md).  So a tile is 8 (w) x 16 (h) x 1 (d) output voxels = 16
// groups of 8 rows, ONE TMA box brings the (8 + 2) x (16 + 2) halo of one input plane (a "plane stage": 180 rows of 128 bytes), and the nine
// (kh, kw) taps of that plane are nine descriptors into it: start row kh * 10 + kw, group stride 10 rows.  The input now crosses L2 -> SM
// 4.2 times (three halo planes per output plane); the weights stream through their own ring, one BN x 32 block per (tap, chunk).
// Stride 2 forward: a plane stage is the four (h, w) parity-class boxes of the plane's d class.  Stride 2 input gradient: per parity class
// of dx a 9 x 17 box of dy, the taps of that class at row shifts {0, 1} x {0, 9}.  Everything is table-driven (HProblem).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHMaxBoxes = 4, kHMaxPlanes = 3, kHMaxTaps = 9;
struct HBox {
  int cls_hw;        // A tensor map = tmA[plane.cls_d * 4 + cls_hw]
  int ow, oh;        // box origin = (w0 + ow, h0 + oh)
  int lw;            // rows per line of the box = group stride of the operand
  int off, bytes;    // byte offset in the A stage, box bytes
};
struct HTap {
  int a_off, sbo;    // A operand: byte offset in the A stage (box + row shift), group stride in bytes
  int wtap0, slot;   // the weight block: taps [wtap0, wtap0 + tps) arrive as ONE box, this tap is slot `slot` of it
  int newgrp;        // 1: this tap starts a new weight box
};
struct HPlane { int cls_d, od, ntaps; HTap taps[kHMaxTaps]; };
struct HClass { int nplanes; HPlane planes[kHMaxPlanes]; };
struct HProblem {
  int batch, tw, th, td;          // tile grid: 8 x 16 x 1 voxels per tile
  int N, chunks, nclass, nbox;
  int a_bytes, a_stage_bytes, a_stages, b_stages;
  int pair;                       // 1: tile pairs (T = 2 instantiation): 16 x 16 voxel work items, halo box 18 x 18
  int dbg;                        // timing experiments (conv3d_gen_set_path bits 3-5): 1 = no MMAs, 2 = no A loads, 4 = no B loads, 8 = no stores
  int tps;                        // taps per weight box (9, 3 or 1): amortises the barrier round trip of a stage over 4 * tps MMAs
  HBox boxes[kHMaxBoxes];
  HClass cls[kMaxClasses];
  CUtensorMap tmA[kMaxClasses];
  CUtensorMap tmD[kMaxClasses];   // box (32 channels, 8, 4, 1, 1): one TMEM lane quarter
  CUtensorMap tmB;
};
constexpr int kHEpiBytes = kEpiWarps * 4096;
constexpr int kHMaxAStages = 4, kHMaxBStages = 16;   // the weight blocks are small (BN x 128 bytes): a deep ring keeps enough bytes in flight

__device__ __forceinline__ uint64_t desc_k_halo(uint32_t addr, uint32_t group_stride_bytes)
{
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(group_stride_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// CH = independent accumulation chains: consecutive taps go to CH different accumulators (summed by the epilogue), so that consecutive
// small-N MMAs do not wait for each other's accumulator
// REG: 0 = table-driven taps; 1 = stride-1 forward, 2 = stride-1 input gradient, 3 = stride-2 forward: the nine taps of a plane are compile-time
// constants and their 36 MMAs are issued from one unrolled block with constant descriptor increments -- the issuing warp executes a
// dependent instruction chain at ~5 clocks per instruction, and a table-driven tap costs it ~350 clocks (profiles/r02_experiments.md)
template <int BN> __host__ __device__ constexpr int halo_tps() { return BN <= 32 ? 9 : BN <= 128 ? 3 : 1; }

// T = 2 (stride 1 only): a work item is TWO tiles side by side in w (16 x 16 voxels, halo box 18 x 18) that share every weight block --
// the weights are what a narrow layer pulls most of from L2 (27 x BN x 128 bytes per 128 voxels and chunk against 23 KB of halo per plane)
template <int BN, bool B_MN, int CH, int REG, int T = 1>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ HProblem p, const float *__restrict__ bias)
{
  constexpr int B_TAP_BYTES = BN * 128;
  const int B_BYTES = p.tps * B_TAP_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = base + p.a_stages * p.a_stage_bytes;
  const uint32_t epi_base = b_base + p.b_stages * B_BYTES;
  const uint32_t bars = epi_base + kHEpiBytes;
  auto afull = [&](int s) { return bars + 8u * s; };
  auto aempty = [&](int s) { return bars + 8u * (kHMaxAStages + s); };
  auto bfull = [&](int s) { return bars + 8u * (2 * kHMaxAStages + s); };
  auto bempty = [&](int s) { return bars + 8u * (2 * kHMaxAStages + kHMaxBStages + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * kHMaxAStages + 2 * kHMaxBStages + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * kHMaxAStages + 2 * kHMaxBStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kHMaxAStages + 2 * kHMaxBStages + 4);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  static_assert(T == 1 || (REG != 0 && CH == 1), "tile pairs: regular taps, one chain");
  constexpr int ACC_COLS = T * CH * BN;                               // one accumulator set: T tiles x CH chains of BN columns
  constexpr int LW = 8 * T + 2;                                       // rows per line of the stride-1 halo box
  static_assert(2 * ACC_COLS <= 512, "TMEM columns");
  constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kHMaxAStages; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < kHMaxBStages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int n_tiles = (p.N + BN - 1) / BN;
  const long long work = (long long)p.nclass * p.batch * p.td * p.th * p.tw * n_tiles;
  // work item -> (class, sample, plane, tile row, tile column, column tile); column tiles fastest (they share the A boxes in L2)
  auto decode_item = [&](long long w, int &cls, int &n, int &d0, int &h0, int &w0, int &nt) {
    nt = (int)(w % n_tiles); w /= n_tiles;
    w0 = (int)(w % p.tw) * (8 * T); w /= p.tw;
    h0 = (int)(w % p.th) * 16; w /= p.th;
    d0 = (int)(w % p.td); w /= p.td;
    n = (int)(w % p.batch);
    cls = (int)(w / p.batch);
  };

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long w = blockIdx.x; w < work; w += gridDim.x) {
        int cls, n, d0, h0, w0, nt;
        decode_item(w, cls, n, d0, h0, w0, nt);
        const HClass &hc = p.cls[cls];
        const int n0 = nt * BN;
        for (int c = 0; c < p.chunks; ++c) {
          for (int pl = 0; pl < hc.nplanes; ++pl) {
            const HPlane &hp = hc.planes[pl];
            mbar_wait(aempty(as), aph ^ 1u);
            const uint32_t sa = base + as * p.a_stage_bytes;
            if (p.dbg & 2) {
              mbar_arrive(afull(as));
            } else {
              mbar_expect_tx(afull(as), (uint32_t)p.a_bytes);
              for (int b = 0; b < p.nbox; ++b) {
                const HBox &bx = p.boxes[b];
                tma_load_5d(sa + bx.off, &p.tmA[hp.cls_d * 4 + bx.cls_hw], afull(as), c * 32, w0 + bx.ow, h0 + bx.oh, d0 + hp.od, n);
              }
            }
            if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            if (REG != 0) {
              constexpr int TPS = halo_tps<BN>();
#pragma unroll
              for (int g = 0; g < 9 / TPS; ++g) {
                // forward: taps pl * 9 + g * TPS ...; input gradient: plane pl holds kd = 2 - pl and its box rows run against the tap order
                const int tap = (REG == 1 || REG == 3) ? pl * 9 + g * TPS : (TPS == 9 ? (2 - pl) * 9 : TPS == 3 ? (2 - pl) * 9 + (2 - g) * 3 : 26 - (pl * 9 + g));
                mbar_wait(bempty(bs), bph ^ 1u);
                if (p.dbg & 4) { mbar_arrive(bfull(bs)); if (++bs == p.b_stages) { bs = 0; bph ^= 1u; } continue; }
                mbar_expect_tx(bfull(bs), (uint32_t)B_BYTES);
                const uint32_t sb = b_base + bs * B_BYTES;
                if (!B_MN) {
                  tma_load_3d(sb, &p.tmB, bfull(bs), c * 32, n0, tap);
                } else {
#pragma unroll
                  for (int j = 0; j < TPS; ++j)
#pragma unroll
                    for (int i = 0; i < BN / 32; ++i) tma_load_3d(sb + j * B_TAP_BYTES + i * kSlabBytes, &p.tmB, bfull(bs), n0 + 32 * i, c * 32, tap + j);
                }
                if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
              }
              continue;
            }
            for (int t = 0; t < hp.ntaps; ++t) {
              if (!hp.taps[t].newgrp) continue;
              mbar_wait(bempty(bs), bph ^ 1u);
              if (p.dbg & 4) { mbar_arrive(bfull(bs)); if (++bs == p.b_stages) { bs = 0; bph ^= 1u; } continue; }
              mbar_expect_tx(bfull(bs), (uint32_t)B_BYTES);
              const uint32_t sb = b_base + bs * B_BYTES;
              const int tap = hp.taps[t].wtap0;
              if (!B_MN) {
                tma_load_3d(sb, &p.tmB, bfull(bs), c * 32, n0, tap);                     // box (32 ci, BN co, tps taps)
              } else {
                for (int j = 0; j < p.tps; ++j)
#pragma unroll
                  for (int i = 0; i < BN / 32; ++i) tma_load_3d(sb + j * B_TAP_BYTES + i * kSlabBytes, &p.tmB, bfull(bs), n0 + 32 * i, c * 32, tap + j);
              }
              if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = instr_desc<BN, false, B_MN, 1, float>();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int as = 0, bs = 0, acs = 0;
    uint32_t aph = 0, bph = 0, acph = 0;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      int cls, n, d0, h0, w0, nt;
      decode_item(w, cls, n, d0, h0, w0, nt);
      const HClass &hc = p.cls[cls];
      mbar_wait(tempty(acs), acph ^ 1u);
      tc_fence_after();
      const uint32_t acc0 = tmem_u + (uint32_t)(acs * ACC_COLS);
      int tapno = 0;
      for (int c = 0; c < p.chunks; ++c) {
        for (int pl = 0; pl < hc.nplanes; ++pl) {
          const HPlane &hp = hc.planes[pl];
          mbar_wait(afull(as), aph);
          tc_fence_after();
          const uint32_t sa = base + as * p.a_stage_bytes;
          if (REG != 0) {
            constexpr int TPS = halo_tps<BN>();
            const uint64_t da_plane = desc_k_halo(sa, LW * 128);
            const uint64_t da_s2_9 = desc_k_halo(sa, 9 * 128), da_s2_8 = desc_k_halo(sa, 8 * 128);      // REG 3: odd-w / even-w class boxes
            const uint32_t first_plane = (c == 0 && pl == 0) ? 1u : 0u;
#pragma unroll
            for (int g = 0; g < 9 / TPS; ++g) {
              mbar_wait(bfull(bs), bph);
              tc_fence_after();
              const uint64_t db_stage = smem_desc<B_MN>(b_base + bs * B_BYTES);
              if (elect_one()) {
                if (!(p.dbg & 1)) {
#pragma unroll
                  for (int j = 0; j < TPS; ++j) {
                    const int t = g * TPS + j;                                 // box rows (jh, jw) = (t / 3, t % 3): compile-time after unrolling
                    const int slot = (REG == 1 || REG == 3) ? j : TPS - 1 - j;
                    const uint32_t later = (t >= CH) ? 1u : (first_plane ^ 1u);
                    if (REG == 3) {
                      // stride-2 forward: tap (kh, kw) = (t / 3, t % 3) reads the (h, w) parity-class box (kh != 1, kw != 1) of the plane stage --
                      // boxes (odd, odd) 9 x 17 rows, (odd, even) 8 x 17, (even, odd) 9 x 16, (even, even) 8 x 16 at 1 KB-aligned offsets --
                      // one line further for kh = 2, one row further for kw = 2
                      constexpr int kOff[4] = {0, 20480, 37888, 56320};
                      const int kh = t / 3, kw = t % 3;
                      const int box = (kh != 1 ? 0 : 2) + (kw != 1 ? 0 : 1), lw = kw != 1 ? 9 : 8;
                      const uint32_t a_off = (uint32_t)(kOff[box] + ((kh == 2 ? lw : 0) + (kw == 2 ? 1 : 0)) * 128);
                      const uint32_t acc = acc0 + (uint32_t)((t % CH) * BN);
#pragma unroll
                      for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_tf32(acc, (kw != 1 ? da_s2_9 : da_s2_8) + (uint64_t)((a_off + k * 32) >> 4),
                                  db_stage + (uint64_t)((slot * B_TAP_BYTES + k * (int)kstep_bytes<B_MN>()) >> 4), idesc, later | (uint32_t)(k > 0));
                    } else {
#pragma unroll
                    for (int hf = 0; hf < T; ++hf) {                           // the tile pair: same weights, A rows 8 voxels further
                      const uint32_t a_off = (uint32_t)(((t / 3) * LW + (t % 3) + 8 * hf) * 128);
                      const uint32_t acc = acc0 + (uint32_t)((hf * CH + t % CH) * BN);
#pragma unroll
                      for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_tf32(acc, da_plane + (uint64_t)((a_off + k * 32) >> 4), db_stage + (uint64_t)((slot * B_TAP_BYTES + k * (int)kstep_bytes<B_MN>()) >> 4),
                                  idesc, later | (uint32_t)(k > 0));
                    }
                    }
                  }
                }
                umma_commit(bempty(bs));
              }
              __syncwarp();
              if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
            }
            tapno += 9;
          } else
          for (int t = 0; t < hp.ntaps; ++t) {
            const HTap &tp = hp.taps[t];
            if (tp.newgrp) {
              mbar_wait(bfull(bs), bph);
              tc_fence_after();
            }
            const uint32_t sb = b_base + bs * B_BYTES + (uint32_t)(tp.slot * B_TAP_BYTES);
            const uint64_t da0 = desc_k_halo(sa + (uint32_t)tp.a_off, (uint32_t)tp.sbo);
            const uint32_t acc = acc0 + (uint32_t)((tapno % CH) * BN);
            const uint32_t later = tapno >= CH ? 1u : 0u;                 // the first tap of each chain overwrites
            const bool last_of_group = t + 1 == hp.ntaps || hp.taps[t + 1].newgrp != 0;
            if (elect_one()) {
              if (!(p.dbg & 1)) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  umma_tf32(acc, da0 + (uint64_t)(k * 2), smem_desc<B_MN>(sb + k * kstep_bytes<B_MN>()), idesc, later | (uint32_t)(k > 0));
                }
              }
              if (last_of_group) umma_commit(bempty(bs));
            }
            __syncwarp();
            ++tapno;
            if (last_of_group) { if (++bs == p.b_stages) { bs = 0; bph ^= 1u; } }
          }
          if (elect_one()) umma_commit(aempty(as));
          __syncwarp();
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(tfull(acs));
      __syncwarp();
      if (++acs == 2) { acs = 0; acph ^= 1u; }
    }
  } else {
    const int q = warp & 3;
    int acs = 0;
    uint32_t acph = 0;
    const uint32_t buf = epi_base + (uint32_t)(warp - 2) * 4096u;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      int cls, n, d0, h0, w0, nt;
      decode_item(w, cls, n, d0, h0, w0, nt);
      int ntaps_total = 0;
      for (int pl = 0; pl < p.cls[cls].nplanes; ++pl) ntaps_total += p.cls[cls].planes[pl].ntaps;
      ntaps_total *= p.chunks;
      mbar_wait(tfull(acs), acph);
      tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < T; ++hf)
#pragma unroll 1
      for (int c0 = 32 * ((warp - 2) >> 2); c0 < BN; c0 += 32 * (kEpiWarps / 4)) {
        float v[32];
        tmem_ld_32x32(tmem_base + (uint32_t)(acs * ACC_COLS + hf * CH * BN + c0) + ((uint32_t)(q * 32) << 16), v);
        const int n0 = nt * BN + c0;
        if (n0 >= p.N) continue;                                  // warp-uniform
        if (CH > 1 && ntaps_total > 1) {                          // (a class with a single tap never touched the second chain)
#pragma unroll
          for (int ch = 1; ch < CH; ++ch) {
            if (ch < ntaps_total) {
              float u[32];
              tmem_ld_32x32(tmem_base + (uint32_t)(acs * ACC_COLS + (hf * CH + ch) * BN + c0) + ((uint32_t)(q * 32) << 16), u);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
          }
        }
        if (bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (n0 + j < p.N) ? __ldg(bias + n0 + j) : 0.f;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous store has read the staging tile
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + (uint32_t)(lane * 128 + ((i ^ (lane & 7)) << 4))), "f"(v[4 * i]),
                       "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !(p.dbg & 8)) {
          asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                       ::"l"(&p.tmD[cls]), "r"(buf), "r"(n0), "r"(w0 + 8 * hf), "r"(h0 + 4 * q), "r"(d0), "r"(n) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(acs));
      if (++acs == 2) { acs = 0; acph ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: dW[co][tap][ci] = sum over output voxels v of dy[v][co] * x[stride * v + tap - 1][ci]
//
// The reduction index is the voxel, so both operands are MN-major (channels contiguous, 128-byte rows of 32 channels, 128-byte swizzle
// with 32-byte atoms).  That operand form is address-based: an operand may start at ANY row of a tile in shared memory, and the "next 32
// M-elements" may be reached with a leading-dimension offset of ONE ROW (tests/test_gpu_conv3d_tc.py::test_mn_major_overlapping_slab_probe).
// So a K-block is a box of 8 x BH x BD output voxels (16 lines of 8 voxels along w) whose x HALO tile arrives by one TMA box per
// parity class, and for every line ONE MMA (M = 128, N = BN, K = 8 voxels) multiplies
//     A = the x rows of that line shifted by kh lines, its four M slabs = the SAME rows shifted by 0, 1, 2 (, 3) voxels = the kw taps
//     B = the 8 dy rows of the line, N slabs = 32-channel chunks of dy (one TMA box each)
// into the accumulator of (kh): the x tile is read from L2 once per (kd, ci chunk), not once per tap.  A work item is
// (kd, ci chunk, co tile, voxel range); its accumulators (3 for stride 1, 6 for stride 2) stay in TMEM over the whole range and are then
// added into the zero-initialised gradient with red.global.add.f32 (channels-last weight layout).
// Stride 2: tap k reads parity class (k + 1) % 2 of x at output index o - (k == 0), so the taps {0, 2} of an axis are two shifts of the
// ODD class and tap 1 is the EVEN class: four (h, w)-class boxes per stage, two MMAs per (kh, line) (odd-w class: kw 0 and 2 as
// overlapping slabs; even-w class: kw 1), six accumulators.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWMaxBoxes = 4, kWMaxGroups = 6, kWLines = 16;

struct WBox {
  int cls_hw;        // tensor map = tmX[pd * 4 + cls_hw]
  int ow, oh;        // box origin = (w0 + ow, h0 + oh)
  int lw, lh;        // rows per line, lines per plane of the box
  int off;           // byte offset in the stage
};
struct WGroup {
  int box, oh;       // A rows of output line (h, d): box row ((d * lh + h + oh) * lw)
  int tap[4];        // M slab j (rows shifted by j voxels) -> kh * 3 + kw, or -1 (unused slab)
};
struct WProblem {
  int batch, tw, th, td;          // grid of K-blocks over the OUTPUT volume (dy)
  int BH, BD;                     // K-block = 8 x BH x BD voxels, BH * BD == 16
  int CI, CO, chunks;             // chunks = ceil(CI / 32)
  int BN, n_tiles;                // co tile (multiple of 32) and their number
  int splits;
  long long kb_per_split;
  int nbox, ngroups;
  int kd_cls[3], kd_off[3];       // tap kd reads d-class kd_cls at d0 + kd_off
  int x_bytes, dy_off, stage_bytes, stages;
  int dbg;                        // timing experiments (conv3d_gen_set_path bits 3-5): 1 = no MMAs, 2 = no loads, 4 = no reductions
  float *dw;                      // [CO][27][CI], zero-initialised
  WBox boxes[kWMaxBoxes];
  WGroup groups[kWMaxGroups];
  CUtensorMap tmX[kMaxClasses];   // per parity class (pd, ph, pw); stride 1: only [0]
  CUtensorMap tmDy;               // box (32 co, 8, BH, BD, 1)
};

constexpr int kThreadsW = 192;    // warp 0 producer, warp 1 MMA, warps 2-5 epilogue (one per TMEM lane quarter = one M slab)
constexpr int kWMaxStages = 4;   // ring depth is WProblem.stages: as many stages as fit (a stage is 57 - 110 KB; with two, every load latency is exposed)
constexpr int kWDyChunkBytes = 128 * 128;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes)
{
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

// Box geometry of a stage as compile-time constants (the host builds WProblem.boxes / groups from the same numbers): with GEO != 0 every
// descriptor of the 48 / 96 MMAs of a stage is the stage's base descriptor plus a constant, so the issuing thread spends ~4 instructions per
// MMA instead of ~10 (it runs a dependent chain at ~5 clocks per instruction; a 128 x 64 x 8 MMA takes 48 clocks).
template <int BH> struct WGeo {
  static constexpr int BD = 16 / BH;
  __host__ __device__ static constexpr int align1k(int b) { return (b + 1023) / 1024 * 1024; }
  // stride 2: boxes (odd h, odd w), (odd h, even w), (even h, odd w), (even h, even w)
  __host__ __device__ static constexpr int lw(int b) { return (b & 1) ? 8 : 9; }
  __host__ __device__ static constexpr int lh(int b) { return b < 2 ? BH + 1 : BH; }
  __host__ __device__ static constexpr int bytes(int b) { return lw(b) * lh(b) * BD * 128; }
  __host__ __device__ static constexpr int off(int b) { return b == 0 ? 0 : align1k(off(b - 1) + bytes(b - 1)); }
  __host__ __device__ static constexpr int gbox(int g) { return g < 4 ? (g & 1) : 2 + (g & 1); }     // groups: (kh 0: box 0, 1), (kh 2: box 0, 1), (kh 1: box 2, 3)
  __host__ __device__ static constexpr int goh(int g) { return (g == 2 || g == 3) ? 1 : 0; }
};

// BHL = log2(BH): the line -> (h, d) split is a compile-time shift, the 16 MMAs of a group are unrolled with constant multipliers.
// GEO: 0 = geometry from the tables, 1 = stride 1 (one 10 x (BH + 2) x BD box, groups = kh), 2 = stride 2 (WGeo).
template <int BHL, int GEO = 0>
__global__ void __launch_bounds__(kThreadsW, 1)
conv_wgrad_kernel(const __grid_constant__ WProblem p)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + p.stages * p.stage_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kWMaxStages + s); };
  const uint32_t tfull = bars + 8u * (2 * kWMaxStages), tempty = bars + 8u * (2 * kWMaxStages + 1);
  const uint32_t tmem_slot = bars + 8u * (2 * kWMaxStages + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 128);
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

  const long long kblocks = (long long)p.batch * p.td * p.th * p.tw;
  const int per_split = 3 * p.chunks * p.n_tiles;                       // (kd, chunk, co tile) fastest: concurrent CTAs read the same voxel range
  const long long work = (long long)per_split * p.splits;
  const int ndy = p.BN / 32;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = blockIdx.x; w < work; w += gridDim.x) {
        const int r = (int)(w % per_split), sp = (int)(w / per_split);
        const int kd = r % 3, ch = (r / 3) % p.chunks, nt = r / (3 * p.chunks);
        const long long kb0 = sp * p.kb_per_split, kb1 = min(kblocks, kb0 + p.kb_per_split);
        // block coordinates: divided out once per work item, then counted up (this thread's loop is on the critical path and a 64-bit
        // division costs it ~500 clocks)
        long long t0 = kb0;
        int iw = (int)(t0 % p.tw); t0 /= p.tw;
        int ih = (int)(t0 % p.th); t0 /= p.th;
        int id = (int)(t0 % p.td);
        int in = (int)(t0 / p.td);
        for (long long kb = kb0; kb < kb1; ++kb) {
          const int w0 = iw * 8, h0 = ih * p.BH, d0 = id * p.BD, n = in;
          if (++iw == p.tw) { iw = 0; if (++ih == p.th) { ih = 0; if (++id == p.td) { id = 0; ++in; } } }
          mbar_wait(empty(stage), phase ^ 1u);
          if (p.dbg & 2) { mbar_arrive(full(stage)); if (++stage == p.stages) { stage = 0; phase ^= 1u; } continue; }
          mbar_expect_tx(full(stage), (uint32_t)(p.x_bytes + ndy * kWDyChunkBytes));
          const uint32_t sb = base + stage * p.stage_bytes;
          for (int b = 0; b < p.nbox; ++b) {
            const WBox &bx = p.boxes[b];
            tma_load_5d(sb + bx.off, &p.tmX[p.kd_cls[kd] * 4 + bx.cls_hw], full(stage), ch * 32, w0 + bx.ow, h0 + bx.oh, d0 + p.kd_off[kd], n);
          }
          for (int i = 0; i < ndy; ++i) tma_load_5d(sb + p.dy_off + i * kWDyChunkBytes, &p.tmDy, full(stage), nt * p.BN + 32 * i, w0, h0, d0, n);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // M = 128, N = BN, both operands MN-major (bits 15 / 16), TF32 in, fp32 out
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0;
    uint32_t phase = 0, aphase = 0;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      const int sp = (int)(w / per_split);
      const long long kb0 = sp * p.kb_per_split, kb1 = min(kblocks, kb0 + p.kb_per_split);
      mbar_wait(tempty, aphase ^ 1u);
      tc_fence_after();
      for (long long kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t sb = base + stage * p.stage_bytes;
        if (elect_one()) {
          if (!(p.dbg & 1)) {
          const uint64_t db0 = desc_mn(sb + p.dy_off, kWDyChunkBytes);
          constexpr int BH = 1 << BHL;
          if (GEO == 1) {
            const uint64_t da_stage = desc_mn(sb, 128);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
              const uint32_t acc = tmem_u + (uint32_t)(g * p.BN);
#pragma unroll
              for (int line = 0; line < kWLines; ++line) {
                const int h = line & (BH - 1), d = line >> BHL;
                umma_tf32(acc, da_stage + (uint64_t)(((d * (BH + 2) + h + g) * 10 * 128) >> 4), db0 + (uint64_t)(line * 8 * 128 >> 4), idesc,
                          (kb > kb0 || line > 0) ? 1u : 0u);
              }
            }
          } else if (GEO == 2) {
            using G = WGeo<BH>;
            const uint64_t da_stage = desc_mn(sb, 128);
#pragma unroll
            for (int g = 0; g < 6; ++g) {
              const uint32_t acc = tmem_u + (uint32_t)(g * p.BN);
#pragma unroll
              for (int line = 0; line < kWLines; ++line) {
                const int h = line & (BH - 1), d = line >> BHL;
                umma_tf32(acc, da_stage + (uint64_t)((G::off(G::gbox(g)) + (d * G::lh(G::gbox(g)) + h + G::goh(g)) * G::lw(G::gbox(g)) * 128) >> 4),
                          db0 + (uint64_t)(line * 8 * 128 >> 4), idesc, (kb > kb0 || line > 0) ? 1u : 0u);
              }
            }
          } else
          for (int g = 0; g < p.ngroups; ++g) {
            const WGroup &gr = p.groups[g];
            const WBox &bx = p.boxes[gr.box];
            const uint32_t step_h = (uint32_t)(bx.lw * 128) >> 4, step_d = (uint32_t)(bx.lh * bx.lw * 128) >> 4;
            const uint64_t da0 = desc_mn(sb + bx.off + (uint32_t)(gr.oh * bx.lw * 128), 128);
            const uint32_t acc = tmem_u + (uint32_t)(g * p.BN);
#pragma unroll
            for (int line = 0; line < kWLines; ++line) {
              const int h = line & (BH - 1), d = line >> BHL;
              umma_tf32(acc, da0 + (uint64_t)(d * step_d + h * step_h), db0 + (uint64_t)(line * 8 * 128 >> 4), idesc, (kb > kb0 || line > 0) ? 1u : 0u);
            }
          }
          }
          umma_commit(empty(stage));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
      aphase ^= 1u;
    }
  } else {
    const int q = warp & 3;                                        // TMEM lane quarter = M slab
    uint32_t aphase = 0;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      const int r = (int)(w % per_split);
      const int kd = r % 3, ch = (r / 3) % p.chunks, nt = r / (3 * p.chunks);
      mbar_wait(tfull, aphase);
      tc_fence_after();
      const int ci = ch * 32 + lane;
      for (int g = 0; g < p.ngroups; ++g) {
        const int tl = p.groups[g].tap[q];
        if (tl < 0) continue;                                      // warp-uniform: unused M slab
        const int tap = kd * 9 + tl;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          float v[32];
          tmem_ld_32x32(tmem_base + (uint32_t)(g * p.BN + c0) + ((uint32_t)(q * 32) << 16), v);
          const int co0 = nt * p.BN + c0;
          float *dst = p.dw + ((long long)co0 * 27 + tap) * p.CI + ci;           // lanes = consecutive ci: coalesced reductions
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (co0 + j < p.CO && ci < p.CI && !(p.dbg & 4)) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)j * 27 * p.CI), "f"(v[j]) : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
      aphase ^= 1u;
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
// Probe (tests / experiments only): can a K-major, 128-byte-swizzled A operand start at a row that is NOT aligned to the 8-row
// (1024-byte) swizzle atom, with 8-row groups that are NOT 1024 bytes apart?  That is what reading the kw / kh taps of a convolution
// straight out of one halo tile in shared memory needs.  X [176 rows][32] is stored as TMA would store it (16-byte chunk c of the row at
// byte address a lands at chunk c ^ ((a >> 7) & 7): the swizzle is a function of the ABSOLUTE shared-memory address), then ONE
// 128 x 32 x 8 MMA reads rows row0 + (m / 8) * gs + m % 8.  mode bit 0: put (start >> 7) & 7 into the descriptor's base-offset field.
//   D[m][n] = sum_{k < 8} X[row0 + (m / 8) * gs + m % 8][k] * Y[n][k]
// ---------------------------------------------------------------------------------------------------------------
constexpr int kProbeRows = 176;
static __global__ void __launch_bounds__(128) k_sw128_probe_kernel(const float *__restrict__ X, const float *__restrict__ Y, float *__restrict__ Dout,
                                                                   int row0, int gs, int mode)
{
  __shared__ __align__(1024) uint8_t sX[kProbeRows * 128];
  __shared__ __align__(1024) uint8_t sY[32 * 128];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  auto swz = [](uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); };
  auto tf32 = [](float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); };
  for (int i = t; i < kProbeRows * 32; i += 128) *reinterpret_cast<float *>(sX + swz((uint32_t)i * 4u)) = tf32(X[i]);
  for (int i = t; i < 32 * 32; i += 128) *reinterpret_cast<float *>(sY + swz((uint32_t)i * 4u)) = tf32(Y[i]);
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
    constexpr uint32_t idesc = instr_desc<32, false, false, 1, float>();
    const uint32_t a0 = smem_u32(sX) + (uint32_t)row0 * 128u;
    uint64_t da = (uint64_t)((a0 & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((uint32_t)(gs * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
    if (mode & 1) da |= (uint64_t)((a0 >> 7) & 7u) << 49;
    umma_tf32(tm, da, smem_desc<false>(smem_u32(sY)), idesc, 0u);
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
// Probe (experiments only, tools/probe_mma_rate.py): how many clocks does ONE tcgen05.mma.kind::tf32 of shape 128 x N x 8 take when issued
// back to back, as a function of the shared-memory operand layout?  layout 0 = K-major 128-byte swizzle (rows of 128 bytes, the MMA reads 32
// bytes of each), 1 = K-major 32-byte swizzle (rows of 32 bytes: a 128 x 8 operand is 4 KB contiguous), 2 = MN-major (128-byte swizzle, 32-byte
// atoms).  The operands are zeros; `iters` MMAs accumulate into one TMEM tile; out[0] = clocks from the first issue to the commit's arrival.
// ---------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128) mma_rate_probe_kernel(int layout, int N, int iters, long long *out, int rotate)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < (64 * 1024) / 4; i += 128) reinterpret_cast<float *>(smem_raw)[i] = 0.f;
  if (t == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t a_addr = base, b_addr = base + 32 * 1024;
    uint64_t da, db;
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (layout == 0) {            // K-major SW128: SBO 1024
      da = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      db = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    } else if (layout == 1) {     // K-major SW32: rows of 32 bytes, 8-row groups 256 bytes apart
      da = (uint64_t)((a_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
      db = (uint64_t)((b_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
    } else {                      // MN-major, 32-byte atoms: slabs of 32 mn-elements, LBO = slab (32 rows x 128 B), SBO 512
      idesc |= (1u << 15) | (1u << 16);
      da = desc_mn(a_addr, 4096);
      db = desc_mn(b_addr, 4096);
    }
    const long long t0 = clock64();
    if (elect_one()) {
      // rotate > 1: consecutive MMAs read `rotate` different operand tiles (4 KB apart) instead of the same one
      for (int i = 0; i < iters; ++i) {
        const uint64_t sh = (uint64_t)((i & (rotate - 1)) * (4096 >> 4));       // rotate is a power of two (1 = always the same tile)
        umma_tf32(tm, da + sh, db + sh, idesc, i > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (t == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

}  // namespace convgen
