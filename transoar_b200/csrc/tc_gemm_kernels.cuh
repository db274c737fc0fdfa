// tc_gemm_kernels.cuh -- TF32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA) for the dense contractions of the hot path: the linear projections of MSDeformAttn
// (transoar/models/ops/modules/ms_deform_attn.py:58-61,109-140), the FFNs of DefAttnLayer / FocusedDecoderLayer
// (backbones/decoder_blocks.py:156-160, necks/focused_decoder.py:131-135), the K/V/output projections of FocusedAttn
// (necks/focused_decoder.py:213-218) and -- through the same kernel -- their gradients.
//
//   D[m, n] (+)= sum_r A(m, r) * B(n, r)  (+ bias[n])  (ReLU)
//
// Each operand is either "K-major" (the reduction index r is the contiguous one: X[M,K] of a Linear, W[N,K]) or "MN-major" (the
// m / n index is contiguous: what the gradient GEMMs dX = dY W and dW = dY^T X need) -- the tensor cores read both
// layouts from shared memory, so no operand is ever transposed in HBM.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer: one elected lane issues cp.async.bulk.tensor boxes (128-byte swizzle) into a ring of stages,
//               completion counted on the stage's "full" mbarrier
//   warp 1      owns TMEM (tcgen05.alloc / dealloc); lane 0 issues tcgen05.mma.kind::tf32 128 x BN x 8, four per stage, and
//               tcgen05.commit's the stage back to the producer ("empty") and the finished tile to the epilogue ("tmem full")
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias / ReLU -> global (or red.add for split-K)
//   The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the MMAs of tile i+1.
// fp32 operands are fed to the tensor cores as they are: kind::tf32 reads the upper 19 bits.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tcgemm {

constexpr int BM = 128;               // rows of D per tile = TMEM lanes
constexpr int BK = 32;                // reduction elements per stage = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;             // tf32
constexpr int kThreads = 192;
constexpr int kSlabBytes = BK * 128;  // MN-major operands: one 32(mn) x BK(r) slab per TMA box
constexpr unsigned kSpinLimit = 1u << 28;

template <int BN> struct Cfg {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;            // 6 (BN=128) / 5 (BN=192) / 4 (BN=256)
  static constexpr int TMEM_COLS = BN <= 128 ? 256 : 512;              // two accumulators of BN columns, rounded to a power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// A barrier that never completes would hang the GPU; trap instead (the host sees a launch failure).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32])
{
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptors (sm_100 version field = 1).
//   K-major : 128-byte swizzle (layout type 2).  Rows of 128 bytes (32 reduction elements), 8-row groups 1024 bytes apart
//             (SBO); LBO unused (1).  TMA: CU_TENSOR_MAP_SWIZZLE_128B.
//   MN-major: for 32-bit operands the tensor core only takes the "128-byte swizzle with 32-byte atoms" form (layout type 1:
//             32-byte chunks of a 128-byte row XORed with row % 4).  Slabs of 32 mn-elements (128 bytes) x BK reduction rows;
//             4-row groups 512 bytes apart (SBO), the next 32 mn-elements one slab (kSlabBytes) further (LBO).
//             TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
template <bool MN> __device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
  const uint64_t lbo = MN ? (uint64_t)(kSlabBytes >> 4) : 1ull;
  const uint64_t sbo = MN ? (uint64_t)(512 >> 4) : (uint64_t)(1024 >> 4);
  const uint64_t type = MN ? 1ull : 2ull;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}
template <bool MN> __device__ __forceinline__ uint32_t kstep_bytes() { return MN ? 1024u : (uint32_t)(UMMA_K * 4); }

template <int BN, bool A_MN, bool B_MN> __host__ __device__ constexpr uint32_t instr_desc()
{
  // c_format F32 (bits 4-5 = 1), a/b format TF32 (bits 7-9 / 10-12 = 2), a/b major (bits 15 / 16), N >> 3 (bits 17-22), M >> 4 (24-28)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct Problem {
  int M, N, R;                 // D is M x N, reduction length R
  long long ldd;
  int splits, rb_per_split;    // split-K: split s reduces r-blocks [s * rb_per_split, ...)
  int relu, atomic;            // atomic: D += (split-K or accumulate into an existing gradient)
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ D,
                 const float *__restrict__ bias, const Problem p)
{
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;                 // 128-byte swizzle atoms are 1024-byte aligned
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int r_blocks = (p.R + BK - 1) / BK;
  const long long work = (long long)m_tiles * n_tiles * p.splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long w = blockIdx.x; w < work; w += gridDim.x) {
        const int nt = (int)(w % n_tiles), mt = (int)((w / n_tiles) % m_tiles), sp = (int)(w / ((long long)n_tiles * m_tiles));
        const int kb0 = sp * p.rb_per_split, kb1 = min(r_blocks, kb0 + p.rb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), C::STAGE_BYTES);
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, full(stage), kb * BK, mt * BM);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 32; ++i) tma_load_2d(sa + i * kSlabBytes, &tmA, full(stage), mt * BM + 32 * i, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, full(stage), kb * BK, nt * BN);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) tma_load_2d(sb + i * kSlabBytes, &tmB, full(stage), nt * BN + 32 * i, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc<BN, A_MN, B_MN>();
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (long long w = blockIdx.x; w < work; w += gridDim.x) {
        const int sp = (int)(w / ((long long)n_tiles * m_tiles));
        const int kb0 = sp * p.rb_per_split, kb1 = min(r_blocks, kb0 + p.rb_per_split);
        mbar_wait(tempty(as), aphase ^ 1u);                       // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_tf32(acc, smem_desc<A_MN>(sa + k * kstep_bytes<A_MN>()), smem_desc<B_MN>(sb + k * kstep_bytes<B_MN>()), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(empty(stage));                              // smem slot free once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull(as));                                   // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;                                       // TMEM lane quarter this warp may read
    const int row_in_tile = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (long long w = blockIdx.x; w < work; w += gridDim.x) {
      const int nt = (int)(w % n_tiles), mt = (int)((w / n_tiles) % m_tiles);
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
      const int m = mt * BM + row_in_tile;
      float *drow = D + (long long)m * p.ldd;
      const bool vec_ok = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem_base + (uint32_t)(as * BN + c0) + ((uint32_t)(q * 32) << 16), v);
        const int n0 = nt * BN + c0;
        if (m < p.M && n0 < p.N) {
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += (n0 + i < p.N) ? __ldg(bias + n0 + i) : 0.f;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int n = n0 + i;
            if (vec_ok && n + 3 < p.N) {
              if (p.atomic)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(drow + n), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                             "f"(v[i + 3]) : "memory");
              else
                *reinterpret_cast<float4 *>(drow + n) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (n + j < p.N) {
                  if (p.atomic) atomicAdd(drow + n + j, v[i + j]);
                  else drow[n + j] = v[i + j];
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

}  // namespace tcgemm
