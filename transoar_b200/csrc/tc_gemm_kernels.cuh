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
// Structure (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 0      TMA producer: one elected lane issues cp.async.bulk.tensor boxes (128-byte swizzle) into a ring of stages,
//               completion counted on the stage's "full" mbarrier
//   warp 1      owns TMEM (tcgen05.alloc / dealloc); lane 0 issues tcgen05.mma.kind::tf32 128 x BN x 8, four per stage, and
//               tcgen05.commit's the stage back to the producer ("empty") and the finished tile to the epilogue ("tmem full")
//   warps 2-9   epilogue (two warps per TMEM lane quarter, alternating 32-column chunks): tcgen05.ld 32 lanes x 32 columns ->
//               registers -> shared-memory turn-around -> bias / ReLU / dropout / gate -> global (or red.add for split-K)
//   The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the MMAs of tile i+1.
// fp32 operands are fed to the tensor cores as they are: kind::tf32 reads the upper 19 bits.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hash_rng.cuh"

namespace tcgemm {

constexpr int BM = 128;               // rows of D per tile = TMEM lanes
constexpr int BK = 32;                // fp32 / TF32: reduction elements per stage = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;             // tf32
constexpr int kEpiWarps = 8;           // two epilogue warps per TMEM lane quarter, alternating 32-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSlabBytes = BK * 128;  // TF32 MN-major operands: one 32(mn) x BK(r) slab per TMA box
constexpr unsigned kSpinLimit = 1u << 28;

// Operand element types.  A stage row is always 128 bytes (one swizzle row) and an MMA always consumes 32 bytes of reduction
// per row, so the pipeline geometry in BYTES is the same for both; what changes is how many elements that is, the MN-major
// shared-memory form the tensor core accepts, and the instruction kind.
//   float (kind::tf32)          : 32 elements per stage row, UMMA_K = 8;  MN-major = 128B swizzle with 32-byte atoms (layout type 1),
//                                 slabs of 32 mn-elements, 4-row groups 512 bytes apart
//   __nv_bfloat16 (kind::f16)   : 64 elements per stage row, UMMA_K = 16; MN-major = the plain 128-byte swizzle (layout type 2),
//                                 slabs of 64 mn-elements, 8-row groups 1024 bytes apart
template <typename ET> struct Elem;
template <> struct Elem<float> {
  static constexpr int BKE = 32, MMA_K = 8, SLAB_MN = 32, MN_SBO = 512, MN_TYPE = 1, FMT = 2, IS_BF16 = 0;
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int BKE = 64, MMA_K = 16, SLAB_MN = 64, MN_SBO = 1024, MN_TYPE = 2, FMT = 1, IS_BF16 = 1;
};
template <typename ET> __host__ __device__ constexpr int slab_bytes() { return Elem<ET>::BKE * 128; }     // one MN-major TMA box: SLAB_MN mn-elements (128 B) x BKE rows

// CTAS = 1: one CTA computes a 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, one TPC) computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the B tile, so the operand bytes an SM pulls
// from L2 per flop drop by (128 + BN/2) / (128 + BN) (0.7 at BN = 192).
template <int BN, int CTAS> struct Cfg {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = (BN / CTAS) * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (160 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;      // 1 CTA: 5 / 4 / 3 (BN = 128 / 192 / 256); pair: 6 / 5 / 5
  static constexpr int TMEM_COLS = BN <= 128 ? 256 : 512;              // two accumulators of BN columns, rounded to a power of two
  static constexpr int EPI_WARP_BYTES = 2 * 4096;                      // per epilogue warp: two 32 x 32 fp32 chunks (TMA-store staging, double-buffered);
  static constexpr int EPI_BYTES = kEpiWarps * EPI_WARP_BYTES;         // the register-store path uses the first 32 x 36 floats of it
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
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

// Pair variant: the copy lands in this CTA's shared memory, its bytes are counted on the LEADER CTA's mbarrier (same offset,
// CTA-rank bit of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank)
{
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(bar), "r"(rank) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Pair variant: arrives on the barrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
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
template <bool MN, typename ET = float> __device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
  const uint64_t lbo = MN ? (uint64_t)(slab_bytes<ET>() >> 4) : 1ull;
  const uint64_t sbo = MN ? (uint64_t)(Elem<ET>::MN_SBO >> 4) : (uint64_t)(1024 >> 4);
  const uint64_t type = MN ? (uint64_t)Elem<ET>::MN_TYPE : 2ull;
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}
// bytes between the operand windows of consecutive MMAs of a stage: MN-major = MMA_K reduction rows of 128 bytes, K-major = 32 bytes
template <bool MN, typename ET = float> __device__ __forceinline__ uint32_t kstep_bytes() { return MN ? (uint32_t)(Elem<ET>::MMA_K * 128) : 32u; }

template <int BN, bool A_MN, bool B_MN, int CTAS, typename ET = float> __host__ __device__ constexpr uint32_t instr_desc()
{
  // c_format F32 (bits 4-5 = 1), a/b format (bits 7-9 / 10-12: kind::tf32 TF32 = 2; kind::f16 BF16 = 1), a/b major (bits 15 / 16),
  // N >> 3 (bits 17-22), M >> 4 (24-28)
  return (1u << 4) | ((uint32_t)Elem<ET>::FMT << 7) | ((uint32_t)Elem<ET>::FMT << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CTAS) >> 4) << 24);
}

struct Problem {
  int M, N, R;                 // D is M x N, reduction length R
  long long ldd;
  int splits, rb_per_split;    // split-K: split s reduces r-blocks [s * rb_per_split, ...)
  int relu, atomic;            // atomic: D += (split-K or accumulate into an existing gradient)
  // fused epilogue extras (NULL / 0 = off); both need the vector path (ldd % 4 == 0, N % 4 == 0, 16-byte aligned D / gate)
  const void *gate;            // D = result * (gate[m, n] > 0 ? gate_scale : 0): gradient of ReLU (+ dropout) taken from the saved activation (element type of D)
  float gate_scale;
  unsigned int drop_thresh;    // dropout after bias / ReLU: keep iff 16-bit hash(seed, element) >= drop_thresh, kept values * drop_scale
  float drop_scale;
  unsigned long long seed;
  const unsigned long long *epoch;   // hashrng::with_epoch: device counter folded into seed (graph replay), or NULL
  unsigned long long *prof;    // diagnostics (tc_gemm_debug_profile): per CTA 8 cycle counters of the three roles' waits, or NULL
  int tma_store;               // 1: the epilogue writes D with cp.async.bulk.tensor stores through tmD (fp32 D, no gate, no accumulation)
  alignas(64) CUtensorMap tmD; // D as a 2-D tensor [M][N], box 32 columns (128 bytes) x 32 rows, 128-byte swizzle
};

// wait + (diagnostics only) the cycles it took
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, const unsigned long long *prof, unsigned long long &acc)
{
  if (prof == nullptr) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

// four values of the output / gate element type as fp32, and back
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16 *p)
{
  const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&t.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&t.y);
  return make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(b.x), __bfloat162float(b.y));
}
__device__ __forceinline__ void st4(float *p, float4 o) { *reinterpret_cast<float4 *>(p) = o; }
__device__ __forceinline__ void st4(__nv_bfloat16 *p, float4 o)
{
  uint2 t;
  *reinterpret_cast<__nv_bfloat162 *>(&t.x) = __floats2bfloat162_rn(o.x, o.y);
  *reinterpret_cast<__nv_bfloat162 *>(&t.y) = __floats2bfloat162_rn(o.z, o.w);
  *reinterpret_cast<uint2 *>(p) = t;
}
__device__ __forceinline__ void st1(float *p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// ET: operand element type (float = TF32 multiply, __nv_bfloat16 = kind::f16); OT: element type of D (float, or bf16 for the
// forward / grad-input GEMMs of the bf16 route; accumulating launches always write fp32).
template <int BN, bool A_MN, bool B_MN, int CTAS, typename ET = float, typename OT = float>
__device__ __forceinline__ void gemm_tf32_body(const CUtensorMap &tmA, const CUtensorMap &tmB, OT *__restrict__ D,
                 const float *__restrict__ bias, const Problem &p)
{
  using C = Cfg<BN, CTAS>;
  using E = Elem<ET>;
  static_assert(!B_MN || (BN / CTAS) % E::SLAB_MN == 0, "MN-major B: the CTA's share of the tile must be whole slabs");
  constexpr int BKE = E::BKE;                                                   // reduction ELEMENTS per stage (128 bytes)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;                 // 128-byte swizzle atoms are 1024-byte aligned
  const uint32_t epi_base = base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bars = epi_base + C::EPI_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for the compiler
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;          // rank 0 = leader: owns the "full" / "tmem empty" barriers, issues the MMAs

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 32 * kEpiWarps * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int m_tiles = (p.M + BM * CTAS - 1) / (BM * CTAS), n_tiles = (p.N + BN - 1) / BN;      // a tile is 128 * CTAS rows
  const int r_blocks = (p.R + BKE - 1) / BKE;
  const long long work = (long long)m_tiles * n_tiles * p.splits;
  const long long w_first = blockIdx.x / CTAS, w_step = gridDim.x / CTAS;                     // both CTAs of a pair walk the same tiles

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long t_wait = 0;
      const long long t_begin = clock64();
      for (long long w = w_first; w < work; w += w_step) {
        const int nt = (int)(w % n_tiles), mt = (int)((w / n_tiles) % m_tiles), sp = (int)(w / ((long long)n_tiles * m_tiles));
        const int kb0 = sp * p.rb_per_split, kb1 = min(r_blocks, kb0 + p.rb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_timed(empty(stage), phase ^ 1u, p.prof, t_wait);
          if (rank == 0) mbar_expect_tx(full(stage), C::STAGE_BYTES * CTAS);            // the leader's barrier counts both CTAs' bytes
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          const int m0 = (mt * CTAS + (int)rank) * BM, n0 = nt * BN + (int)rank * (BN / CTAS);
          auto load = [&](uint32_t dst, const CUtensorMap *map, int c0, int c1) {
            if (CTAS == 2) tma_load_2d_pair(dst, map, full(stage), c0, c1);
            else tma_load_2d(dst, map, full(stage), c0, c1);
          };
          if (!A_MN) {
            load(sa, &tmA, kb * BKE, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / E::SLAB_MN; ++i) load(sa + i * slab_bytes<ET>(), &tmA, m0 + E::SLAB_MN * i, kb * BKE);
          }
          if (!B_MN) {
            load(sb, &tmB, kb * BKE, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / CTAS / E::SLAB_MN; ++i) load(sb + i * slab_bytes<ET>(), &tmB, n0 + E::SLAB_MN * i, kb * BKE);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.prof != nullptr) { p.prof[blockIdx.x * 8 + 0] = t_wait; p.prof[blockIdx.x * 8 + 1] = (unsigned long long)(clock64() - t_begin); }
    }
  } else if (warp == 1) {
    // The whole warp walks the loop and one elected lane issues: descriptors, TMEM and barrier addresses are then warp-uniform for
    // the compiler (uniform registers); under `if (lane == 0)` every tcgen05.mma paid a vector-to-uniform waterfall of ~12 instructions.
    if (rank == 0) {
      constexpr uint32_t idesc = instr_desc<BN, A_MN, B_MN, CTAS, ET>();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      unsigned long long t_full = 0, t_tempty = 0;
      const long long t_begin = clock64();
      for (long long w = w_first; w < work; w += w_step) {
        const int sp = (int)(w / ((long long)n_tiles * m_tiles));
        const int kb0 = sp * p.rb_per_split, kb1 = min(r_blocks, kb0 + p.rb_per_split);
        mbar_wait_timed(tempty(as), aphase ^ 1u, p.prof, t_tempty);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_u + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_timed(full(stage), phase, p.prof, t_full);
          tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BKE / E::MMA_K; ++k) {
              const uint64_t da = smem_desc<A_MN, ET>(sa + k * kstep_bytes<A_MN, ET>()), db = smem_desc<B_MN, ET>(sb + k * kstep_bytes<B_MN, ET>());
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if (E::IS_BF16) { if (CTAS == 2) umma_f16_pair(acc, da, db, idesc, accum); else umma_f16(acc, da, db, idesc, accum); }
              else { if (CTAS == 2) umma_tf32_pair(acc, da, db, idesc, accum); else umma_tf32(acc, da, db, idesc, accum); }
            }
            if (CTAS == 2) umma_commit_pair(empty(stage)); else umma_commit(empty(stage));   // smem slot(s) free once these MMAs have read them
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) { if (CTAS == 2) umma_commit_pair(tfull(as)); else umma_commit(tfull(as)); }   // accumulator complete (in both CTAs)
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
      if (p.prof != nullptr && lane == 0) {
        p.prof[blockIdx.x * 8 + 2] = t_full; p.prof[blockIdx.x * 8 + 3] = t_tempty;
        p.prof[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_begin);
      }
    }
  } else {
    const int q = warp & 3;                                       // TMEM lane quarter this warp may read
    int as = 0, tma_buf = 0;
    uint32_t aphase = 0;
    unsigned long long t_tfull = 0, t_ph[3] = {0, 0, 0};
    const long long t_begin = clock64();
    for (long long w = w_first; w < work; w += w_step) {
      const int nt = (int)(w % n_tiles), mt = (int)((w / n_tiles) % m_tiles);
      mbar_wait_timed(tfull(as), aphase, p.prof, t_tfull);
      tc_fence_after();
      // The accumulator comes out of TMEM one row per thread; a row-per-thread store would touch 32 different lines per
      // instruction.  Each warp therefore turns its 32 x 32 chunk around in shared memory (rows padded to 36 floats: the
      // 16-byte accesses of both phases are conflict-free) and stores 4 full 128-byte row segments per instruction.
      const uint32_t my_epi = epi_base + (uint32_t)(warp - 2) * C::EPI_WARP_BYTES;
      const int sub_row = lane >> 3, quad = lane & 7;
      const bool vec_ok = (p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & (4 * sizeof(OT) - 1)) == 0);
      const bool timed = p.prof != nullptr && warp == 2 && lane == 0;
#pragma unroll 1
      for (int c0 = 32 * ((warp - 2) >> 2); c0 < BN; c0 += 32 * (kEpiWarps / 4)) {   // the quarter's warps take alternating chunks
        const long long tp0 = timed ? clock64() : 0;
        // gate values of this lane's eight (row, 4-column) pieces: loaded first so that their latency hides behind the TMEM read
        float4 g4[8];
        if (p.gate != nullptr) {
          const int ng = nt * BN + c0 + 4 * quad;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int mg = (mt * CTAS + (int)rank) * BM + q * 32 + 4 * i + sub_row;
            g4[i] = (mg < p.M && ng + 3 < p.N) ? ld4(reinterpret_cast<const OT *>(p.gate) + (long long)mg * p.ldd + ng) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float v[32];
        tmem_ld_32x32(tmem_base + (uint32_t)(as * BN + c0) + ((uint32_t)(q * 32) << 16), v);
        const int n0 = nt * BN + c0;
        if (n0 >= p.N) continue;                                  // warp-uniform
        const long long tp1 = timed ? clock64() : 0;
        if constexpr (sizeof(OT) == 4) {
          if (p.tma_store) {
            // TMA-store epilogue: lane r holds row r of the 32 x 32 chunk; bias / ReLU / dropout in registers, the row goes to shared memory
            // in the 128-byte-swizzled layout of the D tensor map (16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4): conflict-free
            // for row-per-lane writes), and ONE bulk tensor store per chunk moves 4 KB to global memory asynchronously (SASS UTMASTG).
            // Rows / columns past M / N are clipped by the copy engine.  Two staging buffers per warp: the store of chunk i overlaps the
            // TMEM read and the arithmetic of chunk i + 1.
            const int m_row = (mt * CTAS + (int)rank) * BM + q * 32 + lane;
            if (bias != nullptr) {
              if (n0 + 31 < p.N && (reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + n0) + j);
                  v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += (n0 + j < p.N) ? __ldg(bias + n0 + j) : 0.f;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.drop_thresh != 0u) {
              const unsigned long long sd = hashrng::with_epoch(p.seed, p.epoch);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float keep[4];
                hashrng::keep4(sd, ((unsigned long long)m_row * p.N + n0 + j) >> 2, p.drop_thresh, p.drop_scale, keep);
                v[j] *= keep[0]; v[j + 1] *= keep[1]; v[j + 2] *= keep[2]; v[j + 3] *= keep[3];
              }
            }
            const uint32_t buf = epi_base + (uint32_t)(warp - 2) * C::EPI_WARP_BYTES + (uint32_t)(tma_buf * 4096);
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the store issued from this buffer two chunks ago has read it
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + (uint32_t)(lane * 128 + ((i ^ (lane & 7)) << 4))), "f"(v[4 * i]),
                           "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&p.tmD), "r"(buf), "r"(n0), "r"((mt * CTAS + (int)rank) * BM + q * 32) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            tma_buf ^= 1;
            if (timed) { const long long tp3 = clock64(); t_ph[0] += (unsigned long long)(tp1 - tp0); t_ph[2] += (unsigned long long)(tp3 - tp1); }
            continue;
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(my_epi + (uint32_t)(lane * 36 + 4 * i) * 4), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                       "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
        __syncwarp();
        const long long tp2 = timed ? clock64() : 0;
        // bias / ReLU on the store side: a lane owns the same four columns for all rows of the chunk
        const int n = n0 + 4 * quad;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) {
          if (n + 3 < p.N && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
            b4 = __ldg(reinterpret_cast<const float4 *>(bias + n));
          } else {
            b4.x = n < p.N ? __ldg(bias + n) : 0.f; b4.y = n + 1 < p.N ? __ldg(bias + n + 1) : 0.f;
            b4.z = n + 2 < p.N ? __ldg(bias + n + 2) : 0.f; b4.w = n + 3 < p.N ? __ldg(bias + n + 3) : 0.f;
          }
        }
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 4) {
          const int r = r0 + sub_row;
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(my_epi + (uint32_t)(r * 36 + 4 * quad) * 4) : "memory");
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          const int m = (mt * CTAS + (int)rank) * BM + q * 32 + r;
          if (m >= p.M || n >= p.N) continue;
          OT *dst = D + (long long)m * p.ldd + n;
          if (p.drop_thresh != 0u) {                              // host guarantees the vector path: N % 4 == 0
            float keep[4];
            hashrng::keep4(hashrng::with_epoch(p.seed, p.epoch), ((unsigned long long)m * p.N + n) >> 2, p.drop_thresh, p.drop_scale, keep);
            o.x *= keep[0]; o.y *= keep[1]; o.z *= keep[2]; o.w *= keep[3];
          }
          if (p.gate != nullptr) {
            const float4 gq = g4[r0 >> 2];
            o.x = gq.x > 0.f ? o.x * p.gate_scale : 0.f; o.y = gq.y > 0.f ? o.y * p.gate_scale : 0.f;
            o.z = gq.z > 0.f ? o.z * p.gate_scale : 0.f; o.w = gq.w > 0.f ? o.w * p.gate_scale : 0.f;
          }
          if (vec_ok && n + 3 < p.N) {
            if (p.atomic) {                                           // host guarantees OT == float for accumulating launches
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            } else {
              st4(dst, o);                                            // (evict-first .cs stores measured 20 % slower here)
            }
          } else {
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n + j < p.N) {
                if (p.atomic) atomicAdd(reinterpret_cast<float *>(dst) + j, ov[j]);
                else st1(dst + j, ov[j]);
              }
            }
          }
        }
        if (timed) {
          const long long tp3 = clock64();
          t_ph[0] += (unsigned long long)(tp1 - tp0); t_ph[1] += (unsigned long long)(tp2 - tp1); t_ph[2] += (unsigned long long)(tp3 - tp2);
        }
      }
      tc_fence_before();
      if (CTAS == 2) mbar_arrive_cluster(tempty(as), 0u); else mbar_arrive(tempty(as));     // the leader's MMA warp waits for both CTAs' epilogues
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // all bulk stores of this warp have completed
    if (p.prof != nullptr && warp == 2 && lane == 0) {
      p.prof[blockIdx.x * 8 + 5] = t_tfull; p.prof[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - t_begin);
      p.prof[296 * 8 + blockIdx.x * 4 + 0] = t_ph[0]; p.prof[296 * 8 + blockIdx.x * 4 + 1] = t_ph[1]; p.prof[296 * 8 + blockIdx.x * 4 + 2] = t_ph[2];
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();         // nobody touches the peer's barriers / TMEM after this point
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ D,
                 const float *__restrict__ bias, const __grid_constant__ Problem p)
{
  gemm_tf32_body<BN, A_MN, B_MN, 1>(tmA, tmB, D, bias, p);
}

// bf16 operands (tcgen05.mma kind::f16, fp32 accumulation in TMEM), D in bf16 (forward / grad-input) or fp32 (weight gradients, split-K)
template <int BN, bool A_MN, bool B_MN, typename OT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OT *__restrict__ D,
                 const float *__restrict__ bias, const __grid_constant__ Problem p)
{
  gemm_tf32_body<BN, A_MN, B_MN, 1, __nv_bfloat16, OT>(tmA, tmB, D, bias, p);
}

template <int BN, bool A_MN, bool B_MN, typename OT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OT *__restrict__ D,
                      const float *__restrict__ bias, const __grid_constant__ Problem p)
{
  gemm_tf32_body<BN, A_MN, B_MN, 2, __nv_bfloat16, OT>(tmA, tmB, D, bias, p);
}

// CTA-pair variant: cluster of two CTAs on one TPC, 256 x BN tiles, tcgen05.mma.cta_group::2.
template <int BN, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *__restrict__ D,
                      const float *__restrict__ bias, const __grid_constant__ Problem p)
{
  gemm_tf32_body<BN, A_MN, B_MN, 2>(tmA, tmB, D, bias, p);
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums of a row-major [rows, C] matrix: the bias gradient of a Linear layer (grad_bias = sum over tokens of grad_output).
// HBM-bound single pass: thread t owns the float4 column group t % (C/4) and walks rows t / (C/4), t / (C/4) + rows_per_pass, ...;
// the row lanes of a CTA are combined in shared memory, one partial per CTA, summed by colsum_finalize_kernel.  C % 4 == 0, C <= 1024.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kColsumThreads = 256;

static __global__ void __launch_bounds__(kColsumThreads)
colsum_partial_kernel(const float *__restrict__ x, long long rows, int C, long long ld, float *__restrict__ part)
{
  __shared__ float4 red[kColsumThreads];
  const int cg = C >> 2, lanes = kColsumThreads / cg, t = threadIdx.x;
  const int g = t % cg, r = t / cg;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < lanes) {
    const long long step = (long long)gridDim.x * lanes;
    long long row = (long long)blockIdx.x * lanes + r;
    for (; row + 3 * step < rows; row += 4 * step) {               // four independent loads in flight
      const float4 a = __ldg(reinterpret_cast<const float4 *>(x + row * ld) + g);
      const float4 b = __ldg(reinterpret_cast<const float4 *>(x + (row + step) * ld) + g);
      const float4 c = __ldg(reinterpret_cast<const float4 *>(x + (row + 2 * step) * ld) + g);
      const float4 d = __ldg(reinterpret_cast<const float4 *>(x + (row + 3 * step) * ld) + g);
      acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
      acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; row < rows; row += step) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(x + row * ld) + g);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
  }
  red[t] = acc;
  __syncthreads();
  if (t < cg) {
    float4 s = red[t];
    for (int k = 1; k < lanes; ++k) {
      const float4 o = red[t + k * cg];
      s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
    }
    reinterpret_cast<float4 *>(part + (long long)blockIdx.x * C)[t] = s;
  }
}

// 256 threads per 32 columns: warp w sums partials w, w + 8, ... of its columns, the eight warps are combined in shared memory
// (a single thread per column walking all CTAs' partials is latency-bound: 45 us for 592 partials)
static __global__ void __launch_bounds__(256) colsum_finalize_kernel(const float *__restrict__ part, int ctas, int C, float *__restrict__ out)
{
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (c < C)
    for (int b = w; b < ctas; b += 8) s += part[(long long)b * C + c];
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][lane];
    out[c] = (float)s;
  }
}

}  // namespace tcgemm
