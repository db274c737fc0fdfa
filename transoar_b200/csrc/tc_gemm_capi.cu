// tc_gemm_capi.cu -- C ABI of the TF32 tcgen05 GEMM (include/tc_gemm.h): tensor-map construction, tile-shape choice, launch.
#include "tc_gemm_kernels.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "../../include/msda3d.h"
#include "../../include/tc_gemm.h"

extern std::atomic<unsigned long long> g_msda3d_launches;
extern std::atomic<const unsigned long long *> g_hashrng_epoch;   // fused_ln_capi.cu (hash_rng_set_epoch)

namespace {

std::atomic<unsigned long long *> g_prof{nullptr};          // diagnostics only (tc_gemm_debug_profile)

using EncodeTiled = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled lives in libcuda; fetched through the runtime so the library has no link-time driver dependency.
// cuTensorMapEncodeTiled is a driver-API call and needs a context current on the CALLING thread.  A fresh host thread (autograd's
// backward thread on device 0: torch skips cudaSetDevice when the device index already matches) has none until its first runtime
// call binds the primary context -- bind it explicitly, once per thread (cudaSetDevice is legal during stream capture).
void ensure_context_on_this_thread()
{
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);
    bound = true;
  }
}

EncodeTiled encode_fn()
{
  static EncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled>(p);
  });
  return fn;
}

// 2-D fp32 tensor [outer, inner] (inner contiguous, `ld` elements between outer rows), box = 32 inner elements (128 bytes,
// one swizzle row) x box_outer rows; the swizzle form is the one the tensor core expects for that operand layout (kernels.cuh).  Out-of-bounds parts of a box are zero-filled, which is what makes ragged M / N / R work.
int make_map(CUtensorMap *map, const float *ptr, long long inner, long long outer, long long ld, int box_outer, bool mn_major)
{
  ensure_context_on_this_thread();
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)tcgemm::BK, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  // TFLOAT32: the copy engine rounds each fp32 value to TF32 (nearest) on its way into shared memory, so the tensor core's
  // read of the upper 19 bits is a rounding, not a truncation (TC_GEMM_TMA_ROUND=0 keeps raw fp32 words: diagnostics only).
  static const bool round_tf32 = [] { const char *e = getenv("TC_GEMM_TMA_ROUND"); return e == nullptr || e[0] != '0'; }();
  const CUresult r = enc(map, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MSDA3D_EINVAL;
}

// D [M][N] fp32 for the TMA-store epilogue: box = 32 columns (128 bytes) x 32 rows, 128-byte swizzle (kernels.cuh); parts of a box past
// M / N are clipped on the way out.
int make_map_out(CUtensorMap *map, float *ptr, long long N, long long M, long long ldd)
{
  ensure_context_on_this_thread();
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)ldd * sizeof(float)};
  const cuuint32_t box[2] = {32u, 32u};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MSDA3D_EINVAL;
}

// 2-D bf16 tensor: box = 64 inner elements (128 bytes) x box_outer rows, plain 128-byte swizzle for both operand layouts (kernels.cuh, Elem<bf16>).
int make_map_bf16(CUtensorMap *map, const void *ptr, long long inner, long long outer, long long ld, int box_outer)
{
  ensure_context_on_this_thread();
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("TC_GEMM_DEBUG") != nullptr)
    fprintf(stderr, "  cuTensorMapEncodeTiled(bf16) -> %d: ptr=%p inner=%lld outer=%lld ld=%lld box_outer=%d\n", (int)r, ptr, inner, outer, ld, box_outer);
  return r == CUDA_SUCCESS ? 0 : MSDA3D_EINVAL;
}

int sm_count()
{
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

// co-resident CTA pairs (clusters of 2) of the pair kernel on the current device; 0 if clusters cannot be scheduled
template <int BN, bool A_MN, bool B_MN> int max_pairs()
{
  static int pairs[64];
  static bool known[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!known[dev]) {
    using C = tcgemm::Cfg<BN, 2>;
    auto kern = tcgemm::gemm_tf32_pair_kernel<BN, A_MN, B_MN>;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sm_count());
      cfg.blockDim = dim3(tcgemm::kThreads);
      cfg.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    (void)cudaGetLastError();
    pairs[dev] = n;
    known[dev] = true;
  }
  return pairs[dev];
}

template <int BN, bool A_MN, bool B_MN>
int launch(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, float *D, const float *bias, const tcgemm::Problem &p, bool pair)
{
  if (pair) {
    const int pairs = max_pairs<BN, A_MN, B_MN>();
    if (pairs > 0) {
      using C = tcgemm::Cfg<BN, 2>;
      const long long work = (long long)((p.M + 2 * tcgemm::BM - 1) / (2 * tcgemm::BM)) * ((p.N + BN - 1) / BN) * p.splits;
      const int grid = 2 * (int)(work < pairs ? work : pairs);
      tcgemm::gemm_tf32_pair_kernel<BN, A_MN, B_MN><<<grid, tcgemm::kThreads, C::SMEM_BYTES, st>>>(ma, mb, D, bias, p);
      ++g_msda3d_launches;
      return (int)cudaGetLastError();
    }
    return MSDA3D_ENODEV;                                          // the B tensor map was built for half tiles: no silent switch
  }
  using C = tcgemm::Cfg<BN, 1>;
  auto kern = tcgemm::gemm_tf32_kernel<BN, A_MN, B_MN>;
  static std::once_flag once;                                    // one flag per instantiation
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return (int)attr_err;
  const long long work = (long long)((p.M + tcgemm::BM - 1) / tcgemm::BM) * ((p.N + BN - 1) / BN) * p.splits;
  const int grid = (int)(work < sm_count() ? work : sm_count());
  kern<<<grid, tcgemm::kThreads, C::SMEM_BYTES, st>>>(ma, mb, D, bias, p);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

template <int BN>
int dispatch(cudaStream_t st, bool a_mn, bool b_mn, const CUtensorMap &ma, const CUtensorMap &mb, float *D, const float *bias,
             const tcgemm::Problem &p, bool pair)
{
  if (!a_mn && !b_mn) return launch<BN, false, false>(st, ma, mb, D, bias, p, pair);
  if (!a_mn && b_mn) return launch<BN, false, true>(st, ma, mb, D, bias, p, pair);
  if (a_mn && !b_mn) return launch<BN, true, false>(st, ma, mb, D, bias, p, pair);
  return launch<BN, true, true>(st, ma, mb, D, bias, p, pair);
}

// ---- bf16 operands ----
template <int BN, bool A_MN, bool B_MN, typename OT> int max_pairs_bf16()
{
  static int pairs = -1;
  if (pairs < 0) {
    using C = tcgemm::Cfg<BN, 2>;
    auto kern = tcgemm::gemm_bf16_pair_kernel<BN, A_MN, B_MN, OT>;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sm_count());
      cfg.blockDim = dim3(tcgemm::kThreads);
      cfg.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    (void)cudaGetLastError();
    pairs = n;
  }
  return pairs;
}

template <int BN, bool A_MN, bool B_MN, typename OT>
int launch_bf16(cudaStream_t st, const CUtensorMap &ma, const CUtensorMap &mb, OT *D, const float *bias, const tcgemm::Problem &p, bool pair)
{
  if constexpr (!B_MN || (BN / 2) % 64 == 0) {
    if (pair) {
      const int pairs = max_pairs_bf16<BN, A_MN, B_MN, OT>();
      if (pairs <= 0) return MSDA3D_ENODEV;
      using C = tcgemm::Cfg<BN, 2>;
      const long long work = (long long)((p.M + 2 * tcgemm::BM - 1) / (2 * tcgemm::BM)) * ((p.N + BN - 1) / BN) * p.splits;
      const int grid = 2 * (int)(work < pairs ? work : pairs);
      tcgemm::gemm_bf16_pair_kernel<BN, A_MN, B_MN, OT><<<grid, tcgemm::kThreads, C::SMEM_BYTES, st>>>(ma, mb, D, bias, p);
      ++g_msda3d_launches;
      return (int)cudaGetLastError();
    }
  } else {
    if (pair) return MSDA3D_EINVAL;
  }
  using C = tcgemm::Cfg<BN, 1>;
  auto kern = tcgemm::gemm_bf16_kernel<BN, A_MN, B_MN, OT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
  if (attr_err != cudaSuccess) return (int)attr_err;
  const long long work = (long long)((p.M + tcgemm::BM - 1) / tcgemm::BM) * ((p.N + BN - 1) / BN) * p.splits;
  const int grid = (int)(work < sm_count() ? work : sm_count());
  kern<<<grid, tcgemm::kThreads, C::SMEM_BYTES, st>>>(ma, mb, D, bias, p);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

template <int BN, typename OT>
int dispatch_bf16(cudaStream_t st, bool a_mn, bool b_mn, const CUtensorMap &ma, const CUtensorMap &mb, OT *D, const float *bias,
                  const tcgemm::Problem &p, bool pair)
{
  if (!a_mn && !b_mn) return launch_bf16<BN, false, false, OT>(st, ma, mb, D, bias, p, pair);
  if (!a_mn && b_mn) return launch_bf16<BN, false, true, OT>(st, ma, mb, D, bias, p, pair);
  if (a_mn && !b_mn) return launch_bf16<BN, true, false, OT>(st, ma, mb, D, bias, p, pair);
  return launch_bf16<BN, true, true, OT>(st, ma, mb, D, bias, p, pair);
}

}  // namespace

constexpr int kColsumCtas = 148 * 4;

extern "C" long long tc_colsum_workspace_floats(int channels) { return channels > 0 ? (long long)kColsumCtas * channels : 0; }

extern "C" int tc_colsum(void *stream, const float *x, long long rows, int channels, long long ld, float *out, float *workspace)
{
  if (!x || !out || !workspace || rows <= 0 || channels <= 0 || channels % 4 != 0 || channels > 1024 || ld < channels || ld % 4 != 0) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(workspace)) & 15) return MSDA3D_EALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int lanes = tcgemm::kColsumThreads / (channels / 4);
  const long long need = (rows + lanes - 1) / lanes;
  const int grid = (int)(need < kColsumCtas ? need : kColsumCtas);
  tcgemm::colsum_partial_kernel<<<grid, tcgemm::kColsumThreads, 0, st>>>(x, rows, channels, ld, workspace);
  tcgemm::colsum_finalize_kernel<<<(channels + 31) / 32, 256, 0, st>>>(workspace, grid, channels, out);
  g_msda3d_launches += 2;
  return (int)cudaGetLastError();
}

extern "C" void tc_gemm_debug_profile(unsigned long long *device_counters) { g_prof.store(device_counters); }

extern "C" int tc_gemm_tf32(void *stream, const float *A, int a_mn_major, long long lda, const float *B, int b_mn_major, long long ldb,
                            float *D, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k)
{
  return tc_gemm_tf32_ex(stream, A, a_mn_major, lda, B, b_mn_major, ldb, D, ldd, bias, M, N, R, relu, accumulate, split_k, nullptr, 1.f, 0.f, 0ull);
}

extern "C" int tc_gemm_tf32_ex(void *stream, const float *A, int a_mn_major, long long lda, const float *B, int b_mn_major, long long ldb,
                               float *D, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k,
                               const float *gate, float gate_scale, float p_drop, unsigned long long seed)
{
  if (A == nullptr || B == nullptr || D == nullptr || M <= 0 || N <= 0 || R <= 0 || lda <= 0 || ldb <= 0 || ldd < N) return MSDA3D_EINVAL;
  if (lda % 4 != 0 || ldb % 4 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(D) & 3))
    return MSDA3D_EALIGN;
  if (split_k < 0 || (split_k != 1 && !accumulate)) return MSDA3D_EINVAL;
  if (gate != nullptr || p_drop > 0.f) {                         // fused gate / dropout: vector epilogue only, never with accumulation
    if (accumulate || p_drop < 0.f || p_drop >= 1.f || N % 4 != 0 || ldd % 4 != 0) return MSDA3D_EINVAL;
    if ((reinterpret_cast<uintptr_t>(D) & 15) || (reinterpret_cast<uintptr_t>(gate) & 15)) return MSDA3D_EALIGN;
  }

  // tile width: fewest column tiles first (every extra one re-reads the whole A operand), then least padding
  int BN = 128;
  {
    int best_tiles = 1 << 30, best_pad = 1 << 30;
    for (int cand : {128, 192, 256}) {
      const int tiles = (N + cand - 1) / cand, pad = tiles * cand - N;
      if (tiles < best_tiles || (tiles == best_tiles && pad < best_pad)) { best_tiles = tiles; best_pad = pad; BN = cand; }
    }
  }
  const int m_tiles = (M + tcgemm::BM - 1) / tcgemm::BM, n_tiles = (N + BN - 1) / BN, r_blocks = (R + tcgemm::BK - 1) / tcgemm::BK;
  int splits = split_k;
  if (splits == 0) {                                             // fill the machine: one work item per SM where the reduction allows
    const long long tiles = (long long)m_tiles * n_tiles;
    splits = (int)(tiles >= sm_count() ? 1 : sm_count() / tiles);
  }
  if (splits > r_blocks) splits = r_blocks;
  tcgemm::Problem p;
  p.M = M; p.N = N; p.R = R; p.ldd = ldd; p.relu = relu; p.atomic = accumulate ? 1 : 0;
  p.gate = gate; p.gate_scale = gate_scale; p.seed = seed; p.epoch = g_hashrng_epoch.load();
  p.drop_thresh = p_drop > 0.f ? (unsigned int)(p_drop * 65536.f + 0.5f) : 0u;
  p.drop_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  p.prof = g_prof.load();
  p.rb_per_split = (r_blocks + splits - 1) / splits;
  p.splits = (r_blocks + p.rb_per_split - 1) / p.rb_per_split;   // no empty split
  // TMA-store epilogue for plain (non-accumulating, un-gated) results: forward and grad-input GEMMs; TC_GEMM_TMA_STORE=0 keeps the
  // register-store epilogue (diagnostics / A-B timing)
  static const bool tma_store_enabled = [] { const char *e = getenv("TC_GEMM_TMA_STORE"); return e == nullptr || e[0] != '0'; }();
  p.tma_store = 0;
  if (tma_store_enabled && !accumulate && gate == nullptr && ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0) {
    if (int rc = make_map_out(&p.tmD, D, N, M, ldd)) return rc;
    p.tma_store = 1;
  }

  // CTA pairs (256-row tiles, B tile split over the pair) when there are enough row tiles to fill the machine with pairs;
  // TC_GEMM_PAIR=0 forces the single-CTA kernel (diagnostics).
  static const bool pair_enabled = [] { const char *e = getenv("TC_GEMM_PAIR"); return e == nullptr || e[0] != '0'; }();
  const bool pair = pair_enabled && !accumulate && splits == 1 && (long long)(M / 256) * n_tiles >= sm_count() / 2;
  CUtensorMap ma, mb;
  int rc = a_mn_major ? make_map(&ma, A, M, R, lda, tcgemm::BK, true) : make_map(&ma, A, R, M, lda, tcgemm::BM, false);
  if (rc != 0) return rc;
  rc = b_mn_major ? make_map(&mb, B, N, R, ldb, tcgemm::BK, true) : make_map(&mb, B, R, N, ldb, pair ? BN / 2 : BN, false);
  if (rc != 0) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (BN == 256) return dispatch<256>(st, a_mn_major != 0, b_mn_major != 0, ma, mb, D, bias, p, pair);
  if (BN == 192) return dispatch<192>(st, a_mn_major != 0, b_mn_major != 0, ma, mb, D, bias, p, pair);
  return dispatch<128>(st, a_mn_major != 0, b_mn_major != 0, ma, mb, D, bias, p, pair);
}

// bf16 operands, fp32 accumulation; D is bf16 (out_fp32 == 0) or fp32 (out_fp32 != 0; required for accumulate / split-K).
extern "C" int tc_gemm_bf16(void *stream, const void *A, int a_mn_major, long long lda, const void *B, int b_mn_major, long long ldb,
                            void *D, int out_fp32, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k,
                            const void *gate, float gate_scale, float p_drop, unsigned long long seed)
{
  static const bool dbg = getenv("TC_GEMM_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "tc_gemm_bf16 A=%p amn=%d lda=%lld B=%p bmn=%d ldb=%lld D=%p f32=%d ldd=%lld M=%d N=%d R=%d relu=%d acc=%d split=%d gate=%p\n", A, a_mn_major, lda, B,
                   b_mn_major, ldb, D, out_fp32, ldd, M, N, R, relu, accumulate, split_k, gate);
  if (A == nullptr || B == nullptr || D == nullptr || M <= 0 || N <= 0 || R <= 0 || lda <= 0 || ldb <= 0 || ldd < N) return MSDA3D_EINVAL;
  if (lda % 8 != 0 || ldb % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(D) & (out_fp32 ? 3 : 1)))
    return MSDA3D_EALIGN;
  if (split_k < 0 || (split_k != 1 && !accumulate) || (accumulate && !out_fp32)) return MSDA3D_EINVAL;
  if (gate != nullptr || p_drop > 0.f) {
    if (accumulate || p_drop < 0.f || p_drop >= 1.f || N % 4 != 0 || ldd % 4 != 0) return MSDA3D_EINVAL;
    if ((reinterpret_cast<uintptr_t>(D) & (out_fp32 ? 15 : 7)) || (reinterpret_cast<uintptr_t>(gate) & (out_fp32 ? 15 : 7))) return MSDA3D_EALIGN;
  }
  int BN = 128;
  {
    int best_tiles = 1 << 30, best_pad = 1 << 30;
    for (int cand : {128, 192, 256}) {
      const int tiles = (N + cand - 1) / cand, pad = tiles * cand - N;
      if (tiles < best_tiles || (tiles == best_tiles && pad < best_pad)) { best_tiles = tiles; best_pad = pad; BN = cand; }
    }
  }
  const int m_tiles = (M + tcgemm::BM - 1) / tcgemm::BM, n_tiles = (N + BN - 1) / BN, r_blocks = (R + 63) / 64;
  int splits = split_k;
  if (splits == 0) {
    const long long tiles = (long long)m_tiles * n_tiles;
    splits = (int)(tiles >= sm_count() ? 1 : sm_count() / tiles);
  }
  if (splits > r_blocks) splits = r_blocks;
  tcgemm::Problem p;
  p.M = M; p.N = N; p.R = R; p.ldd = ldd; p.relu = relu; p.atomic = accumulate ? 1 : 0;
  p.gate = gate; p.gate_scale = gate_scale; p.seed = seed; p.epoch = g_hashrng_epoch.load();
  p.drop_thresh = p_drop > 0.f ? (unsigned int)(p_drop * 65536.f + 0.5f) : 0u;
  p.drop_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  p.prof = nullptr;
  p.tma_store = 0;
  p.rb_per_split = (r_blocks + splits - 1) / splits;
  p.splits = (r_blocks + p.rb_per_split - 1) / p.rb_per_split;
  static const bool pair_enabled = [] { const char *e = getenv("TC_GEMM_PAIR"); return e == nullptr || e[0] != '0'; }();
  bool pair = pair_enabled && !accumulate && splits == 1 && (long long)(M / 256) * n_tiles >= sm_count() / 2;
  if (pair && b_mn_major && (BN / 2) % 64 != 0) pair = false;   // an MN-major B operand comes in slabs of 64 columns: 192 / 2 = 96 is not whole slabs
  CUtensorMap ma, mb;
  int rc = a_mn_major ? make_map_bf16(&ma, A, M, R, lda, 64) : make_map_bf16(&ma, A, R, M, lda, tcgemm::BM);
  if (dbg) fprintf(stderr, "  BN=%d splits=%d pair=%d mapA rc=%d\n", BN, p.splits, (int)pair, rc);
  if (rc != 0) return rc;
  rc = b_mn_major ? make_map_bf16(&mb, B, N, R, ldb, 64) : make_map_bf16(&mb, B, R, N, ldb, pair ? BN / 2 : BN);
  if (dbg) fprintf(stderr, "  mapB rc=%d\n", rc);
  if (rc != 0) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool am = a_mn_major != 0, bm = b_mn_major != 0;
  if (out_fp32) {
    float *Df = reinterpret_cast<float *>(D);
    if (BN == 256) return dispatch_bf16<256, float>(st, am, bm, ma, mb, Df, bias, p, pair);
    if (BN == 192) return dispatch_bf16<192, float>(st, am, bm, ma, mb, Df, bias, p, pair);
    return dispatch_bf16<128, float>(st, am, bm, ma, mb, Df, bias, p, pair);
  }
  __nv_bfloat16 *Dh = reinterpret_cast<__nv_bfloat16 *>(D);
  if (BN == 256) return dispatch_bf16<256, __nv_bfloat16>(st, am, bm, ma, mb, Dh, bias, p, pair);
  if (BN == 192) return dispatch_bf16<192, __nv_bfloat16>(st, am, bm, ma, mb, Dh, bias, p, pair);
  return dispatch_bf16<128, __nv_bfloat16>(st, am, bm, ma, mb, Dh, bias, p, pair);
}
