// fused_ln_capi.cu -- C ABI of the fused residual + dropout + LayerNorm (include/fused_ln.h).
#include "fused_ln_kernels.cuh"

#include <atomic>
#include <mutex>

#include "../../include/fused_ln.h"
#include "../../include/msda3d.h"

extern std::atomic<unsigned long long> g_msda3d_launches;
std::atomic<const unsigned long long *> g_hashrng_epoch{nullptr};   // shared with tc_gemm_capi.cu

namespace {

constexpr int kCtas = 148 * 6;

bool bad(long long rows, int C) { return rows <= 0 || C <= 0 || C % 4 != 0 || C > 4 * 32 * fusedln::kMaxNV; }
int nv_of(int C) { return (C / 4 + 31) / 32; }
uint32_t thresh_of(float p) { return p <= 0.f ? 0u : (uint32_t)(p * 65536.f + 0.5f); }
int grid_of(long long rows) { const long long need = (rows + fusedln::kWarps - 1) / fusedln::kWarps; return (int)(need < kCtas ? need : kCtas); }

template <int NV, typename TB>
int fwd(cudaStream_t st, const float *a, const TB *b, const float *gamma, const float *beta, long long rows, int C, float eps, float p,
        unsigned long long seed, float *z, float *y, float *mean, float *rstd)
{
  fusedln::fwd_kernel<NV, TB><<<grid_of(rows), fusedln::kThreads, 0, st>>>(a, b, gamma, beta, rows, C, eps, thresh_of(p), p > 0.f ? 1.f / (1.f - p) : 1.f,
                                                                        seed, g_hashrng_epoch.load(), z, y, mean, rstd);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

template <int NV, typename TB>
int bwd(cudaStream_t st, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd, long long rows, int C, float p,
        unsigned long long seed, float *da, TB *db, float *dgamma, float *dbeta, float *ws)
{
  const int grid = grid_of(rows);
  const size_t smem = (size_t)fusedln::kWarps * 2 * C * sizeof(float);
  auto kern = fusedln::bwd_kernel<NV, TB>;
  static std::once_flag once;                                      // C = 1024: 64 KB of reduction scratch
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fusedln::kWarps * 2 * 4 * 32 * fusedln::kMaxNV * (int)sizeof(float)); });
  if (err != cudaSuccess) return (int)err;
  kern<<<grid, fusedln::kThreads, smem, st>>>(dy, z, gamma, mean, rstd, rows, C, thresh_of(p), p > 0.f ? 1.f / (1.f - p) : 1.f, seed, g_hashrng_epoch.load(), da, db, ws);
  fusedln::bwd_finalize_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(ws, grid, C, dgamma, dbeta);
  g_msda3d_launches += 2;
  return (int)cudaGetLastError();
}

}  // namespace

template <typename TB>
static int forward_impl(void *stream, const float *a, const TB *b, const float *gamma, const float *beta, long long rows, int channels,
                 float eps, float p_drop, unsigned long long seed, float *z, float *y, float *mean, float *rstd)
{
  if (!a || !gamma || !beta || !y || !mean || !rstd || bad(rows, channels) || p_drop < 0.f || p_drop >= 1.f) return MSDA3D_EINVAL;
  if (b != nullptr && z == nullptr) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(y) |
       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15)
    return MSDA3D_EALIGN;
  if (reinterpret_cast<uintptr_t>(b) & (4 * sizeof(TB) - 1)) return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
#define FWD_CALL(NV) fwd<NV, TB>(st, a, b, gamma, beta, rows, channels, eps, b ? p_drop : 0.f, seed, z, y, mean, rstd)
  switch (nv_of(channels)) {
    case 1: return FWD_CALL(1);
    case 2: return FWD_CALL(2);
    case 3: return FWD_CALL(3);
    case 4: return FWD_CALL(4);
    case 5: case 6: return FWD_CALL(6);
    default: return FWD_CALL(8);
  }
#undef FWD_CALL
}

template <typename TB>
static int backward_impl(void *stream, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd, long long rows,
                  int channels, float p_drop, unsigned long long seed, float *da, TB *db, float *dgamma, float *dbeta, float *workspace)
{
  if (!dy || !z || !gamma || !mean || !rstd || !da || !dgamma || !dbeta || !workspace || bad(rows, channels) || p_drop < 0.f || p_drop >= 1.f)
    return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(gamma)) & 15)
    return MSDA3D_EALIGN;
  if (reinterpret_cast<uintptr_t>(db) & (4 * sizeof(TB) - 1)) return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
#define BWD_CALL(NV) bwd<NV, TB>(st, dy, z, gamma, mean, rstd, rows, channels, db ? p_drop : 0.f, seed, da, db, dgamma, dbeta, workspace)
  switch (nv_of(channels)) {
    case 1: return BWD_CALL(1);
    case 2: return BWD_CALL(2);
    case 3: return BWD_CALL(3);
    case 4: return BWD_CALL(4);
    case 5: case 6: return BWD_CALL(6);
    default: return BWD_CALL(8);
  }
#undef BWD_CALL
}

extern "C" void hash_rng_set_epoch(const unsigned long long *device_counter) { g_hashrng_epoch.store(device_counter); }

extern "C" {

long long fused_ln_workspace_floats(int channels) { return channels > 0 ? (long long)kCtas * 2 * channels : 0; }

int fused_ln_forward(void *stream, const float *a, const float *b, const float *gamma, const float *beta, long long rows, int channels,
                     float eps, float p_drop, unsigned long long seed, float *z, float *y, float *mean, float *rstd)
{
  return forward_impl<float>(stream, a, b, gamma, beta, rows, channels, eps, p_drop, seed, z, y, mean, rstd);
}

int fused_ln_backward(void *stream, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd, long long rows,
                      int channels, float p_drop, unsigned long long seed, float *da, float *db, float *dgamma, float *dbeta, float *workspace)
{
  if (db != nullptr && p_drop <= 0.f) return MSDA3D_EINVAL;          // without dropout db == da: pass NULL
  return backward_impl<float>(stream, dy, z, gamma, mean, rstd, rows, channels, p_drop, seed, da, db, dgamma, dbeta, workspace);
}

/* bf16 branch: b and db are bf16 (8-byte aligned rows), everything else as above.  db is always written (it cannot alias the fp32 da). */
int fused_ln_forward_bf16b(void *stream, const float *a, const void *b, const float *gamma, const float *beta, long long rows, int channels,
                           float eps, float p_drop, unsigned long long seed, float *z, float *y, float *mean, float *rstd)
{
  if (b == nullptr) return MSDA3D_EINVAL;
  return forward_impl<__nv_bfloat16>(stream, a, (const __nv_bfloat16 *)b, gamma, beta, rows, channels, eps, p_drop, seed, z, y, mean, rstd);
}

int fused_ln_backward_bf16b(void *stream, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd, long long rows,
                            int channels, float p_drop, unsigned long long seed, float *da, void *db, float *dgamma, float *dbeta, float *workspace)
{
  if (db == nullptr) return MSDA3D_EINVAL;
  return backward_impl<__nv_bfloat16>(stream, dy, z, gamma, mean, rstd, rows, channels, p_drop, seed, da, (__nv_bfloat16 *)db, dgamma, dbeta, workspace);
}

}  // extern "C"
