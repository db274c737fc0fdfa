// win_attn_capi.cu -- C ABI of the Swin3D window attention (include/win_attn.h).
#include "win_attn_kernels.cuh"

#include <atomic>

#include "../../include/msda3d.h"
#include "../../include/win_attn.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

namespace {
bool bad(int Bw, int n, int H, int hd, int nW, const void *mask)
{
  return Bw <= 0 || n <= 0 || n > winattn::kThreads || H <= 0 || H > 65535 || hd != winattn::HD || (mask != nullptr && (nW <= 0 || Bw % nW != 0));
}
int sm_count_()
{
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}
}  // namespace

extern "C" {

int win_attn_supported(int tokens, int head_dim) { return tokens > 0 && tokens <= winattn::kThreads && head_dim == winattn::HD; }

int win_attn_forward(void *stream, const float *qkv, const float *bias_t, const float *mask, int windows, int tokens, int heads, int head_dim,
                     int mask_windows, float scale, float *out, float *lse)
{
  if (!qkv || !bias_t || !out || !lse || bad(windows, tokens, heads, head_dim, mask_windows, mask)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) return MSDA3D_EALIGN;
  winattn::fwd_kernel<<<dim3(windows, heads), winattn::kThreads, 0, (cudaStream_t)stream>>>(qkv, bias_t, mask, tokens, heads,
                                                                                           mask ? mask_windows : 1, scale, out, lse);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

int win_attn_backward(void *stream, const float *qkv, const float *bias, const float *bias_t, const float *mask, const float *out,
                      const float *dout, const float *lse, int windows, int tokens, int heads, int head_dim, int mask_windows, float scale,
                      float *dqkv, float *dbias)
{
  if (!qkv || !bias || !bias_t || !out || !dout || !lse || !dqkv || !dbias || bad(windows, tokens, heads, head_dim, mask_windows, mask))
    return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15)
    return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(4 * winattn::kThreads * winattn::HD + 2 * winattn::kThreads + tokens * (tokens | 1)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(winattn::bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(dbias, 0, (size_t)heads * tokens * tokens * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  // enough CTAs to fill the machine twice over (2 CTAs of ~96 KB fit an SM), never more than there are windows
  long long chunks = (2LL * sm_count_() + heads - 1) / heads;
  if (chunks > windows) chunks = windows;
  if (chunks < 1) chunks = 1;
  winattn::bwd_kernel<<<dim3((unsigned)chunks, heads), winattn::kBwdThreads, smem, st>>>(qkv, bias, bias_t, mask, out, dout, lse, windows, tokens, heads,
                                                                                      mask ? mask_windows : 1, scale, dqkv, dbias);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

}  // extern "C"
