// instnorm_capi.cu -- C ABI of the fused InstanceNorm3d + ReLU (include/instnorm.h).
#include "instnorm_kernels.cuh"

#include <atomic>

#include "../../include/instnorm.h"
#include "../../include/msda3d.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

namespace {

template <typename T> int chunks_of(long long V) { return (int)((V + instnorm::chunk_elems<T>() - 1) / instnorm::chunk_elems<T>()); }

template <typename T> bool vec_ok(long long V, const void *a, const void *b, const void *c, const void *d)
{
  auto al = [](const void *p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  return V % instnorm::Pack<T>::N == 0 && al(a) && al(b) && al(c) && al(d);
}

template <typename T>
int fwd(cudaStream_t st, const void *x, const float *gamma, const float *beta, int B, int C, long long V, float eps, void *y, float *mean,
        float *rstd, float *ws)
{
  const int I = B * C, ch = chunks_of<T>(V), ok = vec_ok<T>(V, x, y, nullptr, nullptr);
  const dim3 grid(ch, I);
  instnorm::stats_partial_kernel<T><<<grid, instnorm::kThreads, 0, st>>>((const T *)x, V, ch, ok, ws);
  instnorm::stats_finalize_kernel<<<(I + 3) / 4, 128, 0, st>>>(ws, ch, I, eps, mean, rstd);
  instnorm::apply_kernel<T><<<grid, instnorm::kThreads, 0, st>>>((const T *)x, gamma, beta, mean, rstd, V, C, ok, (T *)y);
  g_msda3d_launches += 3;
  return (int)cudaGetLastError();
}

template <typename T>
int bwd(cudaStream_t st, const void *dy, const void *x, const void *y, const float *gamma, const float *mean, const float *rstd, int B, int C,
        long long V, void *dx, float *dgamma, float *dbeta, float *ws)
{
  const int I = B * C, ch = chunks_of<T>(V), ok = vec_ok<T>(V, dy, x, y, dx);
  const dim3 grid(ch, I);
  float *sums = ws + (long long)I * ch * 2;
  instnorm::bwd_partial_kernel<T><<<grid, instnorm::kThreads, 0, st>>>((const T *)dy, (const T *)x, (const T *)y, mean, rstd, V, ch, ok, ws);
  instnorm::bwd_finalize_kernel<<<(I + 3) / 4, 128, 0, st>>>(ws, ch, I, sums);
  instnorm::bwd_param_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, B, C, dgamma, dbeta);
  instnorm::bwd_apply_kernel<T><<<grid, instnorm::kThreads, 0, st>>>((const T *)dy, (const T *)x, (const T *)y, gamma, mean, rstd, sums, V, C,
                                                                      ok, (T *)dx);
  g_msda3d_launches += 4;
  return (int)cudaGetLastError();
}

// channels-last: chunks per sample chosen so that the grid holds ~8 CTAs per SM; every chunk is a whole number of voxel rows
struct ClPlan { int chunks; long long chunk_vox; };
ClPlan cl_plan(int B, int C, long long V)
{
  const int rows = instnorm::kClThreads / (C / 4);
  const long long min_vox = (long long)rows * instnorm::kClUnroll;
  long long chunks = (V + min_vox - 1) / min_vox;
  const long long target = (148 * 8 + B - 1) / B;
  if (chunks > target) chunks = target;
  if (chunks < 1) chunks = 1;
  long long cv = (V + chunks - 1) / chunks;
  cv = (cv + rows - 1) / rows * rows;
  ClPlan p;
  p.chunk_vox = cv;
  p.chunks = (int)((V + cv - 1) / cv);
  return p;
}

bool cl_bad(int B, int C, long long V) { return B <= 0 || V <= 0 || !instnorm::cl_supported(C) || (long long)B * C > 65535; }

bool bad(int dtype, int B, int C, long long V) { return (dtype != MSDA3D_F32 && dtype != MSDA3D_BF16) || B <= 0 || C <= 0 || V <= 0 || (long long)B * C > 65535; }

}  // namespace

template <typename T>
static int cl_forward(void *stream, const T *x, const float *gamma, const float *beta, int batch, int channels, long long voxels,
                      float eps, T *y, float *mean, float *rstd, float *workspace)
{
  if (!x || !gamma || !beta || !y || !mean || !rstd || !workspace || cl_bad(batch, channels, voxels)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & (4 * sizeof(T) - 1)) return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const ClPlan p = cl_plan(batch, channels, voxels);
  const int I = batch * channels;
  const dim3 grid(p.chunks, batch);
  instnorm::cl_stats_partial_kernel<T><<<grid, instnorm::kClThreads, 0, st>>>(x, voxels, channels, p.chunks, p.chunk_vox, workspace);
  instnorm::stats_finalize_kernel<<<(I + 3) / 4, 128, 0, st>>>(workspace, p.chunks, I, eps, mean, rstd);
  instnorm::cl_apply_kernel<T><<<grid, instnorm::kClThreads, 0, st>>>(x, gamma, beta, mean, rstd, voxels, channels, p.chunk_vox, y);
  g_msda3d_launches += 3;
  return (int)cudaGetLastError();
}

template <typename T>
static int cl_backward(void *stream, const T *dy, const T *x, const float *gamma, const float *beta, const float *mean, const float *rstd,
                       int batch, int channels, long long voxels, T *dx, float *dgamma, float *dbeta, float *workspace)
{
  if (!dy || !x || !gamma || !beta || !mean || !rstd || !dx || !dgamma || !dbeta || !workspace || cl_bad(batch, channels, voxels))
    return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & (4 * sizeof(T) - 1))
    return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const ClPlan p = cl_plan(batch, channels, voxels);
  const int I = batch * channels;
  const dim3 grid(p.chunks, batch);
  float *sums = workspace + (long long)I * p.chunks * 2;
  instnorm::cl_bwd_partial_kernel<T><<<grid, instnorm::kClThreads, 0, st>>>(dy, x, gamma, beta, mean, rstd, voxels, channels, p.chunks, p.chunk_vox, workspace);
  instnorm::bwd_finalize_kernel<<<(I + 3) / 4, 128, 0, st>>>(workspace, p.chunks, I, sums);
  instnorm::bwd_param_kernel<<<(channels + 127) / 128, 128, 0, st>>>(sums, batch, channels, dgamma, dbeta);
  instnorm::cl_bwd_apply_kernel<T><<<grid, instnorm::kClThreads, 0, st>>>(dy, x, gamma, beta, mean, rstd, sums, voxels, channels, p.chunk_vox, dx);
  g_msda3d_launches += 4;
  return (int)cudaGetLastError();
}

extern "C" {

long long instnorm_workspace_floats(int dtype, int batch, int channels, long long voxels)
{
  if (bad(dtype, batch, channels, voxels)) return 0;
  const long long ch = dtype == MSDA3D_F32 ? chunks_of<float>(voxels) : chunks_of<__nv_bfloat16>(voxels);
  return (long long)batch * channels * (3 * ch + 2);
}

int instnorm_relu_forward(void *stream, int dtype, const void *x, const float *gamma, const float *beta, int batch, int channels,
                          long long voxels, float eps, void *y, float *mean, float *rstd, float *workspace)
{
  if (!x || !gamma || !beta || !y || !mean || !rstd || !workspace) return MSDA3D_EINVAL;
  if (bad(dtype, batch, channels, voxels)) return MSDA3D_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == MSDA3D_F32 ? fwd<float>(st, x, gamma, beta, batch, channels, voxels, eps, y, mean, rstd, workspace)
                             : fwd<__nv_bfloat16>(st, x, gamma, beta, batch, channels, voxels, eps, y, mean, rstd, workspace);
}

int instnorm_relu_backward(void *stream, int dtype, const void *dy, const void *x, const void *y, const float *gamma, const float *mean,
                           const float *rstd, int batch, int channels, long long voxels, void *dx, float *dgamma, float *dbeta,
                           float *workspace)
{
  if (!dy || !x || !y || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta || !workspace) return MSDA3D_EINVAL;
  if (bad(dtype, batch, channels, voxels)) return MSDA3D_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == MSDA3D_F32 ? bwd<float>(st, dy, x, y, gamma, mean, rstd, batch, channels, voxels, dx, dgamma, dbeta, workspace)
                             : bwd<__nv_bfloat16>(st, dy, x, y, gamma, mean, rstd, batch, channels, voxels, dx, dgamma, dbeta, workspace);
}

long long instnorm_ndhwc_workspace_floats(int batch, int channels, long long voxels)
{
  if (cl_bad(batch, channels, voxels)) return 0;
  return (long long)batch * channels * (3LL * cl_plan(batch, channels, voxels).chunks + 2);
}

int instnorm_relu_forward_ndhwc(void *stream, const float *x, const float *gamma, const float *beta, int batch, int channels, long long voxels,
                                float eps, float *y, float *mean, float *rstd, float *workspace)
{
  return cl_forward<float>(stream, x, gamma, beta, batch, channels, voxels, eps, y, mean, rstd, workspace);
}

int instnorm_relu_backward_ndhwc(void *stream, const float *dy, const float *x, const float *gamma, const float *beta, const float *mean,
                                 const float *rstd, int batch, int channels, long long voxels, float *dx, float *dgamma, float *dbeta,
                                 float *workspace)
{
  return cl_backward<float>(stream, dy, x, gamma, beta, mean, rstd, batch, channels, voxels, dx, dgamma, dbeta, workspace);
}

/* bf16 storage (x, y, dy, dx), fp32 parameters / statistics / arithmetic: the activations of the encoder under the bf16 autocast route */
int instnorm_relu_forward_ndhwc_bf16(void *stream, const void *x, const float *gamma, const float *beta, int batch, int channels, long long voxels,
                                     float eps, void *y, float *mean, float *rstd, float *workspace)
{
  return cl_forward<__nv_bfloat16>(stream, (const __nv_bfloat16 *)x, gamma, beta, batch, channels, voxels, eps, (__nv_bfloat16 *)y, mean, rstd, workspace);
}

int instnorm_relu_backward_ndhwc_bf16(void *stream, const void *dy, const void *x, const float *gamma, const float *beta, const float *mean,
                                      const float *rstd, int batch, int channels, long long voxels, void *dx, float *dgamma, float *dbeta,
                                      float *workspace)
{
  return cl_backward<__nv_bfloat16>(stream, (const __nv_bfloat16 *)dy, (const __nv_bfloat16 *)x, gamma, beta, mean, rstd, batch, channels, voxels,
                                    (__nv_bfloat16 *)dx, dgamma, dbeta, workspace);
}

}  // extern "C"
