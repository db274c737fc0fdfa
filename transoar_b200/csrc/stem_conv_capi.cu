// stem_conv_capi.cu -- C ABI of the encoder's first convolution (include/stem_conv.h).
#include "stem_conv_kernels.cuh"

#include <atomic>

#include "../../include/msda3d.h"
#include "../../include/stem_conv.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

namespace {

constexpr int kWgCtas = 148 * 4;

bool co_ok(int co) { return co == 16 || co == 24 || co == 32; }

template <int CO> int fwd(cudaStream_t st, const float *x, const float *w, int N, int D, int H, int W, float *y)
{
  const long long jobs = (long long)N * D * H * ((W + stemconv::kFwdVox - 1) / stemconv::kFwdVox);
  const int grid = (int)(jobs < 148LL * 32 ? jobs : 148LL * 32);
  stemconv::fwd_kernel<CO><<<grid, stemconv::kFwdThreads, 0, st>>>(x, w, N, D, H, W, y);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

template <int CO> int wgrad(cudaStream_t st, const float *dy, const float *x, int N, int D, int H, int W, float *dw, float *ws)
{
  stemconv::wgrad_partial_kernel<CO><<<kWgCtas, stemconv::kWgThreads, 0, st>>>(dy, x, N, D, H, W, ws);
  stemconv::wgrad_finalize_kernel<<<(27 * CO + 127) / 128, 128, 0, st>>>(ws, kWgCtas, CO, dw);
  g_msda3d_launches += 2;
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

long long stem_conv3d_workspace_floats(int out_channels) { return co_ok(out_channels) ? (long long)kWgCtas * 27 * out_channels : 0; }

int stem_conv3d_forward(void *stream, const float *x, const float *weight, int batch, int depth, int height, int width, int out_channels,
                        float *y)
{
  if (!x || !weight || !y || batch <= 0 || depth <= 0 || height <= 0 || width <= 0 || !co_ok(out_channels)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(x) & 3)) return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_channels) {
    case 16: return fwd<16>(st, x, weight, batch, depth, height, width, y);
    case 24: return fwd<24>(st, x, weight, batch, depth, height, width, y);
    default: return fwd<32>(st, x, weight, batch, depth, height, width, y);
  }
}

int stem_conv3d_wgrad(void *stream, const float *dy, const float *x, int batch, int depth, int height, int width, int out_channels,
                      float *dweight, float *workspace)
{
  if (!dy || !x || !dweight || !workspace || batch <= 0 || depth <= 0 || height <= 0 || width <= 0 || !co_ok(out_channels)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(x) & 3)) return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_channels) {
    case 16: return wgrad<16>(st, dy, x, batch, depth, height, width, dweight, workspace);
    case 24: return wgrad<24>(st, dy, x, batch, depth, height, width, dweight, workspace);
    default: return wgrad<32>(st, dy, x, batch, depth, height, width, dweight, workspace);
  }
}

}  // extern "C"
