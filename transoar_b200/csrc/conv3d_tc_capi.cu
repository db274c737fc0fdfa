// conv3d_tc_capi.cu -- C ABI of the tcgen05 3x3x3 convolution (include/conv3d_tc.h): 5-D tensor map over the NDHWC input, launch.
#include "conv3d_tc_kernels.cuh"

#include <atomic>
#include <mutex>

#include "../../include/conv3d_tc.h"
#include "../../include/msda3d.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

namespace {

std::atomic<int> g_conv_diag{0};

using EncodeTiled = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

// the ring of halo planes + the weights of all 27 taps (rows padded to 32 channels) fill the 227 KB of shared memory; 8, 16 or 24 input channels
bool ci_ok(int ci) { return ci == 8 || ci == 16 || ci == 24; }
static_assert(convtc::ConvCfg<24>::SMEM_BYTES <= 227 * 1024, "shared-memory budget");

template <int CI>
int launch(cudaStream_t st, const CUtensorMap &mx, const float *w, float *y, int N, int D, int H, int W, int CO)
{
  using C = convtc::ConvCfg<CI>;
  auto kern = convtc::conv3d_k3_kernel<CI>;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES); });
  if (err != cudaSuccess) return (int)err;
  const long long tiles = (long long)N * ((D + convtc::DSEG - 1) / convtc::DSEG) * ((H + convtc::FTH - 1) / convtc::FTH) * ((W + convtc::FWO - 1) / convtc::FWO);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, convtc::kThreadsConv, C::SMEM_BYTES, st>>>(mx, w, y, N, D, H, W, CO, g_conv_diag.load());
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" void conv3d_tc_debug_mode(int mode) { g_conv_diag.store(mode); }

extern "C" int conv3d_tc_supported(int in_channels, int out_channels)
{
  return ci_ok(in_channels) && out_channels > 0 && out_channels % 4 == 0 && out_channels <= convtc::NPAD;
}

extern "C" int conv3d_tc_k3_forward(void *stream, const float *x, const float *w_taps, int batch, int depth, int height, int width,
                                    int in_channels, int out_channels, float *y)
{
  if (!x || !w_taps || !y || batch <= 0 || depth <= 0 || height <= 0 || width <= 0) return MSDA3D_EINVAL;
  if (!conv3d_tc_supported(in_channels, out_channels)) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w_taps)) & 15) return MSDA3D_EALIGN;
  ensure_context_on_this_thread();
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t C = (cuuint64_t)in_channels, W = (cuuint64_t)width, H = (cuuint64_t)height, D = (cuuint64_t)depth;
  const cuuint64_t gdim[5] = {C, W, H, D, (cuuint64_t)batch};
  const cuuint64_t gstride[4] = {C * 4, W * C * 4, H * W * C * 4, D * H * W * C * 4};
  // one depth plane of the halo per copy, rows of 32 channels (the copy engine zero-fills channels >= CI), 128-byte swizzle
  const cuuint32_t box[5] = {32, (cuuint32_t)convtc::FWI, (cuuint32_t)convtc::FHH, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap mx;
  // TFLOAT32: the copy engine rounds the activations to TF32; out-of-bounds voxels of the halo box are zero-filled (= padding 1)
  if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float *>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MSDA3D_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (in_channels) {
    case 8: return launch<8>(st, mx, w_taps, y, batch, depth, height, width, out_channels);
    case 16: return launch<16>(st, mx, w_taps, y, batch, depth, height, width, out_channels);
    default: return launch<24>(st, mx, w_taps, y, batch, depth, height, width, out_channels);
  }
}

namespace {
// channels-last fp32 volume [N, D, H, W, C] as rows of 32 channels (C <= 32: the copy engine zero-fills the rest), box = a tile of
// bw x bh x bd voxels, 128-byte swizzle with 32-byte atoms (the MN-major TF32 operand layout)
int make_row_map(CUtensorMap *map, const float *ptr, int N, int D, int H, int W, int C, int bw, int bh, int bd)
{
  ensure_context_on_this_thread();
  EncodeTiled enc = encode_fn();
  if (enc == nullptr) return MSDA3D_ENODEV;
  const cuuint64_t c = (cuuint64_t)C, w = (cuuint64_t)W, h = (cuuint64_t)H, d = (cuuint64_t)D;
  const cuuint64_t gdim[5] = {c, w, h, d, (cuuint64_t)N};
  const cuuint64_t gstride[4] = {c * 4, w * c * 4, h * w * c * 4, d * h * w * c * 4};
  const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 5, const_cast<float *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : MSDA3D_EINVAL;
}
constexpr int kWgCtas = 148;
}  // namespace

extern "C" long long conv3d_tc_wgrad_workspace_floats(void) { return (long long)kWgCtas * 9 * 128 * 32; }

extern "C" int conv3d_tc_k3_wgrad(void *stream, const float *x, const float *dy, int batch, int depth, int height, int width, int in_channels,
                                  int out_channels, float *dweight, float *workspace)
{
  if (!x || !dy || !dweight || !workspace || batch <= 0 || depth <= 0 || height <= 0 || width <= 0) return MSDA3D_EINVAL;
  if (in_channels <= 0 || in_channels > 32 || in_channels % 4 || out_channels <= 0 || out_channels > 32 || out_channels % 4) return MSDA3D_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(workspace)) & 15) return MSDA3D_EALIGN;
  CUtensorMap mx, my;
  int rc = make_row_map(&mx, x, batch, depth, height, width, in_channels, convtc::HW, convtc::HH, 3);
  if (rc != 0) return rc;
  rc = make_row_map(&my, dy, batch, depth, height, width, out_channels, convtc::TW, convtc::TH, 1);
  if (rc != 0) return rc;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [&] { err = cudaFuncSetAttribute(convtc::conv3d_k3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, convtc::kWgSmemBytes); });
  if (err != cudaSuccess) return (int)err;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long tiles = (long long)batch * depth * ((height + convtc::TH - 1) / convtc::TH) * ((width + convtc::TW - 1) / convtc::TW);
  const int grid = (int)(tiles < kWgCtas ? tiles : kWgCtas);
  convtc::conv3d_k3_wgrad_kernel<<<grid, convtc::kThreadsConv, convtc::kWgSmemBytes, st>>>(mx, my, workspace, batch, depth, height, width);
  const int total = out_channels * in_channels * 27;
  convtc::conv3d_k3_wgrad_finalize_kernel<<<(total + 31) / 32, 256, 0, st>>>(workspace, grid, in_channels, out_channels, dweight);
  g_msda3d_launches += 2;
  return (int)cudaGetLastError();
}

// tests only: one 128 x 32 x 8 MMA whose A operand is four overlapping 32-channel slabs of a [24][32] row matrix (see the kernel)
extern "C" int conv3d_tc_debug_mn_probe(void *stream, const float *X, const float *Y, float *D, int row0)
{
  if (!X || !Y || !D || row0 < 0 || row0 + 3 + 8 > 24) return MSDA3D_EINVAL;
  convtc::mn_sw32_probe_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(X, Y, D, row0);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}
