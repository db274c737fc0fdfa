// msda3d_capi.cu -- the C ABI declared in include/msda3d.h: argument checks, kernel selection, launches, and the
// host-buffer staging path.  No torch types anywhere; linked against the static CUDA runtime only.
//
// Mirrors (and replaces) the reference's host side:
//   transoar/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80,83-154   (asserts, shapes, im2col_step loop, zero-init)
//   transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:1094-1125   (forward launcher)
//   transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:1127-1507   (backward launcher + channel dispatch table)
#include "msda3d_kernels.cuh"

#include <atomic>
#include <type_traits>
#include <mutex>
#include <string>

#include "../../include/msda3d.h"

using namespace msda3d;

std::atomic<unsigned long long> g_msda3d_launches{0};   // shared with roi_attn_capi.cu
extern std::atomic<int> g_roi_splits;                    // roi_attn_capi.cu: token splits per box (msda3d_set_tuning "roi_splits")

namespace {

std::atomic<unsigned long long> &g_launches = g_msda3d_launches;
std::atomic<int> g_diag_skip_red{0};   // diagnostics only (msda3d_set_tuning): drop the grad_value reductions

struct Dims {
  int N, S, M, C, L, Lq, P;
};

int check_dims(const Dims &d)
{
  if (d.N <= 0 || d.S <= 0 || d.M <= 0 || d.C <= 0 || d.L <= 0 || d.Lq <= 0 || d.P <= 0) return MSDA3D_EINVAL;
  // generic kernels index with 64-bit arithmetic; only the per-sample count must fit the loop counters
  const long long lp = (long long)d.L * d.P;
  if (lp > (1 << 20)) return MSDA3D_ERANGE;
  return MSDA3D_OK;
}

size_t elem_size(int dtype) { return dtype == MSDA3D_F64 ? 8 : dtype == MSDA3D_F32 ? 4 : 2; }
size_t aux_size(int dtype) { return dtype == MSDA3D_F64 ? 8 : 4; }   // loc / aw / grad_loc / grad_aw / grad_value(16-bit)

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int grid_for(long long units, int units_per_block)
{
  long long g = (units + units_per_block - 1) / units_per_block;
  const long long cap = 148LL * 8 * 64;   // grid-stride loops cover the rest; keeps index math away from 2^31
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Vector-kernel eligibility: C = G * NV * VEC with G a power of two <= 32, <= 16 levels, 16-byte aligned slabs and
// 32-bit element offsets over the whole value tensor.  NV (16-byte vectors per lane) is 1 unless the channel count
// needs 2 to fit a warp (measured on B200, profiles/r01_variants.txt: at C = 64 both are within 2 %, NV = 1 keeps
// registers at 64 / 80 and four / three CTAs per SM resident).
std::atomic<int> g_tune_nv{0};          // 0 = automatic, 1 / 2 = forced (msda3d_set_tuning "nv")
std::atomic<int> g_tune_grid_mult{0};   // 0 = automatic: CTAs per SM for the persistent grid
std::atomic<int> g_tune_order{0};       // 0 = automatic (brick order when Lq == S), 1 = linear, 2 = brick
std::atomic<int> g_tune_stage{0};       // experiment: forward with the coarsest level staged in shared memory by TMA bulk copies (2 / 3 / 4 = CTAs per SM)
std::atomic<int> g_tune_rot{0};         // backward: sample-order rotation 1 = per warp, 2 = per unit; + 4 = consecutive CTAs on different (batch, head) slabs (kernels.cuh, ROT)
std::atomic<int> g_tune_duo{1};         // backward with two w-neighbouring queries per lane group (kernels.cuh, bwd_duo_kernel): 1 = on (default), 0 = bwd_vec_kernel
std::atomic<int> g_tune_duo_cfg{0};     // experiment: CTA shape of bwd_duo_kernel: 0 = 256 threads x 2 per SM, 1 = 128 x 5 (96 registers), 2 = 128 x 6 (80)
std::atomic<int> g_tune_pair{0};        // 1 = pair-combining backward (kernels.cuh, PAIR) in brick order for 16-lane fp32 units; measured slower, off by default

template <typename VT> bool vec_shape(int C, int &G, int &NV)
{
  constexpr int VEC = Vec16<VT>::N;
  if (C % VEC) return false;
  const int q = C / VEC;
  const int forced = g_tune_nv.load();
  for (int nv = (forced ? forced : 1); nv <= 2; ++nv) {
    if (q % nv == 0) {
      const int g = q / nv;
      if (g <= 32 && (g & (g - 1)) == 0) { G = g; NV = nv; return true; }
    }
    if (forced) break;
  }
  return false;
}

template <typename VT>
bool vec_ok(const Dims &d, const void *a, const void *b, int &G, int &NV)
{
  if (!vec_shape<VT>(d.C, G, NV)) return false;
  if (d.L > kMaxLevels) return false;
  if ((long long)d.N * d.S * d.M * d.C >= (1LL << 31)) return false;
  if ((long long)d.L * d.P * 3 >= (1LL << 20)) return false;
  return aligned(a, 4 * sizeof(VT)) && aligned(b, 4 * sizeof(VT));
}

int use_brick(const Dims &d)
{
  const int o = g_tune_order.load();
  return (o == 1) ? 0 : (d.Lq == d.S ? 1 : 0);
}

int sm_count()
{
  static int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (!cached[dev]) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cached[dev];
}

// Grid: enough CTAs for every warp slot, capped at a multiple of the SM count (grid-stride loop covers the rest).
int vec_grid(long long units, int G)
{
  const int upb = (kThreads / 32) * (32 / G);
  long long g = (units + upb - 1) / upb;
  const int mult = g_tune_grid_mult.load();
  const long long cap = (long long)sm_count() * (mult > 0 ? mult : 64);
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

#define VEC_CASE(G_, NV_, ...) \
  case (G_) * 4 + (NV_): { constexpr int G = G_, NV = NV_; __VA_ARGS__; } break;
#define VEC_DISPATCH(g, nv, ...)                                                                              \
  switch ((g) * 4 + (nv)) {                                                                                   \
    VEC_CASE(1, 1, __VA_ARGS__) VEC_CASE(2, 1, __VA_ARGS__) VEC_CASE(4, 1, __VA_ARGS__) VEC_CASE(8, 1, __VA_ARGS__)   \
    VEC_CASE(16, 1, __VA_ARGS__) VEC_CASE(32, 1, __VA_ARGS__)                                                 \
    VEC_CASE(1, 2, __VA_ARGS__) VEC_CASE(2, 2, __VA_ARGS__) VEC_CASE(4, 2, __VA_ARGS__) VEC_CASE(8, 2, __VA_ARGS__)   \
    VEC_CASE(16, 2, __VA_ARGS__) VEC_CASE(32, 2, __VA_ARGS__)                                                 \
    default: return MSDA3D_EINVAL;                                                                            \
  }

// Resident CTAs per SM the kernels are compiled for (register caps 64 / 80 / 128 / 255 per thread).  Measured on B200
// (profiles/r01_variants.txt): the fp32 forward gains 20 % going from 2-3 to 4 CTAs per SM (latency hiding), the
// backward is bound by the reduction traffic into L2 and only needs 3.  Wider per-lane vectors get a looser cap so
// that ptxas does not spill.
template <typename VT, int NV> struct MinBlocks {
  static constexpr int fwd = NV == 1 ? 4 : 3;
  static constexpr int bwd = NV == 1 ? 3 : 2;
};


template <typename VT>
int forward_half_or_float(cudaStream_t st, const Dims &d, const void *value, const int64_t *shapes, const int64_t *starts,
                          const void *loc, const void *aw, void *out)
{
  const long long units = (long long)d.N * d.Lq * d.M;
  int g_ = 0, nv_ = 0;
  if (vec_ok<VT>(d, value, out, g_, nv_)) {
    const int grid = vec_grid(units, g_);
    const int stage = g_tune_stage.load();
    if (stage && std::is_same<VT, float>::value && g_ == 16 && nv_ == 1 && use_brick(d) && d.L >= 1) {
      // experiment: coarsest level staged in shared memory by TMA bulk copies (kernels.cuh, fwd_stage_kernel); needs the level's shape on the host
      int64_t shp[3];                                                  // experiment only: the level's extent is read back per call (a host sync)
      if (cudaMemcpyAsync(shp, shapes + 3 * (d.L - 1), sizeof(shp), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
        return (int)cudaGetLastError();
      const size_t smem = (size_t)shp[0] * shp[1] * shp[2] * d.C * sizeof(float);
      if (smem <= 100 * 1024) {
        auto launch = [&](auto kern) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          kern<<<grid, kThreads, smem, st>>>((const float *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.L,
                                             d.Lq, d.P, (float *)out);
        };
        if (stage == 2) launch(fwd_stage_kernel<2>); else if (stage == 4) launch(fwd_stage_kernel<4>); else launch(fwd_stage_kernel<3>);
        ++g_launches;
        return (int)cudaGetLastError();
      }
    }
    VEC_DISPATCH(g_, nv_, fwd_vec_kernel<VT, G, NV, MinBlocks<VT, NV>::fwd><<<grid, kThreads, 0, st>>>(
                              (const VT *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.L, d.Lq,
                              d.P, (VT *)out, use_brick(d)));
  } else {
    fwd_generic_kernel<VT, float><<<grid_for(units, kThreads / 32), kThreads, 0, st>>>(
        (const VT *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.C, d.L, d.Lq, d.P, (VT *)out);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

template <typename VT>
int backward_half_or_float(cudaStream_t st, const Dims &d, const void *gout, const void *value, const int64_t *shapes,
                           const int64_t *starts, const void *loc, const void *aw, void *gv, void *gl, void *ga)
{
  const long long units = (long long)d.N * d.Lq * d.M;
  cudaError_t e = cudaMemsetAsync(gv, 0, (size_t)d.N * d.S * d.M * d.C * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  int g_ = 0, nv_ = 0;
  if (vec_ok<VT>(d, value, gout, g_, nv_) && aligned(gv, 16)) {
    const int grid = vec_grid(units, g_);
    const int skip = g_diag_skip_red.load();
    if (skip >= 2 && skip <= 4 && g_ == 16 && nv_ == 1 && std::is_same<VT, float>::value) {
      // diagnostics: 2 = every fourth sample reduces, 3 = all but the coarsest level, 4 = all but the two coarsest levels
#define MSDA3D_DIAG(SK)                                                                                                             \
  bwd_vec_kernel<float, 16, 1, 3, SK><<<grid, kThreads, 0, st>>>((const float *)gout, (const float *)value, shapes, starts,       \
                                                                 (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.L, d.Lq,  \
                                                                 d.P, (float *)gv, (float *)gl, (float *)ga, use_brick(d))
      if (skip == 2) MSDA3D_DIAG(2); else if (skip == 3) MSDA3D_DIAG(3); else MSDA3D_DIAG(4);
#undef MSDA3D_DIAG
    } else if (g_diag_skip_red.load() && !(std::is_same<VT, float>::value && g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_duo.load() == 1)) {
      VEC_DISPATCH(g_, nv_, bwd_vec_kernel<VT, G, NV, MinBlocks<VT, NV>::bwd, 1><<<grid, kThreads, 0, st>>>(
                                (const VT *)gout, (const VT *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N,
                                d.S, d.M, d.L, d.Lq, d.P, (float *)gv, (float *)gl, (float *)ga, use_brick(d)));
    } else if (std::is_same<VT, float>::value && g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_duo.load() == 1 && skip <= 1) {
      // two w-neighbouring queries per lane group share corner rows and reductions (kernels.cuh, bwd_duo_kernel)
#define MSDA3D_DUO(R, SK)                                                                                                              \
  bwd_duo_kernel<0, R, SK><<<grid, kThreads, 0, st>>>((const float *)gout, (const float *)value, shapes, starts, (const float *)loc,   \
                                                      (const float *)aw, d.N, d.S, d.M, d.L, d.Lq, d.P, (float *)gv, (float *)gl,      \
                                                      (float *)ga, nullptr, 0, 0, 0)
      const int r = g_tune_rot.load();
      if (skip == 1) MSDA3D_DUO(6, 1);
      else if (r >= 5) MSDA3D_DUO(6, 0); else if (r == 4) MSDA3D_DUO(4, 0); else if (r >= 1) MSDA3D_DUO(2, 0); else MSDA3D_DUO(0, 0);
#undef MSDA3D_DUO
    } else if (std::is_same<VT, float>::value && g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_pair.load() == 1) {
      // w-neighbouring units of a warp combine their grad_value contributions before the reductions (kernels.cuh, PAIR)
      bwd_vec_kernel<float, 16, 1, 2, 0, 0, 1><<<grid, kThreads, 0, st>>>(
          (const float *)gout, (const float *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.L, d.Lq, d.P,
          (float *)gv, (float *)gl, (float *)ga, 1);
    } else if (std::is_same<VT, float>::value && g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_rot.load() != 0) {
      // the units of a CTA walk their samples in rotated order (kernels.cuh, ROT)
#define MSDA3D_ROT(R)                                                                                                                     \
  bwd_vec_kernel<float, 16, 1, 3, 0, 0, 0, R><<<grid, kThreads, 0, st>>>((const float *)gout, (const float *)value, shapes, starts,      \
                                                                         (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.L, d.Lq, \
                                                                         d.P, (float *)gv, (float *)gl, (float *)ga, 1)
      switch (g_tune_rot.load()) { case 1: MSDA3D_ROT(1); break; case 2: MSDA3D_ROT(2); break; case 4: MSDA3D_ROT(4); break; case 5: MSDA3D_ROT(5); break; default: MSDA3D_ROT(6); break; }
#undef MSDA3D_ROT
    } else {
      VEC_DISPATCH(g_, nv_, bwd_vec_kernel<VT, G, NV, MinBlocks<VT, NV>::bwd><<<grid, kThreads, 0, st>>>(
                                (const VT *)gout, (const VT *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N,
                                d.S, d.M, d.L, d.Lq, d.P, (float *)gv, (float *)gl, (float *)ga, use_brick(d)));
    }
  } else {
    bwd_generic_kernel<VT, float><<<grid_for(units, kThreads / 32), kThreads, 0, st>>>(
        (const VT *)gout, (const VT *)value, shapes, starts, (const float *)loc, (const float *)aw, d.N, d.S, d.M, d.C, d.L, d.Lq,
        d.P, (float *)gv, (float *)gl, (float *)ga);
  }
  ++g_launches;
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int msda3d_abi_version(void) { return MSDA3D_ABI_VERSION; }

unsigned long long msda3d_launch_count(void) { return g_launches.load(); }

int msda3d_set_tuning(const char *key, int value)
{
  if (!key) return MSDA3D_EINVAL;
  const std::string k(key);
  if (k == "diag_bwd_skip_red") { g_diag_skip_red = value; return MSDA3D_OK; }
  if (k == "nv" && value >= 0 && value <= 2) { g_tune_nv = value; return MSDA3D_OK; }
  if (k == "grid_mult" && value >= 0) { g_tune_grid_mult = value; return MSDA3D_OK; }
  if (k == "stage" && value >= 0 && value <= 4) { g_tune_stage = value; return MSDA3D_OK; }
  if (k == "order" && value >= 0 && value <= 2) { g_tune_order = value; return MSDA3D_OK; }
  if (k == "pair" && value >= 0 && value <= 1) { g_tune_pair = value; return MSDA3D_OK; }
  if (k == "duo" && value >= 0 && value <= 1) { g_tune_duo = value; return MSDA3D_OK; }
  if (k == "roi_splits" && value >= 0 && value <= 16) { g_roi_splits = value; return MSDA3D_OK; }
  if (k == "duo_cfg" && value >= 0 && value <= 2) { g_tune_duo_cfg = value; return MSDA3D_OK; }
  if (k == "rot" && value >= 0 && value <= 6 && value != 3) { g_tune_rot = value; return MSDA3D_OK; }
  return MSDA3D_EINVAL;
}

const char *msda3d_error_string(int code)
{
  switch (code) {
    case MSDA3D_OK: return "ok";
    case MSDA3D_EINVAL: return "msda3d: invalid argument (null pointer, non-positive dimension or unknown dtype)";
    case MSDA3D_ERANGE: return "msda3d: dimension product out of range";
    case MSDA3D_EALIGN: return "msda3d: pointer not aligned to its element type";
    case MSDA3D_ENODEV: return "msda3d: no usable CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "msda3d: unknown error";
  }
}

int msda3d_forward(void *stream, int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                   const void *sampling_loc, const void *attn_weight, int batch, int spatial_size, int num_heads, int channels,
                   int num_levels, int num_query, int num_point, void *output)
{
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output) return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int rc = check_dims(d)) return rc;
  if (dtype < MSDA3D_F32 || dtype > MSDA3D_F16) return MSDA3D_EINVAL;
  if (!aligned(value, elem_size(dtype)) || !aligned(output, elem_size(dtype)) || !aligned(sampling_loc, aux_size(dtype)) ||
      !aligned(attn_weight, aux_size(dtype)) || !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MSDA3D_F32:
      return forward_half_or_float<float>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
    case MSDA3D_BF16:
      return forward_half_or_float<__nv_bfloat16>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
    case MSDA3D_F16:
      return forward_half_or_float<__half>(st, d, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output);
    default: {
      const long long units = (long long)d.N * d.Lq * d.M;
      fwd_generic_kernel<double, double><<<grid_for(units, kThreads / 32), kThreads, 0, st>>>(
          (const double *)value, spatial_shapes, level_start_index, (const double *)sampling_loc, (const double *)attn_weight, d.N,
          d.S, d.M, d.C, d.L, d.Lq, d.P, (double *)output);
      ++g_launches;
      return (int)cudaGetLastError();
    }
  }
}

int msda3d_backward(void *stream, int dtype, const void *grad_output, const void *value, const int64_t *spatial_shapes,
                    const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                    int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point, void *grad_value,
                    void *grad_sampling_loc, void *grad_attn_weight)
{
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_value ||
      !grad_sampling_loc || !grad_attn_weight)
    return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int rc = check_dims(d)) return rc;
  if (dtype < MSDA3D_F32 || dtype > MSDA3D_F16) return MSDA3D_EINVAL;
  if (!aligned(value, elem_size(dtype)) || !aligned(grad_output, elem_size(dtype)) || !aligned(sampling_loc, aux_size(dtype)) ||
      !aligned(attn_weight, aux_size(dtype)) || !aligned(grad_value, aux_size(dtype)) || !aligned(grad_sampling_loc, aux_size(dtype)) ||
      !aligned(grad_attn_weight, aux_size(dtype)) || !aligned(spatial_shapes, 8) || !aligned(level_start_index, 8))
    return MSDA3D_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MSDA3D_F32:
      return backward_half_or_float<float>(st, d, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                           grad_value, grad_sampling_loc, grad_attn_weight);
    case MSDA3D_BF16:
      return backward_half_or_float<__nv_bfloat16>(st, d, grad_output, value, spatial_shapes, level_start_index, sampling_loc,
                                                   attn_weight, grad_value, grad_sampling_loc, grad_attn_weight);
    case MSDA3D_F16:
      return backward_half_or_float<__half>(st, d, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                            grad_value, grad_sampling_loc, grad_attn_weight);
    default: {
      const long long units = (long long)d.N * d.Lq * d.M;
      cudaError_t e = cudaMemsetAsync(grad_value, 0, (size_t)d.N * d.S * d.M * d.C * sizeof(double), st);
      if (e != cudaSuccess) return (int)e;
      bwd_generic_kernel<double, double><<<grid_for(units, kThreads / 32), kThreads, 0, st>>>(
          (const double *)grad_output, (const double *)value, spatial_shapes, level_start_index, (const double *)sampling_loc,
          (const double *)attn_weight, d.N, d.S, d.M, d.C, d.L, d.Lq, d.P, (double *)grad_value, (double *)grad_sampling_loc,
          (double *)grad_attn_weight);
      ++g_launches;
      return (int)cudaGetLastError();
    }
  }
}

// Fused prologue variants (fp32, vector kernels only, L * P <= G lanes of a unit): see include/msda3d.h.
static int fused_shape(const Dims &d, const void *value, const void *other, int &G, int &NV)
{
  if (!vec_ok<float>(d, value, other, G, NV)) return MSDA3D_EINVAL;
  if (d.L * d.P > G) return MSDA3D_EINVAL;
  return MSDA3D_OK;
}

int msda3d_fused_supported(int channels, int num_levels, int num_point)
{
  int G = 0, NV = 0;
  if (channels <= 0 || num_levels <= 0 || num_point <= 0 || num_levels > kMaxLevels) return 0;
  return vec_shape<float>(channels, G, NV) && num_levels * num_point <= G;
}

// merged_ld > 0: offsets and logits are columns [0, 3*M*L*P) and [3*M*L*P, 4*M*L*P) of one row-major [N*Lq, merged_ld] tensor
static int merged_ok(const Dims &d, long long merged_ld) { return merged_ld == 0 || merged_ld >= 4LL * d.M * d.L * d.P; }

int msda3d_forward_fused(void *stream, const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                         const float *reference_points, int ref_batch, const float *sampling_offsets, const float *attn_logits, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point, float *output)
{
  return msda3d_forward_fused_ld(stream, value, spatial_shapes, level_start_index, reference_points, ref_batch, sampling_offsets, attn_logits, 0,
                                 batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, output);
}

int msda3d_forward_fused_ld(void *stream, const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *reference_points, int ref_batch, const float *sampling_offsets, const float *attn_logits,
                            long long merged_ld, int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                            int num_point, float *output)
{
  if (!value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || (!attn_logits && merged_ld == 0) || !output)
    return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int rc = check_dims(d)) return rc;
  if (ref_batch != 1 && ref_batch != batch) return MSDA3D_EINVAL;
  if (!merged_ok(d, merged_ld)) return MSDA3D_EINVAL;
  if (!aligned(reference_points, 4) || !aligned(sampling_offsets, 4) || !aligned(attn_logits, 4) || !aligned(spatial_shapes, 8) ||
      !aligned(level_start_index, 8))
    return MSDA3D_EALIGN;
  int g_ = 0, nv_ = 0;
  if (int rc = fused_shape(d, value, output, g_, nv_)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long units = (long long)d.N * d.Lq * d.M, rb = ref_batch == 1 ? 0 : (long long)d.Lq * d.L * 3;
  const int grid = vec_grid(units, g_);
  VEC_DISPATCH(g_, nv_, fwd_vec_kernel<float, G, NV, MinBlocks<float, NV>::fwd, 1><<<grid, kThreads, 0, st>>>(
                            value, spatial_shapes, level_start_index, sampling_offsets, merged_ld ? sampling_offsets : attn_logits, d.N, d.S, d.M,
                            d.L, d.Lq, d.P, output, use_brick(d), reference_points, rb, merged_ld, 3 * d.M * d.L * d.P));
  ++g_launches;
  return (int)cudaGetLastError();
}

int msda3d_backward_fused(void *stream, const float *grad_output, const float *value, const int64_t *spatial_shapes,
                          const int64_t *level_start_index, const float *reference_points, int ref_batch, const float *sampling_offsets,
                          const float *attn_logits, int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                          int num_point, float *grad_value, float *grad_sampling_offsets, float *grad_attn_logits)
{
  return msda3d_backward_fused_ld(stream, grad_output, value, spatial_shapes, level_start_index, reference_points, ref_batch, sampling_offsets,
                                  attn_logits, 0, batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, grad_value,
                                  grad_sampling_offsets, grad_attn_logits);
}

int msda3d_backward_fused_ld(void *stream, const float *grad_output, const float *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *reference_points, int ref_batch, const float *sampling_offsets,
                             const float *attn_logits, long long merged_ld, int batch, int spatial_size, int num_heads, int channels,
                             int num_levels, int num_query, int num_point, float *grad_value, float *grad_sampling_offsets,
                             float *grad_attn_logits)
{
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || !grad_value ||
      !grad_sampling_offsets || (merged_ld == 0 && (!attn_logits || !grad_attn_logits)))
    return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  if (int rc = check_dims(d)) return rc;
  if (ref_batch != 1 && ref_batch != batch) return MSDA3D_EINVAL;
  if (!merged_ok(d, merged_ld)) return MSDA3D_EINVAL;
  if (!aligned(grad_value, 16) || !aligned(grad_sampling_offsets, 4) || !aligned(grad_attn_logits, 4) || !aligned(reference_points, 4) ||
      !aligned(sampling_offsets, 4) || !aligned(attn_logits, 4))
    return MSDA3D_EALIGN;
  int g_ = 0, nv_ = 0;
  if (int rc = fused_shape(d, value, grad_output, g_, nv_)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(grad_value, 0, (size_t)d.N * d.S * d.M * d.C * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  const long long units = (long long)d.N * d.Lq * d.M, rb = ref_batch == 1 ? 0 : (long long)d.Lq * d.L * 3;
  const int grid = vec_grid(units, g_);
  if (g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_duo.load() == 1) {
#define MSDA3D_DUO(R, SK)                                                                                                                    \
  bwd_duo_kernel<1, R, SK><<<grid, kThreads, 0, st>>>(grad_output, value, spatial_shapes, level_start_index, sampling_offsets,               \
                                                      merged_ld ? sampling_offsets : attn_logits, d.N, d.S, d.M, d.L, d.Lq, d.P, grad_value, \
                                                      grad_sampling_offsets, merged_ld ? grad_sampling_offsets : grad_attn_logits,           \
                                                      reference_points, rb, merged_ld, 3 * d.M * d.L * d.P)
    const int r = g_tune_rot.load();
    if (g_tune_duo_cfg.load() != 0) {                              // experiment: smaller CTAs, more resident warps
      const int grid2 = vec_grid(units, 32);
      auto kern = g_tune_duo_cfg.load() == 1 ? bwd_duo_kernel<1, 0, 0, 128, 5> : bwd_duo_kernel<1, 0, 0, 128, 6>;
      if (g_diag_skip_red.load() == 1) kern = g_tune_duo_cfg.load() == 1 ? bwd_duo_kernel<1, 0, 1, 128, 5> : bwd_duo_kernel<1, 0, 1, 128, 6>;
      kern<<<grid2, 128, 0, st>>>(grad_output, value, spatial_shapes, level_start_index, sampling_offsets, merged_ld ? sampling_offsets : attn_logits,
                                  d.N, d.S, d.M, d.L, d.Lq, d.P, grad_value, grad_sampling_offsets,
                                  merged_ld ? grad_sampling_offsets : grad_attn_logits, reference_points, rb, merged_ld, 3 * d.M * d.L * d.P);
    } else
    if (g_diag_skip_red.load() == 1) MSDA3D_DUO(6, 1);
    else if (r >= 5) MSDA3D_DUO(6, 0); else if (r == 4) MSDA3D_DUO(4, 0); else if (r >= 1) MSDA3D_DUO(2, 0); else MSDA3D_DUO(0, 0);
#undef MSDA3D_DUO
    ++g_launches;
    return (int)cudaGetLastError();
  }
  if (g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_pair.load() == 1) {
    bwd_vec_kernel<float, 16, 1, 2, 0, 1, 1><<<grid, kThreads, 0, st>>>(
        grad_output, value, spatial_shapes, level_start_index, sampling_offsets, merged_ld ? sampling_offsets : attn_logits, d.N, d.S, d.M, d.L,
        d.Lq, d.P, grad_value, grad_sampling_offsets, merged_ld ? grad_sampling_offsets : grad_attn_logits, 1, reference_points, rb, merged_ld,
        3 * d.M * d.L * d.P);
    ++g_launches;
    return (int)cudaGetLastError();
  }
  if (g_ == 16 && nv_ == 1 && use_brick(d) && g_tune_rot.load() != 0) {
#define MSDA3D_ROT(R)                                                                                                                         \
  bwd_vec_kernel<float, 16, 1, 3, 0, 1, 0, R><<<grid, kThreads, 0, st>>>(                                                                    \
      grad_output, value, spatial_shapes, level_start_index, sampling_offsets, merged_ld ? sampling_offsets : attn_logits, d.N, d.S, d.M, d.L, \
      d.Lq, d.P, grad_value, grad_sampling_offsets, merged_ld ? grad_sampling_offsets : grad_attn_logits, 1, reference_points, rb, merged_ld,  \
      3 * d.M * d.L * d.P)
    switch (g_tune_rot.load()) { case 1: MSDA3D_ROT(1); break; case 2: MSDA3D_ROT(2); break; case 4: MSDA3D_ROT(4); break; case 5: MSDA3D_ROT(5); break; default: MSDA3D_ROT(6); break; }
#undef MSDA3D_ROT
    ++g_launches;
    return (int)cudaGetLastError();
  }
  VEC_DISPATCH(g_, nv_, bwd_vec_kernel<float, G, NV, MinBlocks<float, NV>::bwd, 0, 1><<<grid, kThreads, 0, st>>>(
                            grad_output, value, spatial_shapes, level_start_index, sampling_offsets, merged_ld ? sampling_offsets : attn_logits,
                            d.N, d.S, d.M, d.L, d.Lq, d.P, grad_value, grad_sampling_offsets,
                            merged_ld ? grad_sampling_offsets : grad_attn_logits, use_brick(d), reference_points, rb, merged_ld,
                            3 * d.M * d.L * d.P));
  ++g_launches;
  return (int)cudaGetLastError();
}

int msda3d_debug_indices(void *stream, int dtype, const int64_t *spatial_shapes, const void *sampling_loc, int batch, int num_heads,
                         int num_levels, int num_query, int num_point, int32_t *idx, void *frac)
{
  if (!spatial_shapes || !sampling_loc || !idx || !frac) return MSDA3D_EINVAL;
  if (batch <= 0 || num_heads <= 0 || num_levels <= 0 || num_query <= 0 || num_point <= 0) return MSDA3D_EINVAL;
  const long long T = (long long)batch * num_query * num_heads * num_levels * num_point;
  const int grid = (int)((T + 255) / 256 > 148 * 64 ? 148 * 64 : (T + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSDA3D_F64)
    indices_kernel<double><<<grid, 256, 0, st>>>(spatial_shapes, (const double *)sampling_loc, T, num_levels, num_point, idx, (double *)frac);
  else
    indices_kernel<float><<<grid, 256, 0, st>>>(spatial_shapes, (const float *)sampling_loc, T, num_levels, num_point, idx, (float *)frac);
  ++g_launches;
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Host-buffer path
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct HostCtx {
  int device = -1;
  cudaStream_t stream = nullptr, stream_cmp = nullptr, stream_out = nullptr;   // H2D / compute / D2H
  cudaEvent_t ev_in = nullptr, ev_done = nullptr;
  // grow-only device scratch: 0 value, 1 loc, 2 aw, 3 gout, 4 out, 5 gv, 6 gloc, 7 gaw, 8 shapes, 9 starts
  void *buf[10] = {nullptr};
  size_t cap[10] = {0};
};

struct DeviceGuard {
  int prev = 0;
  DeviceGuard() { cudaGetDevice(&prev); }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

std::mutex g_host_mu;
HostCtx g_host[16];

int ensure(HostCtx &c, int i, size_t bytes)
{
  if (c.cap[i] >= bytes) return 0;
  if (c.buf[i]) cudaFree(c.buf[i]);
  c.buf[i] = nullptr;
  c.cap[i] = 0;
  cudaError_t e = cudaMalloc(&c.buf[i], bytes);
  if (e != cudaSuccess) return (int)e;
  c.cap[i] = bytes;
  return 0;
}

#define CU(x)                                \
  do {                                       \
    cudaError_t e_ = (x);                    \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

int host_run(int device, int dtype, bool do_fwd, bool do_bwd, const void *gout, const void *value, const int64_t *shapes,
             const int64_t *starts, const void *loc, const void *aw, const Dims &d, void *out, void *gv, void *gl, void *ga)
{
  if (device < 0 || device >= 16) return MSDA3D_ENODEV;
  if (dtype < MSDA3D_F32 || dtype > MSDA3D_F16) return MSDA3D_EINVAL;
  if (int rc = check_dims(d)) return rc;
  std::lock_guard<std::mutex> lock(g_host_mu);
  HostCtx &c = g_host[device];
  DeviceGuard guard;
  CU(cudaSetDevice(device));
  if (!c.stream) {
    CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c.stream_cmp, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c.stream_out, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c.ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
    c.device = device;
  }
  const size_t es = elem_size(dtype), as = aux_size(dtype);
  const size_t n_val = (size_t)d.N * d.S * d.M * d.C, n_out = (size_t)d.N * d.Lq * d.M * d.C;
  const size_t n_aw = (size_t)d.N * d.Lq * d.M * d.L * d.P, n_loc = n_aw * 3;
  int rc = 0;
  if ((rc = ensure(c, 0, n_val * es)) || (rc = ensure(c, 1, n_loc * as)) || (rc = ensure(c, 2, n_aw * as)) ||
      (rc = ensure(c, 8, (size_t)d.L * 3 * 8)) || (rc = ensure(c, 9, (size_t)d.L * 8)))
    return rc;
  if (do_fwd && (rc = ensure(c, 4, n_out * es))) return rc;
  if (do_bwd && ((rc = ensure(c, 3, n_out * es)) || (rc = ensure(c, 5, n_val * as)) || (rc = ensure(c, 6, n_loc * as)) ||
                 (rc = ensure(c, 7, n_aw * as))))
    return rc;
  // Three-stage pipeline over the batch elements (they are independent): H2D of element b+1 and D2H of element b-1 run on
  // their own streams while element b computes, so the PCIe link is busy in both directions at once.
  cudaStream_t s_in = c.stream, s_cmp = c.stream_cmp, s_out = c.stream_out;
  CU(cudaMemcpyAsync(c.buf[8], shapes, (size_t)d.L * 3 * 8, cudaMemcpyHostToDevice, s_in));
  CU(cudaMemcpyAsync(c.buf[9], starts, (size_t)d.L * 8, cudaMemcpyHostToDevice, s_in));
  const size_t v1 = n_val / d.N, o1 = n_out / d.N, a1 = n_aw / d.N, l1 = n_loc / d.N;
  auto at = [](const void *p, size_t bytes) { return (const void *)((const char *)p + bytes); };
  auto atw = [](void *p, size_t bytes) { return (void *)((char *)p + bytes); };
  for (int b = 0; b < d.N; ++b) {
    CU(cudaMemcpyAsync(atw(c.buf[0], b * v1 * es), at(value, b * v1 * es), v1 * es, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(atw(c.buf[1], b * l1 * as), at(loc, b * l1 * as), l1 * as, cudaMemcpyHostToDevice, s_in));
    CU(cudaMemcpyAsync(atw(c.buf[2], b * a1 * as), at(aw, b * a1 * as), a1 * as, cudaMemcpyHostToDevice, s_in));
    if (do_bwd) CU(cudaMemcpyAsync(atw(c.buf[3], b * o1 * es), at(gout, b * o1 * es), o1 * es, cudaMemcpyHostToDevice, s_in));
    CU(cudaEventRecord(c.ev_in, s_in));
    CU(cudaStreamWaitEvent(s_cmp, c.ev_in, 0));
    if (do_fwd) {
      rc = msda3d_forward(s_cmp, dtype, at(c.buf[0], b * v1 * es), (const int64_t *)c.buf[8], (const int64_t *)c.buf[9],
                          at(c.buf[1], b * l1 * as), at(c.buf[2], b * a1 * as), 1, d.S, d.M, d.C, d.L, d.Lq, d.P, atw(c.buf[4], b * o1 * es));
      if (rc) return rc;
    }
    if (do_bwd) {
      rc = msda3d_backward(s_cmp, dtype, at(c.buf[3], b * o1 * es), at(c.buf[0], b * v1 * es), (const int64_t *)c.buf[8],
                           (const int64_t *)c.buf[9], at(c.buf[1], b * l1 * as), at(c.buf[2], b * a1 * as), 1, d.S, d.M, d.C, d.L, d.Lq,
                           d.P, atw(c.buf[5], b * v1 * as), atw(c.buf[6], b * l1 * as), atw(c.buf[7], b * a1 * as));
      if (rc) return rc;
    }
    CU(cudaEventRecord(c.ev_done, s_cmp));
    CU(cudaStreamWaitEvent(s_out, c.ev_done, 0));
    if (do_fwd) CU(cudaMemcpyAsync(atw(out, b * o1 * es), at(c.buf[4], b * o1 * es), o1 * es, cudaMemcpyDeviceToHost, s_out));
    if (do_bwd) {
      CU(cudaMemcpyAsync(atw(gv, b * v1 * as), at(c.buf[5], b * v1 * as), v1 * as, cudaMemcpyDeviceToHost, s_out));
      CU(cudaMemcpyAsync(atw(gl, b * l1 * as), at(c.buf[6], b * l1 * as), l1 * as, cudaMemcpyDeviceToHost, s_out));
      CU(cudaMemcpyAsync(atw(ga, b * a1 * as), at(c.buf[7], b * a1 * as), a1 * as, cudaMemcpyDeviceToHost, s_out));
    }
  }
  CU(cudaStreamSynchronize(s_out));
  CU(cudaStreamSynchronize(s_cmp));
  CU(cudaStreamSynchronize(s_in));
  return 0;
}

}  // namespace

int msda3d_forward_host(int device, int dtype, const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                        const void *sampling_loc, const void *attn_weight, int batch, int spatial_size, int num_heads, int channels,
                        int num_levels, int num_query, int num_point, void *output)
{
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output) return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  return host_run(device, dtype, true, false, nullptr, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, d, output,
                  nullptr, nullptr, nullptr);
}

int msda3d_backward_host(int device, int dtype, const void *grad_output, const void *value, const int64_t *spatial_shapes,
                         const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                         int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point, void *grad_value,
                         void *grad_sampling_loc, void *grad_attn_weight)
{
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_value ||
      !grad_sampling_loc || !grad_attn_weight)
    return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  return host_run(device, dtype, false, true, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, d,
                  nullptr, grad_value, grad_sampling_loc, grad_attn_weight);
}

int msda3d_forward_backward_host(int device, int dtype, const void *grad_output, const void *value, const int64_t *spatial_shapes,
                                 const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                                 int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                                 void *output, void *grad_value, void *grad_sampling_loc, void *grad_attn_weight)
{
  if (!grad_output || !value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output || !grad_value ||
      !grad_sampling_loc || !grad_attn_weight)
    return MSDA3D_EINVAL;
  const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
  return host_run(device, dtype, true, true, grad_output, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, d,
                  output, grad_value, grad_sampling_loc, grad_attn_weight);
}

void msda3d_host_release(void)
{
  std::lock_guard<std::mutex> lock(g_host_mu);
  for (HostCtx &c : g_host) {
    if (c.device < 0) continue;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c.device);
    for (int i = 0; i < 10; ++i) {
      if (c.buf[i]) cudaFree(c.buf[i]);
      c.buf[i] = nullptr;
      c.cap[i] = 0;
    }
    if (c.stream) { cudaStreamDestroy(c.stream); cudaStreamDestroy(c.stream_cmp); cudaStreamDestroy(c.stream_out); }
    if (c.ev_in) { cudaEventDestroy(c.ev_in); cudaEventDestroy(c.ev_done); }
    c.stream = c.stream_cmp = c.stream_out = nullptr;
    c.ev_in = c.ev_done = nullptr;
    c.device = -1;
    cudaSetDevice(prev);
  }
}

}  // extern "C"
