// roi_attn_capi.cu -- C ABI of the fused RoI attention (include/roi_attn.h).
#include "roi_attn_tc_kernels.cuh"

#include <atomic>

#include "../../include/msda3d.h"
#include "../../include/roi_attn.h"

extern std::atomic<unsigned long long> g_msda3d_launches;
std::atomic<int> g_roi_splits{0};        // token splits per box forced by msda3d_set_tuning("roi_splits", n); 0 = pick_splits' own choice

namespace {

template <int HD, bool TC>
int launch_fwd(cudaStream_t st, int G, int B, const float *q, const float *k, const float *v, const int *groups, int Nq, int Nkv, int H,
               int Y, int Z, float *out, float *lse, int S, float *part)
{
  auto kern = TC ? roiattn::fwd_tc_kernel<HD> : roiattn::fwd_kernel<HD>;
  constexpr size_t smem = TC ? roiattn::fwd_tc_smem_bytes<HD>() : roiattn::fwd_smem_bytes<HD>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(G * S, H, B), roiattn::kThreads, smem, st>>>(q, k, v, groups, Nq, Nkv, H, Y, Z, out, lse, S, part);
  if (S > 1) {
    const int rows = B * H * Nq;
    roiattn::combine_kernel<HD><<<(rows + 3) / 4, 128, 0, st>>>(part, rows, Nq, H, S, out, lse);
  }
  return (int)cudaGetLastError();
}

template <int HD, bool TC>
int launch_bwd(cudaStream_t st, int G, int B, const float *q, const float *k, const float *v, const int *groups, const float *out,
               const float *dout, const float *lse, int Nq, int Nkv, int H, int Y, int Z, float *dq, float *dk, float *dv, int S)
{
  auto kern = TC ? roiattn::bwd_tc_kernel<HD> : roiattn::bwd_kernel<HD>;
  constexpr size_t smem = TC ? roiattn::bwd_tc_smem_bytes<HD>() : roiattn::bwd_smem_bytes<HD>();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(G * S, H, B), roiattn::kThreads, smem, st>>>(q, k, v, groups, out, dout, lse, Nq, Nkv, H, Y, Z, dq, dk, dv, S);
  return (int)cudaGetLastError();
}

// Token splits per box: enough CTAs for ~20 per SM, at most 16 (each split re-reads Q and adds a partial state).  The boxes of an atlas
// differ 2.7x in size (3024 ... 8265 tokens in the VISCERAL config) and a CTA walks its share serially, so with one wave of CTAs
// (the earlier "~4 per SM": 2 splits) the largest box set the time; several small waves balance themselves.  Sweep on B200
// (tools/exp_roi_splits.py, profiles/r04i_roi_splits.txt): 2 splits 0.416 / 1.225 ms (forward / backward), 10 splits 0.286 / 0.978 ms.
int pick_splits(int G, int H, int B)
{
  const int forced = g_roi_splits.load();                         // msda3d_set_tuning("roi_splits", n): experiment switch, 0 = automatic
  if (forced > 0) return forced > 16 ? 16 : forced;
  const long long ctas = (long long)G * H * B;
  long long s = (148LL * 20 + ctas - 1) / ctas;
  return (int)(s < 1 ? 1 : s > 16 ? 16 : s);
}

bool bad_dims(int G, int B, int Nq, int Nkv, int H, int Y, int Z)
{
  return G <= 0 || B <= 0 || Nq <= 0 || Nkv <= 0 || H <= 0 || Y <= 0 || Z <= 0 || Nkv % (Y * Z) != 0 || G > 65535 * 32 || H > 65535 ||
         B > 65535;
}

}  // namespace

#define HD_DISPATCH(hd, ...)                                  \
  switch (hd) {                                               \
    case 16: { constexpr int HD = 16; __VA_ARGS__; } break;   \
    case 32: { constexpr int HD = 32; __VA_ARGS__; } break;   \
    case 48: { constexpr int HD = 48; __VA_ARGS__; } break;   \
    case 64: { constexpr int HD = 64; __VA_ARGS__; } break;   \
    case 96: { constexpr int HD = 96; __VA_ARGS__; } break;   \
    case 128: { constexpr int HD = 128; __VA_ARGS__; } break; \
    default: return MSDA3D_EINVAL;                            \
  }

template <bool TC>
static int roi_forward_impl(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                     int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, float *out, float *lse,
                     float *workspace, long long workspace_floats)
{
  if (!q || !k || !v || !groups || !out || !lse) return MSDA3D_EINVAL;
  if (bad_dims(num_groups, batch, num_query, num_kv, num_heads, grid_y, grid_z)) return MSDA3D_EINVAL;
  int S = pick_splits(num_groups, num_heads, batch);
  const long long need = (long long)batch * num_heads * num_query * (head_dim + 2);
  while (S > 1 && (!workspace || need * S > workspace_floats)) --S;       // no / small workspace: fewer splits, still correct
  int rc = 0;
  HD_DISPATCH(head_dim, rc = launch_fwd<HD, TC>((cudaStream_t)stream, num_groups, batch, q, k, v, groups, num_query, num_kv, num_heads,
                                            grid_y, grid_z, out, lse, S, workspace));
  g_msda3d_launches += (S > 1) ? 2 : 1;
  return rc;
}


template <bool TC>
static int roi_backward_impl(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                      int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, const float *out,
                      const float *dout, const float *lse, float *dq, float *dk, float *dv)
{
  if (!q || !k || !v || !groups || !out || !dout || !lse || !dq || !dk || !dv) return MSDA3D_EINVAL;
  if (bad_dims(num_groups, batch, num_query, num_kv, num_heads, grid_y, grid_z)) return MSDA3D_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t kv_bytes = (size_t)batch * num_kv * num_heads * head_dim * sizeof(float);
  const int S = pick_splits(num_groups, num_heads, batch);
  cudaError_t e = cudaMemsetAsync(dk, 0, kv_bytes, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(dv, 0, kv_bytes, st);
  if (e == cudaSuccess && S > 1) e = cudaMemsetAsync(dq, 0, (size_t)batch * num_query * num_heads * head_dim * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  int rc = 0;
  HD_DISPATCH(head_dim, rc = launch_bwd<HD, TC>(st, num_groups, batch, q, k, v, groups, out, dout, lse, num_query, num_kv, num_heads,
                                            grid_y, grid_z, dq, dk, dv, S));
  ++g_msda3d_launches;
  return rc;
}


extern "C" {

int roi_attn_forward(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                     int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, float *out, float *lse,
                     float *workspace, long long workspace_floats)
{
  return roi_forward_impl<false>(stream, q, k, v, groups, num_groups, batch, num_query, num_kv, num_heads, head_dim, grid_y, grid_z, out, lse,
                                 workspace, workspace_floats);
}

int roi_attn_forward_tf32(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                          int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, float *out, float *lse,
                          float *workspace, long long workspace_floats)
{
  return roi_forward_impl<true>(stream, q, k, v, groups, num_groups, batch, num_query, num_kv, num_heads, head_dim, grid_y, grid_z, out, lse,
                                workspace, workspace_floats);
}

long long roi_attn_workspace_floats(int num_groups, int batch, int num_query, int num_heads, int head_dim)
{
  if (num_groups <= 0 || batch <= 0 || num_query <= 0 || num_heads <= 0 || head_dim <= 0) return 0;
  const int S = pick_splits(num_groups, num_heads, batch);
  return S > 1 ? (long long)batch * num_heads * num_query * (head_dim + 2) * S : 0;
}

int roi_attn_backward(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                      int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, const float *out,
                      const float *dout, const float *lse, float *dq, float *dk, float *dv)
{
  return roi_backward_impl<false>(stream, q, k, v, groups, num_groups, batch, num_query, num_kv, num_heads, head_dim, grid_y, grid_z, out, dout,
                                  lse, dq, dk, dv);
}

int roi_attn_backward_tf32(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                           int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, const float *out,
                           const float *dout, const float *lse, float *dq, float *dk, float *dv)
{
  return roi_backward_impl<true>(stream, q, k, v, groups, num_groups, batch, num_query, num_kv, num_heads, head_dim, grid_y, grid_z, out, dout,
                                 lse, dq, dk, dv);
}

}  // extern "C"
