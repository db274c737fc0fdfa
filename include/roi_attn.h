/*
 * roi_attn.h -- C ABI of the fused RoI-restricted cross-attention in libmsda3d.so (sm_100a).
 *
 * Replaces the attention core of the reference's FocusedAttn.forward
 * (transoar/models/necks/focused_decoder.py:238-254: q @ k^T, additive -inf RoI mask, softmax, @ v) and its autograd
 * gradient.  The reference builds its masks from axis-aligned boxes shared by `num_queries_per_organ` consecutive
 * queries (generate_attn_masks, focused_decoder.py:138-159); the boxes themselves are the interface here.
 *
 *   q       fp32 [B, Nq, H, HD]   already projected and scaled (focused_decoder.py:235-236)
 *   k, v    fp32 [B, Nkv, H, HD]  Nkv = X*Y*Z tokens of a [B, C, X, Y, Z] map flattened row-major
 *   groups  int32 [G, 8]          {q0, nq (1..32), x1, y1, z1, x2, y2, z2}: queries q0..q0+nq-1 attend to the voxels
 *                                 [x1,x2) x [y1,y2) x [z1,z2); groups must cover every query exactly once
 *   out     fp32 [B, Nq, H*HD]    lse fp32 [B, H, Nq]
 * All pointers are device pointers.  HD in {16, 32, 48, 64, 96, 128}.  Returns 0, a negative MSDA3D_E* code or a
 * positive cudaError_t (msda3d_error_string explains all of them).
 */
#ifndef ROI_ATTN_H_
#define ROI_ATTN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The tokens of a box are split over several CTAs when (groups x heads x batch) would leave SMs idle; the partial softmax
 * states live in `workspace` (device, fp32).  roi_attn_workspace_floats() says how many floats the best split needs
 * (0 = no split); a NULL or smaller workspace only reduces the split, never the correctness. */
long long roi_attn_workspace_floats(int num_groups, int batch, int num_query, int num_heads, int head_dim);

int roi_attn_forward(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups,
                     int batch, int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, float *out,
                     float *lse, float *workspace, long long workspace_floats);

/* dq is fully overwritten (zero-filled + atomics when boxes are split); dk and dv are zero-filled here (cudaMemsetAsync on `stream`) and then accumulated into with
 * atomics, because the boxes of different groups overlap. */
int roi_attn_backward(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups,
                      int batch, int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z,
                      const float *out, const float *dout, const float *lse, float *dq, float *dk, float *dv);

/* The same two entry points with the five contractions (Q K^T, P V; dO V^T, P^T dO, dS^T Q, dS K) on the tensor cores: mma.sync m16n8k8
 * with TF32 operands and fp32 accumulation (transoar_b200/csrc/roi_attn_tc_kernels.cuh).  TF32 is what the reference's torch.matmul calls
 * in FocusedAttn.forward (focused_decoder.py:238,254) run at when torch.backends.cuda.matmul.allow_tf32 is on (torch 1.10's default);
 * the python module selects these when that flag is set and the fp32 CUDA-core kernels above otherwise.  Arguments, workspace, partial
 * state and error codes are identical. */
int roi_attn_forward_tf32(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups, int batch,
                          int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z, float *out, float *lse,
                          float *workspace, long long workspace_floats);

int roi_attn_backward_tf32(void *stream, const float *q, const float *k, const float *v, const int32_t *groups, int num_groups,
                           int batch, int num_query, int num_kv, int num_heads, int head_dim, int grid_y, int grid_z,
                           const float *out, const float *dout, const float *lse, float *dq, float *dk, float *dv);

#ifdef __cplusplus
}
#endif
#endif /* ROI_ATTN_H_ */
