/*
 * msda3d.h -- C ABI of libmsda3d.so: 3D multi-scale deformable attention (forward + gradient) for NVIDIA B200
 * (sm_100a).  This is the drop-in boundary for the one native component of bwittmann/transoar: every entry point
 * below names the reference interface it replaces (paths relative to the reference checkout).
 *
 * Tensor conventions (identical to the reference, transoar/models/ops/src/cuda/ms_deform_attn_cuda.cu:40-60):
 *   value            [N, S, M, C]         S = sum_l D_l*H_l*W_l, contiguous
 *   spatial_shapes   int64 [L, 3]         (D, H, W) per level
 *   level_start_index int64 [L]           first voxel of each level inside S
 *   sampling_loc     [N, Lq, M, L, P, 3]  normalised (x, y, z) = (W, H, D) order, contiguous
 *   attn_weight      [N, Lq, M, L, P]
 *   output           [N, Lq, M*C]
 * Pixel coordinate = fma(loc, size, -0.5) (what nvcc makes of cuh:424-426), trilinear, per-corner
 * zero padding, sample skipped unless -1 < coord < size on all three axes (cuh:428).
 *
 * dtype selects the storage type of value / output / grad_output:
 *   MSDA3D_F32, MSDA3D_F64 : every floating tensor has that type (the reference's AT_DISPATCH_FLOATING_TYPES set).
 *   MSDA3D_BF16, MSDA3D_F16: value / output / grad_output are 16-bit; sampling_loc, attn_weight, grad_sampling_loc,
 *                            grad_attn_weight AND grad_value are fp32 (the reference cannot run under its own
 *                            trainer's autocast -- SURVEY.md D7 -- this is the documented extension).
 *
 * All functions return 0 on success or a negative MSDA3D_E* / positive cudaError_t code; msda3d_error_string()
 * explains either.  Unlike the reference (cuh:1119-1123,1501-1505: launch errors are printf'd and ignored) errors
 * are always returned.  No global mutable state; re-entrant; launches only on the stream passed in.
 */
#ifndef MSDA3D_H_
#define MSDA3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA3D_ABI_VERSION 1

enum { MSDA3D_F32 = 0, MSDA3D_F64 = 1, MSDA3D_BF16 = 2, MSDA3D_F16 = 3 };

enum {
  MSDA3D_OK = 0,
  MSDA3D_EINVAL = -1,   /* null pointer / non-positive dimension / unknown dtype        */
  MSDA3D_ERANGE = -2,   /* a dimension product exceeds what the index arithmetic covers  */
  MSDA3D_EALIGN = -3,   /* a pointer is not aligned to its element type                  */
  MSDA3D_ENODEV = -4    /* no CUDA device / not an sm_100 device                         */
};

int msda3d_abi_version(void);
const char *msda3d_error_string(int code);

/* Replaces ms_deformable_im2col_cuda<scalar_t>(stream, ...) -- cuda/ms_deform_im2col_cuda.cuh:1094-1125 -- and the
 * per-im2col_step loop around it in ms_deform_attn_cuda_forward, cuda/ms_deform_attn_cuda.cu:56-75 (one launch covers
 * the whole batch).  All pointers are DEVICE pointers, including the two int64 arrays.  `output` is fully overwritten
 * (no zero-fill needed, cf. at::zeros at ms_deform_attn_cuda.cu:54). */
int msda3d_forward(void *stream, int dtype, const void *value, const int64_t *spatial_shapes,
                   const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                   int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                   void *output);

/* Replaces ms_deformable_col2im_cuda<scalar_t>(stream, ...) -- cuh:1127-1507, all seven kernel variants -- and the
 * zero-initialisation + loop of ms_deform_attn_cuda_backward, ms_deform_attn_cuda.cu:122-149.  grad_value is
 * zero-filled here (cudaMemsetAsync on `stream`) and then accumulated into; grad_sampling_loc / grad_attn_weight are
 * fully overwritten. */
int msda3d_backward(void *stream, int dtype, const void *grad_output, const void *value, const int64_t *spatial_shapes,
                    const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                    int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                    void *grad_value, void *grad_sampling_loc, void *grad_attn_weight);

/* Host-buffer variants: what MSDA.ms_deform_attn_forward / _backward (vision.cpp:13-16) look like to a caller whose
 * tensors live in host memory.  Every pointer is a HOST pointer (pinned memory gives asynchronous copies); the call
 * stages inputs to the device, runs the kernels above on an internal stream, copies results back and synchronises.
 * `device` is the CUDA ordinal.  Scratch device memory is cached per device and released by msda3d_host_release(). */
int msda3d_forward_host(int device, int dtype, const void *value, const int64_t *spatial_shapes,
                        const int64_t *level_start_index, const void *sampling_loc, const void *attn_weight, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                        void *output);
int msda3d_backward_host(int device, int dtype, const void *grad_output, const void *value,
                         const int64_t *spatial_shapes, const int64_t *level_start_index, const void *sampling_loc,
                         const void *attn_weight, int batch, int spatial_size, int num_heads, int channels,
                         int num_levels, int num_query, int num_point, void *grad_value, void *grad_sampling_loc,
                         void *grad_attn_weight);
/* One pass of the training hot path for a caller with host tensors: forward, then backward with `grad_output`,
 * inputs staged once. */
int msda3d_forward_backward_host(int device, int dtype, const void *grad_output, const void *value,
                                 const int64_t *spatial_shapes, const int64_t *level_start_index,
                                 const void *sampling_loc, const void *attn_weight, int batch, int spatial_size,
                                 int num_heads, int channels, int num_levels, int num_query, int num_point,
                                 void *output, void *grad_value, void *grad_sampling_loc, void *grad_attn_weight);
void msda3d_host_release(void);

/* Fused prologue: the part of MSDeformAttn.forward between the two small Linear layers and the op
 * (transoar/models/ops/modules/ms_deform_attn.py:115-126)
 *     attention_weights = softmax(logits.view(N, Lq, M, L*P), -1)
 *     sampling_locations = reference_points[:, :, None, :, None, :] + sampling_offsets / (W, H, D)_l
 * is done inside the kernels: they read the RAW sampling offsets [N, Lq, M, L, P, 3] and attention logits [N, Lq, M, L, P] (the
 * outputs of the sampling_offsets / attention_weights Linear layers) plus reference_points [ref_batch (1 or N), Lq, L, 3], and the
 * backward returns the gradients with respect to the raw offsets and logits (softmax backward and the 1 / (W, H, D) scale applied
 * in-kernel).  sampling_loc and attn_weight -- 540 MB per layer at VISCERAL -- never exist in HBM, and six ATen kernels per layer
 * (softmax, div, add and their gradients) disappear.  The location is computed as fadd(ref, fdiv(offset, size)) -- the two fp32
 * roundings ATen performs -- so the sampled voxels are the same as on the unfused route; the softmax differs from ATen's in the
 * last ulp.  fp32 only; requires the vector kernels and L * P <= C / 4 lanes (msda3d_fused_supported says so), e.g. C = 64,
 * L * P <= 16: the reference's configuration.  Otherwise MSDA3D_EINVAL -- callers use the unfused entry points. */
int msda3d_fused_supported(int channels, int num_levels, int num_point);

int msda3d_forward_fused(void *stream, const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                         const float *reference_points, int ref_batch, const float *sampling_offsets, const float *attn_logits,
                         int batch, int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                         float *output);

int msda3d_backward_fused(void *stream, const float *grad_output, const float *value, const int64_t *spatial_shapes,
                          const int64_t *level_start_index, const float *reference_points, int ref_batch,
                          const float *sampling_offsets, const float *attn_logits, int batch, int spatial_size, int num_heads,
                          int channels, int num_levels, int num_query, int num_point, float *grad_value,
                          float *grad_sampling_offsets, float *grad_attn_logits);

/* Same, with offsets and logits taken from ONE row-major tensor [N*Lq, merged_ld] (merged_ld >= 4*M*L*P): columns [0, 3*M*L*P) are the
 * offsets of a query laid out [M][L][P][3], columns [3*M*L*P, 4*M*L*P) its logits [M][L][P] -- the output of a single Linear layer
 * whose weight is the concatenation of sampling_offsets.weight and attention_weights.weight (one GEMM instead of two).  `merged`
 * replaces both pointers; the backward writes both gradients into `grad_merged` of the same shape (columns past 4*M*L*P untouched).
 * merged_ld == 0 means the dense two-array form above (attn_logits / grad_attn_logits are then used). */
int msda3d_forward_fused_ld(void *stream, const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *reference_points, int ref_batch, const float *sampling_offsets_or_merged,
                            const float *attn_logits, long long merged_ld, int batch, int spatial_size, int num_heads, int channels,
                            int num_levels, int num_query, int num_point, float *output);

int msda3d_backward_fused_ld(void *stream, const float *grad_output, const float *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *reference_points, int ref_batch,
                             const float *sampling_offsets_or_merged, const float *attn_logits, long long merged_ld, int batch,
                             int spatial_size, int num_heads, int channels, int num_levels, int num_query, int num_point,
                             float *grad_value, float *grad_sampling_offsets_or_merged, float *grad_attn_logits);

/* Test hook: the sampling-index arithmetic of the production kernels, one record per sample (N*Lq*M*L*P):
 * idx int32[4] = {in_range, d_low, h_low, w_low}, frac[3] = {ld, lh, lw} (fp32 for F32/BF16/F16, fp64 for F64).
 * Device pointers.  Compared bit-for-bit with the oracle in tests/. */
int msda3d_debug_indices(void *stream, int dtype, const int64_t *spatial_shapes, const void *sampling_loc, int batch,
                         int num_heads, int num_levels, int num_query, int num_point, int32_t *idx, void *frac);

/* Diagnostics / tuning knobs for profiling sessions; never needed for correct results.  Unknown keys return
 * MSDA3D_EINVAL.  "nv" = 0|1|2: 16-byte vectors per lane in the vector kernels (0 = automatic);  "grid_mult" = CTAs
 * per SM cap of the launch (0 = automatic);  "order" = 0 automatic | 1 linear | 2 brick unit order (brick needs
 * num_query == spatial_size);  "diag_bwd_skip_red" = 1 drops the grad_value reductions
 * (WRONG results; isolates their cost);  "duo" = 1 (default) | 0: backward of fp32 / 64-channel / brick-order problems with two
 * w-neighbouring queries per lane group sharing corner rows and reductions (bwd_duo_kernel) or the one-unit kernel;
 * "rot" = 0 | 1 | 2 | 4 | 5 | 6: sample-order rotation per warp / per unit (+4: consecutive CTAs on different (batch, head)
 * slabs) in the one-unit backward;  "duo_cfg" = 0 | 1 | 2: CTA shape of bwd_duo_kernel (256 x 2, 128 x 5, 128 x 6 per SM);
 * "pair", "stage": earlier experiments.  The environment variable TRANSOAR_B200_TUNING="key=value,..."
 * applies them when the python package loads the library. */
int msda3d_set_tuning(const char *key, int value);

/* Number of kernel launches this library has issued in the calling process (bench.py's gpu_launches). */
unsigned long long msda3d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MSDA3D_H_ */
