/*
 * instnorm.h -- C ABI of the fused InstanceNorm3d(affine) + ReLU in libmsda3d.so (sm_100a), forward and gradient.
 *
 * Replaces the `nn.InstanceNorm3d(affine=True, eps=1e-5) -> nn.ReLU(inplace=True)` pairs of the reference's
 * EncoderCnnBlock (transoar/models/backbones/encoder_blocks.py:28-46), i.e. torch.nn.functional.instance_norm (ATen:
 * cuDNN batch-norm kernels over N*C "batches") followed by ReLU, and their autograd gradients.
 *
 *   x, y, dy, dx   [N, C, D, H, W] contiguous (V = D*H*W voxels per instance), fp32 (MSDA3D_F32) or bf16 (MSDA3D_BF16)
 *   gamma, beta    fp32 [C]   (the module's weight / bias);  dgamma, dbeta fp32 [C]
 *   mean, rstd     fp32 [N*C] (saved by the forward for the backward; biased variance, rstd = 1/sqrt(var + eps))
 *   workspace      fp32, instnorm_workspace_floats(...) floats, contents undefined on entry
 * Device pointers, work enqueued on `stream`, no allocation, no synchronisation.  Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef INSTNORM_H_
#define INSTNORM_H_

#ifdef __cplusplus
extern "C" {
#endif

long long instnorm_workspace_floats(int dtype, int batch, int channels, long long voxels);

int instnorm_relu_forward(void *stream, int dtype, const void *x, const float *gamma, const float *beta, int batch, int channels,
                          long long voxels, float eps, void *y, float *mean, float *rstd, float *workspace);

int instnorm_relu_backward(void *stream, int dtype, const void *dy, const void *x, const void *y, const float *gamma,
                           const float *mean, const float *rstd, int batch, int channels, long long voxels, void *dx,
                           float *dgamma, float *dbeta, float *workspace);

/* Channels-last variants: x, y, dy, dx are [N, D, H, W, C] in memory (torch.channels_last_3d), fp32, 16-byte aligned; channels must be
 * a multiple of 4 with (channels / 4) dividing 192 (24, 48, 96, 192, 384, 768, ...).  Same statistics, same outputs; they let the
 * encoder stay in the layout cuDNN's / this library's tensor-core convolutions use, with no NCDHW <-> NDHWC transposes.  The backward
 * takes beta instead of y: it recomputes the ReLU mask from x with the forward's arithmetic (two tensor reads per pass, not three). */
long long instnorm_ndhwc_workspace_floats(int batch, int channels, long long voxels);

int instnorm_relu_forward_ndhwc(void *stream, const float *x, const float *gamma, const float *beta, int batch, int channels,
                                long long voxels, float eps, float *y, float *mean, float *rstd, float *workspace);

int instnorm_relu_backward_ndhwc(void *stream, const float *dy, const float *x, const float *gamma, const float *beta, const float *mean,
                                 const float *rstd, int batch, int channels, long long voxels, float *dx, float *dgamma, float *dbeta,
                                 float *workspace);

/* The NDHWC kernels with bf16 storage for x / y / dy / dx (8-byte aligned); gamma / beta / statistics / workspace stay fp32 and all
 * arithmetic is fp32.  Used when the encoder runs under the bf16 autocast region of the trainer (transoar/trainer.py:67-69). */
int instnorm_relu_forward_ndhwc_bf16(void *stream, const void *x, const float *gamma, const float *beta, int batch, int channels,
                                     long long voxels, float eps, void *y, float *mean, float *rstd, float *workspace);

int instnorm_relu_backward_ndhwc_bf16(void *stream, const void *dy, const void *x, const float *gamma, const float *beta, const float *mean,
                                      const float *rstd, int batch, int channels, long long voxels, void *dx, float *dgamma, float *dbeta,
                                      float *workspace);

#ifdef __cplusplus
}
#endif
#endif /* INSTNORM_H_ */
