/*
 * conv3d_tc.h -- C ABI of the tcgen05 implicit-GEMM 3x3x3 convolution in libmsda3d.so (sm_100a): stride 1, zero padding 1, no bias,
 * channels-last fp32 tensors, TF32 multiply / fp32 accumulate.  Replaces nn.Conv3d(C, C, 3, 1, 1, bias=False) of the narrow
 * full-resolution encoder stage (EncoderCnnBlock._block[3] of stage 0, transoar/models/backbones/encoder_blocks.py:34-40 via
 * attn_fpn.py:170-182: 24 -> 24 channels at 160x160x256) -- cuDNN implicit-GEMM kernels in the reference -- and, called with the
 * flipped / transposed weights, autograd's gradient with respect to the input (cudnn_convolution_backward_input).
 *
 *   x       fp32 [N, D, H, W, CI]   channels-last (torch.channels_last_3d), 16-byte aligned; CI in {8, 16, 24}
 *   w_taps  fp32 [27, CO, CI]       tap-major weights: w_taps[(kd*3 + kh)*3 + kw][co][ci] = weight[co][ci][kd][kh][kw]
 *                                   (input gradient: w_taps[t][ci][co] = weight[co][ci][2-kd][2-kh][2-kw], channel roles swapped)
 *   y       fp32 [N, D, H, W, CO]   channels-last; CO % 4 == 0, CO <= 32
 * Cross-correlation, as torch.  Device pointers, work enqueued on `stream`, no allocation, no synchronisation.
 * Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef CONV3D_TC_H_
#define CONV3D_TC_H_

#ifdef __cplusplus
extern "C" {
#endif

int conv3d_tc_supported(int in_channels, int out_channels);

int conv3d_tc_k3_forward(void *stream, const float *x, const float *w_taps, int batch, int depth, int height, int width, int in_channels,
                         int out_channels, float *y);

/* Test hook: one 128 x 32 x 8 TF32 MMA whose operands are MN-major in the no-swizzle canonical layout (what a tensor-core weight
 * gradient over channels-last volumes needs): D[m][n] = sum_k At[k][m] * Bt[k][n], At [8][128], Bt [8][32], D [128][32], device pointers. */
int conv3d_tc_debug_mn_probe(void *stream, const float *At, const float *Bt, float *D);

#ifdef __cplusplus
}
#endif
#endif /* CONV3D_TC_H_ */
