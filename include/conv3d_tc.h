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

/* Weight gradient of the same convolution (autograd's cudnn_convolution_backward_weight): dweight [CO, CI, 3, 3, 3] (contiguous) =
 * sum over voxels of dy[v][co] * x[v + tap][ci]; x [N, D, H, W, CI], dy [N, D, H, W, CO] channels-last fp32, CI, CO <= 32 and % 4 == 0.
 * Both operands are MN-major for the tensor core (128-byte rows of 32 channels, padded by the copy engine); the three kw taps are the
 * overlapping slabs of one MMA operand.  workspace: conv3d_tc_wgrad_workspace_floats() floats (per-CTA partial sums). */
long long conv3d_tc_wgrad_workspace_floats(void);
int conv3d_tc_k3_wgrad(void *stream, const float *x, const float *dy, int batch, int depth, int height, int width, int in_channels,
                       int out_channels, float *dweight, float *workspace);

/* Timing experiments only (tools/exp_conv.py): bit 0 = the forward kernel's epilogue neither reads the accumulators nor stores, bit 1 = no
 * MMAs are issued, bit 2 = no stores.  Results are wrong while a mode is set; 0 (the default) is the product. */
void conv3d_tc_debug_mode(int mode);

/* Test hook for the weight-gradient formulation: one 128 x 32 x 8 TF32 MMA with MN-major operands in the 128-byte swizzle /
 * 32-byte atom layout, the A operand being four OVERLAPPING 32-channel slabs (leading offset = one 128-byte row) starting at row `row0`:
 * D[j * 32 + c][n] = sum_{k < 8} X[row0 + j + k][c] * Y[k][n], X [24][32], Y [8][32], D [128][32], device pointers. */
int conv3d_tc_debug_mn_probe(void *stream, const float *X, const float *Y, float *D, int row0);

#ifdef __cplusplus
}
#endif
#endif /* CONV3D_TC_H_ */
