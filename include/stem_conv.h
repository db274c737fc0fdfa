/*
 * stem_conv.h -- C ABI of the encoder's first convolution in libmsda3d.so (sm_100a): 1 input channel -> CO feature maps, kernel
 * 3x3x3, stride 1, zero padding 1, no bias.  Replaces nn.Conv3d(1, start_channels, 3, 1, 1, bias=False) of EncoderCnnBlock
 * stage 0 (transoar/models/backbones/encoder_blocks.py:28-33, built by attn_fpn.py:170-182) -- in the reference an ATen / cuDNN
 * implicit-GEMM call -- and its weight gradient (autograd's cudnn_convolution_backward_weight).  Cross-correlation, as torch.
 *
 *   x        fp32 [N, D, H, W]          (one channel: NCDHW and NDHWC coincide)
 *   weight   fp32 [CO, 1, 3, 3, 3]      contiguous; CO in {16, 24, 32}
 *   y, dy    fp32 [N, D, H, W, CO]      channels-last (torch.channels_last_3d), 16-byte aligned
 *   dweight  fp32 [CO, 1, 3, 3, 3]
 *   workspace fp32, stem_conv3d_workspace_floats(CO) floats
 * fp32 FMA arithmetic (no TF32 rounding).  The gradient with respect to x is not provided (x is the CT volume).
 * Device pointers, work enqueued on `stream`, no allocation, no synchronisation.  Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef STEM_CONV_H_
#define STEM_CONV_H_

#ifdef __cplusplus
extern "C" {
#endif

long long stem_conv3d_workspace_floats(int out_channels);

int stem_conv3d_forward(void *stream, const float *x, const float *weight, int batch, int depth, int height, int width,
                        int out_channels, float *y);

int stem_conv3d_wgrad(void *stream, const float *dy, const float *x, int batch, int depth, int height, int width, int out_channels,
                      float *dweight, float *workspace);

#ifdef __cplusplus
}
#endif
#endif /* STEM_CONV_H_ */
