/*
 * conv3d_gen.h -- C ABI of the general tcgen05 implicit-GEMM 3x3x3 convolution in libmsda3d.so (sm_100a): zero padding 1, stride 1 or 2,
 * any channel counts that are multiples of 4, channels-last fp32 tensors, TF32 multiply / fp32 accumulate, operands staged by TMA,
 * accumulators in TMEM.  Replaces the cuDNN calls behind
 *   nn.Conv3d(cin, cout, 3, stride, 1, bias=False)   EncoderCnnBlock._block[0] / [3] of stages 1-5
 *                                                    (transoar/models/backbones/encoder_blocks.py:28-46 via attn_fpn.py:170-182)
 *   nn.Conv3d(cin, fpn_channels, 3, padding=1)       Decoder._out (transoar/models/backbones/attn_fpn.py:65-74), with bias
 * and autograd's cudnn_convolution_backward_input / _weight of both.  (The 24 -> 24 full-resolution layer keeps its specialised
 * kernel, include/conv3d_tc.h; the 1-channel stem its stencil, include/stem_conv.h.)
 *
 *   x    fp32 [N, D, H, W, CI]      channels-last (torch.channels_last_3d memory of an [N, CI, D, H, W] tensor), 16-byte aligned
 *   w    fp32 [27, CO, CI]          tap-major weights: w[(kd*3 + kh)*3 + kw][co][ci] = weight[co][ci][kd][kh][kw] -- one (tap, 32-channel)
 *                                   block is then CO rows CI * 4 bytes apart, which the copy engine fetches as whole L2 lines; the SAME
 *                                   tensor serves the forward (rows = co, K-major) and the input gradient (rows = the reduction index co,
 *                                   MN-major 32 x 32 slabs), so no transposed copy exists
 *   y    fp32 [N, OD, OH, OW, CO]   OD = ceil(D / stride) etc. (= floor((D + 2 - 3) / stride) + 1)
 *   bias fp32 [CO] or NULL
 * Cross-correlation, as torch.  depth / height / width are always the INPUT volume's.  Device pointers, work enqueued on `stream`, no
 * allocation, no synchronisation (capturable).  Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef CONV3D_GEN_H_
#define CONV3D_GEN_H_

#ifdef __cplusplus
extern "C" {
#endif

int conv3d_gen_supported(int in_channels, int out_channels, int stride);

int conv3d_gen_forward(void *stream, const float *x, const float *w, const float *bias, int batch, int depth, int height, int width,
                       int in_channels, int out_channels, int stride, float *y);

/* dx [N, D, H, W, CI] = gradient of the convolution with respect to its input, from dy [N, OD, OH, OW, CO].  Every element of dx is written. */
int conv3d_gen_dgrad(void *stream, const float *dy, const float *w, int batch, int depth, int height, int width, int in_channels,
                     int out_channels, int stride, float *dx);

/* The same stride-2 input gradient for narrow layers (CI <= 64) with the eight parity classes of dx folded into the column extent of one
 * implicit GEMM.  w_fold fp32 [8, 8 * CIP, CO], CIP = CI rounded up to 32: w_fold[(dd*2 + dh)*2 + dw][cls * CIP + ci][co] = weight[co][ci][kd][kh][kw]
 * with, per axis, k = 1 for (class bit 0, delta 0), k = 2 for (1, 0), k = 0 for (1, 1), and zeros for (0, 1) and for ci >= CI
 * (transoar_b200/conv3d_gen.py::fold_stride2_weights builds it).  Every element of dx is written. */
int conv3d_gen_dgrad_s2_folded(void *stream, const float *dy, const float *w_fold, int batch, int depth, int height, int width, int in_channels,
                               int out_channels, float *dx);

/* dw [CO, 27, CI] (channels-last weight memory) = sum over output voxels of dy[v][co] * x[stride * v + tap - 1][ci].  dw is zero-filled by the
 * call, then accumulated with fp32 reductions (split over voxel ranges: the summation order is not deterministic). */
int conv3d_gen_wgrad(void *stream, const float *x, const float *dy, int batch, int depth, int height, int width, int in_channels,
                     int out_channels, int stride, float *dw);

/* Forward and input gradient have two kernels: "halo" (8 x 16 voxel tiles whose (kh, kw) taps are read out of one halo tile in shared memory;
 * the fast one for volumes with >= ~16 rows) and "tap" (one TMA box per tap; any tile box, used for the small coarse levels).  0 = choose per
 * problem (default), 1 = tap, 2 = halo wherever it fits.  Tests and timing experiments only. */
void conv3d_gen_set_path(int path);

/* Experiment hook (tools/probe_kshift.py): one 128 x 32 x 8 TF32 MMA whose K-major, 128-byte-swizzled A operand starts at row `row0` of a
 * [176][32] matrix held in shared memory with 8-row groups `group_stride_rows` rows apart; mode bit 0 sets the descriptor's base-offset field.
 * D[m][n] = sum_{k < 8} X[row0 + (m / 8) * group_stride_rows + m % 8][k] * Y[n][k], Y [32][32], D [128][32], device pointers. */
int conv3d_gen_debug_k_probe(void *stream, const float *X, const float *Y, float *D, int row0, int group_stride_rows, int mode);

/* Experiment hook (tools/probe_mma_rate.py): clocks for `iters` back-to-back 128 x n x 8 TF32 MMAs with zero operands in shared-memory layout
 * 0 (K-major, 128-byte swizzle), 1 (K-major, 32-byte swizzle) or 2 (MN-major, 32-byte atoms); out_clocks: one int64 on the device. */
int conv3d_gen_debug_mma_rate(void *stream, int layout, int n, int iters, long long *out_clocks);

#ifdef __cplusplus
}
#endif
#endif /* CONV3D_GEN_H_ */
