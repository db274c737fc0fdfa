/*
 * fused_ln.h -- C ABI of y = LayerNorm(a + dropout(b)) in libmsda3d.so (sm_100a), forward and gradient.
 *
 * Replaces the dropout -> residual add -> nn.LayerNorm triples of the reference's post-norm blocks:
 *   DefAttnLayer.forward           transoar/models/backbones/decoder_blocks.py:163-177  (norm1(src + dropout1(attn)), norm2(src + dropout3(ffn)))
 *   FocusedDecoderLayer.forward    transoar/models/necks/focused_decoder.py:166-189     (norm2 / norm1 / norm3)
 * i.e. ATen's fused_dropout, add, layer_norm kernels and their autograd gradients (layer_norm_backward incl. the gamma / beta
 * reduction, masked_scale, gradient accumulation).
 *
 *   a, b, z, y, dy, da, db   fp32 [rows, C], C % 4 == 0, C <= 1024, 16-byte aligned, contiguous
 *   gamma, beta, dgamma, dbeta  fp32 [C];   mean, rstd  fp32 [rows]  (biased variance, rstd = 1 / sqrt(var + eps))
 *   z = a + dropout(b)   (saved by the forward for the backward);  b may be NULL (plain LayerNorm(a), z may then be NULL too:
 *                         the backward takes a as z)
 *   dropout: element e of b is kept with probability 1 - p and scaled by 1 / (1 - p); the keep decision is a counter-based hash
 *            of (seed, e) -- no mask tensor exists, the backward re-evaluates it from the same seed.  p = 0 disables it.
 *   backward: da = d(loss)/da; db = d(loss)/db (NULL when b was NULL or p == 0: then db == da, write it once).
 *   workspace: fused_ln_workspace_floats(C) floats (per-CTA partial sums of the gamma / beta gradients).
 * Device pointers, work enqueued on `stream`, no allocation, no synchronisation.  Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef FUSED_LN_H_
#define FUSED_LN_H_

#ifdef __cplusplus
extern "C" {
#endif

long long fused_ln_workspace_floats(int channels);

int fused_ln_forward(void *stream, const float *a, const float *b, const float *gamma, const float *beta, long long rows, int channels,
                     float eps, float p_drop, unsigned long long seed, float *z, float *y, float *mean, float *rstd);

int fused_ln_backward(void *stream, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd,
                      long long rows, int channels, float p_drop, unsigned long long seed, float *da, float *db, float *dgamma,
                      float *dbeta, float *workspace);

/* Dropout masks under CUDA-graph replay.  The seeds above are passed by value, so a captured launch would repeat its mask on every
 * replay.  While a device counter is installed here, every kernel of this library that draws a mask (fused_ln_forward/backward,
 * tc_gemm_tf32_ex with p_drop > 0) folds *device_counter into its seed at run time: advance the counter once per step (inside the
 * graph) and each replay draws fresh masks, forward and backward of one step still agreeing.  NULL (the default) removes it.
 * Process-global; the pointer must stay valid while installed. */
void hash_rng_set_epoch(const unsigned long long *device_counter);

/* The bf16 route: the branch b (and its gradient db) is the bf16 output of a bf16 GEMM, the residual stream a / z / y / dy / da stays fp32 --
 * the dtypes torch.autocast(bfloat16) produces around `norm(x + dropout(branch))` (its layer_norm runs and returns fp32).  b / db are
 * 8-byte aligned; db is always written (it cannot alias the fp32 da), with p_drop == 0 it is the bf16 rounding of da. */
int fused_ln_forward_bf16b(void *stream, const float *a, const void *b, const float *gamma, const float *beta, long long rows, int channels,
                           float eps, float p_drop, unsigned long long seed, float *z, float *y, float *mean, float *rstd);

int fused_ln_backward_bf16b(void *stream, const float *dy, const float *z, const float *gamma, const float *mean, const float *rstd, long long rows,
                            int channels, float p_drop, unsigned long long seed, float *da, void *db, float *dgamma, float *dbeta,
                            float *workspace);

#ifdef __cplusplus
}
#endif
#endif /* FUSED_LN_H_ */
