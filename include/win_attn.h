/*
 * win_attn.h -- C ABI of the fused Swin3D window attention in libmsda3d.so (sm_100a).
 *
 * Replaces the attention core of WindowAttention3D.forward (transoar/models/backbones/encoder_blocks.py:259-285):
 *     attn = softmax((q * scale) @ k^T + relative_position_bias[head] (+ mask[window]))  ;  x = attn @ v
 * for the windows of one Swin block.  The reference materialises the [B*nW, heads, n, n] score tensor (n = 125) and touches it five
 * times; here a CTA owns one (window, head) and keeps K / V in shared memory.
 *
 *   qkv    [windows][tokens][3][heads][head_dim]   the qkv Linear's output as it lies in memory (encoder_blocks.py:262: the reshape, before
 *                                                  the permute -- no copy is made)
 *   bias   [heads][tokens][tokens]                 relative_position_bias_table[relative_position_index] (encoder_blocks.py:267-270)
 *   bias_t [heads][tokens][tokens]                 the same, last two axes swapped (coalesced reads for the row-owner phases)
 *   mask   [mask_windows][tokens][tokens] or NULL  the shifted-window mask (0 / -100, symmetric; encoder_blocks.py:387-400); window w of
 *                                                  the batch uses mask[w % mask_windows], as attn.view(B_ // nW, nW, ...) does (:275-277)
 *   out    [windows][tokens][heads*head_dim]       what (attn @ v).transpose(1, 2).reshape(B_, N, C) yields (:283)
 *   lse    [windows][heads][tokens]                row log-sum-exp, kept for the backward
 *   dqkv like qkv; dbias like bias (summed over windows; zero-filled by the callee).
 *
 * tokens <= 128, head_dim == 16 (every Swin stage of the reference: 48/3 ... 384/24); win_attn_supported() answers for other shapes.
 * fp32, device pointers 16-byte aligned, work enqueued on `stream`, no allocation, no synchronisation; returns 0 / MSDA3D_E* / cudaError_t.
 * Attention dropout is not fused (attn_drop_rate is 0 in every reference config, config/attn_fpn_foc_dec_*.yaml:64).
 */
#ifndef WIN_ATTN_H_
#define WIN_ATTN_H_

#ifdef __cplusplus
extern "C" {
#endif

int win_attn_supported(int tokens, int head_dim);

int win_attn_forward(void *stream, const float *qkv, const float *bias_t, const float *mask, int windows, int tokens, int heads, int head_dim,
                     int mask_windows, float scale, float *out, float *lse);

int win_attn_backward(void *stream, const float *qkv, const float *bias, const float *bias_t, const float *mask, const float *out,
                      const float *dout, const float *lse, int windows, int tokens, int heads, int head_dim, int mask_windows, float scale,
                      float *dqkv, float *dbias);

#ifdef __cplusplus
}
#endif
#endif /* WIN_ATTN_H_ */
