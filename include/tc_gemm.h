/*
 * tc_gemm.h -- C ABI of the TF32 tensor-core GEMM in libmsda3d.so (sm_100a: tcgen05.mma + TMEM accumulators + TMA operand
 * staging).  It carries the dense contractions of the hot path and their gradients:
 *
 *   torch.nn.Linear forward  (F.linear -> cuBLAS sgemm/tf32 in the reference) used by
 *     MSDeformAttn.value_proj / sampling_offsets / attention_weights / output_proj  transoar/models/ops/modules/ms_deform_attn.py:58-61,109-140
 *     DefAttnLayer.linear1 / linear2                                                transoar/models/backbones/decoder_blocks.py:156-160,172-175
 *     FocusedAttn.k_proj / v_proj / proj                                            transoar/models/necks/focused_decoder.py:213-218,233-257
 *     FocusedDecoderLayer.linear1 / linear2                                         transoar/models/necks/focused_decoder.py:131-135,186-187
 *   and autograd's two gradient GEMMs of each of them (grad_input = grad_output W, grad_weight = grad_output^T input).
 *
 *   D[m, n] (+)= sum_{r < R} A(m, r) * B(n, r)   (+ bias[n])   (ReLU)          m < M, n < N, fp32 everywhere, TF32 multiply
 *
 *   a_mn_major == 0 : A(m, r) = A[m * lda + r]   ("K-major": the reduction index is contiguous, e.g. X[M, K], W[N, K])
 *   a_mn_major == 1 : A(m, r) = A[r * lda + m]   ("MN-major": the m index is contiguous, e.g. grad_output^T)
 *   b_mn_major likewise for B(n, r).  D is row-major with leading dimension ldd.
 *   accumulate != 0 : D += result (red.global.add); required when split_k > 1.  split_k == 0 picks a split automatically
 *                     (only when accumulate != 0; D must then hold the value to add to, e.g. zeros).
 *
 * Preconditions (TC_GEMM / MSDA3D_EINVAL otherwise): device pointers, A and B 16-byte aligned, lda % 4 == 0, ldb % 4 == 0
 * (TMA global strides are multiples of 16 bytes); M, N, R > 0.  bias may be NULL.  Work is enqueued on `stream`; no
 * allocation, no synchronisation.  Returns 0 / MSDA3D_E* / cudaError_t (msda3d_error_string explains all three).
 */
#ifndef TC_GEMM_H_
#define TC_GEMM_H_

#ifdef __cplusplus
extern "C" {
#endif

int tc_gemm_tf32(void *stream, const float *A, int a_mn_major, long long lda, const float *B, int b_mn_major, long long ldb,
                 float *D, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k);

/* Same GEMM with two more epilogue stages, for the FFN of DefAttnLayer / FocusedDecoderLayer (linear2(dropout(relu(linear1(x)))),
 * decoder_blocks.py:172-175, focused_decoder.py:186-187):
 *   p_drop > 0 : after bias / ReLU, element (m, n) is dropped with probability p_drop (kept values scaled by 1 / (1 - p_drop)); the keep
 *                decision is the counter-based hash of (seed, m * N + n) also used by fused_ln.h -- no mask tensor is stored.
 *   gate       : D = result * (gate[m, n] > 0 ? gate_scale : 0), gate laid out like D: the backward of ReLU + dropout taken from the
 *                saved activation h = dropout(relu(.)) (h > 0 exactly where the unit was active and kept), fused into the
 *                grad_input GEMM of the second Linear.
 * Both need N % 4 == 0, ldd % 4 == 0, 16-byte aligned D / gate and accumulate == 0. */
int tc_gemm_tf32_ex(void *stream, const float *A, int a_mn_major, long long lda, const float *B, int b_mn_major, long long ldb,
                    float *D, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k,
                    const float *gate, float gate_scale, float p_drop, unsigned long long seed);

/* The same GEMM with bf16 operands (tcgen05.mma kind::f16, fp32 accumulation in TMEM) -- the Linear layers under the reference
 * trainer's autocast region (transoar/trainer.py:67-69; bf16 here, BASELINE configs[2] / [3]).  A / B are bf16 in either layout
 * (lda % 8 == 0, ldb % 8 == 0: TMA strides are multiples of 16 bytes); D is bf16 (out_fp32 == 0) or fp32 (out_fp32 != 0, required when
 * accumulate != 0: weight gradients go straight into fp32).  bias is fp32; gate has D's element type.  Same epilogue options, error
 * codes and stream semantics as tc_gemm_tf32_ex. */
int tc_gemm_bf16(void *stream, const void *A, int a_mn_major, long long lda, const void *B, int b_mn_major, long long ldb,
                 void *D, int out_fp32, long long ldd, const float *bias, int M, int N, int R, int relu, int accumulate, int split_k,
                 const void *gate, float gate_scale, float p_drop, unsigned long long seed);

/* Column sums out[c] = sum_rows x[row * ld + c] of a row-major fp32 matrix: the bias gradient of a Linear layer (autograd's
 * grad_output.sum(0), an ATen reduce kernel in the reference).  channels % 4 == 0, channels <= 1024, ld % 4 == 0, x and workspace
 * 16-byte aligned; workspace holds tc_colsum_workspace_floats(channels) floats.  One HBM pass + a tiny finalize kernel. */
long long tc_colsum_workspace_floats(int channels);
int tc_colsum(void *stream, const float *x, long long rows, int channels, long long ld, float *out, float *workspace);

/* Diagnostics: while a device buffer of 12 x 296 uint64 is installed, every launch writes per CTA the cycles its three roles spent
 * waiting: [0] producer on "empty", [1] producer total, [2] MMA on "full", [3] MMA on "tmem empty", [4] MMA total, [5] epilogue on
 * "tmem full", [6] epilogue total; then at 8 * 296 + 4 * cta: epilogue cycles in TMEM load / shared-memory stage / store.  NULL (the default) removes it.  Not thread-safe; tools/exp_gemm_roles.py uses it. */
void tc_gemm_debug_profile(unsigned long long *device_counters);

#ifdef __cplusplus
}
#endif
#endif /* TC_GEMM_H_ */
