/*
 * ref_cuda_shim.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * Gives the REFERENCE's own CUDA launchers a C ABI so tests/ and bench.py can run "the reference's compiled
 * op" on the GPU box without torch's extension machinery.  The kernels are not restated here: the header is
 * compiled from where it lies in /root/reference (oracle/Makefile passes -I<reference>/transoar/models/ops/src).
 *
 *   ms_deformable_im2col_cuda<T>   transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:1094-1125
 *   ms_deformable_col2im_cuda<T>   transoar/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:1127-1507
 *
 * All pointers are device pointers (including the int64 shape arrays), exactly as in
 * ms_deform_attn_cuda.cu:64-73,135-147.  The callee does NOT zero outputs (the reference's host wrapper does,
 * ms_deform_attn_cuda.cu:54,122-124): callers must pass zero-filled out / grad_* buffers.
 */
#include "cuda/ms_deform_im2col_cuda.cuh"

#define DEFINE(SUF, T)                                                                                         \
  extern "C" int msda3d_refcuda_forward_##SUF(void *stream, const T *value, const int64_t *shapes,            \
                                              const int64_t *starts, const T *loc, const T *aw, int N, int S, \
                                              int M, int C, int L, int Lq, int P, T *out)                     \
  {                                                                                                            \
    ms_deformable_im2col_cuda<T>((cudaStream_t)stream, value, shapes, starts, loc, aw, N, S, M, C, L, Lq, P,  \
                                 out);                                                                         \
    return (int)cudaGetLastError();                                                                            \
  }                                                                                                            \
  extern "C" int msda3d_refcuda_backward_##SUF(void *stream, const T *grad_out, const T *value,               \
                                               const int64_t *shapes, const int64_t *starts, const T *loc,    \
                                               const T *aw, int N, int S, int M, int C, int L, int Lq, int P, \
                                               T *grad_value, T *grad_loc, T *grad_aw)                        \
  {                                                                                                            \
    ms_deformable_col2im_cuda<T>((cudaStream_t)stream, grad_out, value, shapes, starts, loc, aw, N, S, M, C,  \
                                 L, Lq, P, grad_value, grad_loc, grad_aw);                                    \
    return (int)cudaGetLastError();                                                                            \
  }

DEFINE(f32, float)
DEFINE(f64, double)

extern "C" int msda3d_refcuda_abi_version(void) { return 1; }
