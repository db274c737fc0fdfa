/*
 * msda3d_oracle.c -- CPU restatement of the reference's 3D multi-scale deformable attention op.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (transoar_b200/, libmsda3d.so) never links, imports or calls it.
 *
 * Restates (file:line under /root/reference/transoar/models/ops/src/):
 *   cuda/ms_deform_im2col_cuda.cuh:31-114    trilinear sample, corner + weight order
 *   cuda/ms_deform_im2col_cuda.cuh:116-241   backward of one sample (grad_value / grad_loc / grad_attn)
 *   cuda/ms_deform_im2col_cuda.cuh:370-439   forward kernel: level/point loop, pixel coords, range test
 *   cuda/ms_deform_im2col_cuda.cuh:551-661   backward kernel loop (C = 64 variant; all variants share the math)
 *   cuda/ms_deform_attn_cuda.cu:40-60,122-124 shape conventions, zero-initialised outputs
 *
 * Parity pin: the reference holds no golden vectors (ops/test.py uses unseeded random inputs and checks
 * CUDA == ms_deform_attn_core_pytorch).  This oracle is pinned against
 *   (1) the tests/golden npz fixtures -- outputs of the reference's own Python path (imported from /root/reference by
 *       tests/golden/make_golden.py in the build container), fp64 allclose + fp32 rtol 1e-2/atol 1e-3
 *       exactly as ops/test.py:69-97, plus autograd gradients of that path;
 *   (2) on the GPU box, oracle/_ref/libmsda3d_refcuda.so -- the reference's own CUDA kernels compiled from
 *       where they lie (oracle/Makefile) -- bit-exact forward in `contract` mode.
 */
#include <math.h>
#include <stdint.h>

#define T float
#define SUF f32
#define FMA_ fmaf
#define FLOOR_ floorf
#include "msda3d_oracle_impl.h"
#undef T
#undef SUF
#undef FMA_
#undef FLOOR_

#define T double
#define SUF f64
#define FMA_ fma
#define FLOOR_ floor
#include "msda3d_oracle_impl.h"
#undef T
#undef SUF
#undef FMA_
#undef FLOOR_

int msda3d_oracle_abi_version(void) { return 1; }
