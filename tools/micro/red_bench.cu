// Microbenchmark: throughput of fp32 reductions into global memory on B200.
//   mode 0: LSU path  -- red.global.add.v4.f32, 16 lanes x 16 B = one 256-byte segment per half-warp (what bwd_vec_kernel does)
//   mode 1: TMA path  -- cp.reduce.async.bulk.global.shared::cta.add.f32 of one 256-byte segment staged in shared memory
// Segments are pseudo-random 256-byte aligned chunks of a buffer of `span` bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(256) k_lsu(float* buf, uint32_t nseg, int iters)
{
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  const uint32_t warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t seg = hash32((warp * 2 + half) * 9973u + it * 8 + k) % nseg;
      float* p = buf + (size_t)seg * 64 + l16 * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
  }
}

__global__ void __launch_bounds__(256) k_tma(float* buf, uint32_t nseg, int iters)
{
  extern __shared__ __align__(128) float stage[];        // per warp: 2 buffers x 16 segments x 64 floats
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15, w = threadIdx.x >> 5;
  const uint32_t warp = blockIdx.x * 8 + w;
  float* mine = stage + w * (2 * 16 * 64);
  for (int it = 0; it < iters; ++it) {
    float* sb = mine + (it & 1) * (16 * 64);
    // buffer (it&1) was last used two iterations ago: wait until that bulk group has finished READING shared memory
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k)
      *reinterpret_cast<float4*>(sb + (half * 8 + k) * 64 + l16 * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < 16) {
      const int h = lane >> 3, k = lane & 7;
      const uint32_t seg = hash32((warp * 2 + h) * 9973u + it * 8 + k) % nseg;
      float* g = buf + (size_t)seg * 64;
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(sb + (h * 8 + k) * 64);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 256;" ::"l"(g), "r"(s) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
  const size_t spans[] = {360ull << 20, 60ull << 20, 8ull << 20};
  float* buf;
  cudaMalloc(&buf, spans[0]);
  cudaMemset(buf, 0, spans[0]);
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 16 * 64 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 64, grid = 148 * 24;
  for (size_t span : spans) {
    const uint32_t nseg = (uint32_t)(span / 256);
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_lsu<<<grid, 256>>>(buf, nseg, iters);
        else k_tma<<<grid, 256, 8 * 2 * 16 * 64 * 4>>>(buf, nseg, iters);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * 8 * iters * 16 * 256;
        if (rep == 2) printf("span %4zu MB  %s  %8.3f ms  %7.2f TB/s payload  (%s)\n", span >> 20, mode ? "TMA cp.reduce.async.bulk 256B" : "LSU red.v4.f32 16x16B       ", ms, bytes / ms / 1e9, cudaGetErrorString(err));
      }
    }
  }
  // sanity: sum of buffer must equal total added
  return 0;
}
