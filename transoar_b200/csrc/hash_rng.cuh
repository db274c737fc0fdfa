// hash_rng.cuh -- counter-based dropout masks shared by the fused LayerNorm (fused_ln_kernels.cuh) and the GEMM epilogue
// (tc_gemm_kernels.cuh): the keep decision of an element is a pure function of (seed, element index), so no mask tensor is ever
// stored -- the backward pass evaluates the same function again.
#pragma once

#include <stdint.h>

namespace hashrng {

// splitmix64 finaliser over (seed, index of the float4): four 16-bit uniforms per call, one per element of the float4
__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t idx)
{
  uint64_t x = idx * 0x9E3779B97F4A7C15ull + seed;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
// Under CUDA-graph replay a seed passed by value is frozen into the graph.  `epoch` (NULL = none) points to a device counter the
// caller advances between replays (hash_rng_set_epoch, include/fused_ln.h); it is folded into the seed at run time, so every replay
// draws new masks while a forward / backward pair inside one replay still sees the same ones.
__device__ __forceinline__ uint64_t with_epoch(uint64_t seed, const unsigned long long *epoch)
{
  return epoch == nullptr ? seed : seed + (uint64_t)(*epoch) * 0xD1342543DE82EF95ull;
}
// keep element j (0..3) of float4 number `idx` iff its 16-bit uniform >= thresh (thresh = round(p * 65536))
__device__ __forceinline__ void keep4(uint64_t seed, uint64_t idx, uint32_t thresh, float scale, float (&m)[4])
{
  const uint64_t r = mix64(seed, idx);
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = ((uint32_t)(r >> (16 * j)) & 0xFFFFu) >= thresh ? scale : 0.f;
}

}  // namespace hashrng
