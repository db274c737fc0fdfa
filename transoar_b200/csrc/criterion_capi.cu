// criterion_capi.cu -- C ABI of the fused matcher + criterion kernel (include/criterion.h).
#include "criterion_kernels.cuh"

#include <atomic>

#include "../../include/criterion.h"
#include "../../include/msda3d.h"

extern std::atomic<unsigned long long> g_msda3d_launches;

extern "C" int criterion_fused_supported(int queries_per_class, int layers) { return queries_per_class >= 1 && queries_per_class <= 32 && layers >= 1 && layers <= crit::kMaxLayers; }

extern "C" int criterion_fused(void *stream, const float *logits_layers, const float *final_logits, const float *final_boxes, const float *anchors,
                               const float *tgt_boxes, const unsigned char *tgt_valid, int layers, int batch, int classes, int queries_per_class,
                               float cost_class, float cost_bbox, float cost_giou, float *losses, float *grad_logits, float *grad_boxes, int *best)
{
  if (!logits_layers || !final_logits || !final_boxes || !anchors || !tgt_boxes || !tgt_valid || !losses || !grad_logits || !grad_boxes) return MSDA3D_EINVAL;
  if (batch <= 0 || classes <= 0 || !criterion_fused_supported(queries_per_class, layers)) return MSDA3D_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const cudaError_t e = cudaMemsetAsync(losses, 0, sizeof(float) * 3 * layers, st);
  if (e != cudaSuccess) return (int)e;
  crit::criterion_kernel<<<batch * classes, 32, 0, st>>>(logits_layers, final_logits, final_boxes, anchors, tgt_boxes, tgt_valid, layers, batch, classes,
                                                        queries_per_class, cost_class, cost_bbox, cost_giou, losses, grad_logits, grad_boxes, best);
  ++g_msda3d_launches;
  return (int)cudaGetLastError();
}
