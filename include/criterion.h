/*
 * criterion.h -- C ABI of the fused matcher + detection criterion in libmsda3d.so (sm_100a): the per-class matching of
 * transoar/models/matcher.py:22-65 (anchor matching, one match per class, soft labels) and the three losses of
 * transoar/models/criterion.py:40-77,92-125 (BCE on the soft labels, L1 and GIoU of the matched query, normalised by the number of target
 * boxes) for the final and the auxiliary decoder layers, together with the gradients of every loss with respect to the final layer's
 * logits and boxes -- one kernel launch instead of the ~300 small kernels of the batched torch mirror.
 *
 *   logits_layers fp32 [L, B, Nq]   the logits the matcher sees: layer 0 = final layer, 1.. = auxiliary layers (criterion.py:113-120)
 *   final_logits  fp32 [B, Nq]      what the losses are evaluated on (criterion.py:118-119 passes the FINAL outputs for every layer)
 *   final_boxes   fp32 [B, Nq, 6]   (cx, cy, cz, w, h, d) in [0, 1]
 *   anchors       fp32 [Nq, 6]      the box costs use the anchors (matcher.py:27-28); Nq = classes * queries_per_class, class c owns queries
 *                                   [c * Q, (c + 1) * Q)
 *   tgt_boxes     fp32 [B, O, 6], tgt_valid uint8 [B, O]    one box per class and sample, 0 where the class is absent
 *   losses        fp32 [3, L]       rows cls / bbox / giou, one column per layer; zero-filled by the call
 *   grad_logits   fp32 [B, Nq]      d cls_l / d final_logits (the same for every l: the soft labels do not depend on the layer)
 *   grad_boxes    fp32 [2, L, B, Nq, 6]   d bbox_l / d final_boxes and d giou_l / d final_boxes; every element is written
 *   best          int32 [L, B, O] or NULL: the matched query of each class inside its group
 * queries_per_class <= 32, layers <= 8.  Device pointers, enqueued on `stream`, no allocation, no synchronisation.  Returns 0 / MSDA3D_E* / cudaError_t.
 */
#ifndef CRITERION_H_
#define CRITERION_H_

#ifdef __cplusplus
extern "C" {
#endif

int criterion_fused_supported(int queries_per_class, int layers);

int criterion_fused(void *stream, const float *logits_layers, const float *final_logits, const float *final_boxes, const float *anchors,
                    const float *tgt_boxes, const unsigned char *tgt_valid, int layers, int batch, int classes, int queries_per_class,
                    float cost_class, float cost_bbox, float cost_giou, float *losses, float *grad_logits, float *grad_boxes, int *best);

#ifdef __cplusplus
}
#endif
#endif /* CRITERION_H_ */
