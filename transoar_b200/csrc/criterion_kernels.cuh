// criterion_kernels.cuh -- matcher + detection losses + their gradients of the Focused-Decoder model in ONE kernel.
//
// Reference: transoar/models/matcher.py:22-65 (Matcher.forward: per (batch, class) the query with the smallest
// cost_class * (-sigmoid(logit)) + cost_bbox * L1(anchor, target) + cost_giou * (-GIoU(anchor, target)) inside the class's static query group;
// soft labels = the GIoU cost min-max normalised inside the group and clipped at 0, -1 for classes absent from the sample) and
// transoar/models/criterion.py:40-77,92-125 (TransoarCriterion: BCE-with-logits on the soft labels, L1 and GIoU of the matched query against the
// target, normalised by the number of target boxes; the auxiliary layers re-run the matcher on THEIR logits but all losses are evaluated on the
// FINAL layer's predictions, criterion.py:118-119).  The batched torch mirror of that (transoar_b200/criterion.py) is ~300 kernels of 2-3 us per
// training step; the problem is 2 x 20 x 27 queries.
//
// One CTA of 32 threads per (batch, class); lane q = query q of the class's group (Q <= 32).  Anchor matching only (the box costs use the anchors,
// matcher.py:27-28: they are the same for every decoder layer, so the soft labels and the classification loss are too).  Every CTA recomputes
// the two global normalisers (target boxes in the batch, kept logits) from the [B, O] validity mask, adds its loss terms into the zero-filled
// loss array and writes the gradient rows of its own queries -- every element of the gradient arrays is written, none needs a memset.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace crit {

struct Box { float lo[3], hi[3]; };

// (cx, cy, cz, w, h, d), clamped at 0 as the mirrors do before the conversion, -> corners
__device__ __forceinline__ Box corners(const float *b, bool clamp)
{
  Box r;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float c = clamp ? fmaxf(b[i], 0.f) : b[i], s = clamp ? fmaxf(b[3 + i], 0.f) : b[3 + i];
    r.lo[i] = c - 0.5f * s;
    r.hi[i] = c + 0.5f * s;
  }
  return r;
}

// GIoU of a against b (criterion.py paired_giou_3d = utils/bboxes.py:6-29,99-136) and, optionally, its gradient with respect to a's corners
__device__ __forceinline__ float giou(const Box &a, const Box &b, float *dlo, float *dhi)
{
  float sa[3], sb[3], in[3], hu[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sa[i] = a.hi[i] - a.lo[i];
    sb[i] = b.hi[i] - b.lo[i];
    in[i] = fmaxf(fminf(a.hi[i], b.hi[i]) - fmaxf(a.lo[i], b.lo[i]), 0.f);
    hu[i] = fmaxf(fmaxf(a.hi[i], b.hi[i]) - fminf(a.lo[i], b.lo[i]), 0.f);
  }
  const float va = sa[0] * sa[1] * sa[2], vb = sb[0] * sb[1] * sb[2];
  const float I = in[0] * in[1] * in[2], U = va + vb - I, H = hu[0] * hu[1] * hu[2];
  const float g = I / U - (H - U) / H;
  if (dlo != nullptr) {
    // g = I / U + U / H - 1:  dg = dI / U + dU (1 / H - I / U^2) - U / H^2 dH,  dU = dVa - dI
    const float cI = 1.f / U, cU = 1.f / H - I / (U * U), cH = -U / (H * H);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      const float dva = sa[j] * sa[k];                                   // d Va / d a.hi[i]  (= -d Va / d a.lo[i])
      const float din = in[i] > 0.f ? in[j] * in[k] : 0.f, dhu = hu[i] > 0.f ? hu[j] * hu[k] : 0.f;
      const float dI_hi = a.hi[i] < b.hi[i] ? din : (a.hi[i] == b.hi[i] ? 0.5f * din : 0.f);
      const float dI_lo = a.lo[i] > b.lo[i] ? -din : (a.lo[i] == b.lo[i] ? -0.5f * din : 0.f);
      const float dH_hi = a.hi[i] > b.hi[i] ? dhu : (a.hi[i] == b.hi[i] ? 0.5f * dhu : 0.f);
      const float dH_lo = a.lo[i] < b.lo[i] ? -dhu : (a.lo[i] == b.lo[i] ? -0.5f * dhu : 0.f);
      dhi[i] = cI * dI_hi + cU * (dva - dI_hi) + cH * dH_hi;
      dlo[i] = cI * dI_lo + cU * (-dva - dI_lo) + cH * dH_lo;
    }
  }
  return g;
}

constexpr int kMaxLayers = 8;

// logits_layers [L, B, Nq]: the logits the matcher sees (layer 0 = the final layer, then the auxiliary ones); final_logits [B, Nq] / final_boxes
// [B, Nq, 6]: what every loss is evaluated on; anchors [Nq, 6]; tgt_boxes [B, O, 6]; tgt_valid [B, O] (bytes).
// losses [3, L] (rows: cls, bbox, giou), zero-filled by the caller; grad_logits [B, Nq] = d cls / d final_logits (the same for every layer);
// grad_boxes [2, L, B, Nq, 6] = d bbox_l / d final_boxes, d giou_l / d final_boxes; best [L, B, O] (int32) = matched query inside the group.
static __global__ void __launch_bounds__(32)
criterion_kernel(const float *__restrict__ logits_layers, const float *__restrict__ final_logits, const float *__restrict__ final_boxes,
                 const float *__restrict__ anchors, const float *__restrict__ tgt_boxes, const unsigned char *__restrict__ tgt_valid, int L, int B, int O, int Q,
                 float cost_class, float cost_bbox, float cost_giou, float *__restrict__ losses, float *__restrict__ grad_logits,
                 float *__restrict__ grad_boxes, int *__restrict__ best_out)
{
  const int bo = blockIdx.x, b = bo / O, o = bo % O, q = threadIdx.x, Nq = O * Q;
  const bool active = q < Q;
  const unsigned full = 0xffffffffu;
  // global normalisers (criterion.py:98-99: number of target boxes, clamped at 1; kept logits = Q per present class)
  int nvalid = 0;
  for (int i = q; i < B * O; i += 32) nvalid += tgt_valid[i] ? 1 : 0;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) nvalid += __shfl_xor_sync(full, nvalid, s);
  const float num_boxes = fmaxf((float)nvalid, 1.f), kept = (float)nvalid * (float)Q;
  const bool valid = tgt_valid[bo] != 0;

  float tgt[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) tgt[i] = tgt_boxes[(long long)bo * 6 + i];
  const Box tb = corners(tgt, false);
  const long long row = (long long)b * Nq + o * Q + (active ? q : 0);      // this lane's query in [B, Nq]

  // layer-independent costs of this query's ANCHOR
  float c_bbox = 0.f, c_giou = 0.f;
  {
    float an[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { an[i] = anchors[(long long)(o * Q + (active ? q : 0)) * 6 + i]; c_bbox += fabsf(an[i] - tgt[i]); }
    c_giou = -giou(corners(an, true), tb, nullptr, nullptr);
  }
  // soft labels: min-max normalised GIoU cost inside the group, clipped at 0 (matcher.py:59); -1 for an absent class (matcher.py:44-45)
  float hi = active ? c_giou : -INFINITY, lo = active ? c_giou : INFINITY;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) { hi = fmaxf(hi, __shfl_xor_sync(full, hi, s)); lo = fminf(lo, __shfl_xor_sync(full, lo, s)); }
  const float soft = Q == 1 ? 1.f : fmaxf((c_giou - hi) / (lo - hi), 0.f);

  // classification: BCE with logits against the soft label, over the present classes (criterion.py:40-49); identical for every layer
  float cls_term = 0.f;
  if (active) {
    const float x = final_logits[row];
    float gx = 0.f;
    if (valid) {
      cls_term = (fmaxf(x, 0.f) - x * soft + log1pf(expf(-fabsf(x)))) / kept;
      gx = (1.f / (1.f + expf(-x)) - soft) / kept;
    }
    grad_logits[row] = gx;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) cls_term += __shfl_xor_sync(full, cls_term, s);

  for (int l = 0; l < L; ++l) {
    // the match of layer l: smallest cost, first query on ties (matcher.py:54)
    float cost = INFINITY;
    if (active) {
      const float x = logits_layers[((long long)l * B + b) * Nq + o * Q + q];
      cost = cost_bbox * c_bbox + cost_class * (-1.f / (1.f + expf(-x))) + cost_giou * c_giou;
    }
    int arg = q;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const float oc = __shfl_xor_sync(full, cost, s);
      const int oa = __shfl_xor_sync(full, arg, s);
      if (oc < cost || (oc == cost && oa < arg)) { cost = oc; arg = oa; }
    }
    if (q == 0 && best_out != nullptr) best_out[((long long)l * B + b) * O + o] = arg;
    // box losses of the matched query on the FINAL predictions (criterion.py:52-77)
    float gb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active && q == arg && valid) {
      float m[6], l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        m[i] = final_boxes[row * 6 + i];
        const float d = m[i] - tgt[i];
        l1 += fabsf(d);
        gb[i] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / num_boxes;
      }
      float dlo[3], dhi[3];
      const float g = giou(corners(m, true), tb, dlo, dhi);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // corners = c -+ s / 2 of the clamped box; the clamp passes gradients where the prediction is >= 0; loss = 1 - giou
        gg[i] = m[i] >= 0.f ? -(dlo[i] + dhi[i]) / num_boxes : 0.f;
        gg[3 + i] = m[3 + i] >= 0.f ? -0.5f * (dhi[i] - dlo[i]) / num_boxes : 0.f;
      }
      atomicAdd(losses + 1 * L + l, l1 / num_boxes);
      atomicAdd(losses + 2 * L + l, (1.f - g) / num_boxes);
    }
    if (active) {
      float *pb = grad_boxes + (((long long)(0 * L + l) * B * Nq) + row) * 6, *pg = grad_boxes + (((long long)(1 * L + l) * B * Nq) + row) * 6;
#pragma unroll
      for (int i = 0; i < 6; ++i) { pb[i] = gb[i]; pg[i] = gg[i]; }
    }
    if (q == 0) atomicAdd(losses + 0 * L + l, cls_term);
  }
}

}  // namespace crit
