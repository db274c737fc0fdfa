"""Matcher + loss of the Focused-Decoder model, batched on the device -- mirrors of transoar/models/matcher.py:9-65
(``Matcher``) and transoar/models/criterion.py:9-125 (``TransoarCriterion``), SURVEY 8(f) rank 2.

The reference moves logits / boxes / targets to the CPU and loops over (batch, class) in Python with a ``topk`` per class
(matcher.py:30-63: 2 x 20 iterations and four host syncs per training step at VISCERAL).  Here the per-class costs of all
(batch, class) pairs are one [B, organs, queries-per-organ] tensor, the match is an ``argmin`` over its last axis and nothing
leaves the device, so a training step has no host synchronisation before the optimiser.

Behaviour kept from the reference on purpose:
  * queries are statically split into ``num_organs`` groups; class c (1-based label) only competes inside group c-1;
  * with ``anchor_matching`` the box costs use the ANCHORS, not the predictions (matcher.py:27-28);
  * soft labels = GIoU cost of each query against the class's target, min-max normalised inside the group and clipped at 0
    (matcher.py:59); a class absent from the sample gets soft label -1 and is dropped from the BCE (criterion.py:45-48);
  * the auxiliary-layer losses re-run the matcher on the auxiliary logits but evaluate the losses on the FINAL layer's
    predictions (criterion.py:118-119 pass ``outputs``, not ``aux_outputs``);
  * box losses are normalised by the number of target boxes in the batch.
Differences: ``num_top_queries`` is fixed at 1 (the only value the reference's criterion ever passes); on exact cost ties the
first query wins (the reference inherits whatever CPU ``topk`` returns); targets whose labels repeat inside a sample keep the
last box (the reference keeps the last one in its dict too, matcher.py:34)."""
import os

import torch
import torch.nn.functional as F
from torch import nn


def box_cxcyczwhd_to_xyzxyz(b):
    c, s = b[..., :3], b[..., 3:]
    return torch.cat((c - 0.5 * s, c + 0.5 * s), dim=-1)


def paired_giou_3d(a, b):
    """Generalised IoU of box pairs (xyzxyz, broadcastable leading dims) -- utils/bboxes.py:6-29,99-136 for matched pairs."""
    vol = lambda x: (x[..., 3] - x[..., 0]) * (x[..., 4] - x[..., 1]) * (x[..., 5] - x[..., 2])
    # the product of the three extents written out: autograd's prod() backward looks for zeros with nonzero(), a host synchronisation
    # (and not capturable into a CUDA graph)
    prod3 = lambda e: e[..., 0] * e[..., 1] * e[..., 2]
    inter = prod3((torch.minimum(a[..., 3:], b[..., 3:]) - torch.maximum(a[..., :3], b[..., :3])).clamp(min=0))
    union = vol(a) + vol(b) - inter
    hull = prod3((torch.maximum(a[..., 3:], b[..., 3:]) - torch.minimum(a[..., :3], b[..., :3])).clamp(min=0))
    return inter / union - (hull - union) / hull


def dense_targets(targets, num_organs, device):
    """List of {'boxes': [n,6], 'labels': [n] (1-based)} -> (boxes [B, organs, 6], valid [B, organs]); no host sync."""
    B = len(targets)
    boxes = torch.zeros(B, num_organs, 6, dtype=torch.float32, device=device)
    valid = torch.zeros(B, num_organs, dtype=torch.bool, device=device)
    for b, t in enumerate(targets):
        idx = t["labels"].to(device=device, dtype=torch.long) - 1
        boxes[b].index_copy_(0, idx, t["boxes"].to(device=device, dtype=torch.float32))
        valid[b].index_fill_(0, idx, True)
    return boxes, valid


class Matcher(nn.Module):
    """matcher.py:9-65.  ``forward`` returns (matched query index [B, organs] (long), soft_labels [B, organs, Q]) plus the
    reference's dense 0/1 ``matches`` tensor on request."""

    def __init__(self, cost_class=1, cost_bbox=1, cost_giou=1, anchor_matching=True, num_organs=None):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs can't be 0"
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.anchor_matching, self.num_organs = anchor_matching, num_organs

    @torch.no_grad()
    def forward(self, outputs, tgt_boxes, tgt_valid, anchors):
        best, soft = self.match_layers([outputs], tgt_boxes, tgt_valid, anchors)
        return best[0], soft[0]

    @torch.no_grad()
    def match_layers(self, layers, tgt_boxes, tgt_valid, anchors):
        """The matcher for several decoder layers at once (the final layer + the auxiliary ones, criterion.py:113-120): everything carries
        a leading layer axis, so the ~40 tiny kernels of one match are launched once, not once per layer.  Returns best [L, B, organs],
        soft_labels [L, B, organs, Q]."""
        logits = torch.stack([o["pred_logits"] for o in layers])                            # [L, B, Nq, 1]
        L, B, Nq, _ = logits.shape
        O = self.num_organs
        Q = Nq // O
        if self.anchor_matching:
            boxes = anchors[None, None].expand(1, B, -1, -1)                                 # the same for every layer: computed once
        else:
            boxes = torch.stack([o["pred_boxes"] for o in layers])
        boxes = boxes.reshape(boxes.shape[0], B, O, Q, -1).float()
        logits = logits.reshape(L, B, O, Q).float()
        tgt = tgt_boxes[None, :, :, None, :]                                                 # [1, B, O, 1, 6]
        cost_class = -logits.sigmoid()
        cost_bbox = (boxes - tgt).abs().sum(-1)                                              # cdist(p=1), matcher.py:50
        cost_giou = -paired_giou_3d(box_cxcyczwhd_to_xyzxyz(boxes.clamp(min=0)), box_cxcyczwhd_to_xyzxyz(tgt))
        cost = self.cost_bbox * cost_bbox + self.cost_class * cost_class + self.cost_giou * cost_giou
        best = cost.argmin(-1)                                                               # topk(C, 1, largest=False), matcher.py:54
        if Q == 1:                                                                           # matcher.py:60-62 (0-dim topk result)
            soft = torch.ones_like(cost_giou)
        else:
            hi, lo = cost_giou.amax(-1, keepdim=True), cost_giou.amin(-1, keepdim=True)
            soft = ((cost_giou - hi) / (lo - hi)).clamp(min=0)                               # matcher.py:59
        soft = torch.where(tgt_valid[None, :, :, None], soft, torch.full_like(soft, -1.0))   # matcher.py:44-45
        return best, soft.expand(L, -1, -1, -1)

    @staticmethod
    def dense_matches(best, tgt_valid, Q):
        """The reference's [B, organs, Q] 0/1 tensor (matcher.py:40,57)."""
        return (F.one_hot(best, Q) * tgt_valid[:, :, None]).long()


class SoftDiceLoss(nn.Module):
    """criterion.py:127-166 (batch dice, softmax non-linearity, background dropped)."""

    def __init__(self, smooth_nom=1e-5, smooth_denom=1e-5):
        super().__init__()
        self.smooth_nom, self.smooth_denom = smooth_nom, smooth_denom

    def forward(self, inp, target):
        p = inp.softmax(1)
        onehot = torch.zeros_like(p).scatter_(1, target[:, None].long(), 1)
        axes = [0] + list(range(2, p.dim()))
        tp = (p * onehot).sum(axes)
        fp = (p * (1 - onehot)).sum(axes)
        fn = ((1 - p) * onehot).sum(axes)
        dc = (2 * tp + self.smooth_nom) / (2 * tp + fp + fn + self.smooth_denom)
        return 1 - dc[1:].mean()


class FusedCriterionFunction(torch.autograd.Function):
    """Matcher + the three losses of every decoder layer + their gradients in ONE kernel launch (include/criterion.h).  Returns ``losses``
    [3, L] (rows cls / bbox / giou; column 0 = final layer); gradients flow to the final layer's logits and boxes only -- the auxiliary
    layers' logits feed the (non-differentiable) matcher alone, as in the reference (criterion.py:113-120)."""

    @staticmethod
    def forward(ctx, logits_layers, final_logits, final_boxes, anchors, tgt_boxes, tgt_valid, num_organs, cost_class, cost_bbox, cost_giou):
        import ctypes
        from . import _lib
        L, B, Nq = logits_layers.shape
        Q = Nq // num_organs
        f32 = lambda t: t.detach().float().contiguous()
        ll, fl, fb, an, tb = f32(logits_layers), f32(final_logits).reshape(B, Nq), f32(final_boxes), f32(anchors), f32(tgt_boxes)
        tv = tgt_valid.to(torch.uint8).contiguous()
        dev = fl.device
        losses = torch.empty(3, L, dtype=torch.float32, device=dev)
        g_logits = torch.empty(B, Nq, dtype=torch.float32, device=dev)
        g_boxes = torch.empty(2, L, B, Nq, 6, dtype=torch.float32, device=dev)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rc = _lib.lib().criterion_fused(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), p(ll), p(fl), p(fb), p(an), p(tb), p(tv), L, B,
                                            num_organs, Q, float(cost_class), float(cost_bbox), float(cost_giou), p(losses), p(g_logits), p(g_boxes), None)
        _lib.check(rc, "criterion_fused")
        ctx.save_for_backward(g_logits, g_boxes)
        ctx.shapes = (final_logits.shape, final_logits.dtype, final_boxes.dtype)
        return losses

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g_logits, g_boxes = ctx.saved_tensors
        shape, ldt, bdt = ctx.shapes
        d_logits = (g_logits * g[0].sum()).reshape(shape).to(ldt) if ctx.needs_input_grad[1] else None
        d_boxes = torch.einsum("kl,klbnd->bnd", g[1:], g_boxes).to(bdt) if ctx.needs_input_grad[2] else None
        return None, d_logits, d_boxes, None, None, None, None, None, None, None


class TransoarCriterion(nn.Module):
    """criterion.py:9-125.  ``forward(outputs, targets, seg_targets, anchors)`` -> the reference's loss dict."""

    def __init__(self, num_classes, matcher, seg_proxy, seg_fg_bg):
        super().__init__()
        self.num_classes, self.matcher = num_classes, matcher
        self._seg_proxy, self._seg_fg_bg = seg_proxy, seg_fg_bg
        if seg_proxy:
            self._dice_loss = SoftDiceLoss()

    # the single-kernel route (include/criterion.h) on CUDA; False = the batched torch route (tests compare the two; TRANSOAR_B200_CRITERION=torch
    # selects it for A/B timing)
    fused = os.environ.get("TRANSOAR_B200_CRITERION", "fused") != "torch"

    def _fused_ok(self, outputs, layers):
        lg = outputs["pred_logits"]
        if not (self.fused and lg.is_cuda and self.matcher.anchor_matching and not self._seg_proxy and self.matcher.num_organs == self.num_classes):
            return False
        from . import _lib
        return lg.shape[1] % self.num_classes == 0 and bool(_lib.lib().criterion_fused_supported(lg.shape[1] // self.num_classes, len(layers)))

    def loss_class(self, outputs, soft_labels):
        """BCE-with-logits over the queries of the classes present in the sample (criterion.py:40-49).  ``soft_labels`` may carry a leading
        layer axis ([L, B, organs, Q]): one loss per layer, all evaluated on these logits."""
        lead = soft_labels.shape[:-3]
        logits = outputs["pred_logits"].flatten().float()
        labels = soft_labels.reshape(*lead, -1)
        keep = (labels != -1).float()
        per = F.binary_cross_entropy_with_logits(logits.expand_as(labels), labels.clamp(min=0), reduction="none")
        return (per * keep).sum(-1) / keep.sum(-1)

    def loss_bboxes(self, outputs, tgt_boxes, tgt_valid, best, num_boxes):
        """L1 + GIoU of the matched query of every present class against its target (criterion.py:52-77).  ``best`` may carry a leading
        layer axis ([L, B, organs]): one pair of losses per layer."""
        B, Nq, _ = outputs["pred_boxes"].shape
        O = self.num_classes
        lead = best.shape[:-2]
        preds = outputs["pred_boxes"].reshape(B, O, Nq // O, -1).float().expand(*lead, -1, -1, -1, -1)
        matched = torch.gather(preds, -2, best[..., None, None].expand(*best.shape, 1, preds.shape[-1])).squeeze(-2)   # [(L,) B, O, 6]
        w = tgt_valid.float()
        loss_bbox = ((matched - tgt_boxes).abs().sum(-1) * w).sum((-2, -1)) / num_boxes
        # classes absent from a sample have a zero target box (0/0 GIoU): give them a harmless stand-in box, so that neither the value nor --
        # through `where`'s zero gradient times an infinite derivative -- the gradient can turn into NaN, then mask them out of the sum
        safe_tgt = torch.where(tgt_valid[..., None], tgt_boxes, torch.full_like(tgt_boxes, 0.5))
        giou = paired_giou_3d(box_cxcyczwhd_to_xyzxyz(matched.clamp(min=0)), box_cxcyczwhd_to_xyzxyz(safe_tgt))
        loss_giou = (torch.where(tgt_valid, 1 - giou, torch.zeros_like(giou))).sum((-2, -1)) / num_boxes
        return loss_bbox, loss_giou

    def loss_segmentation(self, outputs, targets):
        if self._seg_fg_bg:
            targets = (targets > 0).to(targets.dtype)
        targets = targets.squeeze(1).long()
        return F.cross_entropy(outputs["pred_seg"], targets), self._dice_loss(outputs["pred_seg"], targets)

    def forward(self, outputs, targets, seg_targets, anchors):
        dev = outputs["pred_logits"].device
        tgt_boxes, tgt_valid = targets if isinstance(targets, tuple) else dense_targets(targets, self.num_classes, dev)
        num_boxes = tgt_valid.sum().clamp(min=1).float()
        # the final layer and the auxiliary layers in ONE pass (leading layer axis): the reference re-runs the matcher on every auxiliary
        # layer's logits but evaluates all losses on the FINAL layer's predictions (criterion.py:118-119 pass `outputs`, not `aux_outputs`)
        layers = [outputs] + list(outputs.get("aux_outputs", []))
        if self._fused_ok(outputs, layers):
            logits_layers = torch.stack([o["pred_logits"].detach() for o in layers]).squeeze(-1)
            m = self.matcher
            loss = FusedCriterionFunction.apply(logits_layers, outputs["pred_logits"], outputs["pred_boxes"], anchors, tgt_boxes, tgt_valid,
                                                self.num_classes, m.cost_class, m.cost_bbox, m.cost_giou)
            zero = torch.zeros((), device=dev)
            losses = {"bbox": loss[1, 0], "giou": loss[2, 0], "cls": loss[0, 0], "segce": zero, "segdice": zero}
            for i in range(len(layers) - 1):
                losses[f"bbox_{i}"], losses[f"giou_{i}"], losses[f"cls_{i}"] = loss[1, i + 1], loss[2, i + 1], loss[0, i + 1]
            return losses
        best, soft = self.matcher.match_layers(layers, tgt_boxes, tgt_valid, anchors)
        loss_bbox, loss_giou = self.loss_bboxes(outputs, tgt_boxes, tgt_valid, best, num_boxes)
        loss_cls = self.loss_class(outputs, soft)
        zero = torch.zeros((), device=dev)
        losses = {"bbox": loss_bbox[0], "giou": loss_giou[0], "cls": loss_cls[0], "segce": zero, "segdice": zero}
        if self._seg_proxy:
            losses["segce"], losses["segdice"] = self.loss_segmentation(outputs, seg_targets)
        for i in range(len(layers) - 1):
            losses[f"bbox_{i}"], losses[f"giou_{i}"], losses[f"cls_{i}"] = loss_bbox[i + 1], loss_giou[i + 1], loss_cls[i + 1]
        return losses


def total_loss(loss_dict, loss_coefs):
    """trainer.py:71-74: sum of loss * coefficient, coefficient looked up by the part of the key before '_'.  One stacked dot product on the
    device (two kernels) instead of a multiply and an add per entry."""
    keys = tuple(loss_dict)
    vals = torch.stack([loss_dict[k].float() for k in keys])
    cache_key = (keys, tuple(sorted(loss_coefs.items())), vals.device)
    coefs = _COEF_CACHE.get(cache_key)
    if coefs is None:                                    # built once (outside any graph capture: the eager warm-up steps come first)
        coefs = _COEF_CACHE[cache_key] = torch.tensor([float(loss_coefs[k.split("_")[0]]) for k in keys], dtype=torch.float32, device=vals.device)
    return (vals * coefs).sum()


_COEF_CACHE = {}


VISCERAL_LOSS_COEFS = {"cls": 2, "bbox": 5, "giou": 2, "segce": 2, "segdice": 2}      # config/attn_fpn_foc_dec_visceral.yaml:36-41


def build_criterion(config):
    """models/build.py:32-48."""
    matcher = Matcher(cost_class=config["set_cost_class"], cost_bbox=config["set_cost_bbox"], cost_giou=config["set_cost_giou"],
                      anchor_matching=config["anchor_matching"], num_organs=config["neck"]["num_organs"])
    return TransoarCriterion(num_classes=config["num_classes"], matcher=matcher, seg_proxy=config["backbone"]["use_seg_proxy_loss"],
                             seg_fg_bg=config["backbone"]["fg_bg"])
