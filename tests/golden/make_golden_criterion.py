"""Golden fixtures for the matcher + criterion from the REFERENCE's Matcher / TransoarCriterion on CPU.
Build-container only:   python tests/golden/make_golden_criterion.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def case(seed, B, O, Q, missing, anchor_matching, costs, n_aux):
    g = torch.Generator().manual_seed(seed)
    Nq = O * Q
    rnd = lambda *s: torch.rand(*s, generator=g)
    box = lambda *s: torch.cat((rnd(*s, 3) * 0.5 + 0.25, rnd(*s, 3) * 0.3 + 0.05), -1)
    anchors = box(Nq)
    outs = {"pred_logits": torch.randn(B, Nq, 1, generator=g), "pred_boxes": box(B, Nq)}
    outs["aux_outputs"] = [{"pred_logits": torch.randn(B, Nq, 1, generator=g), "pred_boxes": box(B, Nq)} for _ in range(n_aux)]
    targets = []
    for b in range(B):
        labels = torch.tensor([c for c in range(1, O + 1) if (b, c) not in missing])
        targets.append({"labels": labels, "boxes": box(len(labels))})
    return dict(anchors=anchors, outs=outs, targets=targets, O=O, Q=Q, anchor_matching=anchor_matching, costs=costs)


CASES = {
    "visceral_like": dict(seed=1, B=2, O=20, Q=27, missing={(0, 3), (1, 20), (1, 7)}, anchor_matching=True, costs=(1, 0, 0), n_aux=2),
    "pred_matching": dict(seed=2, B=3, O=4, Q=7, missing={(2, 1)}, anchor_matching=False, costs=(1, 5, 2), n_aux=1),
    "one_query": dict(seed=3, B=2, O=5, Q=1, missing=set(), anchor_matching=True, costs=(1, 1, 1), n_aux=0),
}


def main():
    sys.path.insert(0, "/root/reference")
    torch.Tensor.cuda = lambda self, *a, **k: self                 # criterion.py:48 calls .cuda() on the labels
    from transoar.models.criterion import TransoarCriterion
    from transoar.models.matcher import Matcher
    blob = {}
    for name, spec in CASES.items():
        c = case(**spec)
        m = Matcher(*c["costs"], anchor_matching=c["anchor_matching"], num_organs=c["O"])
        crit = TransoarCriterion(c["O"], m, seg_proxy=False, seg_fg_bg=True)
        outs = c["outs"]
        leaves = [outs["pred_logits"], outs["pred_boxes"]]
        for t in leaves:
            t.requires_grad_(True)
        matches, soft = m(outs, c["targets"], c["anchors"])
        losses = crit(outs, c["targets"], None, c["anchors"])
        coefs = {"cls": 2, "bbox": 5, "giou": 2, "segce": 2, "segdice": 2}
        total = sum(v * coefs[k.split("_")[0]] for k, v in losses.items())
        total.backward()
        blob[f"{name}.matches"] = matches.numpy()
        blob[f"{name}.soft"] = soft.numpy()
        blob[f"{name}.total"] = total.detach().numpy()
        blob[f"{name}.grad_logits"] = outs["pred_logits"].grad.numpy()
        blob[f"{name}.grad_boxes"] = outs["pred_boxes"].grad.numpy()
        for k, v in losses.items():
            blob[f"{name}.loss.{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "criterion.npz"), **blob)
    print("wrote criterion.npz:", len(blob), "arrays")


if __name__ == "__main__":
    main()
