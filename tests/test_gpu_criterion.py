"""The single-kernel matcher + criterion (include/criterion.h) against the batched torch route of transoar_b200/criterion.py -- which is
pinned to the reference's Matcher / TransoarCriterion by tests/test_criterion_cpu.py -- on the same inputs: every loss of every decoder layer
and the gradients reaching the final layer's logits and boxes."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, O, Q, L, seed, absent=0.3):
    g = torch.Generator().manual_seed(seed)
    Nq = O * Q
    rnd = lambda *s: torch.rand(*s, generator=g)
    anchors = torch.cat((rnd(Nq, 3) * 0.6 + 0.2, rnd(Nq, 3) * 0.3 + 0.05), -1)
    tgt_boxes = torch.cat((rnd(B, O, 3) * 0.6 + 0.2, rnd(B, O, 3) * 0.3 + 0.05), -1)
    tgt_valid = rnd(B, O) > absent
    tgt_boxes = tgt_boxes * tgt_valid[..., None]
    layers = []
    for _ in range(L):
        layers.append({"pred_logits": torch.randn(B, Nq, 1, generator=g).to(DEV), "pred_boxes": (anchors[None] + 0.1 * torch.randn(B, Nq, 6, generator=g)).to(DEV)})
    # a few negative box components exercise the clamp's gradient mask
    layers[0]["pred_boxes"][0, :3, 3] = -0.01
    return layers, anchors.to(DEV), tgt_boxes.to(DEV), tgt_valid.to(DEV)


@pytest.mark.parametrize("B,O,Q,L", [(2, 20, 27, 3), (1, 15, 27, 1), (3, 4, 7, 2), (2, 5, 1, 3), (2, 3, 32, 4)])
def test_fused_criterion_matches_the_torch_route(B, O, Q, L):
    from transoar_b200.criterion import Matcher, TransoarCriterion, VISCERAL_LOSS_COEFS, total_loss
    layers, anchors, tgt_boxes, tgt_valid = _inputs(B, O, Q, L, seed=B * 100 + O * 10 + Q)
    crit = TransoarCriterion(O, Matcher(cost_class=2, cost_bbox=5, cost_giou=2, anchor_matching=True, num_organs=O), seg_proxy=False, seg_fg_bg=True)
    got = {}
    for fused in (True, False):
        crit.fused = fused
        out = {k: v.detach().clone().requires_grad_(True) for k, v in layers[0].items()}
        out["aux_outputs"] = [{k: v.detach().clone().requires_grad_(True) for k, v in lay.items()} for lay in layers[1:]]
        losses = crit(out, (tgt_boxes, tgt_valid), None, anchors)
        total_loss(losses, VISCERAL_LOSS_COEFS).backward()
        got[fused] = (losses, out["pred_logits"].grad, out["pred_boxes"].grad, [a["pred_logits"].grad for a in out["aux_outputs"]])
    lf, lt = got[True][0], got[False][0]
    assert set(lf) == set(lt)
    for k in lt:
        assert abs(float(lf[k]) - float(lt[k])) <= 1e-5 * max(1.0, abs(float(lt[k]))), (k, float(lf[k]), float(lt[k]))
    for a, b in zip(got[True][1:3], got[False][1:3]):
        assert torch.isfinite(a).all() and torch.isfinite(b).all()
        assert float((a - b).abs().max()) <= 1e-6 + 1e-4 * float(b.abs().max()), float((a - b).abs().max())
    assert all(g is None for g in got[True][3]) and all(g is None for g in got[False][3])       # auxiliary logits only feed the matcher


def test_all_classes_absent_is_finite_where_the_torch_route_is():
    from transoar_b200.criterion import Matcher, TransoarCriterion
    layers, anchors, tgt_boxes, tgt_valid = _inputs(2, 6, 27, 2, seed=3, absent=0.0)
    tgt_valid[0] = False
    tgt_boxes[0] = 0
    crit = TransoarCriterion(6, Matcher(num_organs=6), seg_proxy=False, seg_fg_bg=True)
    out = dict(layers[0], aux_outputs=layers[1:])
    for v in crit(out, (tgt_boxes, tgt_valid), None, anchors).values():
        assert torch.isfinite(v).all()
