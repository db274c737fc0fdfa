"""Matcher + criterion mirror (transoar_b200/criterion.py) against fixtures produced by the reference's Matcher /
TransoarCriterion (tests/golden/make_golden_criterion.py).  The mirror is device-agnostic torch code, so it is checked here
on the CPU; the same code runs on the GPU inside the training step (tests/test_gpu_model_step.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
sys.path.insert(0, GOLDEN)
import make_golden_criterion as G          # noqa: E402  (input generator only; the reference is not imported)
from transoar_b200.criterion import Matcher, TransoarCriterion, VISCERAL_LOSS_COEFS, dense_targets, total_loss


@pytest.mark.parametrize("name", list(G.CASES))
def test_matcher_and_losses_match_the_reference(name):
    z = np.load(os.path.join(GOLDEN, "criterion.npz"))
    c = G.case(**G.CASES[name])
    outs = c["outs"]
    outs["pred_logits"].requires_grad_(True)
    outs["pred_boxes"].requires_grad_(True)
    m = Matcher(*c["costs"], anchor_matching=c["anchor_matching"], num_organs=c["O"])
    crit = TransoarCriterion(c["O"], m, seg_proxy=False, seg_fg_bg=True)
    tb, tv = dense_targets(c["targets"], c["O"], torch.device("cpu"))
    best, soft = m(outs, tb, tv, c["anchors"])
    assert np.array_equal(Matcher.dense_matches(best, tv, c["Q"]).numpy(), z[f"{name}.matches"])
    assert np.allclose(soft.numpy(), z[f"{name}.soft"], atol=1e-6)
    losses = crit(outs, c["targets"], None, c["anchors"])
    for k, v in losses.items():
        assert np.allclose(v.detach().numpy(), z[f"{name}.loss.{k}"], rtol=1e-5, atol=1e-6), k
    total = total_loss(losses, VISCERAL_LOSS_COEFS)
    assert np.allclose(total.detach().numpy(), z[f"{name}.total"], rtol=1e-5)
    total.backward()
    assert np.allclose(outs["pred_logits"].grad.numpy(), z[f"{name}.grad_logits"], rtol=1e-4, atol=1e-7)
    assert np.allclose(outs["pred_boxes"].grad.numpy(), z[f"{name}.grad_boxes"], rtol=1e-4, atol=1e-7)


def test_absent_classes_never_produce_nans():
    c = G.case(seed=9, B=1, O=3, Q=7, missing={(0, 1), (0, 2)}, anchor_matching=True, costs=(1, 0, 0), n_aux=1)
    m = Matcher(1, 0, 0, anchor_matching=True, num_organs=3)
    crit = TransoarCriterion(3, m, False, True)
    c["outs"]["pred_boxes"].requires_grad_(True)
    total = total_loss(crit(c["outs"], c["targets"], None, c["anchors"]), VISCERAL_LOSS_COEFS)
    total.backward()
    assert torch.isfinite(total) and torch.isfinite(c["outs"]["pred_boxes"].grad).all()
