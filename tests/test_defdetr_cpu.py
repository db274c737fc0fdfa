"""The restated 3D Deformable-DETR model (BASELINE configs[2]; transoar_b200/def_detr.py) on the host through the oracle's ATen route
(oracle/model_oracle.cpu_reference_ops swaps the three fused CUDA ops for the reference's use_cuda=False compositions): shapes, the
decoder's learnable reference points receive gradients through the unfused prologue, criterion + optimiser step run."""
import copy

import torch

from oracle.model_oracle import cpu_reference_ops
from transoar_b200.configs import VISCERAL_BACKBONE, synthetic_atlas
from transoar_b200.criterion import VISCERAL_LOSS_COEFS, build_criterion, total_loss
from transoar_b200.engine import build_model, optimizer_param_groups, synthetic_targets


def _tiny_config():
    bb = copy.deepcopy(VISCERAL_BACKBONE)
    bb.update(start_channels=4, fpn_channels=48, hidden_dim=48, dim_feedforward=64, out_fmaps=["P2", "P3", "P4", "P5"],
              feature_levels=["P2", "P3", "P4", "P5"], n_points=2, layers=1)
    neck = dict(name="def_detr", hidden_dim=48, dropout=0.0, nheads=4, dim_feedforward=64, dec_layers=2, n_points=2, num_queries=30,
                num_organs=15, aux_loss=True)
    return dict(model_family="def_detr", backbone=bb, neck=neck, bbox_properties=synthetic_atlas(15, 0), lr=2e-4, lr_backbone=2e-5,
                weight_decay=1e-4, anchor_matching=False, set_cost_class=1, set_cost_bbox=5, set_cost_giou=2,
                loss_coefs=copy.deepcopy(VISCERAL_LOSS_COEFS), num_classes=15)


def test_defdetr_forward_backward_on_the_cpu_route():
    cfg = _tiny_config()
    torch.manual_seed(0)
    net = build_model(cfg).train()
    crit = build_criterion(cfg)
    opt = torch.optim.AdamW(optimizer_param_groups(net, cfg), lr=2e-5, weight_decay=1e-4)
    x = torch.rand(2, 1, 32, 32, 64, generator=torch.Generator().manual_seed(1))
    tg = synthetic_targets(cfg, 2, 0, "cpu")
    losses = []
    with cpu_reference_ops():
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            out = net(x)
            assert out["pred_logits"].shape == (2, 30, 1) and out["pred_boxes"].shape == (2, 30, 6) and len(out["aux_outputs"]) == 1
            assert float(out["pred_boxes"].min()) >= 0 and float(out["pred_boxes"].max()) <= 1
            loss = total_loss(crit(out, tg, None, net._anchors), cfg["loss_coefs"])
            loss.backward()
            opt.step()
            losses.append(float(loss))
    assert all(v == v for v in losses)
    ref = net._neck.reference_points
    assert ref.weight.grad is not None and float(ref.weight.grad.abs().max()) > 0       # gradient reaches the learnable reference points
    cross = net._neck.layers[0].cross_attn
    assert cross.sampling_offsets.weight.grad is not None and cross.value_proj.weight.grad is not None
