"""CPU checks for SURVEY 8 rows a7 / a10 and the model assembly: checkpoint-compatible names, anchors / restrictions."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN
sys.path.insert(0, GOLDEN)
import make_golden_model as G          # noqa: E402  (only its config dicts are used; the reference is not imported)
from transoar_b200.attn_fpn import AttnFPN
from transoar_b200.transoarnet import TransoarNet, generate_anchors


def test_attn_fpn_names_and_shapes_follow_the_reference():
    z = np.load(os.path.join(GOLDEN, "attn_fpn.npz"))
    cfg5 = dict(G.BACKBONE, conv_kernels=[[3, 3, 3]] * 5, strides=[[1, 1, 1]] + [[2, 2, 2]] * 4, feature_levels=["P2", "P3", "P4"], use_cuda=True)
    fpn = AttnFPN(cfg5)
    assert list(fpn.state_dict().keys()) == z["keys"].tolist()
    assert "_encoder._stages.0._block.0.weight" in fpn.state_dict() and "_decoder._refine.level_embed" in fpn.state_dict()
    # visceral yaml channel plan: 24..768 in the encoder, 384 in the FPN
    vis = dict(G.BACKBONE, start_channels=24, fpn_channels=384, hidden_dim=384, dim_feedforward=1024, n_points=4, use_cuda=True)
    big = AttnFPN(vis)
    enc = [s._block[0].out_channels for s in big._encoder._stages]
    assert enc == [24, 48, 96, 192, 384, 768]
    assert [c.in_channels for c in big._decoder._lateral] == [96, 192, 384, 768] and all(c.out_channels == 384 for c in big._decoder._out)
    assert sum(p.numel() for p in big._decoder._refine.parameters()) == 2 * (443520 + 384 * 1024 * 2 + 1024 + 384 + 4 * 384) + 4 * 384


def test_transoarnet_assembly_anchors_and_restrictions():
    z = np.load(os.path.join(GOLDEN, "transoarnet.npz"))
    cfg = {"backbone": dict(G.BACKBONE, start_channels=2, use_cuda=True), "neck": dict(G.NECK, nheads=3), "bbox_properties": G.PROPS}
    net = TransoarNet(cfg)
    assert list(net.state_dict().keys()) == z["keys"].tolist()
    assert np.allclose(net._anchors.numpy(), z["anchors"], atol=1e-7) and np.allclose(net._restrictions.numpy(), z["restrictions"], atol=1e-7)
    assert net._anchors.shape == (14, 6)
    a, r = generate_anchors(dict(G.NECK, num_queries=54, num_organs=2), G.PROPS)       # 27 offsets per organ (transoarnet.py:81-94)
    assert a.shape == (54, 6) and r.shape == (54, 6) and float(a.min()) >= 0 and float(a.max()) <= 1
    # at init all logits are 0 and all boxes are the anchors (transoarnet.py:50-58)
    assert float(net._cls_head.weight.abs().max()) == 0 and float(net._reg_head.layers[-1].weight.abs().max()) == 0
