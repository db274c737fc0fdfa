"""Golden fixtures for SURVEY 8 rows a7 / a10 from the REFERENCE's AttnFPN and TransoarNet (CPU, eval mode, use_cuda=False,
use_decoder_attn=True).  Build-container only:   python tests/golden/make_golden_model.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


BACKBONE = dict(name="attn_fpn", use_encoder_attn=False, conv_kernels=[[3, 3, 3]] * 6, strides=[[1, 1, 1]] + [[2, 2, 2]] * 5,
                in_channels=1, start_channels=4, depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=[5, 5, 5], mlp_ratio=4,
                qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, conv_merging=False,
                use_decoder_attn=True, fpn_channels=48, out_fmaps=["P2"], pos_encoding="sine", feature_levels=["P2", "P3", "P4", "P5"],
                hidden_dim=48, dim_feedforward=64, dropout=0.1, nheads=6, layers=2, n_points=2, use_cuda=False,
                use_seg_proxy_loss=False, fg_bg=True)
NECK = dict(name="foc_attn", pos_encoding="sine", input_levels="P5", hidden_dim=48, dropout=0.1, nheads=2, dim_feedforward=64,
            dec_layers=2, restrict_attn=True, obj_self_attn=False, anchor_gen_dynamic_offset=True, anchor_gen_offset=0.1,
            anchor_offset_pred=True, max_anchor_pred_offset=0.1, num_queries=14, num_organs=2, aux_loss=True)
PROPS = {"1": {"median": [0.30, 0.36, 0.40, 0.20, 0.22, 0.30], "min": [0.25, 0.30, 0.35, 0.16, 0.18, 0.26],
               "max": [0.35, 0.42, 0.45, 0.26, 0.28, 0.36], "attn_area": [0.05, 0.10, 0.0, 0.55, 0.62, 0.8]},
         "2": {"median": [0.70, 0.60, 0.60, 0.30, 0.30, 0.40], "min": [0.65, 0.55, 0.55, 0.26, 0.26, 0.36],
               "max": [0.75, 0.65, 0.65, 0.36, 0.36, 0.46], "attn_area": [0.4, 0.3, 0.2, 1.0, 0.9, 1.0]}}


def _patch_environment():
    """Only inside the generator process: reference on sys.path, timm shim, .cuda() -> no-op."""
    sys.path.insert(0, "/root/reference")
    timm = types.ModuleType("timm"); timm.models = types.ModuleType("timm.models"); timm.models.layers = types.ModuleType("timm.models.layers")
    timm.models.layers.trunc_normal_ = torch.nn.init.trunc_normal_


    class _DropPath(torch.nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()


    timm.models.layers.DropPath = _DropPath
    sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.layers": timm.models.layers})
    torch.Tensor.cuda = lambda self, *a, **k: self


def main():
    _patch_environment()
    from transoar.models.backbones.attn_fpn import AttnFPN
    from transoar.models.transoarnet import TransoarNet
    sys.path.insert(0, HERE)
    from detfill import det_fill_module, det_tensor

    # ---- a7: AttnFPN (5 CNN stages, refine over P2..P4), weights = deterministic fill (not stored)
    cfg5 = dict(BACKBONE, conv_kernels=[[3, 3, 3]] * 5, strides=[[1, 1, 1]] + [[2, 2, 2]] * 4, feature_levels=["P2", "P3", "P4"])
    fpn = det_fill_module(AttnFPN(cfg5).eval())
    x = det_tensor((1, 1, 32, 32, 16), 7, scale=0.5, offset=0.5).requires_grad_(True)
    out = fpn(x)
    g = {k: det_tensor(tuple(v.shape), 11 + i) for i, (k, v) in enumerate(out.items())}
    sum((out[k] * g[k]).sum() for k in out).backward()
    blob = {"grad_x": x.grad.numpy(), "keys": np.array(list(fpn.state_dict().keys()))}
    for k in out:
        blob["out." + k] = out[k].detach().numpy()
    for k, p in fpn.named_parameters():
        if p.numel() <= 2048:
            blob["pg." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "attn_fpn.npz"), **blob)

    # ---- a10 + assembly: TransoarNet on the reference's AMOS table (input level P5 = 8x8x4 needs a 256x256x128 volume)
    cfg = {"backbone": dict(BACKBONE, start_channels=2), "neck": dict(NECK, nheads=3), "bbox_properties": PROPS}
    net = det_fill_module(TransoarNet(cfg).eval())
    x = det_tensor((1, 1, 256, 256, 128), 3, scale=0.5, offset=0.5)
    out = net(x)
    loss = out["pred_logits"].sum() + (out["pred_boxes"] * torch.arange(6.)).sum() + sum(a["pred_boxes"].sum() for a in out["aux_outputs"])
    loss.backward()
    blob = {"pred_logits": out["pred_logits"].detach().numpy(), "pred_boxes": out["pred_boxes"].detach().numpy(),
            "aux0_logits": out["aux_outputs"][0]["pred_logits"].detach().numpy(), "aux0_boxes": out["aux_outputs"][0]["pred_boxes"].detach().numpy(),
            "anchors": net._anchors.numpy(), "restrictions": net._restrictions.numpy(), "keys": np.array(list(net.state_dict().keys()))}
    for k, p in net.named_parameters():
        if p.grad is not None and p.numel() <= 512:
            blob["pg." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "transoarnet.npz"), **blob)
    for f in ("attn_fpn.npz", "transoarnet.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/transoar"):
        sys.exit("reference not mounted")
    main()
