"""Golden fixtures for SURVEY 8 row a8 from the REFERENCE's Swin encoder stages (EncoderSwinBlock / SwinBlock / WindowAttention3D /
PatchMerging, encoder_blocks.py:56-400): CPU, eval mode (drop-path inactive), deterministic weights.
Build-container only:   python tests/golden/make_golden_swin.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# stage plan of the reference Encoder with use_encoder_attn (attn_fpn.py:172-190): dims / heads double per stage
STAGES = [dict(dim=12, heads=3, depth=2), dict(dim=24, heads=6, depth=2), dict(dim=48, heads=12, depth=2)]
SHAPE = (2, 12, 8, 16, 6)        # B, C, D, H, W: D pads 8 -> 10, H 16 -> 20, W 6 -> 10; later stages fall below the window (no shift there)
WINDOW = (5, 5, 5)


def main():
    from make_golden_model import _patch_environment
    _patch_environment()
    from transoar.models.backbones.encoder_blocks import EncoderSwinBlock, PatchMerging, compute_mask
    from detfill import det_fill_module, det_tensor
    blob = {}
    x = det_tensor(SHAPE, 5, scale=0.8).requires_grad_(True)
    feats, cur = [], x
    mods = []
    for i, st in enumerate(STAGES):
        m = det_fill_module(EncoderSwinBlock(dim=st["dim"], depth=st["depth"], num_heads=st["heads"], window_size=WINDOW, mlp_ratio=4,
                                             qkv_bias=True, qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=[0.0, 0.1],
                                             downsample=PatchMerging).eval())
        with torch.no_grad():      # a relative-position bias large enough to matter
            for j, b in enumerate(m.blocks):
                b.attn.relative_position_bias_table.copy_(det_tensor(tuple(b.attn.relative_position_bias_table.shape), 70 + 2 * i + j, scale=0.5))
        cur = m(cur)
        feats.append(cur)
        mods.append(m)
        for k, v in m.state_dict().items():
            blob[f"sd{i}.{k}"] = v.numpy()
    loss = sum((f * det_tensor(tuple(f.shape), 90 + i)).sum() for i, f in enumerate(feats))
    loss.backward()
    blob["x"] = x.detach().numpy()
    blob["grad_x"] = x.grad.numpy()
    for i, f in enumerate(feats):
        blob[f"out{i}"] = f.detach().numpy()
    for i, m in enumerate(mods):
        for k, p in m.named_parameters():
            blob[f"pg{i}.{k}"] = p.grad.numpy()
    blob["mask_10_20_10"] = compute_mask(10, 20, 10, (5, 5, 5), (2, 2, 2), torch.device("cpu")).numpy()
    blob["mask_4_8_5"] = compute_mask(4, 10, 5, (4, 5, 5), (0, 2, 0), torch.device("cpu")).numpy()
    np.savez_compressed(os.path.join(HERE, "swin.npz"), **blob)
    print("wrote swin.npz:", len(blob), "arrays,", [tuple(f.shape) for f in feats])


if __name__ == "__main__":
    main()
