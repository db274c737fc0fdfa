"""Swin encoder stages (transoar_b200/swin.py) against fixtures produced by the reference's EncoderSwinBlock /
SwinBlock / WindowAttention3D / PatchMerging (tests/golden/make_golden_swin.py).  The mirror is device-agnostic torch code (its
Linear layers take the tcgen05 GEMM only on CUDA tensors with TF32 requested), so parity is checked on the CPU here."""
import os
import sys

import numpy as np
import torch

from conftest import GOLDEN
sys.path.insert(0, GOLDEN)
import make_golden_swin as G          # noqa: E402  (constants only; the reference is not imported)
from transoar_b200 import swin


def _rel(a, b):
    b = torch.as_tensor(b)
    return float((a.detach() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_shift_masks_equal_the_reference():
    z = np.load(os.path.join(GOLDEN, "swin.npz"))
    m = swin.shift_mask((10, 20, 10), (5, 5, 5), (2, 2, 2), torch.device("cpu"))
    assert np.array_equal(m.numpy(), z["mask_10_20_10"])
    m2 = swin.shift_mask((4, 10, 5), (4, 5, 5), (0, 2, 0), torch.device("cpu"))
    assert np.array_equal(m2.numpy(), z["mask_4_8_5"])
    assert swin.effective_window((4, 10, 5), (5, 5, 5), (2, 2, 2)) == ((4, 5, 5), (0, 2, 0))


def test_three_stages_match_the_reference_forward_and_backward():
    z = np.load(os.path.join(GOLDEN, "swin.npz"))
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    cur, feats, mods = x, [], []
    for i, st in enumerate(G.STAGES):
        m = swin.EncoderSwinBlock(dim=st["dim"], depth=st["depth"], num_heads=st["heads"], window_size=G.WINDOW, mlp_ratio=4, qkv_bias=True,
                                  qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=[0.0, 0.1], downsample=swin.PatchMerging).eval()
        ref_keys = sorted(k[len(f"sd{i}."):] for k in z.files if k.startswith(f"sd{i}."))
        assert sorted(m.state_dict()) == ref_keys                                              # checkpoint-compatible names, incl. the index buffer
        m.load_state_dict({k: torch.from_numpy(z[f"sd{i}.{k}"]) for k in ref_keys}, strict=True)
        cur = m(cur)
        feats.append(cur)
        mods.append(m)
    from detfill import det_tensor
    sum((f * det_tensor(tuple(f.shape), 90 + i)).sum() for i, f in enumerate(feats)).backward()
    for i, f in enumerate(feats):
        assert f.shape == z[f"out{i}"].shape and _rel(f, z[f"out{i}"]) < 2e-5, i
    assert _rel(x.grad, z["grad_x"]) < 1e-4
    for i, m in enumerate(mods):
        for k, p in m.named_parameters():
            assert _rel(p.grad, z[f"pg{i}.{k}"]) < 2e-4, (i, k)


def test_relative_position_index_and_drop_path():
    a = swin.WindowAttention3D(12, (5, 5, 5), 3, True, None, 0.0, 0.0)
    idx = a.relative_position_index
    assert idx.shape == (125, 125) and int(idx.min()) == 0 and int(idx.max()) == 9 ** 3 - 1
    assert int(idx[0, 0]) == (4 * 9 + 4) * 9 + 4 and torch.equal(idx.diagonal(), torch.full((125,), 364))
    dp = swin.DropPath(0.5).train()
    torch.manual_seed(0)
    y = dp(torch.ones(64, 3, 2))
    kept = (y[:, 0, 0] != 0)
    assert 10 < int(kept.sum()) < 54 and torch.allclose(y[kept], torch.full_like(y[kept], 2.0))
    assert torch.equal(swin.DropPath(0.5).eval()(torch.ones(4, 2)), torch.ones(4, 2))


def test_encoder_builds_swin_stages_like_the_reference():
    from transoar_b200.attn_fpn import Encoder
    from transoar_b200.configs import VISCERAL_BACKBONE
    enc = Encoder(dict(VISCERAL_BACKBONE, use_encoder_attn=True))
    kinds = [type(s).__name__ for s in enc._stages]
    assert kinds == ["EncoderCnnBlock"] * 2 + ["EncoderSwinBlock"] * 4
    assert [s.blocks[0].dim for s in enc._stages[2:]] == [48, 96, 192, 384] and [s.blocks[0].num_heads for s in enc._stages[2:]] == [3, 6, 12, 24]
    assert sum(p.numel() for p in enc.parameters()) > 5e6
