"""CPU checks of the module-level mirrors against fixtures generated from the reference modules
(tests/golden/make_golden_blocks.py): positional encoding values, state_dict key/shape compatibility."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from transoar_b200.position_encoding import PositionEmbeddingSine3D
from transoar_b200.refine import DecoderDefAttnBlock


def test_sine_position_encoding_matches_reference_and_is_cached():
    z = np.load(os.path.join(GOLDEN, "posenc.npz"))
    pe = PositionEmbeddingSine3D(channels=48)
    x = torch.zeros(2, 48, 3, 4, 5)
    got = pe(x)
    assert got.shape == (2, 48, 3, 4, 5)
    assert np.abs(got.numpy() - z["pos_c48"]).max() < 1e-6
    assert pe(x) is got                                            # cached per shape: no recomputation
    pe2 = PositionEmbeddingSine3D(channels=384)
    assert np.abs(pe2(torch.zeros(1, 384, 5, 5, 8)).numpy() - z["pos_c384"]).max() < 1e-6


def test_refine_block_state_dict_is_checkpoint_compatible():
    z = np.load(os.path.join(GOLDEN, "block_defattn.npz"))
    blk = DecoderDefAttnBlock(d_model=48, nhead=6, num_layers=2, dim_feedforward=64, dropout=0.1,
                              feature_levels=["P2", "P3", "P4"], n_points=2)
    ref_keys = sorted(k[3:] for k in z.files if k.startswith("sd."))
    assert sorted(blk.state_dict()) == ref_keys
    sd = {k: torch.from_numpy(z["sd." + k]) for k in ref_keys}
    blk.load_state_dict(sd, strict=True)
    assert all(p.requires_grad for p in blk.parameters())
    # init behaviour (decoder_blocks.py:40-48): MSDeformAttn re-initialises itself after the xavier pass
    fresh = DecoderDefAttnBlock(48, 6, 1, 64, 0.1, ["P2"], 2)
    attn = fresh.refine_def_attn.layers[0].self_attn
    assert float(attn.sampling_offsets.weight.detach().abs().max()) == 0.0
    assert float(attn.attention_weights.weight.detach().abs().max()) == 0.0
    assert float(attn.value_proj.bias.detach().abs().max()) == 0.0


def test_folded_stride2_weights_reproduce_the_input_gradient():
    """transoar_b200.conv3d_gen.fold_stride2_weights (the weight layout of include/conv3d_gen.h::conv3d_gen_dgrad_s2_folded): summing, for every
    parity class of dx, the eight 2x2x2 neighbours of dy against the folded blocks must equal torch's conv3d_input for stride 2."""
    import torch
    from transoar_b200.conv3d_gen import fold_stride2_weights
    co, ci, (N, D, H, W) = 5, 3, (1, 5, 4, 7)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(co, ci, 3, 3, 3, generator=g, dtype=torch.float64)
    od, oh, ow = (D + 1) // 2, (H + 1) // 2, (W + 1) // 2
    dy = torch.randn(N, co, od, oh, ow, generator=g, dtype=torch.float64)
    wf = fold_stride2_weights(w.permute(2, 3, 4, 0, 1).reshape(27, co, ci).contiguous())          # [8, 8 * 32, co]
    assert wf.shape == (8, 8 * 32, co)
    dyp = torch.nn.functional.pad(dy, (0, 1, 0, 1, 0, 1))
    dx = torch.zeros(N, ci, D, H, W, dtype=torch.float64)
    for cls in range(8):
        pd, ph, pw = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
        blocks = wf[:, cls * 32:cls * 32 + ci]                                                  # [8 deltas, ci, co]
        for delta in range(8):
            dd, dh, dw = (delta >> 2) & 1, (delta >> 1) & 1, delta & 1
            contrib = torch.einsum("ic,ncdhw->nidhw", blocks[delta], dyp[:, :, dd:dd + od, dh:dh + oh, dw:dw + ow])
            tgt = dx[:, :, pd::2, ph::2, pw::2]
            tgt += contrib[:, :, :tgt.shape[2], :tgt.shape[3], :tgt.shape[4]]
        assert float(wf[:, cls * 32 + ci:(cls + 1) * 32].abs().max()) == 0.0                       # channel padding is zero
    ref = torch.nn.grad.conv3d_input((N, ci, D, H, W), w, dy, stride=2, padding=1)
    assert float((dx - ref).abs().max()) < 1e-12
