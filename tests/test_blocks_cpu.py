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
