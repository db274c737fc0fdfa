"""CPU checks for SURVEY 8 row a9 (Focused Decoder): box / group construction against the reference's masks, the dense
oracle against the reference FocusedAttn fixture, checkpoint compatibility of the mirrors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.focused_attn_oracle import dense_masked_attention
from transoar_b200 import focused


def _z(name):
    return np.load(os.path.join(GOLDEN, name))


def test_boxes_from_bbox_props_reproduce_the_reference_mask():
    z = _z("focused_layer.npz")
    props = {str(i): {"attn_area": z["props"][i].tolist()} for i in range(2)}
    boxes = focused.boxes_from_bbox_props(props, 14, (8, 8, 4))
    mask = torch.ones(14, 8, 8, 4, dtype=torch.bool)
    for q, (x1, y1, z1, x2, y2, z2) in enumerate(boxes.tolist()):
        mask[q, x1:x2, y1:y2, z1:z2] = False
    assert np.array_equal(mask.flatten(1).numpy(), z["attn_mask"])                 # generate_attn_masks, focused_decoder.py:138-159
    full = focused.boxes_from_bbox_props(props, 14, (8, 8, 4), restrict_attn=False)
    assert full.tolist() == [[0, 0, 0, 8, 8, 4]] * 14                              # :159


def test_boxes_round_trip_through_the_reference_mask_format():
    z = _z("focused_attn.npz")
    boxes = focused.boxes_from_mask(torch.from_numpy(z["mask"]), z["grid"])
    assert np.array_equal(boxes.numpy(), z["boxes"])
    bad = torch.from_numpy(z["mask"]).clone()
    bad[0, 0] = True                                                               # knock a corner out of query 0's box
    with pytest.raises(ValueError, match="not an axis-aligned box"):
        focused.boxes_from_mask(bad, z["grid"])


def test_groups_cover_every_query_once_in_chunks_of_32():
    boxes = torch.tensor([[0, 0, 0, 2, 2, 2]] * 27 + [[1, 1, 1, 3, 4, 5]] * 54 + [[0, 0, 0, 0, 0, 0]], dtype=torch.int32)
    g = focused.groups_from_boxes(boxes)
    assert g[:, 1].sum() == 82 and int(g[:, 1].max()) <= 32
    assert g[:, 0].tolist() == [0, 27, 59, 81] and g[:, 1].tolist() == [27, 32, 22, 1]
    for q0, nq, *box in g.tolist():
        assert all(boxes[q].tolist() == box for q in range(q0, q0 + nq))


def test_dense_oracle_matches_reference_focused_attn():
    """Pins oracle/focused_attn_oracle.py: same projections (incl. k_proj on the query) + dense core == reference module output."""
    z = _z("focused_attn.npz")
    t = lambda k: torch.from_numpy(z[k])
    W = {k[3:]: t(k) for k in z.files if k.startswith("sd.")}
    B, Nq, C, H = 2, 14, 96, 2
    lin = lambda x, n: x @ W[n + ".weight"].T + (W[n + ".bias"] if n + ".bias" in W else 0)
    qp = (lin(t("q"), "k_proj") * (C // H) ** -0.5).reshape(B, Nq, H, C // H)
    kp = lin(t("k"), "k_proj").reshape(B, -1, H, C // H)
    vp = lin(t("v"), "v_proj").reshape(B, -1, H, C // H)
    x = lin(dense_masked_attention(qp, kp, vp, t("boxes"), tuple(z["grid"])), "proj")
    assert torch.allclose(x, t("x"), rtol=1e-5, atol=1e-6)


def test_mirrors_are_checkpoint_compatible_and_keep_the_dead_q_proj():
    z = _z("focused_layer.npz")
    cfg = {"num_queries": 14, "num_organs": 2, "input_levels": "P5", "restrict_attn": True}
    props = {str(i): {"attn_area": z["props"][i].tolist()} for i in range(2)}
    layer = focused.FocusedDecoderLayer(96, 64, 0.1, "relu", 2, cfg, props)
    assert layer.input_shape == (8, 8, 4)                                          # AMOS table / 2^5, focused_decoder.py:108-117
    ref_keys = sorted(k[3:] for k in z.files if k.startswith("sd."))
    assert sorted(layer.state_dict()) == ref_keys
    assert "cross_attn.q_proj.weight" in ref_keys
    layer.load_state_dict({k: torch.from_numpy(z["sd." + k]) for k in ref_keys}, strict=True)
    dec = focused.FocusedDecoder(96, 2, 3, 64, 0.1, "relu", True, props, cfg)
    assert len(dec.decoder.layers) == 3 and dec.decoder.return_intermediate
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        layer(torch.zeros(1, 14, 96), None, None, torch.zeros(1, 256, 96))


def test_key_value_tokens_must_match_the_grid_the_boxes_were_built_for():
    """a feature map other than the one the RoI boxes were built for would index past the key / value tokens.  The reference
    fails at ``attn += mask`` (focused_decoder.py:243-245: [Nq, X*Y*Z] does not broadcast against [B, H, Nq, Nkv]); the mirror raises before
    any device work."""
    boxes = torch.tensor([[0, 0, 0, 2, 2, 2], [1, 1, 1, 3, 3, 3]], dtype=torch.int32)
    attn = focused.FocusedAttn(32, 2, boxes, grid_shape=(3, 3, 3))
    q = torch.zeros(1, 2, 32)
    for nkv in (26, 28, 54):
        with pytest.raises(RuntimeError, match=r"RoI boxes were built for a \(3, 3, 3\) feature map"):
            attn(q, torch.zeros(1, nkv, 32), torch.zeros(1, nkv, 32))
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):         # the right token count reaches the native op
        attn(q, torch.zeros(1, 27, 32), torch.zeros(1, 27, 32))
