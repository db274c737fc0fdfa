"""The reference's model configuration for the VISCERAL 160x160x256 data set (config/attn_fpn_foc_dec_visceral.yaml:47-116)
as a python dict, with the two switches the hot path needs flipped on (`use_decoder_attn`, `use_cuda`; both False as
shipped), plus the synthetic atlas of SURVEY 8d (seeded; schema of preprocessor_visceral.py:95-130)."""
import copy

import torch

VISCERAL_BACKBONE = dict(
    name="attn_fpn", use_encoder_attn=False,
    conv_kernels=[[3, 3, 3]] * 6, strides=[[1, 1, 1]] + [[2, 2, 2]] * 5, in_channels=1, start_channels=24,
    depths=[2, 2, 2, 2], num_heads=[3, 6, 12, 24], window_size=[5, 5, 5], mlp_ratio=4, qkv_bias=True, qk_scale=None,
    drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, conv_merging=False,
    use_decoder_attn=True, fpn_channels=384, out_fmaps=["P2"], pos_encoding="sine", feature_levels=["P2", "P3", "P4", "P5"],
    hidden_dim=384, dim_feedforward=1024, dropout=0.1, nheads=6, layers=2, n_points=4, use_cuda=True,
    use_seg_proxy_loss=False, fg_bg=True)

VISCERAL_NECK = dict(
    name="foc_attn", pos_encoding="sine", input_levels="P2", hidden_dim=384, dropout=0.1, nheads=8, dim_feedforward=1024,
    dec_layers=3, restrict_attn=True, obj_self_attn=False, anchor_gen_dynamic_offset=True, anchor_gen_offset=0.1,
    anchor_offset_pred=True, max_anchor_pred_offset=0.1, num_queries=540, num_organs=20, aux_loss=True)


def synthetic_atlas(num_organs=20, seed=0):
    """bbox_properties: per class median / min / max (cx,cy,cz,w,h,d) and attn_area (x1..z2), all in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    props = {}
    for o in range(num_organs):
        c = torch.rand(3, generator=g) * 0.4 + 0.3
        s = torch.rand(3, generator=g) * 0.2 + 0.1
        med = torch.cat((c, s))
        lo = torch.cat((c - 0.05, s - 0.04)).clamp(0.01, 1)
        hi = torch.cat((c + 0.05, s + 0.06)).clamp(0, 1)
        hull = torch.cat(((lo[:3] - hi[3:] / 2).clamp(0, 1), (hi[:3] + hi[3:] / 2).clamp(0, 1)))
        props[str(o + 1)] = {"median": med.tolist(), "min": lo.tolist(), "max": hi.tolist(), "attn_area": hull.tolist()}
    return props


def visceral_config(seed=0):
    return {"backbone": copy.deepcopy(VISCERAL_BACKBONE), "neck": copy.deepcopy(VISCERAL_NECK),
            "bbox_properties": synthetic_atlas(20, seed)}


def amos_config(seed=0, volume=(256, 256, 128)):
    """config/attn_fpn_foc_dec_amos.yaml: as the VISCERAL file except out_fmaps [P3], feature_levels [P3, P4, P5], input_levels P3,
    405 queries for 15 organs, 256x256x128 patches (diff of the two yamls: lines 73, 77, 93, 114-115, 121).  The RoI grid follows
    the P3 feature map of ``volume`` (the reference hard-codes 256x256x128 for 15 organs, focused_decoder.py:108-117)."""
    bb = copy.deepcopy(VISCERAL_BACKBONE)
    bb.update(out_fmaps=["P3"], feature_levels=["P3", "P4", "P5"])
    neck = copy.deepcopy(VISCERAL_NECK)
    neck.update(input_levels="P3", num_queries=405, num_organs=15)
    return {"backbone": bb, "neck": neck, "bbox_properties": synthetic_atlas(15, seed), "neck_input_shape": tuple(v // 8 for v in volume)}


def defdetr_amos_config(seed=0, volume=(256, 256, 128), queries=300, dec_layers=3):
    """BASELINE.json configs[2]: "3D Deformable DETR (attn-fpn-def-detr), 4-level FPN, 300 queries, synthetic AMOS shapes".  The neck is
    not in the reference tree -- ``transoar_b200.def_detr`` restates it (parity unpinned at neck level, op pinned).  Backbone
    keys are the amos yaml's with all four FPN levels refined and returned; 300 queries = 20 per organ for the 15 AMOS organs."""
    bb = copy.deepcopy(VISCERAL_BACKBONE)
    bb.update(out_fmaps=["P2", "P3", "P4", "P5"], feature_levels=["P2", "P3", "P4", "P5"])
    neck = dict(name="def_detr", hidden_dim=384, dropout=0.1, nheads=8, dim_feedforward=1024, dec_layers=dec_layers, n_points=4,
                num_queries=queries, num_organs=15, aux_loss=True)
    return {"model_family": "def_detr", "backbone": bb, "neck": neck, "bbox_properties": synthetic_atlas(15, seed), "volume": tuple(volume)}


def swin_focused_config(seed=0, volume=(192, 192, 384)):
    """BASELINE.json configs[3]: "SwinFPN backbone (use_encoder_attn=True) + Focused Decoder, VISCERAL-shape 192x192x384".  The visceral yaml
    with the Swin encoder switched on (encoder_blocks.py:56-334); 192x192x384 has no row in the reference's RoI shape table
    (focused_decoder.py:99-117), so the mask grid is derived from the P2 feature map (48x48x96)."""
    cfg = visceral_config(seed)
    cfg["backbone"]["use_encoder_attn"] = True
    cfg["neck_input_shape"] = tuple(v // 4 for v in volume)
    cfg["volume"] = tuple(volume)
    return cfg
