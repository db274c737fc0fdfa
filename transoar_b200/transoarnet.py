"""``TransoarNet`` -- mirror of transoar/models/transoarnet.py:11-171 (model assembly: AttnFPN backbone -> Focused Decoder
-> class / box heads, anchors and offset restrictions from the atlas).  Same module names (``_backbone``, ``_neck``,
``_cls_head``, ``_reg_head``, ``_query_embed``, ``_seg_head``) and the same output dict.  Anchors / restrictions are
registered as (non-persistent) buffers instead of being ``.cuda()``-ed in the constructor, so the model can be built on
any device and moved with ``.to()``; the RoI grid is derived from the feature map."""

import torch
import torch.nn.functional as F
from torch import nn

from .attn_fpn import AttnFPN
from .focused import FocusedDecoder
from .position_encoding import PositionEmbeddingSine3D


class MLP(nn.Module):
    """transoarnet.py:157-171."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x) if i == self.num_layers - 1 else F.relu(layer(x))
        return x


def generate_anchors(neck_cfg, bbox_props):
    """Anchors [Nq,6] (cx,cy,cz,w,h,d) and per-query offset restrictions [Nq,6], restating transoarnet.py:60-116."""
    nq, n_org = neck_cfg["num_queries"], neck_cfg["num_organs"]
    per = int(nq / n_org)
    dyn = neck_cfg["anchor_gen_dynamic_offset"]
    anchors, restr_pos = [], []
    for props in bbox_props.values():
        size = torch.tensor(props["median"])[3:]
        vol = torch.tensor(props["attn_area"])
        centre, whd = (vol[:3] + vol[3:]) / 2, vol[3:] - vol[:3]
        if dyn:
            step = ((whd - size) / 3)[None]
            cand = torch.cat((step, -step, torch.zeros_like(step)), dim=0)                       # [3,3]: +d, -d, 0 per axis
            grid = torch.cartesian_prod(*cand.unbind(dim=-1))
        else:
            o = neck_cfg["anchor_gen_offset"]
            cand = torch.tensor([0, o, -o])
            grid = torch.cartesian_prod(cand, cand, cand)
        if per == 1:
            offs = torch.zeros(3)[None]
        elif per == 7:
            offs = grid[torch.count_nonzero(grid, dim=-1) <= 1]
        else:
            offs = grid
        a = torch.cat((offs, size[None].repeat(offs.shape[0], 1)), dim=-1)
        a[:, :3] += centre
        anchors.append(a)
        restr_pos.append(offs.max(dim=0)[0][None])
    med = torch.tensor([v["median"] for v in bbox_props.values()])[:, 3:]
    lo = med - torch.tensor([v["min"] for v in bbox_props.values()])[:, 3:]
    hi = torch.tensor([v["max"] for v in bbox_props.values()])[:, 3:] - med
    restr = torch.repeat_interleave(torch.cat((torch.cat(restr_pos), torch.max(lo, hi)), dim=-1), per, dim=0)
    return torch.cat(anchors).clamp(min=0, max=1), restr


class TransoarNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        neck, bb = config["neck"], config["backbone"]
        hidden = neck["hidden_dim"]
        self._input_levels = neck["input_levels"]
        self._anchor_offset = neck["anchor_offset_pred"]
        self._aux_loss = neck["aux_loss"]
        self._backbone = AttnFPN(bb)
        anchors, restrictions = generate_anchors(neck, config["bbox_properties"])
        if not neck["anchor_gen_dynamic_offset"]:
            restrictions = torch.full_like(restrictions, float(neck["max_anchor_pred_offset"]))
        restrictions[:, :3] /= 2                                                               # transoarnet.py:29
        self.register_buffer("_anchors", anchors, persistent=False)
        self.register_buffer("_restrictions", restrictions, persistent=False)
        self._neck = FocusedDecoder(d_model=hidden, nhead=neck["nheads"], num_decoder_layers=neck["dec_layers"],
                                    dim_feedforward=neck["dim_feedforward"], dropout=neck["dropout"], activation="relu",
                                    return_intermediate_dec=True, bbox_props=config["bbox_properties"], config=neck,
                                    input_shape=config.get("neck_input_shape"))
        self._cls_head = nn.Linear(hidden, 1)
        self._reg_head = MLP(hidden, hidden, 6, 3)
        self._seg_proxy = bb["use_seg_proxy_loss"]
        if self._seg_proxy:
            self._seg_head = nn.Conv3d(bb["start_channels"], 2 if bb["fg_bg"] else neck["num_organs"] + 1, kernel_size=1, stride=1)
        self._query_embed = nn.Embedding(neck["num_queries"], hidden * 2)
        if neck["pos_encoding"] != "sine":
            raise ValueError("Please select a implemented pos. encoding.")
        self._pos_enc = PositionEmbeddingSine3D(channels=hidden)
        if self._anchor_offset:                                                                # transoarnet.py:50-58
            for t in (self._cls_head.weight, self._cls_head.bias, self._reg_head.layers[-1].weight, self._reg_head.layers[-1].bias):
                nn.init.constant_(t.data, 0)

    def forward(self, x):
        feats = self._backbone(x)
        det_src = feats[self._input_levels]
        hs = self._neck(det_src, self._query_embed.weight, self._pos_enc(det_src))             # [layers, B, Nq, hidden]
        logits, boxes = self._cls_head(hs), self._reg_head(hs)
        if self._anchor_offset:
            boxes = torch.clamp(boxes.tanh() * self._restrictions + self._anchors, min=0, max=1)
        else:
            boxes = boxes.sigmoid()
        out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
               "pred_seg": self._seg_head(feats["P0"]) if self._seg_proxy else 0}
        if self._aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(logits[:-1], boxes[:-1])]
        return out
