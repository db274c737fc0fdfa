"""Golden fixtures for SURVEY 8 row a9 from the REFERENCE's FocusedAttn / FocusedDecoderLayer (CPU, eval mode).
Build-container only:   python tests/golden/make_golden_focused.py
timm (not installed) is only used for trunc_normal_ -> shimmed; the reference's unconditional .cuda() calls are patched to no-ops."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))



def _patch_environment():
    """Only inside the generator process: reference on sys.path, timm shim, .cuda() -> no-op."""
    sys.path.insert(0, "/root/reference")
    timm = types.ModuleType("timm"); timm.models = types.ModuleType("timm.models"); timm.models.layers = types.ModuleType("timm.models.layers")
    timm.models.layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.layers": timm.models.layers})
    torch.Tensor.cuda = lambda self, *a, **k: self


def main():
    _patch_environment()
    from transoar.models.necks.focused_decoder import FocusedAttn, FocusedDecoderLayer

    # ---- FocusedAttn on its own: 2 organs x 7 queries, grid (4,5,6), one box touching the border, one interior
    torch.manual_seed(21)
    X, Y, Z = 4, 5, 6
    boxes = torch.tensor([[0, 0, 0, 2, 3, 6]] * 7 + [[1, 2, 1, 4, 5, 4]] * 7)
    mask = torch.ones(14, X, Y, Z, dtype=torch.bool)
    for qi, (x1, y1, z1, x2, y2, z2) in enumerate(boxes.tolist()):
        mask[qi, x1:x2, y1:y2, z1:z2] = False
    mask = mask.flatten(1)
    m = FocusedAttn(96, 2, mask, proj_drop=0.1).eval()
    q = torch.randn(2, 14, 96, requires_grad=True)
    k = torch.randn(2, X * Y * Z, 96, requires_grad=True)
    v = torch.randn(2, X * Y * Z, 96, requires_grad=True)
    g = torch.randn(2, 14, 96)
    x, w = m(q, k, v, mask=mask.float())
    x.backward(g)
    blob = dict(boxes=boxes.numpy().astype(np.int32), mask=mask.numpy(), q=q.detach().numpy(), k=k.detach().numpy(), v=v.detach().numpy(),
                g=g.numpy(), x=x.detach().numpy(), weights=w.detach().numpy(), grad_q=q.grad.numpy(), grad_k=k.grad.numpy(),
                grad_v=v.grad.numpy(), grid=np.array([X, Y, Z]))
    for name, t in m.state_dict().items():
        blob["sd." + name] = t.numpy()
    for name, p in m.named_parameters():
        blob["pg." + name] = np.zeros(p.shape, np.float32) if p.grad is None else p.grad.numpy()
        blob["pg_none." + name] = np.array(p.grad is None)
    np.savez_compressed(os.path.join(HERE, "focused_attn.npz"), **blob)

    # ---- a whole FocusedDecoderLayer: AMOS table, input level P5 -> (8, 8, 4) grid, 2 organs x 7 queries
    torch.manual_seed(22)
    cfg = {"num_queries": 14, "num_organs": 2, "input_levels": "P5", "restrict_attn": True}
    props = {"1": {"attn_area": [0.05, 0.10, 0.0, 0.55, 0.62, 0.8]}, "2": {"attn_area": [0.4, 0.3, 0.2, 1.0, 0.9, 1.0]}}
    layer = FocusedDecoderLayer(d_model=96, d_ffn=64, dropout=0.1, activation="relu", n_heads=2, config=cfg, bbox_props=props).eval()
    for p in layer.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
    n_tok = 8 * 8 * 4
    tgt = torch.randn(2, 14, 96, requires_grad=True)
    qpos = torch.randn(2, 14, 96)
    src = torch.randn(2, n_tok, 96, requires_grad=True)
    spos = torch.randn(2, n_tok, 96)
    g = torch.randn(2, 14, 96)
    out, _ = layer(tgt, qpos, spos, src)
    out.backward(g)
    blob = dict(tgt=tgt.detach().numpy(), qpos=qpos.numpy(), src=src.detach().numpy(), spos=spos.numpy(), g=g.numpy(),
                out=out.detach().numpy(), grad_tgt=tgt.grad.numpy(), grad_src=src.grad.numpy(),
                attn_mask=layer.attn_mask.numpy(), props=np.array([props["1"]["attn_area"], props["2"]["attn_area"]]))
    for name, t in layer.state_dict().items():
        blob["sd." + name] = t.numpy()
    for name, p in layer.named_parameters():
        blob["pg." + name] = np.zeros(p.shape, np.float32) if p.grad is None else p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "focused_layer.npz"), **blob)
    for f in ("focused_attn.npz", "focused_layer.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/transoar"):
        sys.exit("reference not mounted")
    main()
