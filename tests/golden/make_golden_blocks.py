"""Golden fixtures for the module-level rows, generated from the REFERENCE modules on CPU
(``use_cuda=False`` route).  Build-container only:   python tests/golden/make_golden_blocks.py

Stores, per case: the reference module's state_dict, seeded inputs, eval-mode outputs and the gradients of
sum(out * g) w.r.t. inputs and parameters."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from transoar.models.backbones.decoder_blocks import DecoderDefAttnBlock
    from transoar.models.ops.modules import MSDeformAttn
    from transoar.models.position_encoding import PositionEmbeddingSine3D

    # ---- a10: sine positional encoding
    blob = {}
    for name, (c, shape) in {"c48": (48, (2, 3, 4, 5)), "c384": (384, (1, 5, 5, 8))}.items():
        pe = PositionEmbeddingSine3D(channels=c)
        blob[f"pos_{name}"] = pe(torch.zeros(shape[0], c, *shape[1:])).numpy()
    np.savez_compressed(os.path.join(HERE, "posenc.npz"), **blob)

    # ---- a5: MSDeformAttn module
    torch.manual_seed(11)
    shapes = [(4, 3, 5), (2, 2, 3)]
    S = sum(d * h * w for d, h, w in shapes)
    m = MSDeformAttn(d_model=48, n_levels=2, n_heads=6, n_points=2, use_cuda=False).eval()
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.attention_weights.weight.normal_(0, 0.3)
    ss = torch.as_tensor(shapes, dtype=torch.long)
    starts = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    query = torch.randn(2, 7, 48, requires_grad=True)
    src = torch.randn(2, S, 48, requires_grad=True)
    ref = torch.rand(2, 7, 2, 3)
    g = torch.randn(2, 7, 48)
    out = m(query, ref, src, ss, starts)
    out.backward(g)
    blob = {"shapes": ss.numpy(), "starts": starts.numpy(), "query": query.detach().numpy(), "src": src.detach().numpy(),
            "ref": ref.numpy(), "g": g.numpy(), "out": out.detach().numpy(), "grad_query": query.grad.numpy(),
            "grad_src": src.grad.numpy()}
    for k, v in m.state_dict().items():
        blob["sd." + k] = v.numpy()
    for k, p in m.named_parameters():
        blob["pg." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "module_msdeformattn.npz"), **blob)

    # ---- a6: DecoderDefAttnBlock (2 layers, 3 levels) with sine pos-enc
    torch.manual_seed(12)
    blk = DecoderDefAttnBlock(d_model=48, nhead=6, num_layers=2, dim_feedforward=64, dropout=0.1,
                              feature_levels=["P2", "P3", "P4"], n_points=2, use_cuda=False).eval()
    with torch.no_grad():
        for layer in blk.refine_def_attn.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.self_attn.attention_weights.weight.normal_(0, 0.3)
    lv = [(4, 4, 6), (2, 2, 3), (1, 1, 2)]
    fmaps = [torch.randn(2, 48, *s, requires_grad=True) for s in lv]
    pe = PositionEmbeddingSine3D(channels=48)
    pos = [pe(f) for f in fmaps]
    outs = blk(fmaps, pos)
    gs = [torch.randn_like(o) for o in outs]
    sum((o * g_).sum() for o, g_ in zip(outs, gs)).backward()
    blob = {}
    for i in range(3):
        blob[f"fmap{i}"] = fmaps[i].detach().numpy(); blob[f"g{i}"] = gs[i].numpy()
        blob[f"out{i}"] = outs[i].detach().numpy(); blob[f"grad_fmap{i}"] = fmaps[i].grad.numpy()
    for k, v in blk.state_dict().items():
        blob["sd." + k] = v.numpy()
    for k, p in blk.named_parameters():
        blob["pg." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "block_defattn.npz"), **blob)
    for f in ("posenc.npz", "module_msdeformattn.npz", "block_defattn.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/transoar"):
        sys.exit("reference not mounted")
    main()
