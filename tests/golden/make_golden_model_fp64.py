"""fp64 yardstick for the model-level fixtures (VERDICT r01 "weak" 1 / "next" 5): the REFERENCE's AttnFPN and TransoarNet of
make_golden_model.py run once more in float64 (same deterministic weights: det_fill_module fills fp32 values, ``.double()`` widens
them exactly), so that every tolerance of the model-level parity tests can be read against how far the reference's OWN fp32 CPU
run is from the exact answer.  Stored per tensor: the fp64 result and ``e32.<name>`` = max |fp32 - fp64| / max |fp64| of the
reference's fp32 CPU run (the numbers attn_fpn.npz / transoarnet.npz hold).
Build-container only:   python tests/golden/make_golden_model_fp64.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_model as G  # noqa: E402


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    G._patch_environment()
    torch.set_default_dtype(torch.float64)          # tensors the reference creates inside forward (position encodings, masks) follow
    from transoar.models.backbones.attn_fpn import AttnFPN
    from transoar.models.transoarnet import TransoarNet
    from detfill import det_fill_module, det_tensor

    # ---- AttnFPN fixture in fp64
    z32 = np.load(os.path.join(HERE, "attn_fpn.npz"))
    cfg5 = dict(G.BACKBONE, conv_kernels=[[3, 3, 3]] * 5, strides=[[1, 1, 1]] + [[2, 2, 2]] * 4, feature_levels=["P2", "P3", "P4"])
    torch.set_default_dtype(torch.float32)
    fpn = det_fill_module(AttnFPN(cfg5).eval())
    torch.set_default_dtype(torch.float64)
    fpn = fpn.double()
    x = det_tensor((1, 1, 32, 32, 16), 7, scale=0.5, offset=0.5).double().requires_grad_(True)
    out = fpn(x)
    sum((out[k] * det_tensor(tuple(v.shape), 11 + i).double()).sum() for i, (k, v) in enumerate(out.items())).backward()
    blob = {"grad_x": x.grad.numpy()}
    for k in out:
        assert out[k].dtype == torch.float64
        blob["out." + k] = out[k].detach().numpy()
    for k, p in fpn.named_parameters():
        if "pg." + k in z32.files:
            blob["pg." + k] = p.grad.numpy()
    for k in list(blob):
        blob["e32." + k] = np.float64(_rel(z32[k], blob[k]))
    np.savez_compressed(os.path.join(HERE, "attn_fpn_fp64.npz"), **blob)
    worst = sorted(((float(v), k) for k, v in blob.items() if k.startswith("e32.")), reverse=True)
    print("attn_fpn: reference fp32 CPU vs fp64, worst five:", worst[:5])

    # ---- TransoarNet fixture in fp64
    z32 = np.load(os.path.join(HERE, "transoarnet.npz"))
    cfg = {"backbone": dict(G.BACKBONE, start_channels=2), "neck": dict(G.NECK, nheads=3), "bbox_properties": G.PROPS}
    torch.set_default_dtype(torch.float32)
    net = det_fill_module(TransoarNet(cfg).eval())
    torch.set_default_dtype(torch.float64)
    net = net.double()
    x = det_tensor((1, 1, 256, 256, 128), 3, scale=0.5, offset=0.5).double()
    out = net(x)
    assert out["pred_boxes"].dtype == torch.float64
    loss = out["pred_logits"].sum() + (out["pred_boxes"] * torch.arange(6.)).sum() + sum(a["pred_boxes"].sum() for a in out["aux_outputs"])
    loss.backward()
    blob = {"pred_logits": out["pred_logits"].detach().numpy(), "pred_boxes": out["pred_boxes"].detach().numpy(),
            "aux0_logits": out["aux_outputs"][0]["pred_logits"].detach().numpy(), "aux0_boxes": out["aux_outputs"][0]["pred_boxes"].detach().numpy()}
    for k, p in net.named_parameters():
        if "pg." + k in z32.files:
            blob["pg." + k] = p.grad.numpy()
    for k in list(blob):
        blob["e32." + k] = np.float64(_rel(z32[k], blob[k]))
    np.savez_compressed(os.path.join(HERE, "transoarnet_fp64.npz"), **blob)
    worst = sorted(((float(v), k) for k, v in blob.items() if k.startswith("e32.")), reverse=True)
    print("transoarnet: reference fp32 CPU vs fp64, worst eight:", worst[:8])
    for f in ("attn_fpn_fp64.npz", "transoarnet_fp64.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/transoar"):
        sys.exit("reference not mounted")
    main()
