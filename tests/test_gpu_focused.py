"""GPU parity for SURVEY 8 row a9: the fused RoI attention (include/roi_attn.h) against the dense oracle and against
fixtures from the reference's FocusedAttn / FocusedDecoderLayer."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.focused_attn_oracle import dense_masked_attention
from transoar_b200 import focused

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _random_boxes(n_groups, per, grid, gen):
    X, Y, Z = grid
    out = []
    for _ in range(n_groups):
        lo = [int(torch.randint(0, s, (1,), generator=gen)) for s in grid]
        hi = [int(torch.randint(l + 1, s + 1, (1,), generator=gen)) for l, s in zip(lo, grid)]
        out += [lo + hi] * per
    return torch.tensor(out, dtype=torch.int32)


@pytest.mark.parametrize("hd,H,per,grid", [(48, 8, 27, (6, 7, 9)), (16, 2, 1, (3, 3, 3)), (32, 3, 7, (5, 4, 8)), (64, 2, 54, (4, 6, 5)),
                                           (96, 1, 27, (3, 5, 7)), (128, 1, 5, (2, 3, 9))])
def test_roi_attention_matches_dense_oracle(hd, H, per, grid):
    gen = torch.Generator().manual_seed(hd + per)
    boxes = _random_boxes(3, per, grid, gen)
    boxes[0:per] = torch.tensor([0, 0, 0, *grid], dtype=torch.int32)                 # full grid (restrict_attn=False case)
    boxes[per:2 * per] = torch.tensor([1, 1, 1, 2, 2, 2], dtype=torch.int32)        # a single voxel
    Nq, Nkv, B = boxes.shape[0], grid[0] * grid[1] * grid[2], 2
    mk = lambda *s: torch.randn(*s, generator=gen).to(DEV).requires_grad_(True)
    q, k, v = mk(B, Nq, H, hd), mk(B, Nkv, H, hd), mk(B, Nkv, H, hd)
    g = torch.randn(B, Nq, H * hd, generator=gen).to(DEV)
    groups = focused.groups_from_boxes(boxes).to(DEV)
    out = focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:])
    out.backward(g)
    got = [out.detach(), q.grad.clone(), k.grad.clone(), v.grad.clone()]
    q.grad = k.grad = v.grad = None
    want = dense_masked_attention(q, k, v, boxes, grid)
    want.backward(g)
    for a, b, tol in zip(got, [want.detach(), q.grad, k.grad, v.grad], (2e-5, 1e-4, 1e-4, 1e-4)):
        assert _rel(a, b) < tol


def test_empty_box_gives_nan_rows_like_softmax_of_all_minus_inf():
    grid = (3, 3, 3)
    boxes = torch.tensor([[0, 0, 0, 2, 2, 2], [1, 1, 1, 1, 2, 2]], dtype=torch.int32)  # second box is empty (x1 == x2)
    q, k, v = (torch.randn(1, n, 2, 16, device=DEV) for n in (2, 27, 27))
    out = focused.RoIAttentionFunction.apply(q, k, v, focused.groups_from_boxes(boxes).to(DEV), grid[1:])
    want = dense_masked_attention(q, k, v, boxes, grid)
    assert torch.isnan(out[0, 1]).all() and torch.isnan(want[0, 1]).all()
    assert _rel(out[0, 0], want[0, 0]) < 2e-5


def test_focused_attn_module_against_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "focused_attn.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(DEV)
    m = focused.FocusedAttn(96, 2, torch.from_numpy(z["mask"]), proj_drop=0.1, grid_shape=z["grid"]).to(DEV).eval()
    m.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")})
    q, k, v = (t(n).requires_grad_(True) for n in ("q", "k", "v"))
    x, w = m(q, k, v, mask=None)
    assert w is None                                                                 # dense weights only on request
    x.backward(t("g"))
    assert _rel(x.detach(), t("x")) < 1e-4
    for got, name in ((q.grad, "grad_q"), (k.grad, "grad_k"), (v.grad, "grad_v")):
        assert _rel(got, t(name)) < 2e-4, name
    for name, p in m.named_parameters():
        if bool(z["pg_none." + name]):
            assert p.grad is None and name.startswith("q_proj")                      # q_proj never gets a gradient
        else:
            assert _rel(p.grad, t("pg." + name)) < 2e-4, name
    m.materialize_weights = True
    _, w = m(q.detach(), k.detach(), v.detach())
    assert _rel(w, t("weights")) < 1e-4


def test_focused_decoder_layer_against_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "focused_layer.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(DEV)
    cfg = {"num_queries": 14, "num_organs": 2, "input_levels": "P5", "restrict_attn": True}
    props = {str(i): {"attn_area": z["props"][i].tolist()} for i in range(2)}
    layer = focused.FocusedDecoderLayer(96, 64, 0.1, "relu", 2, cfg, props).to(DEV).eval()
    layer.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")})
    tgt, src = t("tgt").requires_grad_(True), t("src").requires_grad_(True)
    out, _ = layer(tgt, t("qpos"), t("spos"), src)
    out.backward(t("g"))
    assert _rel(out.detach(), t("out")) < 1e-4
    assert _rel(tgt.grad, t("grad_tgt")) < 2e-4 and _rel(src.grad, t("grad_src")) < 2e-4
    for name, p in layer.named_parameters():
        if p.grad is not None:
            assert _rel(p.grad, t("pg." + name)) < 5e-4, name


def test_visceral_shape_against_dense_path():
    """540 queries = 20 organs x 27, P2 grid 40x40x64 = 102 400 tokens, 8 heads x 48 -- the dense path needs a 1.77 GB score tensor."""
    gen = torch.Generator().manual_seed(5)
    grid = (40, 40, 64)
    props = {}
    for o in range(20):
        c = torch.rand(3, generator=gen) * 0.4 + 0.3
        s = torch.rand(3, generator=gen) * 0.2 + 0.1
        props[str(o)] = {"attn_area": torch.cat(((c - s / 2 - 0.05).clamp(0, 1), (c + s / 2 + 0.05).clamp(0, 1))).tolist()}
    boxes = focused.boxes_from_bbox_props(props, 540, grid)
    q = torch.randn(1, 540, 8, 48, generator=gen).to(DEV) * 0.3
    k = torch.randn(1, 102400, 8, 48, generator=gen).to(DEV)
    v = torch.randn(1, 102400, 8, 48, generator=gen).to(DEV)
    out = focused.RoIAttentionFunction.apply(q, k, v, focused.groups_from_boxes(boxes).to(DEV), grid[1:])
    want = dense_masked_attention(q, k, v, boxes, grid)
    assert _rel(out, want) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# Backbone (a7) and whole-model assembly (a10) against fixtures from the reference modules (deterministic weights)
# ---------------------------------------------------------------------------------------------------------------
def _golden_cfgs():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_golden_model as G
    from detfill import det_fill_module, det_tensor
    return G, det_fill_module, det_tensor


def test_attn_fpn_against_reference_fixture():
    from transoar_b200.attn_fpn import AttnFPN
    G, det_fill_module, det_tensor = _golden_cfgs()
    z = np.load(os.path.join(GOLDEN, "attn_fpn.npz"))
    cfg5 = dict(G.BACKBONE, conv_kernels=[[3, 3, 3]] * 5, strides=[[1, 1, 1]] + [[2, 2, 2]] * 4, feature_levels=["P2", "P3", "P4"], use_cuda=True)
    fpn = det_fill_module(AttnFPN(cfg5).eval()).to(DEV)
    x = det_tensor((1, 1, 32, 32, 16), 7, scale=0.5, offset=0.5).to(DEV).requires_grad_(True)
    with torch.backends.cudnn.flags(allow_tf32=False):
        out = fpn(x)
        sum((out[k] * det_tensor(tuple(v.shape), 11 + i).to(DEV)).sum() for i, (k, v) in enumerate(out.items())).backward()
    assert sorted(out) == ["P2", "P3", "P4"]
    for k in out:
        assert _rel(out[k].detach(), torch.from_numpy(z["out." + k]).to(DEV)) < 2e-4, k
    assert _rel(x.grad, torch.from_numpy(z["grad_x"]).to(DEV)) < 1e-3
    for k, p in fpn.named_parameters():
        if "pg." + k in z.files:
            assert _rel(p.grad, torch.from_numpy(z["pg." + k]).to(DEV)) < 2e-3, k


def test_transoarnet_against_reference_fixture():
    from transoar_b200.transoarnet import TransoarNet
    G, det_fill_module, det_tensor = _golden_cfgs()
    z = np.load(os.path.join(GOLDEN, "transoarnet.npz"))
    cfg = {"backbone": dict(G.BACKBONE, start_channels=2, use_cuda=True), "neck": dict(G.NECK, nheads=3), "bbox_properties": G.PROPS}
    net = det_fill_module(TransoarNet(cfg).eval()).to(DEV)
    x = det_tensor((1, 1, 256, 256, 128), 3, scale=0.5, offset=0.5).to(DEV)
    with torch.backends.cudnn.flags(allow_tf32=False):
        out = net(x)
        loss = out["pred_logits"].sum() + (out["pred_boxes"] * torch.arange(6., device=DEV)).sum() + sum(a["pred_boxes"].sum() for a in out["aux_outputs"])
        loss.backward()
    t = lambda k: torch.from_numpy(z[k]).to(DEV)
    errs = {"pred_logits": (_rel(out["pred_logits"], t("pred_logits")), 2e-3), "pred_boxes": (_rel(out["pred_boxes"], t("pred_boxes")), 2e-3),
            "aux0_boxes": (_rel(out["aux_outputs"][0]["pred_boxes"], t("aux0_boxes")), 2e-3)}
    for k, p in net.named_parameters():
        if "pg." + k in z.files and float(np.abs(z["pg." + k]).max()) > 1e-6:
            # encoder parameter gradients are fp32 sums over up to 8.4 M voxels behind six InstanceNorm stages, driven by a
            # loss that only sees 14 queries: heavy cancellation, the CPU fixture and the GPU differ by up to ~7 % of the
            # tensor's max there (measured) purely from accumulation order; everything outside the encoder is tight.
            # Yardstick (tests/golden/transoarnet_fp64.npz, tests/test_model_yardstick_cpu.py): the REFERENCE's own fp32 CPU
            # run is up to 2.7 % of the tensor's max away from its float64 run in exactly these tensors, < 4e-5 in all others
            errs["grad " + k] = (_rel(p.grad, t("pg." + k)), 1e-1 if k.startswith("_backbone._encoder") else 5e-3)
    _report_against_fp64(out, net)
    bad = {k: v for k, v in errs.items() if not v[0] < v[1]}
    assert not bad, f"{len(bad)} of {len(errs)} quantities off: {bad}"


def _report_against_fp64(out, net):
    """Not an assertion: per tensor, this run's distance to the reference's float64 run next to the reference's own fp32 CPU distance
    (``e32.*``), written to gpurun_out/ when that directory exists so the ratio can be read after a GPU run."""
    try:
        import json
        z64 = np.load(os.path.join(GOLDEN, "transoarnet_fp64.npz"))
        got = {"pred_logits": out["pred_logits"], "pred_boxes": out["pred_boxes"], "aux0_boxes": out["aux_outputs"][0]["pred_boxes"]}
        got.update({"pg." + k: p.grad for k, p in net.named_parameters() if p.grad is not None and "pg." + k in z64.files})
        rows = {}
        for k, v in got.items():
            ref = z64[k]
            err = float(np.abs(v.detach().double().cpu().numpy() - ref).max() / max(float(np.abs(ref).max()), 1e-30))
            rows[k] = {"gpu_vs_fp64": err, "reference_cpu_fp32_vs_fp64": float(z64["e32." + k])}
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "model_parity_vs_fp64.json"), "w") as f:
                json.dump(rows, f, indent=1)
    except Exception as exc:                                   # a report must never fail the parity test
        print("fp64 report skipped:", exc)


@pytest.mark.parametrize("hd,H,per,grid", [(48, 8, 27, (6, 7, 9)), (16, 2, 1, (3, 3, 3)), (32, 3, 7, (5, 4, 8)), (64, 2, 54, (4, 6, 5)),
                                           (96, 1, 27, (3, 5, 7)), (128, 1, 5, (2, 3, 9)), (48, 8, 27, (12, 14, 20))])
def test_roi_attention_tensor_core_kernels_match_dense_oracle_at_tf32_tolerance(hd, H, per, grid):
    """The mma.sync TF32 kernels (include/roi_attn.h, *_tf32): same results as the dense fp64-free oracle up to TF32 rounding of the
    operands (10 mantissa bits: relative 1e-3 on products; scores here are O(sqrt(hd)) so 3e-3 of each tensor's maximum), and the
    yardstick -- torch's own TF32 matmul on the dense formulation -- is not closer."""
    gen = torch.Generator().manual_seed(hd + per + 1)
    boxes = _random_boxes(3, per, grid, gen)
    boxes[0:per] = torch.tensor([0, 0, 0, *grid], dtype=torch.int32)
    boxes[per:2 * per] = torch.tensor([1, 1, 1, 2, 2, 2], dtype=torch.int32)
    Nq, Nkv, B = boxes.shape[0], grid[0] * grid[1] * grid[2], 2
    mk = lambda *s: (torch.randn(*s, generator=gen) * 0.5).to(DEV).requires_grad_(True)
    q, k, v = mk(B, Nq, H, hd), mk(B, Nkv, H, hd), mk(B, Nkv, H, hd)
    g = torch.randn(B, Nq, H * hd, generator=gen).to(DEV)
    groups = focused.groups_from_boxes(boxes).to(DEV)
    from transoar_b200 import _lib
    n0 = _lib.lib().msda3d_launch_count()
    out = focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:], True)
    out.backward(g)
    assert _lib.lib().msda3d_launch_count() - n0 in (2, 3)
    got = [out.detach(), q.grad.clone(), k.grad.clone(), v.grad.clone()]
    q.grad = k.grad = v.grad = None
    want = dense_masked_attention(q, k, v, boxes, grid)
    want.backward(g)
    for a, b in zip(got, [want.detach(), q.grad, k.grad, v.grad]):
        assert _rel(a, b) < 4e-3, _rel(a, b)
    # against the strict-fp32 CUDA-core kernels of the same library: the two must differ only at TF32 level too
    q.grad = k.grad = v.grad = None
    out32 = focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:], False)
    out32.backward(g)
    for a, b in zip(got, [out32.detach(), q.grad, k.grad, v.grad]):
        assert _rel(a, b) < 4e-3
