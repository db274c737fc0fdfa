"""baseline/_ref (tools/install_reference.py): the installed tree is the reference's, its model and this repository's mirror share one
state_dict layout, anchors and criterion -- checked on the host (construction only; a forward at the only volume size the reference
supports takes a minute on CPU and runs on the GPU box instead, tests/test_gpu_reference_model.py)."""
import hashlib
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_model as R  # noqa: E402

pytestmark = pytest.mark.skipif(not R.available(), reason="baseline/_ref not installed (python tools/install_reference.py)")


def test_installed_tree_is_verbatim():
    with open(os.path.join(R.REF_ROOT, "MANIFEST.json")) as f:
        man = json.load(f)
    assert len(man["files"]) >= 30
    for rel, digest in man["files"].items():
        with open(os.path.join(R.REF_ROOT, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):                                  # build container: compare with the mounted reference itself
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel


@pytest.fixture(scope="module")
def pair():
    from transoar_b200.transoarnet import TransoarNet
    cfg = R.reference_config(use_cuda=False)
    try:
        ref = R.build_model(cfg, "cpu")
    finally:
        R.leave_cpu_mode()
    ours = TransoarNet({"backbone": dict(cfg["backbone"]), "neck": dict(cfg["neck"]), "bbox_properties": cfg["bbox_properties"]})
    return cfg, ref, ours


def test_mirror_loads_the_reference_state_dict(pair):
    cfg, ref, ours = pair
    res = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert torch.equal(ours._anchors, ref._anchors) and torch.equal(ours._restrictions, ref._restrictions)
    assert sum(p.numel() for p in ref.parameters()) == 54294447


def test_reference_criterion_equals_device_criterion(pair):
    cfg, ref, _ = pair
    from transoar.models.build import build_criterion as ref_build
    from transoar_b200.criterion import build_criterion, dense_targets, total_loss
    from transoar_b200.engine import visceral_train_config
    tcfg = visceral_train_config()
    tg = R.list_targets(cfg, 2, 7, "cpu")
    g = torch.Generator().manual_seed(0)
    out = {"pred_logits": torch.randn(2, 540, 1, generator=g), "pred_boxes": torch.rand(2, 540, 6, generator=g) * 0.5 + 0.2, "pred_seg": 0,
           "aux_outputs": [{"pred_logits": torch.randn(2, 540, 1, generator=g), "pred_boxes": torch.rand(2, 540, 6, generator=g)} for _ in range(2)]}
    R.cpu_mode()
    try:
        rl = ref_build(cfg)(out, tg, None, ref._anchors)
    finally:
        R.leave_cpu_mode()
    rt = sum(v * cfg["loss_coefs"][k.split("_")[0]] for k, v in rl.items())
    ot = total_loss(build_criterion(tcfg)(out, dense_targets(tg, 20, "cpu"), None, ref._anchors), tcfg["loss_coefs"])
    assert abs(float(rt) - float(ot)) <= 1e-5 * abs(float(rt))


def test_reference_cuda_route_needs_an_op_bound():
    """as shipped, use_cuda=True dies with NameError('MSDA'); bind_op supplies the name without editing the file."""
    func = R.bind_op("ours")
    import transoar_b200.MultiScaleDeformableAttention as MSDA
    assert func.MSDA is MSDA
    func = R.bind_op("reference")
    assert func.MSDA is R.RefCudaMSDA
    with pytest.raises(RuntimeError):
        v = torch.zeros(1, 8, 1, 4)
        func.MSDA.ms_deform_attn_forward(v, torch.tensor([[2, 2, 2]]), torch.tensor([0]), torch.zeros(1, 1, 1, 1, 1, 3), torch.zeros(1, 1, 1, 1, 1), 64)


def test_optimizer_state_is_interchangeable_with_the_reference(pair):
    """the frozen q_proj tensors must stay in the AdamW groups, or a reference checkpoint's optimizer_state_dict does not
    load here (and ours not there): same group sizes, same parameter indices (scripts/train.py:52-64)."""
    cfg, ref, ours = pair
    from transoar_b200.engine import optimizer_param_groups, visceral_train_config
    match = lambda n, keys: any(k in n for k in keys)
    ref_groups = [{"params": [p for n, p in ref.named_parameters() if match(n, ["_backbone"]) and p.requires_grad]},
                  {"params": [p for n, p in ref.named_parameters() if not match(n, ["_backbone"]) and p.requires_grad], "lr": float(cfg["lr"])}]
    ref_opt = torch.optim.AdamW(ref_groups, lr=float(cfg["lr_backbone"]), weight_decay=float(cfg["weight_decay"]))
    for n, p in ours.named_parameters():
        if ".q_proj." in n:
            p.requires_grad_(False)                        # what TrainStep does for DDP
    try:
        tcfg = visceral_train_config()
        groups = optimizer_param_groups(ours, tcfg)
        assert [len(g["params"]) for g in groups] == [len(g["params"]) for g in ref_groups]
        opt = torch.optim.AdamW(groups, lr=float(tcfg["lr_backbone"]), weight_decay=float(tcfg["weight_decay"]))
        # one reference step so that its state dict carries exp_avg / exp_avg_sq for every parameter that gets a gradient
        for g in ref_groups:
            for p in g["params"]:
                if ".q_proj." not in [n for n, q in ref.named_parameters() if q is p][0]:
                    p.grad = torch.zeros_like(p)
        ref_opt.step()
        opt.load_state_dict(ref_opt.state_dict())          # raises on a group-size mismatch
        ref_opt.load_state_dict(opt.state_dict())
        assert len(opt.state_dict()["state"]) == len(ref_opt.state_dict()["state"])
    finally:
        for p in ours.parameters():
            p.requires_grad_(True)
        for p in ref.parameters():
            p.grad = None
