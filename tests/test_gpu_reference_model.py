"""The UNMODIFIED reference model on the GPU (VERDICT r01 "what's weak" 1, "next" 1c).

``baseline/_ref`` (tools/install_reference.py) holds the reference's own package; ``oracle/_ref`` its own CUDA kernels compiled for
sm_100a.  Three checks at the real VISCERAL size (the reference hard-codes the 40x40x64 RoI grid, focused_decoder.py:99-117):

* the reference ``TransoarNet`` gives the same logits / boxes / parameter gradients whether ``MSDA`` is its own compiled op or this
  repository's library bound by ``transoar_b200.install_into_reference()``  (north_star: "match the reference's own compiled op on
  identical inputs within 1e-4 fp32");
* this repository's mirror model (``transoar_b200.transoarnet.TransoarNet``) loaded with the reference's ``state_dict`` reproduces the
  reference model's logits / boxes: <= 1e-4 with strict fp32 multiplies, and no worse than the reference itself moves when it is
  switched to TF32 (cuBLAS / cuDNN TF32 as the yardstick) with the tcgen05 TF32 kernels on;
* the reference criterion's loss on the reference outputs equals the device criterion's on ours."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _need():
    from oracle import msda3d_oracle as O
    from oracle import reference_model as R
    if not R.available():
        pytest.skip("baseline/_ref not installed (tools/install_reference.py, build container only)")
    if not O.refcuda_available():
        pytest.skip("oracle/_ref not built")
    return R


@pytest.fixture(scope="module")
def ref_env():
    R = _need()
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    dev = torch.device("cuda", 0)
    cfg = R.reference_config(use_cuda=True)
    R.bind_op("reference")
    torch.manual_seed(3)
    model = R.build_model(cfg, dev).eval()
    # the reference zero-initialises the class head and the last box layer (transoarnet.py:50-58): every query would answer
    # "anchor, logit 0" and the comparison would be vacuous -- give the heads weights
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in (model._cls_head.weight, model._cls_head.bias, model._reg_head.layers[-1].weight, model._reg_head.layers[-1].bias):
            p.copy_((torch.rand(p.shape, generator=gen) - 0.5) * 0.2)
    x = torch.rand(1, 1, 160, 160, 256, generator=torch.Generator().manual_seed(5)).to(dev)
    yield R, model, cfg, x, dev
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved


def _fwd_bwd(model, x):
    for p in model.parameters():
        p.grad = None
    out = model(x)
    w = torch.linspace(0.5, 1.5, 6, device=x.device)
    loss = out["pred_logits"].sum() + (out["pred_boxes"] * w).sum() + sum((a["pred_boxes"] * w).sum() + a["pred_logits"].sum() for a in out["aux_outputs"])
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return out["pred_logits"].detach().clone(), out["pred_boxes"].detach().clone(), grads


def test_reference_model_is_unchanged_by_swapping_in_our_op(ref_env):
    R, model, cfg, x, dev = ref_env
    from transoar_b200 import _lib
    R.bind_op("reference")
    logits_r, boxes_r, grads_r = _fwd_bwd(model, x)
    n0 = _lib.lib().msda3d_launch_count()
    func = R.bind_op("ours")
    import transoar_b200.MultiScaleDeformableAttention as MSDA
    assert func.MSDA is MSDA
    logits_o, boxes_o, grads_o = _fwd_bwd(model, x)
    assert _lib.lib().msda3d_launch_count() == n0 + 2 * 2, "our kernels did not run inside the reference model (2 layers x fwd + bwd)"
    R.bind_op("reference")
    assert boxes_r.std() > 1e-3 and logits_r.std() > 1e-3, "degenerate comparison"
    # forward: our op is bit-identical to the reference's compiled op, everything else is the same code
    assert (logits_o - logits_r).abs().max().item() <= 1e-4 * max(1.0, logits_r.abs().max().item())
    assert (boxes_o - boxes_r).abs().max().item() <= 1e-4
    # gradients: fp32 atomics order differs (both ops are non-deterministic there): 1e-4 of each tensor's largest entry ... of the
    # same magnitude as the reference op's own run-to-run spread
    assert set(grads_o) == set(grads_r)
    worst = max(((grads_o[n] - grads_r[n]).abs().max() / grads_r[n].abs().max().clamp_min(1e-20)).item() for n in grads_r)
    assert worst <= 2e-4, worst


def _mirror(cfg, ref_model, dev, channels_last):
    from transoar_b200.transoarnet import TransoarNet
    mcfg = {"backbone": dict(cfg["backbone"]), "neck": dict(cfg["neck"]), "bbox_properties": cfg["bbox_properties"]}
    ours = TransoarNet(mcfg).to(dev).eval()
    missing, unexpected = ours.load_state_dict(ref_model.state_dict(), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    if channels_last:
        ours = ours.to(memory_format=torch.channels_last_3d)
    return ours


def test_mirror_model_reproduces_reference_model_strict_fp32(ref_env):
    R, model, cfg, x, dev = ref_env
    R.bind_op("reference")
    with torch.no_grad():
        want = model(x)
        ours = _mirror(cfg, model, dev, channels_last=True)
        got = ours(x)
    assert torch.equal(ours._anchors, model._anchors)
    for key, tol in (("pred_logits", 1e-4), ("pred_boxes", 1e-4)):
        err = (got[key] - want[key]).abs().max().item()
        assert err <= tol * max(1.0, want[key].abs().max().item()), (key, err)
    for a, b in zip(got["aux_outputs"], want["aux_outputs"]):
        assert (a["pred_boxes"] - b["pred_boxes"]).abs().max().item() <= 1e-4


def test_mirror_model_tf32_error_is_within_the_reference_models_own_tf32_error(ref_env):
    """The bench runs TF32 tensor-core multiplies (tcgen05 GEMM / convolution kernels).  Yardstick: how far the REFERENCE model moves
    when cuBLAS / cuDNN are switched to TF32.  Our TF32 path must stay within 2x of that, measured against the strict-fp32 reference."""
    R, model, cfg, x, dev = ref_env
    R.bind_op("reference")
    with torch.no_grad():
        exact = model(x)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        try:
            ref_tf32 = model(x)
            ours = _mirror(cfg, model, dev, channels_last=True)
            got = ours(x)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
    for key in ("pred_logits", "pred_boxes"):
        yard = (ref_tf32[key] - exact[key]).abs().max().item()
        err = (got[key] - exact[key]).abs().max().item()
        assert err <= 2.0 * yard + 1e-5, (key, err, yard)


def test_reference_criterion_and_device_criterion_agree_on_gpu(ref_env):
    R, model, cfg, x, dev = ref_env
    from transoar_b200.criterion import build_criterion, dense_targets, total_loss
    from transoar_b200.engine import visceral_train_config
    R.activate()
    from transoar.models.build import build_criterion as ref_build
    tcfg = visceral_train_config()
    targets = R.list_targets(cfg, 1, 7, dev)
    with torch.no_grad():
        out = model(x)
        ref_losses = ref_build(cfg).to(dev)(out, targets, None, model._anchors)
        ref_total = sum(v * cfg["loss_coefs"][k.split("_")[0]] for k, v in ref_losses.items())
        ours = build_criterion(tcfg).to(dev)
        boxes, valid = dense_targets(targets, 20, dev)
        our_total = total_loss(ours(out, (boxes, valid), None, model._anchors), tcfg["loss_coefs"])
    assert abs(float(ref_total) - float(our_total)) <= 1e-5 * max(1.0, abs(float(ref_total)))
