"""The whole training step (transoar_b200.engine.TrainStep) on the GPU against the reference's CPU route of the same step
(oracle/model_oracle.py) from identical weights and inputs, and its behaviour over a few steps at full size."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(shape, tf32):
    from oracle.model_oracle import CpuTrainStep
    from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
    cfg = visceral_train_config()
    cfg["neck_input_shape"] = tuple(s // 4 for s in shape)
    cpu = CpuTrainStep(cfg, shape, seed=3)
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    gpu = TrainStep(cfg, "cuda:0", tf32=tf32)
    if not tf32:
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    gpu.net.load_state_dict(cpu.net.state_dict())
    return cfg, cpu, gpu, synthetic_targets, prev


@pytest.mark.parametrize("tf32,tol", [(False, 2e-3), (True, 3e-2)])
def test_first_step_loss_matches_the_cpu_reference_route(tf32, tol):
    shape = (64, 64, 128)
    cfg, cpu, gpu, synthetic_targets, prev = _pair(shape, tf32)
    try:
        gpu.net.eval(); cpu.net.eval()                      # dropout off: both sides see the same function
        x = torch.rand(1, 1, *shape, generator=torch.Generator().manual_seed(1))
        _, loss_cpu = cpu.step(x, synthetic_targets(cfg, 1, 0, "cpu"))
        loss_gpu = float(gpu.step(x.pin_memory(), synthetic_targets(cfg, 1, 0, "cuda:0")))
        assert abs(loss_gpu - loss_cpu) < tol * abs(loss_cpu), (loss_gpu, loss_cpu)
        # after one optimiser step from identical states the heads have moved identically (AdamW's first step is sign-like; compare logits' head)
        wc, wg = cpu.net._cls_head.weight.detach(), gpu.net._cls_head.weight.detach().cpu()
        agree = float((torch.sign(wc) == torch.sign(wg)).float().mean())
        assert agree > 0.9, agree
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


def test_full_size_steps_run_without_host_sync_and_reduce_the_loss():
    from transoar_b200 import _lib
    from transoar_b200.engine import TrainStep, synthetic_targets, visceral_train_config
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        cfg = visceral_train_config()
        torch.manual_seed(0)
        ts = TrainStep(cfg, "cuda:0")
        x = torch.rand(2, 1, 160, 160, 256, device="cuda:0")
        tg = synthetic_targets(cfg, 2, 0, "cuda:0")
        n0 = _lib.lib().msda3d_launch_count()
        losses = [ts.step(x, tg) for _ in range(4)]
        per_step = (_lib.lib().msda3d_launch_count() - n0) / 4
        losses = [float(l) for l in losses]
        assert all(l == l for l in losses) and losses[-1] < losses[0], losses
        # 2 refinement layers x (msda fwd + bwd) + 3 RoI attention layers + 12 InstanceNorm pairs + the tcgen05 GEMMs
        assert per_step >= 2 * 2 + 3 * 2 + 12 * 7 + 30, per_step
        del ts
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        torch.cuda.empty_cache()
