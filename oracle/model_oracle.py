"""TEST INFRASTRUCTURE ONLY -- the reference's CPU route of the whole training step, restated.

The reference trains on a CPU (or debugs) through ``use_cuda=False``: ``ms_deform_attn_core_pytorch`` (F.grid_sample per level,
transoar/models/ops/functions/ms_deform_attn_func.py:41-65), ``nn.InstanceNorm3d -> nn.ReLU`` (backbones/encoder_blocks.py:28-46)
and the dense, additively masked attention of ``FocusedAttn.forward`` (necks/focused_decoder.py:238-254), all ATen kernels.
This module runs the model mirrors of ``transoar_b200`` (structurally the reference's modules, checkpoint-compatible) with
exactly those three ATen compositions substituted for the package's fused CUDA ops, so the same step can execute on host
cores.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline / ``--impl reference`` arm import it; the
product (``transoar_b200/``) never does and has no CPU path of its own."""
import contextlib
import time

import torch
import torch.nn.functional as F

from . import msda3d_oracle as O
from .focused_attn_oracle import dense_masked_attention


class _GridSampleFunction:
    """apply() of the reference's use_cuda=False branch (ms_deform_attn.py:137-138)."""

    @staticmethod
    def apply(value, shapes, starts, loc, aw, im2col_step):
        return O.gridsample_path(value, [tuple(int(v) for v in s) for s in shapes.tolist()], loc, aw)


class _DenseRoIAttention:
    @staticmethod
    def apply(q, k, v, groups, grid_yz):
        raise RuntimeError("patched per module, see cpu_reference_ops")


def _instance_norm_relu(x, weight, bias, eps=1e-5):
    return F.relu(F.instance_norm(x, weight=weight, bias=bias, eps=eps))


@contextlib.contextmanager
def cpu_reference_ops():
    """Inside the block the mirrors run the reference's ATen route instead of the fused sm_100a kernels."""
    from transoar_b200 import attn_fpn, focused
    from transoar_b200.ops.modules import ms_deform_attn as mod
    saved = (mod.MSDeformAttnFunction, attn_fpn.instance_norm_relu, focused.FocusedAttn.forward)

    def focused_forward(self, q, k, v, mask=None):
        B, Nkv, C = k.shape
        Nq, H = q.shape[1], self.num_heads
        kp = F.linear(k, self.k_proj.weight, self.k_proj.bias).reshape(B, Nkv, H, C // H)
        vp = F.linear(v, self.v_proj.weight, self.v_proj.bias).reshape(B, Nkv, H, C // H)
        qp = F.linear(q, self.k_proj.weight, self.k_proj.bias).reshape(B, Nq, H, C // H) * self.scale     # focused_decoder.py:235-236
        x = dense_masked_attention(qp, kp, vp, self.boxes, self.grid_shape)
        x = self.proj_drop(self.proj(x))
        return (x, None) if self.ret_weights else x

    mod.MSDeformAttnFunction, attn_fpn.instance_norm_relu, focused.FocusedAttn.forward = _GridSampleFunction, _instance_norm_relu, focused_forward
    try:
        yield
    finally:
        mod.MSDeformAttnFunction, attn_fpn.instance_norm_relu, focused.FocusedAttn.forward = saved


class CpuTrainStep:
    """Whole train step (forward, criterion, backward, AdamW) of the VISCERAL Focused-Decoder model on host cores.
    ``shape`` is the input volume; the RoI grid follows the feature map (P2 = shape / 4)."""

    def __init__(self, config, shape, threads=None, seed=0):
        from transoar_b200.criterion import build_criterion
        from transoar_b200.transoarnet import TransoarNet
        if threads:
            torch.set_num_threads(threads)
        cfg = dict(config)
        cfg["neck_input_shape"] = tuple(s // 4 for s in shape)
        torch.manual_seed(seed)
        self.config, self.shape = cfg, tuple(shape)
        self.net = TransoarNet(cfg).train()
        for name, p in self.net.named_parameters():
            if ".q_proj." in name:
                p.requires_grad_(False)
        self.criterion = build_criterion(cfg)
        named = [(n, p) for n, p in self.net.named_parameters() if p.requires_grad]
        groups = [{"params": [p for n, p in named if "_backbone" in n]},
                  {"params": [p for n, p in named if "_backbone" not in n], "lr": float(cfg["lr"])}]
        self.optim = torch.optim.AdamW(groups, lr=float(cfg["lr_backbone"]), weight_decay=float(cfg["weight_decay"]))

    def step(self, volumes, targets):
        from transoar_b200.criterion import total_loss
        t0 = time.perf_counter()
        with cpu_reference_ops():
            self.optim.zero_grad(set_to_none=True)
            out = self.net(volumes)
            loss = total_loss(self.criterion(out, targets, None, self.net._anchors), self.config["loss_coefs"])
            loss.backward()
            self.optim.step()
        return time.perf_counter() - t0, float(loss)
