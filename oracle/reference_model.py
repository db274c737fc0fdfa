"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference model (bwittmann/transoar, installed by tools/install_reference.py into the
git-ignored ``baseline/_ref/``) driven through its own public classes: ``TransoarNet`` (transoar/models/transoarnet.py:11-155),
``build_criterion`` (models/build.py:31-46) and the optimiser / step of ``scripts/train.py:52-64`` + ``trainer.py:50-87``.

Two ways of running it, neither of which edits a reference file:

* **on the GPU with its own compiled op** -- the reference's ``use_cuda=True`` route expects a module named ``MSDA`` inside
  ``ms_deform_attn_func.py`` (the import is commented out there, func.py:18).  ``RefCudaMSDA`` is that module: it restates the
  reference's host wrapper (``ms_deform_attn_cuda.cu:20-80,83-154``: contiguity asserts, zero-filled outputs, the ``im2col_step``
  batch loop) around ``oracle/_ref/libmsda3d_refcuda.so``, i.e. the reference's own kernels compiled from where they lie.
  This is the "reference's own compiled op" comparator of BASELINE.json's north_star (`bench.py`'s ``ref_gpu_model``) and the
  checker of ``tests/test_gpu_reference_model.py``.
* **on host cores through ``use_cuda=False``** -- the reference's pure-PyTorch route (``ms_deform_attn_core_pytorch``); the model's
  unconditional ``.cuda()`` calls (transoarnet.py:27-29, focused_decoder.py:120, criterion.py:48 -- SURVEY D9) are made no-ops
  in this process by ``cpu_mode()``.  This is ``bench.py --impl reference`` and ``cpu_baseline``.

``bind_op("ours")`` runs the same unmodified model on this repository's kernels via ``transoar_b200.install_into_reference()``.
Only ``tests/`` and ``bench.py``'s reference / baseline legs import this file; nothing under ``transoar_b200/`` does."""
from __future__ import annotations

import copy
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "transoar", "models", "transoarnet.py"))


def activate():
    """Put the installed reference (and the timm shim next to it) on sys.path.  Idempotent."""
    if not available():
        raise ImportError("the reference is not installed: run `python tools/install_reference.py` in the build container "
                          "(baseline/_ref/ is git-ignored and ships with the gpurun snapshot)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


_ORIGINAL_CUDA = None


def cpu_mode():
    """the reference calls ``.cuda()`` unconditionally (constructors and every criterion call).  For a host-only run make
    it the identity in this process; ``leave_cpu_mode()`` restores torch's method."""
    global _ORIGINAL_CUDA
    if _ORIGINAL_CUDA is None:
        _ORIGINAL_CUDA = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self


def leave_cpu_mode():
    global _ORIGINAL_CUDA
    if _ORIGINAL_CUDA is not None:
        torch.Tensor.cuda = _ORIGINAL_CUDA
        _ORIGINAL_CUDA = None


# ---------------------------------------------------------------------------------------------------------------------
# MSDA module over the reference's own compiled kernels
# ---------------------------------------------------------------------------------------------------------------------
def _assert_inputs(named):
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")                 # ms_deform_attn_cuda.cu:28-32,93-98
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")                       # ms_deform_attn_cuda.cu:34-38,100-105


def _ref_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    from . import msda3d_oracle as O
    _assert_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)])
    batch = value.size(0)
    step = min(batch, int(im2col_step))                                                # ms_deform_attn_cuda.cu:50
    if batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")          # :52
    outs = [O.refcuda_forward(value[n:n + step], spatial_shapes, level_start_index, sampling_loc[n:n + step], attn_weight[n:n + step])
            for n in range(0, batch, step)]                                            # :56-75 (one launch per chunk, zero-filled output :54)
    return outs[0] if len(outs) == 1 else torch.cat(outs)


def _ref_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    from . import msda3d_oracle as O
    _assert_inputs([("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)])
    batch = value.size(0)
    step = min(batch, int(im2col_step))
    if batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    gv, gl, ga = torch.zeros_like(value), torch.zeros_like(sampling_loc), torch.zeros_like(attn_weight)   # :122-124
    for n in range(0, batch, step):                                                    # :126-149
        O.refcuda_backward(grad_output[n:n + step], value[n:n + step], spatial_shapes, level_start_index, sampling_loc[n:n + step],
                           attn_weight[n:n + step], out=(gv[n:n + step], gl[n:n + step], ga[n:n + step]))
    return [gv, gl, ga]


RefCudaMSDA = types.ModuleType("MultiScaleDeformableAttention_reference_kernels")
RefCudaMSDA.ms_deform_attn_forward = _ref_forward
RefCudaMSDA.ms_deform_attn_backward = _ref_backward


def bind_op(which):
    """Bind the name ``MSDA`` inside the reference's function module: "reference" = its own kernels (oracle/_ref),
    "ours" = this repository's library (transoar_b200.install_into_reference)."""
    activate()
    import importlib
    func = importlib.import_module("transoar.models.ops.functions.ms_deform_attn_func")
    if which == "reference":
        func.MSDA = RefCudaMSDA
    elif which == "ours":
        import transoar_b200
        transoar_b200.install_into_reference()
    else:
        raise ValueError(which)
    return func


# ---------------------------------------------------------------------------------------------------------------------
# Config, targets, model, step
# ---------------------------------------------------------------------------------------------------------------------
def reference_config(use_cuda: bool, atlas_seed: int = 0, yaml_name: str = "attn_fpn_foc_dec_visceral.yaml", num_organs: int = 20):
    """The reference's own yaml (config/attn_fpn_foc_dec_visceral.yaml) with the two hot-path switches on and what
    ``get_config`` merges in from ``dataset/<name>/data_info.json`` (utils/io.py:33-36): ``bbox_properties`` (the synthetic atlas of
    SURVEY 8d -- the same one ``transoar_b200.configs`` uses) and ``num_classes``."""
    import yaml
    from transoar_b200.configs import synthetic_atlas
    with open(os.path.join(REF_ROOT, "config", yaml_name)) as f:
        cfg = yaml.safe_load(f)
    cfg["backbone"]["use_decoder_attn"] = True
    cfg["backbone"]["use_cuda"] = bool(use_cuda)
    cfg["bbox_properties"] = synthetic_atlas(num_organs, atlas_seed)
    cfg["num_classes"] = num_organs
    return cfg


def list_targets(config, batch, seed, device):
    """The same boxes as ``transoar_b200.engine.synthetic_targets`` in the reference's collated form: list of {'boxes','labels'}."""
    g = torch.Generator().manual_seed(seed)
    med = torch.tensor([p["median"] for p in config["bbox_properties"].values()], dtype=torch.float32)
    out = []
    for _ in range(batch):
        boxes = (med + (torch.rand(med.shape, generator=g) - 0.5) * 0.04).clamp(0.01, 0.99)
        out.append({"boxes": boxes.to(device), "labels": torch.arange(1, med.shape[0] + 1).to(device)})
    return out


def build_model(config, device):
    activate()
    if torch.device(device).type == "cpu":
        cpu_mode()
    from transoar.models.transoarnet import TransoarNet
    return TransoarNet(config).to(device=device)


class ReferenceTrainStep:
    """scripts/train.py:38-64 (model, criterion, AdamW with the two parameter groups) + one iteration of
    ``Trainer._train_one_epoch`` (trainer.py:50-87) without the fp16 autocast / GradScaler pair (the reference's CUDA op cannot run
    under autocast; on the CPU route autocast does not apply).  ``op``: "reference" | "ours" | None (use_cuda=False)."""

    def __init__(self, device, op="reference", seed=0, config=None):
        activate()
        self.device = torch.device(device)
        if self.device.type == "cpu":
            cpu_mode()
        self.config = config if config is not None else reference_config(use_cuda=op is not None)
        if op is not None:
            bind_op(op)
        from transoar.models.build import build_criterion
        from transoar.models.transoarnet import TransoarNet
        torch.manual_seed(seed)
        self.model = TransoarNet(self.config).to(device=self.device)                    # train.py:39
        self.criterion = build_criterion(self.config).to(device=self.device)            # train.py:40
        match = lambda n, keys: any(k in n for k in keys)
        groups = [{"params": [p for n, p in self.model.named_parameters() if match(n, ["_backbone"]) and p.requires_grad]},
                  {"params": [p for n, p in self.model.named_parameters() if not match(n, ["_backbone"]) and p.requires_grad],
                   "lr": float(self.config["lr"])}]                                     # train.py:52-60
        self.optim = torch.optim.AdamW(groups, lr=float(self.config["lr_backbone"]), weight_decay=float(self.config["weight_decay"]))
        self.model.train()                                                              # trainer.py:46

    def step(self, volumes, targets):
        """volumes [B,1,160,160,256] (host or device), targets: list of {'boxes','labels'}.  Returns the total loss as a python float
        (the reference reads six ``.item()`` per step, trainer.py:87-92)."""
        data = volumes.to(device=self.device)                                           # trainer.py:56
        det_targets = [{"boxes": t["boxes"].to(dtype=torch.float, device=self.device), "labels": t["labels"].to(device=self.device)}
                       for t in targets]                                                # trainer.py:58-64
        out = self.model(data)                                                          # trainer.py:68
        loss_dict = self.criterion(out, det_targets, None, self.model._anchors)         # trainer.py:69
        loss_abs = 0
        for key, val in loss_dict.items():
            loss_abs = loss_abs + val * self.config["loss_coefs"][key.split("_")[0]]    # trainer.py:72-74
        self.optim.zero_grad()                                                          # trainer.py:76
        loss_abs.backward()
        self.optim.step()
        return float(loss_abs.item())                                                   # trainer.py:87
