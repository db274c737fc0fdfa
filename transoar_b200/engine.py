"""One training step of the Focused-Decoder model on one GPU (or one DDP rank) -- the public call of this package.

Mirrors what ``scripts/train.py:38-63`` builds and what ``Trainer._train_one_epoch`` (transoar/trainer.py:50-87) does per
batch: model forward -> criterion (matcher + losses) -> weighted total loss -> backward -> AdamW step, with the reference's two
parameter groups (backbone at ``lr_backbone``, everything else at ``lr``).  Differences, all in DESIGN.md: fp32 without the
trainer's fp16 autocast (the reference's own CUDA op cannot run under it); TF32 tensor-core multiplies requested
explicitly (torch 1.10, the reference's pin, has them on by default); the three dead ``q_proj`` parameters are frozen so DDP
needs no ``find_unused_parameters``; no host synchronisation inside the step (the criterion stays on the device).

    ts = TrainStep(visceral_train_config(), device)
    loss = ts.step(volumes, targets)          # volumes: [B,1,160,160,256] on the host (pinned) or on the device

Multi-GPU: one process per GPU, ``world > 1`` wraps the model in DistributedDataParallel (NCCL gradient all-reduce overlapped
with the backward); volumes are sharded by rank, nothing else crosses GPUs."""
import copy
import ctypes
import os

import torch

from .configs import visceral_config
from .criterion import VISCERAL_LOSS_COEFS, build_criterion, dense_targets, total_loss
from .transoarnet import TransoarNet


def visceral_train_config(seed=0):
    """config/attn_fpn_foc_dec_visceral.yaml: model part from ``configs.visceral_config`` + the training keys (:11-41)."""
    cfg = visceral_config(seed)
    cfg.update(lr=2e-4, lr_backbone=2e-5, weight_decay=1e-4, clip_max_norm=-1, batch_size=2, anchor_matching=True,
               set_cost_class=1, set_cost_bbox=0, set_cost_giou=0, loss_coefs=copy.deepcopy(VISCERAL_LOSS_COEFS), num_classes=20)
    return cfg


def amos_train_config(seed=0, volume=(256, 256, 128)):
    """config/attn_fpn_foc_dec_amos.yaml (same training keys as the VISCERAL file, 15 classes)."""
    from .configs import amos_config
    cfg = amos_config(seed, volume)
    cfg.update(lr=2e-4, lr_backbone=2e-5, weight_decay=1e-4, clip_max_norm=-1, batch_size=2, anchor_matching=True,
               set_cost_class=1, set_cost_bbox=0, set_cost_giou=0, loss_coefs=copy.deepcopy(VISCERAL_LOSS_COEFS), num_classes=15)
    return cfg


def defdetr_train_config(seed=0, volume=(256, 256, 128)):
    """configs.defdetr_amos_config + the reference's training keys; matching runs on the predicted boxes (there are no atlas anchors)."""
    from .configs import defdetr_amos_config
    cfg = defdetr_amos_config(seed, volume)
    cfg.update(lr=2e-4, lr_backbone=2e-5, weight_decay=1e-4, clip_max_norm=-1, batch_size=2, anchor_matching=False,
               set_cost_class=1, set_cost_bbox=5, set_cost_giou=2, loss_coefs=copy.deepcopy(VISCERAL_LOSS_COEFS), num_classes=15)
    return cfg


def swin_focused_train_config(seed=0, volume=(192, 192, 384)):
    from .configs import swin_focused_config
    cfg = swin_focused_config(seed, volume)
    cfg.update(lr=2e-4, lr_backbone=2e-5, weight_decay=1e-4, clip_max_norm=-1, batch_size=2, anchor_matching=True,
               set_cost_class=1, set_cost_bbox=0, set_cost_giou=0, loss_coefs=copy.deepcopy(VISCERAL_LOSS_COEFS), num_classes=20)
    return cfg


def build_model(config):
    """``TransoarNet`` (Focused Decoder, the reference's shipped model) or the restated 3D Deformable-DETR (config['model_family'] == 'def_detr')."""
    if config.get("model_family") == "def_detr":
        from .def_detr import DefDetrNet
        return DefDetrNet(config)
    return TransoarNet(config)


def synthetic_targets(config, batch, seed, device):
    """One box per organ = the atlas median jittered by a few percent, labels 1..num_organs; dense form."""
    g = torch.Generator().manual_seed(seed)
    med = torch.tensor([p["median"] for p in config["bbox_properties"].values()], dtype=torch.float32)
    targets = []
    for _ in range(batch):
        boxes = (med + (torch.rand(med.shape, generator=g) - 0.5) * 0.04).clamp(0.01, 0.99)
        targets.append({"boxes": boxes, "labels": torch.arange(1, med.shape[0] + 1)})
    return dense_targets(targets, med.shape[0], device)


def optimizer_param_groups(net, config):
    """The reference's two AdamW groups in the reference's order (scripts/train.py:52-60): parameters whose name contains ``_backbone``
    at ``lr_backbone``, everything else at ``lr``.  The three dead ``q_proj`` tensors of every FocusedAttn stay IN their
    group although ``TrainStep`` freezes them for DDP: in the reference they have ``requires_grad=True`` and simply never receive a
    gradient (AdamW skips a parameter whose ``.grad`` is None), so group sizes and parameter indices -- and therefore
    ``optimizer.state_dict()`` / ``load_state_dict`` of a checkpoint -- are interchangeable with the reference in both directions."""
    named = [(n, p) for n, p in net.named_parameters() if p.requires_grad or ".q_proj." in n]
    return [{"params": [p for n, p in named if "_backbone" in n]},
            {"params": [p for n, p in named if "_backbone" not in n], "lr": float(config["lr"])}]


def _all_reduce_mean(flat, world):
    """Mean over the ranks, in place, one collective: NCCL averages inside the reduction (ReduceOp.AVG); gloo sums and we divide."""
    import torch.distributed as dist
    diag = os.environ.get("TRANSOAR_B200_DIAG_ALLREDUCE", "")          # diagnostics only: "skip" = no exchange at all, "local" = no NCCL call
    if diag == "skip" or flat.numel() == 0:
        return
    if diag == "local":
        flat.div_(world)
        return
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.div_(world)


def all_reduce_gradients(parameters, world):
    """Average the gradients of ``parameters`` over the ranks through one temporary flat bucket (``cat`` -> all-reduce -> copy back).
    Used for the calibration step of ``OverlappedGradientAverage`` and by callers whose gradients do not live in a persistent flat
    buffer; the steady state of the graph-captured step does not copy (see ``OverlappedGradientAverage``)."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    if os.environ.get("TRANSOAR_B200_DIAG_ALLREDUCE", "") == "skip":
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    _all_reduce_mean(flat, world)
    torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])


def _dense_view(flat, offset, p):
    """A view of ``flat[offset : offset + p.numel()]`` with the parameter's own size and strides (contiguous or channels-last: any dense
    permutation), so that autograd accumulates into it in place and the fused AdamW sees gradient and parameter in one layout."""
    if p.is_contiguous():
        return flat[offset:offset + p.numel()].view(p.shape)
    strides = sorted(((st, sz) for st, sz in zip(p.stride(), p.shape) if sz > 1), reverse=True)
    dense, expect = True, 1
    for st, sz in reversed(strides):
        dense &= st == expect
        expect *= sz
    if not dense:
        raise ValueError("parameter is neither contiguous nor a dense permutation; cannot alias its gradient into the flat buffer")
    return flat.as_strided(p.shape, p.stride(), offset)


class OverlappedGradientAverage:
    """Gradient averaging for the graph-captured step at ``world > 1``: the gradients LIVE in one persistent flat fp32 buffer
    (``p.grad`` of every parameter is a view into it, laid out [early parameters | late parameters]), so the exchange is two bare
    all-reduces on slices of that buffer -- no ``cat``, no copy back (those copies were ~1 ms of the 2.1 ms per step lost
    at 8 GPUs).  The backward reaches the encoder's full-resolution stages last: they hold under 1 % of the parameters but a quarter
    of the backward's time.  So the moment every *other* parameter has its gradient (counted by post-accumulate hooks) the early
    slice -- 99 % of the bytes -- is all-reduced on a side stream, under the rest of the backward; the small late slice follows on
    the main stream.  Everything is stream-ordered (fork and join by events), so the sequence is capturable.

    Which parameters receive gradients is learnt from one un-overlapped step (``calibrate``: ordinary ``.grad`` tensors, one temporary
    bucket); parameters that never get a gradient keep ``grad = None`` (AdamW skips them, as in the reference) and are not in the
    buffer.  From then on ``zero()`` replaces ``optimizer.zero_grad()``: one memset of the buffer, the views stay."""

    def __init__(self, net, world, late_prefixes=("_backbone._encoder._stages.0.", "_backbone._encoder._stages.1.", "_backbone._encoder._stages.2.")):
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        self.world = world
        self.late = [p for n, p in named if n.startswith(tuple(late_prefixes))]
        self.early = [p for n, p in named if not n.startswith(tuple(late_prefixes))]
        self.expected, self.pending, self.fired = None, 0, False
        self.flat, self.n_early = None, 0
        cuda = bool(named) and named[0][1].is_cuda
        self.side = torch.cuda.Stream(named[0][1].device) if cuda else None
        for p in self.early:
            p.register_post_accumulate_grad_hook(self._on_gradient)

    def _on_gradient(self, _param):
        if self.expected is None:
            return
        self.pending -= 1
        if self.pending == 0 and not self.fired:
            self.fired = True
            if self.side is None:
                _all_reduce_mean(self.flat[:self.n_early], self.world)
                return
            self.side.wait_stream(torch.cuda.current_stream())          # fork: the gradients of the early slice are complete
            with torch.cuda.stream(self.side):
                _all_reduce_mean(self.flat[:self.n_early], self.world)

    def zero(self):
        """Start of a step: gradients to zero (or to None before the calibration step)."""
        if self.flat is None:
            for p in self.early + self.late:
                p.grad = None
        else:
            self.flat.zero_()

    def before_backward(self):
        self.pending, self.fired = (self.expected or 0), False

    def _adopt(self):
        """After the calibration backward: move the gradients into the persistent buffer and alias ``p.grad`` to it."""
        early = [p for p in self.early if p.grad is not None]
        late = [p for p in self.late if p.grad is not None]
        self.expected = len(early)
        self.n_early = sum(p.numel() for p in early)
        total = self.n_early + sum(p.numel() for p in late)
        ref = (early + late)[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        with torch.no_grad():
            for p in early + late:
                view = _dense_view(self.flat, off, p)
                view.copy_(p.grad)
                p.grad = view
                off += p.numel()

    def after_backward(self):
        if self.expected is None:                                       # calibration step: adopt, then reduce everything at once
            self._adopt()
            _all_reduce_mean(self.flat, self.world)
            return
        if not self.fired:                                              # fewer gradients than calibrated: still correct, not overlapped
            _all_reduce_mean(self.flat[:self.n_early], self.world)
        elif self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)          # join
        _all_reduce_mean(self.flat[self.n_early:], self.world)


class TrainStep:
    """``graph=True``: after ``graph_warmup`` eager steps the whole step (forward, criterion, backward, gradient all-reduce, AdamW) is
    captured once into a CUDA graph and every later ``step`` is a copy of the inputs into the graph's static buffers plus one replay --
    the ~1700 launches of a step no longer pay host latency.  What makes the step capturable: no host synchronisation anywhere in it
    (device criterion), all native launches on torch's current stream without allocation, ``capturable`` AdamW, and a device epoch
    counter folded into the hash dropout seeds (``hash_rng_set_epoch``) so that replays draw new masks.  With ``world > 1`` the graph
    path averages the gradients through two flat buckets (``OverlappedGradientAverage``: NCCL all-reduce inside the graph, the large
    one under the tail of the backward) instead of DDP's hooks; parameters are broadcast from rank 0 at construction as DDP would."""

    def __init__(self, config, device, world=1, tf32=True, channels_last=True, cudnn_autotune=True, graph=False, graph_warmup=3, amp_dtype=None):
        self.config, self.device = config, torch.device(device)
        # amp_dtype=torch.bfloat16: forward + criterion inside torch.autocast, as the reference's trainer runs them (fp16 + GradScaler there,
        # trainer.py:67-69; bf16 needs no loss scaling).  Linear layers then take the bf16 tcgen05 GEMM route, the op gets bf16 `value`.
        self.amp_dtype = amp_dtype
        if tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        if cudnn_autotune:
            # the reference pins cudnn.benchmark = False for reproducibility (scripts/train.py:113-114); with fixed shapes the
            # autotuner picks the weight-stationary tensor-core kernels for the 24- and 48-channel stages (105 -> 93 ms per step)
            torch.backends.cudnn.benchmark = True
        self.net = build_model(config).to(self.device).train()
        if channels_last:
            # conv weights NDHWC: the backbone's activations then stay in the layout the tensor-core convolutions use, the fused
            # InstanceNorm kernels follow it, and flatten(2).transpose(1, 2) of a feature map is a view
            self.net = self.net.to(memory_format=torch.channels_last_3d)
        for name, p in self.net.named_parameters():
            if ".q_proj." in name:
                p.requires_grad_(False)
        self.model = self.net
        self.world, self.graph, self.graph_warmup = world, bool(graph), max(int(graph_warmup), 1)
        if world > 1 and not self.graph:
            from torch.nn.parallel import DistributedDataParallel as DDP
            self.model = DDP(self.net, device_ids=[self.device.index], gradient_as_bucket_view=True, static_graph=True)
        elif world > 1:
            import torch.distributed as dist
            with torch.no_grad():
                for t in list(self.net.parameters()) + list(self.net.buffers()):
                    dist.broadcast(t, 0)
            self._averager = OverlappedGradientAverage(self.net, world)
        self.criterion = build_criterion(config).to(self.device)
        groups = optimizer_param_groups(self.net, config)
        self.optim = torch.optim.AdamW(groups, lr=float(config["lr_backbone"]), weight_decay=float(config["weight_decay"]),
                                       capturable=self.graph, fused=True)
        self._staging = None
        self._copy_stream = self._prefetch_buf = self._prefetch_done = self._prefetch_read = self._prefetch_src = None
        self._cuda_graph, self._seen, self._static = None, 0, None
        self._stream = torch.cuda.Stream(self.device) if self.graph else None

    def prefetch(self, volumes):
        """Start the host -> device copy of the NEXT step's volumes (pinned host memory) on a copy stream, so that it runs under the
        current step's kernels instead of in front of the next step.  Call it after ``step`` and before reading the loss; the next
        ``step`` must be given the same host tensor.  Optional: a ``step`` without a matching prefetch copies in line as before."""
        if volumes.is_cuda:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        if self._prefetch_buf is None or self._prefetch_buf.shape != volumes.shape:
            self._prefetch_buf = torch.empty(volumes.shape, dtype=torch.float32, device=self.device)
        if self._prefetch_read is not None:
            self._copy_stream.wait_event(self._prefetch_read)         # the previous step's device-side copy has read the buffer
        with torch.cuda.stream(self._copy_stream):
            self._prefetch_buf.copy_(volumes, non_blocking=True)
            self._prefetch_done = torch.cuda.Event()
            self._prefetch_done.record(self._copy_stream)
        self._prefetch_src = (volumes.data_ptr(), tuple(volumes.shape))

    def to_device(self, volumes):
        """Host volumes (ideally pinned) -> a reused device buffer, asynchronously on the current stream (or, if ``prefetch`` was called
        for this host tensor, a device-to-device copy of the already transferred volumes)."""
        if volumes.is_cuda and not self.graph:
            return volumes
        if self._staging is None or self._staging.shape != volumes.shape:
            self._staging = torch.empty(volumes.shape, dtype=torch.float32, device=self.device)
            self._cuda_graph = None                                   # a graph captured for another shape is void
        if volumes.data_ptr() != self._staging.data_ptr():
            if self._prefetch_src == (volumes.data_ptr(), tuple(volumes.shape)):
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(self._prefetch_done)
                self._staging.copy_(self._prefetch_buf, non_blocking=True)
                self._prefetch_read = torch.cuda.Event()
                self._prefetch_read.record(cur)
                self._prefetch_src = None
            else:
                self._staging.copy_(volumes, non_blocking=True)
        return self._staging

    def invalidate_graph(self):
        """Drop the captured graph (the next ``step`` captures again): needed when something baked into it changes -- the learning
        rate after a scheduler step, a frozen / unfrozen parameter."""
        self._cuda_graph = None

    def close(self):
        """Uninstall this object's dropout epoch counter from the library (see ``_capture``)."""
        if self._static is not None:
            from . import _lib
            _lib.lib().hash_rng_set_epoch(None)
            self._static = self._cuda_graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, x, targets, seg_targets):
        if self.world > 1 and self.graph:
            self._averager.zero()                                       # gradients live in the averager's flat buffer
        else:
            self.optim.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.amp_dtype or torch.bfloat16, enabled=self.amp_dtype is not None):
            out = self.model(x)
            losses = self.criterion(out, targets, seg_targets, self.net._anchors)
            loss = total_loss(losses, self.config["loss_coefs"])
        if self.world > 1 and self.graph:
            self._averager.before_backward()
        loss.backward()
        if self.world > 1 and self.graph:
            self._averager.after_backward()
        if self.config.get("clip_max_norm", -1) > 0:
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.config["clip_max_norm"])
        self.optim.step()
        return loss.detach()

    def _capture(self, x, boxes, valid):
        from . import _lib
        epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._static = {"boxes": boxes.clone(), "valid": valid.clone(), "epoch": epoch}
        _lib.lib().hash_rng_set_epoch(ctypes.c_void_p(epoch.data_ptr()))
        if not (self.world > 1 and self.graph):
            self.optim.zero_grad(set_to_none=True)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):                 # recorded, not executed: the first replay is this step
            epoch.add_(1)
            self._static["loss"] = self._run(x, (self._static["boxes"], self._static["valid"]), None)
        self._cuda_graph = g

    def _eager_on_side_stream(self, x, targets):
        """The eager steps before the capture run on the stream the capture will use, so that autograd's gradient-accumulation nodes
        and the allocator are bound to it from the start (a node bound to the default stream cannot join a capture)."""
        cur = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            loss = self._run(x, targets, None)
        cur.wait_stream(self._stream)
        return loss

    def eager_step(self, volumes, targets):
        """One step launched eagerly (kernel by kernel) even when this object replays a CUDA graph: same model, optimiser state and
        stream.  bench.py uses it to time individual kernels with CUDA events, which cannot be read inside a replayed graph."""
        x = self.to_device(volumes)
        boxes, valid = targets if isinstance(targets, tuple) else dense_targets(targets, self.criterion.num_classes, self.device)
        if self._stream is None:
            return self._run(x, (boxes, valid), None)
        return self._eager_on_side_stream(x, (boxes, valid))

    def step(self, volumes, targets, seg_targets=None):
        """targets: the reference's list of {'boxes','labels'} dicts or the dense (boxes [B,O,6], valid [B,O]) pair.  Returns the total loss (device scalar)."""
        x = self.to_device(volumes)
        if not self.graph or seg_targets is not None:
            return self._run(x, targets, seg_targets)
        boxes, valid = targets if isinstance(targets, tuple) else dense_targets(targets, self.criterion.num_classes, self.device)
        if self._cuda_graph is None:
            self._seen += 1
            if self._seen <= self.graph_warmup:                         # eager: cuDNN autotuning, lazy initialisation, NCCL communicator
                return self._eager_on_side_stream(x, (boxes, valid))
            self._capture(x, boxes, valid)
        self._static["boxes"].copy_(boxes, non_blocking=True)
        self._static["valid"].copy_(valid, non_blocking=True)
        self._cuda_graph.replay()
        return self._static["loss"]
