"""Training launcher around ``engine.TrainStep`` -- what ``scripts/train.py:27-96`` + ``Trainer.run`` /
``_train_one_epoch`` / ``_save_checkpoint`` (transoar/trainer.py:50-110, 203-241) do, one process per GPU:

    python -m transoar_b200.train --config visceral --synthetic 8 --epochs 2                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 -m transoar_b200.train --config my.yaml    # volumes sharded over ranks

Formats kept so the reference's own tools keep working on a run made here:
  * ``runs/<experiment_name>/config.json`` -- the frozen config (yaml keys + ``data_info.json`` keys + meta data, train.py:83-88), which
    ``scripts/test.py:20-21`` loads;
  * ``runs/<experiment_name>/model_last.pt`` -- ``{'epoch', 'metric_max_val', 'model_state_dict', 'optimizer_state_dict',
    'scheduler_state_dict'}`` (trainer.py:235-241); the model keys are the reference's parameter names, so ``scripts/test.py:56-57``
    and ``--resume`` of the reference load it, and this launcher resumes from a reference checkpoint the same way (train.py:66-77).
Not rebuilt (out of scope): the evaluator / mAP (``metric_max_val`` is carried through unchanged), TensorBoard (a JSON-lines log
instead), the MONAI augmentation."""
import argparse
import json
import os
import random
import socket
import sys
from pathlib import Path

import numpy as np
import torch

from .data import SyntheticCaseDataset, detection_targets, get_loader


def load_config(name, config_dir="./config", dataset_root="./dataset"):
    """``get_config`` (utils/io.py:20-38): ``<config_dir>/<name>.yaml`` (or a path to a yaml file) merged with
    ``<dataset_root>/<dataset>/data_info.json`` when the yaml names a dataset.  ``visceral`` / ``amos`` without a yaml file select the
    built-in dictionaries of ``engine`` (the two shipped yaml files restated, with the synthetic atlas)."""
    path = Path(name) if str(name).endswith((".yaml", ".yml")) else Path(config_dir) / f"{name}.yaml"
    if not path.exists():
        from . import engine
        builtin = {"visceral": engine.visceral_train_config, "amos": engine.amos_train_config}
        if str(name) not in builtin:
            raise FileNotFoundError(f"no such config: {path}")
        config = builtin[str(name)]()
        config.setdefault("experiment_name", f"foc_dec_{name}")
        return config
    import yaml
    with open(path) as stream:
        config = yaml.safe_load(stream)
    if "dataset" in config:
        info = Path(dataset_root) / config["dataset"] / "data_info.json"
        if info.exists():
            with open(info) as f:
                config.update(json.load(f))
    return config


TRAINING_DEFAULTS = dict(epochs=1, lr_drop=2500, val_interval=1, debug_mode=False, shuffle=True, num_workers=0, bbox_padding=1,
                         overfit=False, seed=10, dataset="synthetic", augmentation={"use_augmentation": False})


def meta_data():
    """``get_meta_data`` (utils/io.py:156-164) without the git call (a run directory need not be a checkout)."""
    lines = sys.version.splitlines()
    return {"python_version": lines[0], "gcc_version": lines[1] if len(lines) > 1 else "", "pytorch_version": torch.__version__,
            "host_name": socket.gethostname()}


def save_checkpoint(path, epoch, metric_max_val, net, optim, scheduler):
    torch.save({"epoch": epoch, "metric_max_val": metric_max_val, "model_state_dict": net.state_dict(),
                "optimizer_state_dict": optim.state_dict(), "scheduler_state_dict": scheduler.state_dict()}, path)


def load_checkpoint(path, net, optim, scheduler, lr_drop, device="cpu"):
    """train.py:66-77 -- returns (epoch, metric_max_val)."""
    checkpoint = torch.load(Path(path), map_location=device, weights_only=False)
    checkpoint["scheduler_state_dict"]["step_size"] = lr_drop
    net.load_state_dict(checkpoint["model_state_dict"])
    optim.load_state_dict(checkpoint["optimizer_state_dict"])
    scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    return checkpoint["epoch"], checkpoint["metric_max_val"]


def to_jsonable(obj):
    if isinstance(obj, dict):
        return {str(k): to_jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_jsonable(v) for v in obj]
    if isinstance(obj, (np.generic,)):
        return obj.item()
    if isinstance(obj, torch.Tensor):
        return obj.tolist()
    return obj


def train(config, args):
    import torch.distributed as dist
    from .engine import TrainStep
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("transoar_b200.train needs a CUDA device: the package has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    # to get reproducible results (train.py:108-115); identical initial weights on every rank
    torch.manual_seed(config["seed"]); np.random.seed(config["seed"]); random.seed(config["seed"])
    ts = TrainStep(config, device, world=world, graph=args.graph, cudnn_autotune=not args.deterministic)
    if world > 1:
        # weights are identical on every rank now (broadcast from rank 0); the dropout streams must not be: the hash seeds of the fused
        # LayerNorm / FFN kernels are drawn from torch's CPU generator and ATen's dropout from the CUDA generator
        torch.manual_seed(config["seed"] + 7919 * rank)
    scheduler = torch.optim.lr_scheduler.StepLR(ts.optim, config["lr_drop"])
    epoch, metric_max_val = 0, 0
    if args.resume is not None:
        epoch, metric_max_val = load_checkpoint(args.resume, ts.net, ts.optim, scheduler, config["lr_drop"], device)

    if args.synthetic:
        volume = tuple(args.volume)
        train_set = SyntheticCaseDataset(config, args.synthetic, volume, seed=config["seed"])
        val_set = train_set if config["overfit"] else SyntheticCaseDataset(config, max(args.synthetic // 4, config["batch_size"]), volume, seed=config["seed"] + 1)
    else:
        train_set = val_set = None
    train_loader = get_loader(config, "train", root=args.dataset_root, rank=rank, world=world, dataset=train_set)
    val_loader = get_loader(config, "train" if config["overfit"] else "val", root=args.dataset_root, rank=rank, world=world, dataset=val_set)

    path_to_run = Path(args.runs) / config["experiment_name"]
    if rank == 0:
        path_to_run.mkdir(parents=True, exist_ok=True)
        frozen = dict(config)
        frozen.update(meta_data())
        with open(path_to_run / "config.json", "w") as f:
            json.dump(to_jsonable(frozen), f, indent=3)
    log = open(path_to_run / "train_log.jsonl", "a") if rank == 0 else None
    seg = bool(config["backbone"].get("use_seg_proxy_loss", False))

    def mean_over_ranks(total, count):
        t = torch.stack((total.double(), torch.tensor(float(count), device=device, dtype=torch.float64)))
        if world > 1:
            dist.all_reduce(t)
        return float(t[0] / t[1].clamp(min=1))

    for epoch in range(epoch + 1, config["epochs"] + 1):
        ts.net.train()
        train_loader.sampler.set_epoch(epoch)
        total, steps = torch.zeros((), device=device), 0
        for data, _, bboxes, seg_targets in train_loader:
            loss = ts.step(data, detection_targets(bboxes, device), seg_targets.to(device, non_blocking=True) if seg else None)
            total, steps = total + loss, steps + 1                  # accumulated on the device: no host sync inside the epoch
        record = {"epoch": epoch, "train_total_loss": mean_over_ranks(total, steps), "steps_per_rank": steps,
                  "lr_backbone": ts.optim.param_groups[0]["lr"], "lr_neck": ts.optim.param_groups[1]["lr"]}
        if epoch % config["val_interval"] == 0:
            record["val_total_loss"] = validate(ts, val_loader, device, seg, mean_over_ranks)
        lr_before = ts.optim.param_groups[0]["lr"]
        scheduler.step()
        if ts.optim.param_groups[0]["lr"] != lr_before:
            ts.invalidate_graph()                                      # the learning rate is a constant of a captured step
        if rank == 0:
            log.write(json.dumps(record) + "\n"); log.flush()
            print(json.dumps(record), flush=True)
            if not config["debug_mode"]:
                save_checkpoint(path_to_run / "model_last.pt", epoch, metric_max_val, ts.net, ts.optim, scheduler)
    if log is not None:
        log.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return path_to_run


@torch.no_grad()
def validate(ts, loader, device, seg, mean_over_ranks):
    """The loss half of ``Trainer._validate`` (trainer.py:112-165); the detection metrics are the evaluator's (out of scope)."""
    from .criterion import total_loss
    ts.net.eval()
    total, steps = torch.zeros((), device=device), 0
    for data, _, bboxes, seg_targets in loader:
        out = ts.net(ts.to_device(data) if not ts.graph else data.to(device, non_blocking=True))
        losses = ts.criterion(out, detection_targets(bboxes, device), seg_targets.to(device) if seg else None, ts.net._anchors)
        total, steps = total + total_loss(losses, ts.config["loss_coefs"]), steps + 1
    ts.net.train()
    return mean_over_ranks(total, steps)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config", required=True, help="yaml name in --config-dir, a path to a yaml file, or the built-ins visceral / amos")
    ap.add_argument("--resume", default=None, help="path to a checkpoint (this launcher's or the reference's)")
    ap.add_argument("--config-dir", default="./config")
    ap.add_argument("--dataset-root", default="./dataset")
    ap.add_argument("--runs", default="./runs")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--synthetic", type=int, default=0, help="train on this many synthetic cases instead of an npy dataset")
    ap.add_argument("--volume", type=int, nargs=3, default=[160, 160, 256])
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (fixed shapes)")
    ap.add_argument("--deterministic", action="store_true", help="cudnn.benchmark off, as the reference pins it")
    args = ap.parse_args(argv)
    config = load_config(args.config, args.config_dir, args.dataset_root)
    for k, v in TRAINING_DEFAULTS.items():
        config.setdefault(k, v)
    if args.epochs is not None:
        config["epochs"] = args.epochs
    if args.synthetic:                                               # the RoI masks are built for the feature map the neck attends to
        stride = 2 ** int(str(config["neck"]["input_levels"])[1:])
        config["neck_input_shape"] = tuple(s // stride for s in args.volume)
    train(config, args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
