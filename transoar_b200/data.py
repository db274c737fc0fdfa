"""Data side of the training step: the reference's npy case layout, its collator and the box extraction from
label maps, plus synthetic cases in the same format -- mirrors of

  transoar/utils/bboxes.py:45-95      ``segmentation2bbox``
  transoar/data/dataset.py:12-52      ``TransoarDataset``    (here ``NpyCaseDataset``; MONAI augmentation is out of scope)
  transoar/data/dataloader.py:10-58   ``get_loader`` / ``TransoarCollator``

Layout read and written (dataset.py:18-34, utils/io.py:20-38): ``<root>/<dataset>/data_info.json`` and
``<root>/<dataset>/<split>/<case>/{data.npy, label.npy}`` with ``data`` float32 [1, X, Y, Z] and ``label`` [1, X, Y, Z] organ ids
(0 = background).  The two files of a case are told apart as the reference does -- shorter path first (dataset.py:32).

What is re-designed: ``segmentation2bbox`` makes three passes over a label map (one per axis, all organs at once, as presence tables)
instead of one ``nonzero()`` over the whole volume per organ; the loader shards cases over ranks by striding (one process per GPU)
and pins its batches so ``TrainStep.step`` can copy them asynchronously.  Quirks kept: the smallest label value present is treated as
background whatever it is (bboxes.py:53), organs thinner than 5 voxels along any axis are dropped before padding (:60-62), padding
clips to [0, size] -- not size - 1 -- (:65-66), and an image without boxes yields an empty 1-D box tensor (:91-94)."""
import json
import warnings
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Sampler


def segmentation2bbox(segmentation_maps, padding, box_format="cxcyczwhd", normalize=True):
    """[B, 1, X, Y, Z] label maps -> (list of [n_b, 6] float boxes, list of [n_b] int64 class ids), one entry per sample."""
    if box_format not in ("xyzxyz", "xyxyzz", "cxcyczwhd"):
        raise ValueError("Please select a valid box format.")
    batch_boxes, batch_classes = [], []
    for map_ in segmentation_maps:
        assert map_.ndim == 4
        lab = map_[0].long()
        size = torch.tensor(lab.shape, dtype=torch.float32)
        lo_id = int(lab.min())
        rel = (lab - lo_id).reshape(-1)                                  # label ids shifted to start at 0
        n_ids = int(rel.max()) + 1
        lo, hi = [], []
        for axis, length in enumerate(lab.shape):                         # presence[id, coordinate] along this axis
            coord = torch.arange(length).reshape([-1 if a == axis else 1 for a in range(3)]).expand(lab.shape).reshape(-1)
            presence = torch.zeros(n_ids, length, dtype=torch.bool)
            presence[rel, coord] = True
            first = presence.float().argmax(1)
            last = length - 1 - presence.flip(1).float().argmax(1)
            lo.append(first)
            hi.append(last)
        lo, hi = torch.stack(lo, 1).float(), torch.stack(hi, 1).float()    # [ids, 3]
        present = torch.zeros(n_ids, dtype=torch.bool)
        present[rel] = True
        boxes, classes = [], []
        for rid in present.nonzero().flatten().tolist()[1:]:               # the smallest id present is the background (:53)
            mn, mx = lo[rid], hi[rid]
            if bool(((mx - mn) < 5).any()):
                continue
            mn = (mn - padding).clamp(min=0)
            mx = torch.minimum(mx + padding, size)
            assert bool((mn < mx).all())
            if normalize:
                mn, mx = mn / size, mx / size
            if box_format == "xyzxyz":
                boxes.append(torch.cat((mn, mx)))
            elif box_format == "xyxyzz":
                boxes.append(torch.stack((mn[0], mn[1], mx[0], mx[1], mn[2], mx[2])))
            else:
                boxes.append(torch.cat(((mx + mn) / 2, mx - mn)))
            classes.append(rid + lo_id)
        batch_classes.append(torch.tensor(classes, dtype=torch.int64))
        batch_boxes.append(torch.stack(boxes) if boxes else torch.tensor([]))
    return batch_boxes, batch_classes


class NpyCaseDataset(Dataset):
    """``TransoarDataset`` without the MONAI pipeline: one (data, label) tensor pair per case directory.  ``transform`` (optional) is
    called as ``transform(data, label, index)`` in place of the reference's augmentation; cases are listed in sorted order (the
    reference keeps the file system's order, dataset.py:20)."""

    def __init__(self, config, split, root="./dataset", transform=None):
        assert split in ["train", "val", "test"]
        self._config, self._transform = config, transform
        self._path_to_split = Path(root).resolve() / config["dataset"] / split
        self._data = sorted(p.name for p in self._path_to_split.iterdir() if p.is_dir())
        if transform is None and config.get("augmentation", {}).get("use_augmentation", False):
            warnings.warn("transoar_b200.data: the MONAI augmentation pipeline (transforms.py) is outside this package; "
                          "cases are returned as stored.  Pass transform= to apply your own.")

    def __len__(self):
        return len(self._data)

    def __getitem__(self, idx):
        if self._config.get("overfit", False):
            idx = 0
        case = self._path_to_split / self._data[idx]
        data_path, label_path = sorted(case.iterdir(), key=lambda p: len(str(p)))[:2]
        data, label = torch.from_numpy(np.load(data_path)), torch.from_numpy(np.load(label_path))
        if self._transform is not None:
            data, label = self._transform(data, label, idx)
        return data, label


class Collator:
    """``TransoarCollator`` (dataloader.py:42-58): (images [B,1,X,Y,Z], zero masks, [(boxes, classes)] per sample, label maps)."""

    def __init__(self, config):
        self._bbox_padding = config["bbox_padding"]

    def __call__(self, batch):
        images = torch.stack([image for image, _ in batch])
        labels = torch.stack([label for _, label in batch])
        boxes, classes = segmentation2bbox(labels, self._bbox_padding)
        return images, torch.zeros_like(images), list(zip(boxes, classes)), labels


class StridedSampler(Sampler):
    """Rank r of w takes cases r, r + w, r + 2w, ... of a permutation that every rank derives from (seed, epoch): the shards are
    disjoint, equally long (the tail that does not fill a round is dropped, as ``drop_last`` drops the last short batch) and need no
    communication."""

    def __init__(self, length, rank=0, world=1, shuffle=True, seed=0):
        self.length, self.rank, self.world, self.shuffle, self.seed, self.epoch = length, rank, world, shuffle, seed, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return self.length // self.world

    def __iter__(self):
        if self.shuffle:
            order = torch.randperm(self.length, generator=torch.Generator().manual_seed(self.seed + self.epoch)).tolist()
        else:
            order = list(range(self.length))
        return iter(order[self.rank:len(self) * self.world:self.world])


def get_loader(config, split, batch_size=None, root="./dataset", rank=0, world=1, dataset=None, transform=None):
    """``get_loader`` (dataloader.py:10-24) + rank sharding: no shuffling for val / test, ``drop_last``, the reference's collator."""
    dataset = dataset if dataset is not None else NpyCaseDataset(config, split, root, transform)
    shuffle = False if split in ["test", "val"] else config["shuffle"]
    sampler = StridedSampler(len(dataset), rank, world, shuffle, int(config.get("seed", 0)))
    return DataLoader(dataset, batch_size=batch_size or config["batch_size"], sampler=sampler, num_workers=config.get("num_workers", 0),
                      collate_fn=Collator(config), drop_last=True, pin_memory=torch.cuda.is_available())


def detection_targets(bboxes, device):
    """What ``Trainer._train_one_epoch`` builds from the collator's third output (trainer.py:58-64)."""
    return [{"boxes": b.reshape(-1, 6).to(dtype=torch.float, device=device), "labels": c.to(device=device)} for b, c in bboxes]


# ---------------------------------------------------------------------------------------------------------------------------------
# synthetic cases in the reference's format (there is no network for the VISCERAL / AMOS data)
# ---------------------------------------------------------------------------------------------------------------------------------
class SyntheticCaseDataset(Dataset):
    """Volumes of uniform noise in [0, 1] (the range after the reference's intensity scaling, transforms.py:89-93) with one box-shaped
    organ per atlas entry: the atlas median (``config['bbox_properties']``) jittered by a few percent.  Organs are painted in label
    order, so a later organ overwrites an earlier one where they overlap -- the boxes the collator extracts are those of what is
    visible, as with real label maps.  Case ``i`` depends on (seed, i) only."""

    def __init__(self, config, length, volume, seed=0):
        self._config, self._length, self._volume, self._seed = config, int(length), tuple(volume), int(seed)
        props = config["bbox_properties"]
        self._ids = [int(k) for k in props.keys()]
        self._median = torch.tensor([props[k]["median"] for k in props.keys()], dtype=torch.float32)

    def __len__(self):
        return self._length

    def __getitem__(self, idx):
        if self._config.get("overfit", False):
            idx = 0
        g = torch.Generator().manual_seed(self._seed * 1_000_003 + idx)
        data = torch.rand((1,) + self._volume, generator=g)
        label = torch.zeros((1,) + self._volume, dtype=torch.int16)
        size = torch.tensor(self._volume, dtype=torch.float32)
        boxes = (self._median + (torch.rand(self._median.shape, generator=g) - 0.5) * 0.04).clamp(0.01, 0.99)
        for organ, box in zip(self._ids, boxes):
            lo = ((box[:3] - box[3:] / 2).clamp(0, 1) * size).round().long()
            hi = torch.maximum(((box[:3] + box[3:] / 2).clamp(0, 1) * size).round().long(), lo + 1)
            label[0, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = organ
        return data, label


def write_synthetic_dataset(root, config, volume, n_train, n_val, seed=0):
    """Write ``<root>/<config['dataset']>/{data_info.json, train/, val/}`` in the reference's layout.  ``data_info.json`` carries the
    keys ``get_config`` merges into the yaml (utils/io.py:33-36): labels, labels_small / mid / large, num_classes, bbox_properties."""
    base = Path(root) / config["dataset"]
    for split, n, s in (("train", n_train, seed), ("val", n_val, seed + 1)):
        cases = SyntheticCaseDataset(config, n, volume, s)
        for i in range(n):
            case_dir = base / split / f"case_{i:04d}"
            case_dir.mkdir(parents=True, exist_ok=True)
            data, label = cases[i]
            np.save(case_dir / "data.npy", data.numpy())
            np.save(case_dir / "label.npy", label.numpy())
    props = config["bbox_properties"]
    ids = [str(k) for k in props.keys()]
    info = {"labels": config.get("labels", {k: f"organ_{k}" for k in ids}), "labels_small": config.get("labels_small", []),
            "labels_mid": config.get("labels_mid", []), "labels_large": config.get("labels_large", ids),
            "num_classes": len(ids), "bbox_properties": props, "shape_statistics": {"median": list(volume)}}
    with open(base / "data_info.json", "w") as f:
        json.dump(info, f, indent=2)
    return base
