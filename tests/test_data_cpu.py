"""Data side and run formats: box extraction against fixtures made with the reference's ``segmentation2bbox``
(tests/golden/make_golden_data.py), the npy case layout, the collator, rank sharding, the checkpoint / config formats."""
import json
import os

import numpy as np
import pytest
import torch

from transoar_b200 import data as D
from transoar_b200 import train as T
from transoar_b200.configs import visceral_config

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "data.npz"))
PADDING = {"plain": 1, "thin_and_padding3": 3, "background_is_not_zero": 1, "touching_the_border": 2}


@pytest.mark.parametrize("name", sorted(PADDING))
@pytest.mark.parametrize("fmt", ["cxcyczwhd", "xyzxyz", "xyxyzz"])
@pytest.mark.parametrize("normalize", [True, False])
def test_segmentation2bbox_equals_the_reference(name, fmt, normalize):
    maps = torch.from_numpy(GOLD[f"{name}.maps"].astype(np.int64))
    boxes, classes = D.segmentation2bbox(maps, PADDING[name], fmt, normalize)
    assert len(boxes) == maps.shape[0]
    for b in range(maps.shape[0]):
        want_b, want_c = GOLD[f"{name}.{fmt}.{int(normalize)}.{b}.boxes"], GOLD[f"{name}.{fmt}.{int(normalize)}.{b}.classes"]
        assert classes[b].dtype == torch.int64 and np.array_equal(classes[b].numpy(), want_c)
        assert boxes[b].dtype == torch.float32 and np.array_equal(boxes[b].numpy(), want_b)          # bit-equal


def test_segmentation2bbox_edge_cases():
    boxes, classes = D.segmentation2bbox(torch.zeros(1, 1, 16, 16, 16, dtype=torch.int64), 1)
    assert boxes[0].shape == GOLD["empty.boxes"].shape == (0,) and classes[0].numel() == 0
    with pytest.raises(ValueError):
        D.segmentation2bbox(torch.zeros(1, 1, 8, 8, 8, dtype=torch.int64), 1, box_format="zyx")
    with pytest.raises(AssertionError):
        D.segmentation2bbox(torch.zeros(1, 8, 8, 8, dtype=torch.int64), 1)                            # maps must be [B, 1, X, Y, Z]
    assert D.detection_targets(list(zip(boxes, classes)), "cpu")[0]["boxes"].shape == (0, 6)


def _config(**over):
    cfg = visceral_config()
    cfg.update(dataset="synthetic_ct", bbox_padding=1, batch_size=2, shuffle=True, num_workers=0, overfit=False, seed=3,
               augmentation={"use_augmentation": False})
    cfg.update(over)
    return cfg


def test_synthetic_dataset_round_trip_in_the_reference_layout(tmp_path):
    cfg = _config()
    volume = (32, 32, 48)
    base = D.write_synthetic_dataset(tmp_path, cfg, volume, n_train=3, n_val=2, seed=5)
    info = json.load(open(base / "data_info.json"))
    assert {"labels", "labels_small", "labels_mid", "labels_large", "num_classes", "bbox_properties"} <= set(info)
    assert sorted(p.name for p in (base / "train" / "case_0001").iterdir()) == ["data.npy", "label.npy"]
    ds = D.NpyCaseDataset(cfg, "train", root=tmp_path)
    ref = D.SyntheticCaseDataset(cfg, 3, volume, seed=5)
    assert len(ds) == 3 and len(D.NpyCaseDataset(cfg, "val", root=tmp_path)) == 2
    for i in range(3):
        data, label = ds[i]
        assert data.shape == (1,) + volume and data.dtype == torch.float32 and label.shape == (1,) + volume
        assert torch.equal(data, ref[i][0]) and torch.equal(label.long(), ref[i][1].long())
    assert torch.equal(D.NpyCaseDataset(_config(overfit=True), "train", root=tmp_path)[2][0], ds[0][0])      # dataset.py:27-28
    # a transform stands where the reference applies its augmentation
    flipped = D.NpyCaseDataset(cfg, "train", root=tmp_path, transform=lambda d, l, i: (d.flip(1), l.flip(1)))[0]
    assert torch.equal(flipped[0], ds[0][0].flip(1))
    with pytest.warns(UserWarning):
        D.NpyCaseDataset(_config(augmentation={"use_augmentation": True}), "train", root=tmp_path)


def test_collator_builds_what_the_trainer_consumes():
    cfg = _config()
    volume = (64, 64, 96)
    cases = D.SyntheticCaseDataset(cfg, 4, volume, seed=1)
    images, masks, bboxes, labels = D.Collator(cfg)([cases[0], cases[1]])
    assert images.shape == (2, 1) + volume and masks.shape == images.shape and not masks.any() and labels.shape == images.shape
    assert len(bboxes) == 2
    for (boxes, classes), label in zip(bboxes, labels):
        assert boxes.shape == (len(classes), 6) and boxes.min() >= 0 and boxes.max() <= 1
        assert set(classes.tolist()) <= set(range(1, 21)) and len(classes) >= 10
        lo, hi = boxes[:, :3] - boxes[:, 3:] / 2, boxes[:, :3] + boxes[:, 3:] / 2
        size = torch.tensor(volume, dtype=torch.float32)
        for c, l, h in zip(classes.tolist(), lo * size, hi * size):             # the box encloses every voxel of its organ
            idx = (label[0] == c).nonzero().float()
            assert (idx.min(0)[0] >= l - 1e-3).all() and (idx.max(0)[0] <= h + 1e-3).all()
    targets = D.detection_targets(bboxes, "cpu")
    assert targets[0]["boxes"].dtype == torch.float32 and targets[0]["labels"].dtype == torch.int64


def test_strided_sampler_shards_are_disjoint_and_equal():
    shards = [list(D.StridedSampler(11, rank=r, world=3, shuffle=True, seed=2)) for r in range(3)]
    assert all(len(s) == 3 for s in shards) and len(set(sum(shards, []))) == 9
    a, b = D.StridedSampler(11, 0, 3, True, 2), D.StridedSampler(11, 0, 3, True, 2)
    b.set_epoch(1)
    assert list(a) != list(b)
    assert list(D.StridedSampler(6, 1, 2, shuffle=False)) == [1, 3, 5]
    loader = D.get_loader(_config(), "train", rank=1, world=2, dataset=D.SyntheticCaseDataset(_config(), 9, (32, 32, 32)))
    assert len(loader) == 2                                                      # 9 // 2 = 4 cases on this rank, batches of 2


def test_checkpoint_and_config_formats(tmp_path):
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    groups = [{"params": net[0].parameters()}, {"params": net[1].parameters(), "lr": 2e-4}]
    optim = torch.optim.AdamW(groups, lr=2e-5, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(optim, 2500)
    net(torch.randn(5, 4)).sum().backward(); optim.step(); sched.step()
    T.save_checkpoint(tmp_path / "model_last.pt", 7, 0.25, net, optim, sched)
    blob = torch.load(tmp_path / "model_last.pt", weights_only=False)
    assert list(blob) == ["epoch", "metric_max_val", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict"]   # trainer.py:235-241
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    optim2 = torch.optim.AdamW([{"params": net2[0].parameters()}, {"params": net2[1].parameters(), "lr": 2e-4}], lr=2e-5, weight_decay=1e-4)
    sched2 = torch.optim.lr_scheduler.StepLR(optim2, 2500)
    assert T.load_checkpoint(tmp_path / "model_last.pt", net2, optim2, sched2, lr_drop=40) == (7, 0.25)
    assert sched2.step_size == 40 and sched2.last_epoch == 1                      # train.py:68: lr_drop of the current config wins
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    assert optim2.state_dict()["state"][0]["exp_avg"].equal(optim.state_dict()["state"][0]["exp_avg"])


def test_load_config_merges_yaml_and_data_info(tmp_path):
    (tmp_path / "config").mkdir(); (tmp_path / "dataset" / "toy").mkdir(parents=True)
    (tmp_path / "config" / "run.yaml").write_text("experiment_name: toy_run\ndataset: toy\nlr: 2e-4\nbackbone:\n  use_cuda: true\n")
    json.dump({"num_classes": 3, "labels": {"1": "liver"}}, open(tmp_path / "dataset" / "toy" / "data_info.json", "w"))
    cfg = T.load_config("run", tmp_path / "config", tmp_path / "dataset")
    assert cfg["experiment_name"] == "toy_run" and cfg["num_classes"] == 3 and cfg["backbone"]["use_cuda"] is True
    assert T.load_config(str(tmp_path / "config" / "run.yaml"), "/nonexistent", tmp_path / "dataset")["labels"] == {"1": "liver"}
    builtin = T.load_config("visceral", tmp_path / "config")
    assert builtin["neck"]["num_queries"] == 540 and builtin["experiment_name"] == "foc_dec_visceral"
    with pytest.raises(FileNotFoundError):
        T.load_config("nope", tmp_path / "config")
    assert json.dumps(T.to_jsonable({"a": (1, 2), 3: np.float32(1.5), "t": torch.ones(2)}))
