"""Golden fixtures for the data side from the REFERENCE's ``segmentation2bbox`` (transoar/utils/bboxes.py:45-95).
Build-container only:   python tests/golden/make_golden_data.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def label_maps(seed, shape, organs, batch, background=0, thin=False):
    """Random box-shaped organs (later ids overwrite earlier ones), optionally one organ thinner than 5 voxels and one absent."""
    g = torch.Generator().manual_seed(seed)
    maps = torch.full((batch, 1) + shape, background, dtype=torch.int64)
    size = torch.tensor(shape)
    for b in range(batch):
        for organ in range(1, organs + 1):
            if (organ + b) % 7 == 3:
                continue                                            # absent from this sample
            lo = (torch.rand(3, generator=g) * (size - 12).float()).long()
            ext = (torch.rand(3, generator=g) * 10).long() + 6
            if thin and organ == 2:
                ext[1] = 3
            hi = torch.minimum(lo + ext, size)
            maps[b, 0, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = organ + background
    return maps


CASES = {
    "plain": dict(seed=1, shape=(40, 36, 48), organs=6, batch=2),
    "thin_and_padding3": dict(seed=2, shape=(32, 32, 32), organs=9, batch=3, thin=True),
    "background_is_not_zero": dict(seed=3, shape=(24, 40, 28), organs=4, batch=1, background=5),
    "touching_the_border": dict(seed=4, shape=(20, 20, 20), organs=12, batch=2),
}
PADDING = {"plain": 1, "thin_and_padding3": 3, "background_is_not_zero": 1, "touching_the_border": 2}


def main():
    sys.path.insert(0, "/root/reference")
    from transoar.utils.bboxes import segmentation2bbox
    blob = {}
    for name, spec in CASES.items():
        maps = label_maps(**spec)
        blob[f"{name}.maps"] = maps.numpy().astype(np.int16)
        for fmt in ("cxcyczwhd", "xyzxyz", "xyxyzz"):
            for normalize in (True, False):
                boxes, classes = segmentation2bbox(maps, PADDING[name], fmt, normalize)
                for b, (bx, cl) in enumerate(zip(boxes, classes)):
                    blob[f"{name}.{fmt}.{int(normalize)}.{b}.boxes"] = bx.numpy()
                    blob[f"{name}.{fmt}.{int(normalize)}.{b}.classes"] = cl.numpy()
    empty = torch.zeros(1, 1, 16, 16, 16, dtype=torch.int64)
    boxes, classes = segmentation2bbox(empty, 1)
    blob["empty.boxes"], blob["empty.classes"] = boxes[0].numpy(), classes[0].numpy()
    np.savez_compressed(os.path.join(HERE, "data.npz"), **blob)
    print("wrote data.npz with", len(blob), "arrays")


if __name__ == "__main__":
    main()
