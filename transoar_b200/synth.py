"""Seeded synthetic inputs for the 3D multi-scale deformable attention path.

Workload geometry follows the reference configs:
  * VISCERAL 160x160x256, FPN levels P2..P5 -> (40,40,64),(20,20,32),(10,10,16),(5,5,8); 6 heads x 64 ch, 4 points
    (config/attn_fpn_foc_dec_visceral.yaml:73-84, backbones/attn_fpn.py:86-103)
  * AMOS 256x256x128, P3..P5 -> (32,32,16),(16,16,8),(8,8,4)
  * "Small"/"Tiny"/"Medium" are the presets of the reference's only test, transoar/models/ops/test.py:22-42.

Two sampling-location distributions:
  * dist "A": the reference test's own inputs (ops/test.py:54-57) -- uniform locations, worst-case locality.
  * dist "B0": dist "B" without the jitter -- exactly what the freshly initialised model produces.
  * dist "B": what the model feeds the op -- reference point = the query's voxel centre at every level
    (backbones/decoder_blocks.py:107-131) plus offsets ``(dir_m * (p+1) + N(0,1)) / (W_l,H_l,D_l)`` where ``dir_m`` are
    the six axis directions MSDeformAttn initialises its offset bias with (ops/modules/ms_deform_attn.py:63-79).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

import torch


@dataclass(frozen=True)
class Geometry:
    name: str
    shapes: Tuple[Tuple[int, int, int], ...]   # (D, H, W) per level
    heads: int
    channels: int                              # per head
    points: int
    queries: int = 0                           # 0 -> Lq = S (encoder-style self attention)

    @property
    def levels(self) -> int:
        return len(self.shapes)

    @property
    def spatial_size(self) -> int:
        return sum(d * h * w for d, h, w in self.shapes)

    @property
    def num_query(self) -> int:
        return self.queries or self.spatial_size


GEOMETRIES = {
    "visceral_refine": Geometry("visceral_refine", ((40, 40, 64), (20, 20, 32), (10, 10, 16), (5, 5, 8)), 6, 64, 4),
    "amos_refine": Geometry("amos_refine", ((32, 32, 16), (16, 16, 8), (8, 8, 4)), 6, 64, 4),
    "config1": Geometry("config1", ((32, 32, 32),), 4, 32, 4),
    "detr300": Geometry("detr300", ((40, 40, 64), (20, 20, 32), (10, 10, 16), (5, 5, 8)), 6, 64, 4, queries=300),
    "test_small": Geometry("test_small", ((3, 6, 4), (2, 3, 2)), 3, 4, 4, queries=4),
    "test_tiny": Geometry("test_tiny", ((2, 2, 2),), 1, 1, 1, queries=1),
    "test_medium": Geometry("test_medium", ((8, 15, 39), (4, 4, 10), (2, 2, 5)), 16, 16, 4, queries=4860),
}


def level_tensors(shapes: Sequence[Sequence[int]], device="cpu"):
    """(spatial_shapes int64 [L,3], level_start_index int64 [L]) as ops/test.py:44 builds them."""
    ss = torch.as_tensor(list(shapes), dtype=torch.long, device=device).reshape(-1, 3)
    starts = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    return ss, starts


def reference_points(shapes: Sequence[Sequence[int]], device="cpu") -> torch.Tensor:
    """Voxel centres of every level's voxels in (x,y,z) order, [S,3] (decoder_blocks.py:107-131, valid ratio 1)."""
    pts = []
    for d, h, w in shapes:
        z = (torch.arange(d, dtype=torch.float32, device=device) + 0.5) / d
        y = (torch.arange(h, dtype=torch.float32, device=device) + 0.5) / h
        x = (torch.arange(w, dtype=torch.float32, device=device) + 0.5) / w
        zz, yy, xx = torch.meshgrid(z, y, x, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)), -1))
    return torch.cat(pts, 0)


def head_directions(heads: int, device="cpu") -> torch.Tensor:
    """[heads,3] unit steps: the 6 / 26 neighbourhood directions of ms_deform_attn.py:66-73, else a fixed fan."""
    grid = torch.cartesian_prod(*(torch.tensor([-1.0, 0.0, 1.0]),) * 3)
    l1 = grid.abs().sum(1)
    if heads == 6:
        dirs = grid[l1 == 1]
    elif heads == 26:
        dirs = grid[l1 > 0]
    else:
        dirs = grid[l1 > 0][torch.arange(heads) % 26]
    return dirs.to(device)


def make_inputs(geom: Geometry, batch: int, dist: str = "A", seed: int = 1234, device="cpu",
                dtype: torch.dtype = torch.float32, with_grad_output: bool = True):
    """-> dict(value, shapes, starts, loc, aw, grad_out) on ``device``; loc/aw stay >= fp32."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    N, S, M, C, L, P, Lq = batch, geom.spatial_size, geom.heads, geom.channels, geom.levels, geom.points, geom.num_query
    aux_dtype = dtype if dtype in (torch.float32, torch.float64) else torch.float32

    def rnd(*shape, normal=False):
        f = torch.randn if normal else torch.rand
        return f(*shape, generator=g, dtype=torch.float32)

    value = rnd(N, S, M, C) * 0.01
    if dist == "A":
        loc = rnd(N, Lq, M, L, P, 3)
        aw = rnd(N, Lq, M, L, P) + 1e-5
        aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    elif dist in ("B", "B0"):
        ref = reference_points(geom.shapes)                                      # [S,3]
        if Lq != S:
            ref = ref[torch.randint(0, S, (Lq,), generator=g)]
        dirs = head_directions(M)                                                # [M,3]
        steps = torch.arange(1, P + 1, dtype=torch.float32)                      # (p+1)
        off = (dirs[None, None, :, None, None, :] * steps[None, None, None, None, :, None]).expand(N, Lq, M, L, P, 3)
        if dist == "B":                 # "B0" = the untrained model: offsets are exactly the bias pattern, no jitter
            off = off + rnd(N, Lq, M, L, P, 3, normal=True)
        norm = torch.tensor([[w, h, d] for d, h, w in geom.shapes], dtype=torch.float32)      # (W,H,D) per level
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
        aw = torch.softmax(rnd(N, Lq, M, L * P, normal=True), -1).reshape(N, Lq, M, L, P)
    else:
        raise ValueError(f"unknown distribution {dist!r}")
    shapes, starts = level_tensors(geom.shapes)
    out = {
        "value": value.to(dtype).to(device).contiguous(),
        "shapes": shapes.to(device),
        "starts": starts.to(device),
        "loc": loc.to(aux_dtype).to(device).contiguous(),
        "aw": aw.to(aux_dtype).to(device).contiguous(),
    }
    if with_grad_output:
        out["grad_out"] = (rnd(N, Lq, M * C, normal=True) * 0.1).to(dtype).to(device).contiguous()
    return out
