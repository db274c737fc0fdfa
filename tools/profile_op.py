"""Tiny driver for ncu: a few forward+backward launches of the op on the bench workload (no timing here)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from transoar_b200 import MultiScaleDeformableAttention as MSDA
from transoar_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--geom", default="visceral_refine")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--dist", default="B")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--dtype", default="float32")
ap.add_argument("--nv", type=int, default=0)
ap.add_argument("--order", type=int, default=0)
ap.add_argument("--grid-mult", type=int, default=0)
a = ap.parse_args()
from transoar_b200 import _lib
for k, v in (("nv", a.nv), ("order", a.order), ("grid_mult", a.grid_mult)):
    assert _lib.lib().msda3d_set_tuning(k.encode(), v) == 0, k
x = synth.make_inputs(synth.GEOMETRIES[a.geom], a.batch, a.dist, seed=1234, device="cuda:0", dtype=getattr(torch, a.dtype))
for _ in range(a.iters):
    MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
    MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
torch.cuda.synchronize()
print("done")
