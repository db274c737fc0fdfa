"""Experiment: the backward's sample-order rotation (msda3d_set_tuning("rot", 0 | 1 | 2)) on the fused (merged-projection) route, with
the sampling pattern the MODEL produces: offsets = the module's directional bias (+ jitter of sigma voxels: 0 = untrained, 0.03 = a
few optimizer steps in, 1 = synth dist B), logits ~ N(0, 0.1).  Prints ms per launch and the max relative deviation from rot = 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth
from transoar_b200.ops.modules import MSDeformAttn

def timeit(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

dev = "cuda:0"
g = synth.GEOMETRIES["visceral_refine"]
N, S, M, C, L, P = 2, g.spatial_size, g.heads, g.channels, g.levels, g.points
mod = MSDeformAttn(M * C, L, M, P, True)
bias = mod.sampling_offsets.bias.detach().view(1, 1, M * L * P * 3)
gen = torch.Generator().manual_seed(7)
shapes, starts = synth.level_tensors(g.shapes, dev)
ref = synth.reference_points(g.shapes)[None, :, None, :].expand(1, S, L, 3).contiguous().to(dev)
value = (torch.rand(N, S, M, C, generator=gen) * 0.01).to(dev)
gout = (torch.randn(N, S, M * C, generator=gen) * 0.1).to(dev)
for sigma in [float(v) for v in os.environ.get('SIGMAS', '0.03,1.0').split(',')]:
    off = bias + sigma * torch.randn(N, S, M * L * P * 3, generator=gen)
    logit = 0.1 * torch.randn(N, S, M * L * P, generator=gen)
    merged = torch.cat((off, logit), -1).to(dev).contiguous()
    base = None
    for rot in [int(v) for v in os.environ.get('VARIANTS', '0,6,1000,1004,2000').split(',')]:
        _lib.lib().msda3d_set_tuning(b"duo", 1 if rot >= 1000 else 0)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 1 if rot % 10000 >= 2000 else 0)
        _lib.lib().msda3d_set_tuning(b"duo_cfg", rot // 10000)
        _lib.lib().msda3d_set_tuning(b"rot", rot % 100)
        _lib.lib().msda3d_set_tuning(b"grid_mult", int(os.environ.get("GRID_MULT", "0")))
        f = lambda: MSDA.ms_deform_attn_backward_merged(value, shapes, starts, ref, merged, gout, L, P)
        gv, gm = f()
        t = timeit(f)
        if base is None:
            base = (gv, gm); dev_v = dev_m = 0.0
        else:
            dev_v = float((gv - base[0]).abs().max() / base[0].abs().max()); dev_m = float((gm - base[1]).abs().max() / base[1].abs().max())
        print(f"jitter sigma {sigma:4.2f} voxels  rot {rot}: bwd {t:7.3f} ms   max rel dev grad_value {dev_v:.2e} grad_merged {dev_m:.2e}", flush=True)
    _lib.lib().msda3d_set_tuning(b"rot", 0); _lib.lib().msda3d_set_tuning(b"duo", 0); _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 0); _lib.lib().msda3d_set_tuning(b"duo_cfg", 0)
x = synth.make_inputs(g, 2, "B", seed=1, device=dev)
for rot in (0, 1000):
    _lib.lib().msda3d_set_tuning(b"duo", 1 if rot >= 1000 else 0)
    _lib.lib().msda3d_set_tuning(b"rot", rot % 100)
    b = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
    print(f"unfused op, synth dist B  rot {rot}: bwd {timeit(b):7.3f} ms", flush=True)
    gvb, glb, gab = b()
    if rot == 0: base = (gvb, glb, gab)
    else: print("   max rel dev vs rot 0:", [float((a - c).abs().max() / c.abs().max()) for a, c in zip((gvb, glb, gab), base)], flush=True)
_lib.lib().msda3d_set_tuning(b"rot", 0); _lib.lib().msda3d_set_tuning(b"duo", 0)
