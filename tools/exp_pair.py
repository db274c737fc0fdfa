"""Pair-combining backward (msda3d_set_tuning 'pair') on/off: op-level timing on the refinement workload for the three location
distributions, and max gradient difference between the two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA
from transoar_b200 import _lib, synth
g = synth.GEOMETRIES["visceral_refine"]
lib = _lib.lib()
def ms(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for dist in ("B0", "B", "A"):
    x = synth.make_inputs(g, 2, dist, seed=1234, device="cuda:0")
    bwd = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
    res = {}
    for pair in (0, 1):
        assert lib.msda3d_set_tuning(b"pair", pair) == 0
        res[pair] = (ms(bwd), [t.clone() for t in bwd()])
    lib.msda3d_set_tuning(b"pair", 0)
    diff = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(res[0][1], res[1][1]))
    print(f"dist {dist}: backward {res[0][0]:.3f} ms direct -> {res[1][0]:.3f} ms pair-combining   max rel diff {diff:.2e}", flush=True)
