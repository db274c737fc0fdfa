"""bwd with / without shared-memory combining (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth
def timeit(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for gname, N in (("visceral_refine", 2), ("visceral_refine", 1), ("amos_refine", 2)):
    g = synth.GEOMETRIES[gname]
    for d in ("B", "B0", "A"):
        x = synth.make_inputs(g, N, d, seed=1234, device="cuda:0")
        b = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
        f = lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
        row = [f"fwd {timeit(f):6.3f}"]
        for comb in (0, 1):
            _lib.lib().msda3d_set_tuning(b"bwd_combine", comb)
            row.append(f"bwd(combine={comb}) {timeit(b):6.3f}")
        print(f"{gname:16s} N={N} dist {d:2s}: " + "  ".join(row), flush=True)
