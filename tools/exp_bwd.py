"""Experiments: where does the time of the op go?  (diagnostic script, prints a table)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

base = synth.GEOMETRIES["visceral_refine"]
geoms = {
  "visceral L4": base,
  "4x fine (no coarse levels)": synth.Geometry("f4", ((40,40,64),)*4, 6, 64, 4, queries=117000),
  "L1 fine only P16": synth.Geometry("f1", ((40,40,64),), 6, 64, 16, queries=117000),
  "4x coarse (5,5,8)": synth.Geometry("c4", ((5,5,8),)*4, 6, 64, 4, queries=117000),
}
for name, g in geoms.items():
    for dist in ("A", "B"):
        x = synth.make_inputs(g, 2, dist, seed=1, device="cuda:0")
        f = lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
        b = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
        tf, tb = timeit(f), timeit(b)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 1)
        tbn = timeit(b)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 2)
        tbq = timeit(b)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 3)
        tb3 = timeit(b)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 4)
        tb4 = timeit(b)
        _lib.lib().msda3d_set_tuning(b"diag_bwd_skip_red", 0)
        print(f"{name:28s} dist {dist}: fwd {tf:7.3f} ms  bwd {tb:7.3f} ms  bwd(no RED) {tbn:7.3f} ms  bwd(1/4 RED) {tbq:7.3f} ms  "
              f"bwd(no RED coarsest level) {tb3:7.3f} ms  bwd(no RED two coarsest) {tb4:7.3f} ms", flush=True)
        del x
