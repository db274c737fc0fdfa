"""A/B: shipped forward (all corners through LDG / L1) vs the TMA-staged experiment (coarsest level gathered from a shared-memory slab
filled by cp.async.bulk; msda3d_set_tuning("stage", CTAs per SM)) on the refinement workload.  Prints one table (profiles/r02_experiments.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


lib = _lib.lib()
g = synth.GEOMETRIES["visceral_refine"]
for dist in ("B", "A"):
    x = synth.make_inputs(g, 2, dist, seed=1, device="cuda:0")
    f = lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
    base = f()
    row = [f"dist {dist}: shipped (LDG only) {timeit(f):.3f} ms"]
    for ctas in (2, 3, 4):
        lib.msda3d_set_tuning(b"stage", ctas)
        same = torch.equal(f(), base)
        row.append(f"staged coarsest level, {ctas} CTAs/SM {timeit(f):.3f} ms (bit-identical: {same})")
        lib.msda3d_set_tuning(b"stage", 0)
    print(" | ".join(row), flush=True)
