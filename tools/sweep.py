"""BASELINE.json config 5: MSDeformAttn3D fwd / bwd microbenchmark sweep -> markdown table (gpurun_out/r02_sweep.md).

CUDA-event timing, 3 warm-up + 10 timed launches, median; algorithmic bytes per DESIGN.md section 6; fraction of the
measured HBM peak.  Levels are prefixes of the VISCERAL pyramid (40,40,64),(20,20,32),(10,10,16),(5,5,8); M=6, C=64."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from transoar_b200 import MultiScaleDeformableAttention as MSDA, synth

PYR = ((40, 40, 64), (20, 20, 32), (10, 10, 16), (5, 5, 8))
peak = bench.load_peaks()["hbm_gbs"]


def median_ms(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def run(N, L, P, queries, dtype, dist):
    g = synth.Geometry("sweep", PYR[:L], 6, 64, P, queries=queries)
    # one batch element is generated (host RNG) and repeated on the device: the sweep's largest case has 1.4 G samples (17 GB of
    # locations); every batch element still has its own value slab and its own location / weight / gradient memory
    x1 = synth.make_inputs(g, 1, dist, seed=1234, device="cuda:0", dtype=dtype)
    x = {k: (v if k in ("shapes", "starts") else v.repeat(N, *([1] * (v.dim() - 1))).contiguous()) for k, v in x1.items()}
    del x1
    f = lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
    b = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
    tf, tb = median_ms(f), median_ms(b)
    del f, b
    ev = 4 if dtype == torch.float32 else 2
    bf, bb = bench.algorithmic_bytes(N, g.spatial_size, 6, 64, L, g.num_query, P, ev=ev)
    del x
    torch.cuda.empty_cache()
    return tf, tb, bf / tf / 1e6, bb / tb / 1e6


cases = []
# the full batch x levels x points cross of BASELINE.json configs[4] (Lq = S, fp32, model-like sampling pattern)
for N in (1, 2, 4, 8, 16):
    for L in (1, 2, 3, 4):
        for P in (4, 8, 16, 32):
            cases.append((N, L, P, 0, torch.float32, "B"))
# bf16 value storage, decoder-like calls (Lq = 300), and the reference test's uniform locations (dist A) / no jitter (B0)
for N in (1, 2, 4, 8, 16):
    cases.append((N, 4, 4, 0, torch.bfloat16, "B"))
for N in (1, 2, 16):
    for L in (1, 4):
        for P in (4, 32):
            for dt in (torch.float32, torch.bfloat16):
                cases.append((N, L, P, 300, dt, "B"))
for N in (1, 2, 16):
    for P in (4, 32):
        cases.append((N, 4, P, 0, torch.float32, "A"))
for dt in (torch.float32, torch.bfloat16):
    cases.append((2, 4, 4, 0, dt, "A"))
    cases.append((2, 4, 4, 0, dt, "B0"))

out = ["# MSDeformAttn3D microbenchmark sweep on B200 (round 2: full batch x levels x points cross)", "",
       f"M=6, C=64, levels = first L of {PYR}; Lq = S unless stated; median of 10 launches, CUDA events; HBM peak {peak} GB/s (measured).",
       "GB/s = algorithmic (compulsory) bytes / time, DESIGN.md section 6; bwd includes the grad_value zero-fill.", "",
       "| N | L | P | Lq | value dtype | dist | fwd ms | bwd ms | fwd GB/s (frac) | bwd GB/s (frac) |", "|---|---|---|---|---|---|---|---|---|---|"]
for (N, L, P, q, dt, dist) in cases:
    tf, tb, gf, gb = run(N, L, P, q, dt, dist)
    S = sum(d * h * w for d, h, w in PYR[:L])
    line = f"| {N} | {L} | {P} | {q or S} | {str(dt).split('.')[-1]} | {dist} | {tf:.3f} | {tb:.3f} | {gf:.0f} ({gf / peak:.3f}) | {gb:.0f} ({gb / peak:.3f}) |"
    print(line, flush=True)
    out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r02_sweep.md"), "w").write("\n".join(out) + "\n")
