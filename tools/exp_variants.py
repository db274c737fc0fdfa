"""Time fwd / bwd of the bench workload for library build variants and tuning knobs (diagnostic)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth
    tag = sys.argv[2]
    def timeit(fn, n=6):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    g = synth.GEOMETRIES["visceral_refine"]
    xs = {d: synth.make_inputs(g, 2, d, seed=1234, device="cuda:0") for d in ("B", "B0", "A")}
    for nv in (1, 2):
        for gm, order in ((0, 1), (0, 2), (4, 2), (64, 2)):
            _lib.lib().msda3d_set_tuning(b"nv", nv); _lib.lib().msda3d_set_tuning(b"grid_mult", gm); _lib.lib().msda3d_set_tuning(b"order", order)
            row = []
            for d, x in xs.items():
                f = lambda: MSDA.ms_deform_attn_forward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], 64)
                b = lambda: MSDA.ms_deform_attn_backward(x["value"], x["shapes"], x["starts"], x["loc"], x["aw"], x["grad_out"], 64)
                row.append(f"{d}: fwd {timeit(f):6.3f} bwd {timeit(b):6.3f}")
            print(f"{tag:8s} nv={nv} order={order} grid_mult={gm:2d}  " + "   ".join(row), flush=True)
else:
    for tag in ["default", "m3", "m4"]:
        env = dict(os.environ)
        if tag != "default":
            env["MSDA3D_LIB"] = os.path.join(ROOT, "variants", f"libmsda3d_{tag}.so")
        subprocess.run([sys.executable, __file__, "child", tag], env=env)
