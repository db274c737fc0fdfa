"""Per-layer timing of the general tcgen05 convolution (include/conv3d_gen.h) against cuDNN (autotuned, TF32) on the AttnFPN layers of the
VISCERAL step (batch 2, 160x160x256): forward, input gradient, weight gradient.  CUDA events, median of `reps`."""
import ctypes
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from transoar_b200 import _lib  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
DEV = "cuda:0"
# name, CI, CO, stride, input (D, H, W), bias
LAYERS = [("enc1.conv1", 24, 48, 2, (160, 160, 256), False), ("enc1.conv2", 48, 48, 1, (80, 80, 128), False),
          ("enc2.conv1", 48, 96, 2, (80, 80, 128), False), ("enc2.conv2", 96, 96, 1, (40, 40, 64), False),
          ("enc3.conv1", 96, 192, 2, (40, 40, 64), False), ("enc3.conv2", 192, 192, 1, (20, 20, 32), False),
          ("enc4.conv1", 192, 384, 2, (20, 20, 32), False), ("enc4.conv2", 384, 384, 1, (10, 10, 16), False),
          ("enc5.conv1", 384, 768, 2, (10, 10, 16), False), ("enc5.conv2", 768, 768, 1, (5, 5, 8), False),
          ("out.P2", 96, 384, 1, (40, 40, 64), True), ("out.P3", 192, 384, 1, (20, 20, 32), True),
          ("out.P4", 384, 384, 1, (10, 10, 16), True), ("out.P5", 384, 384, 1, (5, 5, 8), True)]


def timed(fn, reps=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    only = [a for a in sys.argv[1:] if not a.startswith("--")] or None
    lib = _lib.lib()
    for a in sys.argv[1:]:
        if a.startswith("--path="):
            lib.conv3d_gen_set_path({"auto": 0, "tap": 1, "halo": 2, "halo1": 6}[a.split("=")[1]])
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last_3d)
    rows = []
    for name, ci, co, s, (D, H, W), bias in LAYERS:
        if only and not any(o in name for o in only):
            continue
        N = 2
        x = cl(torch.randn(N, ci, D, H, W, device=DEV))
        w = cl(torch.randn(co, ci, 3, 3, 3, device=DEV) / (27 * ci) ** 0.5)
        wt = w.permute(2, 3, 4, 0, 1).reshape(27, co, ci).contiguous()
        b = torch.randn(co, device=DEV) if bias else None
        od, oh, ow = ((v + s - 1) // s for v in (D, H, W))
        y = cl(torch.empty(N, co, od, oh, ow, device=DEV))
        dy = cl(torch.randn(N, co, od, oh, ow, device=DEV))
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        ours = {"fwd": lambda: lib.conv3d_gen_forward(st(), p(x), p(wt), p(b), N, D, H, W, ci, co, s, p(y)),
                "dgrad": lambda: lib.conv3d_gen_dgrad(st(), p(dy), p(wt), N, D, H, W, ci, co, s, p(dx)),
                "wgrad": lambda: lib.conv3d_gen_wgrad(st(), p(x), p(dy), N, D, H, W, ci, co, s, p(dw))}
        if s == 2 and ci <= 64:
            from transoar_b200.conv3d_gen import fold_stride2_weights
            wf = fold_stride2_weights(wt)
            ours["dgrad"] = lambda: lib.conv3d_gen_dgrad_s2_folded(st(), p(dy), p(wf), N, D, H, W, ci, co, p(dx))
        cb = lambda mask: torch.ops.aten.convolution_backward(dy, x, w, None, [s] * 3, [1] * 3, [1] * 3, False, [0] * 3, 1, mask)
        theirs = {"fwd": lambda: F.conv3d(x, w, b, s, 1), "dgrad": lambda: cb([True, False, False]), "wgrad": lambda: cb([False, True, False])}
        gf = 2.0 * N * od * oh * ow * 27 * ci * co / 1e9
        row = {"layer": name, "ci": ci, "co": co, "stride": s, "gflop": round(gf, 1)}
        for k in ("fwd", "dgrad", "wgrad"):
            assert ours[k]() == 0
            t_o, t_c = timed(ours[k]), timed(theirs[k])
            row[k] = {"ours_ms": round(t_o, 3), "cudnn_ms": round(t_c, 3), "ours_tflops": round(gf / t_o, 1)}
        # numerics of this very call (random data, TF32): against cuDNN's result
        yr = F.conv3d(x, w, b, s, 1)
        row["fwd_rel_err_vs_cudnn"] = float((y - yr).abs().max() / yr.abs().max())
        rows.append(row)
        print(json.dumps(row), flush=True)
    tot = {k: (sum(r[k]["ours_ms"] for r in rows), sum(r[k]["cudnn_ms"] for r in rows)) for k in ("fwd", "dgrad", "wgrad")}
    print(json.dumps({"total_ms": {k: {"ours": round(a, 3), "cudnn": round(b, 3)} for k, (a, b) in tot.items()}}))


if __name__ == "__main__":
    main()
