"""TF32 tcgen05 GEMM vs cuBLAS (allow_tf32) on the hot path's Linear shapes (VISCERAL refine: 2 x 117000 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.linear import gemm
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda:0"
T = int(sys.argv[1]) if len(sys.argv) > 1 else 234000


def ms(fn, warm=3, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"{'case':34s} {'ours ms':>8s} {'TF/s':>7s} {'cublas ms':>9s} {'TF/s':>7s}")
for K, N in ((384, 384), (384, 1024), (1024, 384), (384, 288), (384, 96)):
    x = torch.randn(T, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    dy = torch.randn(T, N, device=dev)
    y = torch.empty(T, N, device=dev); dx = torch.empty(T, K, device=dev); dw = torch.zeros(N, K, device=dev)
    fl = 2.0 * T * K * N / 1e12
    cases = [
        (f"fwd  [{T}x{K}]x[{N}x{K}]^T+b", lambda: gemm(x, 0, K, w, 0, K, y, T, N, K, bias=b), lambda: torch.addmm(b, x, w.t(), out=y)),
        (f"dX   [{T}x{N}]x[{N}x{K}]", lambda: gemm(dy, 0, N, w, 1, K, dx, T, K, N), lambda: torch.mm(dy, w, out=dx)),
        (f"dW   [{N}x{T}]x[{T}x{K}]", lambda: gemm(dy, 1, N, x, 1, K, dw, N, K, T, accumulate=True, split_k=0), lambda: torch.mm(dy.t(), x, out=dw)),
    ]
    for name, ours, ref in cases:
        a, c = ms(ours), ms(ref)
        print(f"{name:34s} {a:8.3f} {fl / a * 1e3:7.1f} {c:9.3f} {fl / c * 1e3:7.1f}", flush=True)
