"""Capture single native ops at model-like sizes (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = True
from transoar_b200.linear import TCLinear, ffn, gemm, colsum
from transoar_b200.fused_ln import add_dropout_layer_norm
dev = "cuda:0"
stream = torch.cuda.Stream()
def attempt(tag, fn):
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.no_grad(), torch.cuda.graph(g, stream=stream):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("ok     ", tag, flush=True)
    except Exception as e:
        print("BROKEN ", tag, str(e).splitlines()[0][:100], flush=True)
for M in (256, 4680, 37440):
    x = torch.randn(M, 384, device=dev)
    l1, l2 = TCLinear(384, 1024).to(dev), TCLinear(1024, 384).to(dev)
    lin = TCLinear(384, 384).to(dev)
    norm = torch.nn.LayerNorm(384).to(dev)
    attempt(f"linear M={M}", lambda: lin(x))
    attempt(f"ffn M={M}", lambda: ffn(x, l1, l2, 0.1, True))
    attempt(f"fused LN M={M}", lambda: add_dropout_layer_norm(x, x, norm, 0.1, True))
    w = torch.randn(384, 384, device=dev); d = torch.zeros(384, 384, device=dev)
    attempt(f"wgrad gemm (MN-major, split-K) M={M}", lambda: gemm(x, 1, 384, x, 1, 384, d, 384, 384, M, accumulate=True, split_k=0))
    attempt(f"colsum M={M}", lambda: colsum(x))
