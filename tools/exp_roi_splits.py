"""RoI attention (TF32 mma.sync kernels, the step's shape: B 2, 540 queries in 20 organ groups, 8 heads x 48, 40x40x64 tokens, the VISCERAL
atlas boxes of the model) for every token split count per box: msda3d_set_tuning("roi_splits", n); 0 = the library's own choice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import _lib, focused
from transoar_b200.engine import visceral_train_config
from transoar_b200.transoarnet import TransoarNet
DEV = "cuda:0"
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
net = TransoarNet(visceral_train_config())
fa = next(m for m in net.modules() if isinstance(m, focused.FocusedAttn))
groups, grid = fa.groups.to(DEV), fa.grid_shape
gen = torch.Generator().manual_seed(0)
B = 2
q = (torch.randn(B, 540, 8, 48, generator=gen) * 0.3).to(DEV).requires_grad_(True)
k = torch.randn(B, 102400, 8, 48, generator=gen).to(DEV).requires_grad_(True)
v = torch.randn(B, 102400, 8, 48, generator=gen).to(DEV).requires_grad_(True)
g = torch.randn(B, 540, 384, generator=gen).to(DEV)
torch.backends.cuda.matmul.allow_tf32 = True
f = lambda: focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:])
def fb():
    out = f(); out.backward(g)
    r = (out.detach(), q.grad.clone(), k.grad.clone(), v.grad.clone()); q.grad = k.grad = v.grad = None
    return r
base = None
for s in (0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 16):
    assert _lib.lib().msda3d_set_tuning(b"roi_splits", s) == 0
    with torch.no_grad():
        tf = timeit(f)
    tfb = timeit(fb)
    r = fb()
    if base is None: base = r
    dev = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(r, base))
    print(f"splits {s:2d}: fwd {tf:.3f} ms  fwd+bwd {tfb:.3f} ms  bwd {tfb - tf:.3f} ms   max rel dev from automatic {dev:.1e}", flush=True)
_lib.lib().msda3d_set_tuning(b"roi_splits", 0)
