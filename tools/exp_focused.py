"""Timing of the fused RoI attention vs the dense masked path at the VISCERAL shape (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.focused_attn_oracle import dense_masked_attention
from transoar_b200 import focused
DEV = "cuda:0"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gen = torch.Generator().manual_seed(0)
grid = (40, 40, 64)
props = {}
for o in range(20):   # SURVEY 8d synthetic atlas: centre U(.3,.7), size U(.1,.3), attn_area = hull +- margins
    c = torch.rand(3, generator=gen) * 0.4 + 0.3
    s = torch.rand(3, generator=gen) * 0.2 + 0.1
    props[str(o)] = {"attn_area": torch.cat(((c - s / 2 - 0.08).clamp(0, 1), (c + s / 2 + 0.08).clamp(0, 1))).tolist()}
boxes = focused.boxes_from_bbox_props(props, 540, grid)
vol = ((boxes[:, 3] - boxes[:, 0]) * (boxes[:, 4] - boxes[:, 1]) * (boxes[:, 5] - boxes[:, 2])).float()
print(f"unmasked KV fraction: {float(vol.mean()) / 102400:.3f}")
groups = focused.groups_from_boxes(boxes).to(DEV)
for B in (1, 2):
    q = (torch.randn(B, 540, 8, 48, generator=gen) * 0.3).to(DEV).requires_grad_(True)
    k = torch.randn(B, 102400, 8, 48, generator=gen).to(DEV).requires_grad_(True)
    v = torch.randn(B, 102400, 8, 48, generator=gen).to(DEV).requires_grad_(True)
    g = torch.randn(B, 540, 384, generator=gen).to(DEV)
    f = lambda: focused.RoIAttentionFunction.apply(q, k, v, groups, grid[1:])
    def fb():
        out = f(); out.backward(g); q.grad = k.grad = v.grad = None
    d = lambda: dense_masked_attention(q, k, v, boxes, grid)
    def db():
        out = d(); out.backward(g); q.grad = k.grad = v.grad = None
    with torch.no_grad():
        tf, td = timeit(f), timeit(d)
    print(f"B={B}: fused fwd {tf:.3f} ms  fwd+bwd {timeit(fb):.3f} ms | dense torch fwd {td:.3f} ms  fwd+bwd {timeit(db):.3f} ms", flush=True)
