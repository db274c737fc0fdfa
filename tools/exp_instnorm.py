import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from transoar_b200.instnorm import instance_norm_relu
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dt in (torch.float32, torch.bfloat16):
    for shape in ((1, 24, 160, 160, 256), (1, 48, 80, 80, 128), (2, 24, 160, 160, 256)):
        x = torch.randn(*shape, device="cuda:0").to(dt).requires_grad_(True)
        w = torch.ones(shape[1], device="cuda:0", requires_grad=True); b = torch.zeros(shape[1], device="cuda:0", requires_grad=True)
        dy = torch.randn_like(x)
        es = x.element_size(); n = x.numel()
        def ours():
            y = instance_norm_relu(x, w, b); y.backward(dy); x.grad = None
        def ref():
            y = F.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5)); y.backward(dy); x.grad = None
        with torch.no_grad():
            tf, rf = timeit(lambda: instance_norm_relu(x, w, b)), timeit(lambda: F.relu(F.instance_norm(x, weight=w, bias=b, eps=1e-5)))
        tb, rb = timeit(ours), timeit(ref)
        print(f"{str(dt)[6:]:9s} {shape}: fwd ours {tf:.3f} ms ({3 * n * es / tf / 1e6:.0f} GB/s) torch {rf:.3f} ms | fwd+bwd ours {tb:.3f} ms ({8 * n * es / tb / 1e6:.0f} GB/s) torch {rb:.3f} ms", flush=True)
