import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200.linear import gemm
torch.manual_seed(0)
M, K, Hd, N = 3000, 384, 1024, 384
x = torch.randn(M, K).cuda(); w1 = (torch.randn(Hd, K) / math.sqrt(K)).cuda(); b1 = (torch.randn(Hd) * 0.1).cuda()
w2 = (torch.randn(N, Hd) / math.sqrt(Hd)).cuda(); dy = torch.randn(M, N).cuda()
def rep(name, got, want):
    e = (got.double() - want).abs()
    i = int(e.argmax()); r, c = divmod(i, got.shape[1])
    print(f"{name}: max err {float(e.max()):.3e} at ({r},{c}) of max {float(want.abs().max()):.3f}; rows with err>1e-2: {sorted(set((e > 1e-2).nonzero()[:,0].tolist()))[:12]}", flush=True)
for p in (0.0, 0.1):
    h = torch.empty(M, Hd, device="cuda")
    gemm(x, 0, K, w1, 0, K, h, M, Hd, K, bias=b1, relu=True, p_drop=p, seed=77)
    pre = torch.relu(x.double() @ w1.double().t() + b1.double())
    keep = ((h > 0) | (pre <= 1e-4)).double()
    rep(f"p={p} h", h, pre * keep / (1 - p))
    dh = torch.empty(M, Hd, device="cuda")
    gemm(dy, 0, N, w2, 1, Hd, dh, M, Hd, N, gate=h, gate_scale=1 / (1 - p))
    dh_ref = (dy.double() @ w2.double()) * (h > 0).double() / (1 - p)
    rep(f"p={p} dh", dh, dh_ref)
    dx = torch.empty(M, K, device="cuda")
    gemm(dh, 0, Hd, w1, 1, K, dx, M, K, Hd)
    rep(f"p={p} dx", dx, dh.double() @ w1.double())
    dx2 = torch.empty(M, K, device="cuda")
    dhc = dh.clone()
    gemm(dhc, 0, Hd, w1, 1, K, dx2, M, K, Hd)
    print("  dx repeat equal:", bool(torch.equal(dx, dx2)), " cublas-ref err", float((dx.double() - (dh @ w1).double()).abs().max()))
