"""Where do the three warp roles of the tcgen05 GEMM wait?  Uses the tc_gemm_debug_profile hook (cycle counters per CTA)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import _lib
from transoar_b200.linear import gemm
dev = "cuda:0"
T = 234000
buf = torch.zeros(12 * 296, dtype=torch.int64, device=dev)
lib = _lib.lib()
for pair in ("1", "0"):
    os.environ["TC_GEMM_PAIR"] = pair
for K, N in ((384, 384), (384, 1024), (1024, 384)):
    x = torch.randn(T, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    dy = torch.randn(T, N, device=dev); y = torch.empty(T, N, device=dev); dx = torch.empty(T, K, device=dev)
    for name, fn in (("fwd", lambda: gemm(x, 0, K, w, 0, K, y, T, N, K, bias=b)), ("dX", lambda: gemm(dy, 0, N, w, 1, K, dx, T, K, N))):
        fn(); torch.cuda.synchronize()
        buf.zero_()
        lib.tc_gemm_debug_profile(ctypes.c_void_p(buf.data_ptr()))
        fn(); torch.cuda.synchronize()
        lib.tc_gemm_debug_profile(None)
        ph = buf[8 * 296:].view(296, 4).double()
        ph = ph[ph.sum(1) > 0].mean(0)
        c = buf[:8 * 296].view(296, 8).double()
        c = c[c[:, 4] + c[:, 6] > 0]
        m = c.mean(0)
        mma_rows = c[c[:, 4] > 0].mean(0)
        print(f"{name} K={K} N={N}: CTAs {c.shape[0]}  producer wait(empty) {m[0]/max(m[1],1):.2f} of {m[1]/1e3:.0f} kclk | "
              f"MMA wait(full) {mma_rows[2]/mma_rows[4]:.2f} wait(tmem empty) {mma_rows[3]/mma_rows[4]:.2f} of {mma_rows[4]/1e3:.0f} kclk | "
              f"epilogue wait(tmem full) {m[5]/max(m[6],1):.2f} of {m[6]/1e3:.0f} kclk "
              f"[tmem-ld {ph[0]/1e3:.0f} stage {ph[1]/1e3:.0f} store {ph[2]/1e3:.0f} kclk]", flush=True)
