"""Time the tcgen05 convolution kernels at the step's shape (2 x 24 x 160 x 160 x 256, channels-last) against cuDNN (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from transoar_b200.conv3d_tc import conv3d_k3
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
dev = "cuda:0"
x = torch.randn(2, 24, 160, 160, 256, device=dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
w = (torch.randn(24, 24, 3, 3, 3, device=dev) / 25).requires_grad_(True)
dy = torch.randn(2, 24, 160, 160, 256, device=dev).contiguous(memory_format=torch.channels_last_3d)
def ms(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    t_f = ms(lambda: conv3d_k3(x, w))
    t_c = ms(lambda: F.conv3d(x, w, None, 1, 1))
y = conv3d_k3(x, w)
t_b = ms(lambda: torch.autograd.grad(y, (x, w), dy, retain_graph=True))
yc = F.conv3d(x, w, None, 1, 1)
t_cb = ms(lambda: torch.autograd.grad(yc, (x, w), dy, retain_graph=True))
gb = 2 * 24 * 160 * 160 * 256 * 4 * 2 / 1e9
print(f"forward: ours {t_f:.3f} ms ({gb / t_f * 1e3:.0f} GB/s algorithmic, {2 * 27 * 24 * 24 * 2 * 160 * 160 * 256 / t_f / 1e9:.0f} TFLOP/s)  cuDNN {t_c:.3f} ms")
print(f"backward (dgrad + wgrad): ours {t_b:.3f} ms  cuDNN {t_cb:.3f} ms")

from transoar_b200 import _lib
for mode, what in ((1, "epilogue off"), (2, "MMAs off"), (4, "stores off"), (3, "only TMA + barriers")):
    _lib.lib().conv3d_tc_debug_mode(mode)
    with torch.no_grad():
        print(f"forward with {what}: {ms(lambda: conv3d_k3(x, w)):.3f} ms")
_lib.lib().conv3d_tc_debug_mode(0)
