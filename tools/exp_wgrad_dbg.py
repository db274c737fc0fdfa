"""Where does the weight-gradient kernel's time go?  One layer with parts of the kernel switched off (results are wrong in those modes)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from transoar_b200 import _lib
lib = _lib.lib()
DEV = "cuda:0"
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
cl = lambda t: t.contiguous(memory_format=torch.channels_last_3d)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for name, ci, co, s, (D, H, W) in [("enc1.conv2", 48, 48, 1, (80, 80, 128)), ("enc1.conv1", 24, 48, 2, (160, 160, 256)), ("enc2.conv1", 48, 96, 2, (80, 80, 128)), ("out.P2", 96, 384, 1, (40, 40, 64))]:
    N = 2
    od, oh, ow = ((v + s - 1) // s for v in (D, H, W))
    x = cl(torch.randn(N, ci, D, H, W, device=DEV)); dy = cl(torch.randn(N, co, od, oh, ow, device=DEV))
    dw = torch.empty(co, 27, ci, device=DEV)
    for label, dbg in [("full", 0), ("no MMA", 1), ("no loads", 2), ("no loads, no MMA", 3), ("no reductions", 4), ("nothing", 7)]:
        lib.conv3d_gen_set_path(dbg << 3)
        f = lambda: lib.conv3d_gen_wgrad(None, p(x), p(dy), N, D, H, W, ci, co, s, p(dw))
        assert f() == 0
        print(f"{name:11s} {label:20s} {timed(f):7.3f} ms", flush=True)
    lib.conv3d_gen_set_path(0)
