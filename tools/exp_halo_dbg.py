"""Where does the halo kernel's time go?  Forward of one layer with parts of the kernel switched off (results are wrong in those modes)."""
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
for name, ci, co, s, (D, H, W) in [("enc1.conv2", 48, 48, 1, (80, 80, 128)), ("enc2.conv2", 96, 96, 1, (40, 40, 64)), ("out.P2", 96, 384, 1, (40, 40, 64))]:
    N = 2
    x = cl(torch.randn(N, ci, D, H, W, device=DEV)); wt = torch.randn(27, co, ci, device=DEV)
    y = cl(torch.empty(N, co, D, H, W, device=DEV))
    for label, dbg in [("full", 0), ("no MMA", 1), ("no A loads", 2), ("no B loads", 4), ("no loads", 6), ("no loads, no MMA", 7), ("no stores", 8), ("only MMA (no loads, no stores)", 14), ("nothing", 15)]:
        lib.conv3d_gen_set_path(2 | (dbg << 3))
        f = lambda: lib.conv3d_gen_forward(None, p(x), p(wt), None, N, D, H, W, ci, co, s, p(y))
        assert f() == 0
        print(f"{name:11s} {label:32s} {timed(f):7.3f} ms", flush=True)
    lib.conv3d_gen_set_path(0)
