"""Clocks per tcgen05.mma.kind::tf32 (128 x N x 8, issued back to back by one thread) for the three shared-memory operand layouts."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from transoar_b200 import _lib  # noqa: E402

lib = _lib.lib()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
names = {0: "K-major SW128", 1: "K-major SW32 ", 2: "MN-major     "}
for layout, rotate in ((0, 0), (1, 0), (2, 0), (0, 4), (2, 4)):
    for n in (32, 64, 96, 128, 192, 256):
        if rotate and n > 128:
            continue                                   # (the probe's 32 KB operand areas hold four tiles only up to N = 128)
        res = []
        for iters in (64, 1024):
            assert lib.conv3d_gen_debug_mma_rate(None, layout | (rotate << 4), n, iters, ctypes.c_void_p(out.data_ptr())) == 0
            torch.cuda.synchronize()
            res.append(int(out.item()))
        per = (res[1] - res[0]) / (1024 - 64)
        print(f"{names[layout]} {'4 rotating tiles' if rotate else 'same tile       '} N={n:3d}: {per:7.1f} clk / MMA   (math at 1447 FMA/clk: {128 * n * 8 / 1447:6.1f})")
