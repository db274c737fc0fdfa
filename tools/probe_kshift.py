"""Does tcgen05.mma accept a K-major SW128 A operand that starts at an arbitrary row of a halo tile (kw shift) with 8-row groups that are
not 1024 bytes apart (row pitch of a halo line)?  Prints the max error of one MMA per (row0, group stride, base-offset mode)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from transoar_b200 import _lib  # noqa: E402

lib = _lib.lib()
g = torch.Generator().manual_seed(0)
X = torch.randint(-4, 5, (176, 32), generator=g).float().cuda()
Y = torch.randint(-3, 4, (32, 32), generator=g).float().cuda()
D = torch.empty(128, 32, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
m = torch.arange(128, device="cuda")
for gs in (8, 10, 18, 9):
    for row0 in (0, 1, 2, 3, 5, 8, 11):
        for mode in (0, 1):
            if row0 + 15 * gs + 8 > 176:
                continue
            assert lib.conv3d_gen_debug_k_probe(None, p(X), p(Y), p(D), row0, gs, mode) == 0
            torch.cuda.synchronize()
            rows = row0 + (m // 8) * gs + m % 8
            want = X[rows][:, :8] @ Y[:, :8].t()
            err = float((D - want).abs().max())
            print(f"gs={gs:2d} row0={row0:2d} base_offset={mode}  max_err={err:g}  {'OK' if err == 0 else 'WRONG'}")
