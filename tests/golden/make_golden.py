"""Generate tests/golden/*.npz from the REFERENCE's own Python path.  Run in the build container only:

    python tests/golden/make_golden.py

Imports ``ms_deform_attn_core_pytorch`` straight from /root/reference (file loaded by path, nothing copied),
runs it in fp64 and fp32 with autograd on seeded inputs, and stores inputs + outputs + gradients.  The reference
has no golden vectors of its own (ops/test.py is unseeded), so these fixtures are what pins oracle/ (SURVEY 8c).
Cases follow ops/test.py's presets and input distribution (test.py:35-42,54-57) plus an out-of-range case that
exercises the zero-padding and range-test branches (ms_deform_im2col_cuda.cuh:60-107,428).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/transoar/models/ops/functions/ms_deform_attn_func.py"

CASES = {
    # name: (N, M, C, Lq, P, shapes, loc_lo, loc_hi)
    "small": (2, 3, 4, 4, 4, [(3, 6, 4), (2, 3, 2)], 0.0, 1.0),          # ops/test.py:35-37
    "tiny": (1, 1, 1, 1, 1, [(2, 2, 2)], 0.0, 1.0),                      # ops/test.py:40-42
    "border": (2, 2, 5, 37, 3, [(4, 5, 6), (3, 2, 4), (1, 1, 1)], -0.3, 1.3),
    "heads6": (1, 6, 8, 64, 4, [(6, 6, 8), (3, 3, 4), (2, 2, 2), (1, 1, 1)], -0.05, 1.05),
}


def main():
    spec = importlib.util.spec_from_file_location("ref_func", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for ci, (name, (N, M, C, Lq, P, shapes, lo, hi)) in enumerate(CASES.items()):
        torch.manual_seed(1234 + ci)
        L = len(shapes)
        S = sum(d * h * w for d, h, w in shapes)
        ss = torch.as_tensor(shapes, dtype=torch.long)
        starts = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
        value = torch.rand(N, S, M, C, dtype=torch.float64) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 3, dtype=torch.float64) * (hi - lo) + lo
        aw = torch.rand(N, Lq, M, L, P, dtype=torch.float64) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
        gout = torch.randn(N, Lq, M * C, dtype=torch.float64)
        blob = dict(shapes=ss.numpy(), starts=starts.numpy(), value=value.numpy(), loc=loc.numpy(), aw=aw.numpy(),
                    grad_out=gout.numpy())
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            v, l, a = (t.to(dt).clone().requires_grad_(True) for t in (value, loc, aw))
            out = ref.ms_deform_attn_core_pytorch(v, ss, l, a)
            out.backward(gout.to(dt))
            blob[f"out_{tag}"] = out.detach().numpy()
            blob[f"grad_value_{tag}"] = v.grad.numpy()
            blob[f"grad_loc_{tag}"] = l.grad.numpy()
            blob[f"grad_aw_{tag}"] = a.grad.numpy()
        path = os.path.join(HERE, f"msda3d_{name}.npz")
        np.savez_compressed(path, **blob)
        print(name, "S", S, "->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("reference not mounted; golden fixtures can only be regenerated in the build container")
    main()
