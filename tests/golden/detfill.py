"""Deterministic parameter / input fill shared by the fixture generators and the tests, so that large state_dicts need
not be stored: parameter k of a module (state_dict order) becomes a scaled sine sequence that depends only on its name
index and shape."""
import math

import torch


def det_tensor(shape, k, scale=1.0, offset=0.0):
    n = int(torch.tensor(shape).prod()) if len(shape) else 1
    t = torch.sin(torch.arange(n, dtype=torch.float64) * 0.37 + 1.3 * k) * scale + offset
    return t.reshape(shape).float()


@torch.no_grad()
def det_fill_module(module):
    for k, (name, p) in enumerate(module.state_dict().items()):
        if not p.dtype.is_floating_point:
            continue
        if p.dim() > 1:
            fan_in = p[0].numel()
            p.copy_(det_tensor(p.shape, k, scale=1.0 / math.sqrt(fan_in)))
        elif name.endswith("weight"):                      # norm scales
            p.copy_(det_tensor(p.shape, k, scale=0.1, offset=1.0))
        else:                                              # biases
            p.copy_(det_tensor(p.shape, k, scale=0.05))
    return module
