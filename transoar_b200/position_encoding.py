"""3D sine positional encoding -- mirror of transoar/models/position_encoding.py:10-51 (`PositionEmbeddingSine3D`).

Same numbers as the reference (per-axis channel count ceil(C/6)*2, cumsum coordinates (k-0.5)/(n+1e-6)*2*pi, concat order
(y, x, z) in the reference's naming = (axis 2, axis 1, axis 3) of a [N,C,D,H,W] map, truncated to C channels).  The encoding
depends only on the map's shape, so it is computed once per (shape, device, dtype) and cached -- the
reference recomputes cumsum + sin/cos over every FPN level in every forward."""
import math

import torch
from torch import nn


class PositionEmbeddingSine3D(nn.Module):
    def __init__(self, channels=64, temperature=10000, normalize=True, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.orig_channels = channels
        self.channels = int(math.ceil(channels / 6) * 2)
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = {}

    def _axis(self, n, device):
        """Encoding of one axis of length n: [n, channels] = [sin of the even frequencies | cos of the odd ones]."""
        coord = torch.arange(1, n + 1, dtype=torch.float32, device=device)          # cumsum of ones (pe.py:29-32)
        if self.normalize:
            coord = (coord - 0.5) / (coord[-1:] + 1e-6) * self.scale                 # pe.py:34-38
        k = torch.arange(self.channels, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * torch.div(k, 2, rounding_mode="trunc") / self.channels)   # pe.py:40-41
        ang = coord[:, None] / dim_t
        # pe.py:47-49 stacks on dim 4 of a 5-D tensor, i.e. BEFORE the channel axis: all sines first, then all cosines
        return torch.cat((ang[:, 0::2].sin(), ang[:, 1::2].cos()), dim=1)

    def forward(self, src):
        N, _, A, B, C3 = src.shape                       # reference names: x = axis 1 (A), y = axis 2 (B), z = axis 3 (C3)
        key = (A, B, C3, src.device, N)
        pos = self._cache.get(key)
        if pos is None:
            ea, eb, ec = self._axis(A, src.device), self._axis(B, src.device), self._axis(C3, src.device)
            ch = self.channels
            full = torch.empty(A, B, C3, 3 * ch, dtype=torch.float32, device=src.device)
            full[..., :ch] = eb[None, :, None, :]           # pos_y first (pe.py:50)
            full[..., ch:2 * ch] = ea[:, None, None, :]     # then pos_x
            full[..., 2 * ch:] = ec[None, None, :, :]       # then pos_z
            pos = full.permute(3, 0, 1, 2)[None, :self.orig_channels].expand(N, -1, -1, -1, -1)
            self._cache[key] = pos
        return pos
