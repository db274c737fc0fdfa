"""One launch each of the fused backward in the model's sampling state (offsets = directional bias + 0.03 voxel jitter) for ncu:
bwd_vec_kernel (rot 6), bwd_duo_kernel (rot 0), bwd_duo_kernel without reductions.  Run under
  ncu --set full --clock-control none -k regex:bwd_ -o gpurun_out/<name> python tools/profile_duo.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transoar_b200 import MultiScaleDeformableAttention as MSDA, _lib, synth
from transoar_b200.ops.modules import MSDeformAttn

dev = "cuda:0"
g = synth.GEOMETRIES["visceral_refine"]
N, S, M, C, L, P = 2, g.spatial_size, g.heads, g.channels, g.levels, g.points
bias = MSDeformAttn(M * C, L, M, P, True).sampling_offsets.bias.detach().view(1, 1, M * L * P * 3)
gen = torch.Generator().manual_seed(7)
shapes, starts = synth.level_tensors(g.shapes, dev)
ref = synth.reference_points(g.shapes)[None, :, None, :].expand(1, S, L, 3).contiguous().to(dev)
value = (torch.rand(N, S, M, C, generator=gen) * 0.01).to(dev)
gout = (torch.randn(N, S, M * C, generator=gen) * 0.1).to(dev)
sigma = float(os.environ.get("SIGMA", "0.03"))
merged = torch.cat((bias + sigma * torch.randn(N, S, M * L * P * 3, generator=gen), 0.1 * torch.randn(N, S, M * L * P, generator=gen)), -1).to(dev).contiguous()
T = _lib.lib().msda3d_set_tuning
for duo, rot, skip in ((0, 6, 0), (1, 0, 0), (1, 0, 1)):
    T(b"duo", duo); T(b"rot", rot); T(b"diag_bwd_skip_red", skip)
    MSDA.ms_deform_attn_backward_merged(value, shapes, starts, ref, merged, gout, L, P)
    torch.cuda.synchronize()
