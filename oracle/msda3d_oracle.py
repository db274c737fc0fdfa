"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker or the timed CPU baseline.  Nothing under ``transoar_b200/`` imports it.

Three things live here:

* ctypes bindings of ``libmsda3d_oracle.so`` (``msda3d_oracle.c``: the C restatement of the reference's CUDA
  kernel arithmetic, ``ms_deform_im2col_cuda.cuh:31-241,370-439``),
* ``gridsample_path`` -- a restatement of the reference's ``use_cuda=False`` route
  (``transoar/models/ops/functions/ms_deform_attn_func.py:41-65``: per level ``F.grid_sample`` on ``2*loc-1`` with
  ``align_corners=False`` and zero padding, weighted by the attention weights and summed over levels x points),
* ctypes bindings of ``oracle/_ref/libmsda3d_refcuda.so`` -- the reference's OWN CUDA kernels compiled from where
  they lie (``oracle/Makefile``), callable on the GPU box only.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REFCUDA = None

_c_int = ctypes.c_int
_c_void_p = ctypes.c_void_p


def build(with_ref: bool | None = None) -> None:
    """Compile the C oracle; compile oracle/_ref too when /root/reference is present (build container only)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if with_ref is None:
        with_ref = os.path.isdir("/root/reference/transoar/models/ops/src")
    if with_ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsda3d_oracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        _LIB = ctypes.CDLL(path)
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_c_void_p)


def _prep(value, shapes, starts, loc, aw):
    value = np.ascontiguousarray(value)
    dt = value.dtype
    if dt not in (np.float32, np.float64):
        raise TypeError("oracle supports float32/float64 only (reference: AT_DISPATCH_FLOATING_TYPES)")
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    loc = np.ascontiguousarray(loc, dtype=dt)
    aw = np.ascontiguousarray(aw, dtype=dt)
    N, S, M, C = value.shape
    _, Lq, _, L, P, _ = loc.shape
    assert loc.shape == (N, Lq, M, L, P, 3) and aw.shape == (N, Lq, M, L, P) and shapes.shape == (L, 3)
    assert int((shapes.prod(1)).sum()) == S
    return value, shapes, starts, loc, aw, (N, S, M, C, L, Lq, P), ("f32" if dt == np.float32 else "f64")


def forward(value, shapes, starts, loc, aw, contract: bool = True) -> np.ndarray:
    """out[N,Lq,M*C]; ``contract`` selects the compiled-reference (FMA) or source-literal rounding of loc*size-0.5."""
    value, shapes, starts, loc, aw, dims, suf = _prep(value, shapes, starts, loc, aw)
    N, S, M, C, L, Lq, P = dims
    out = np.empty((N, Lq, M * C), dtype=value.dtype)
    fn = getattr(lib(), f"msda3d_oracle_forward_{suf}")
    fn.restype = None
    fn(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(aw), *map(_c_int, dims), _ptr(out), _c_int(int(contract)))
    return out


def backward(grad_out, value, shapes, starts, loc, aw, contract: bool = True):
    """-> (grad_value[N,S,M,C], grad_loc[N,Lq,M,L,P,3], grad_aw[N,Lq,M,L,P])."""
    value, shapes, starts, loc, aw, dims, suf = _prep(value, shapes, starts, loc, aw)
    N, S, M, C, L, Lq, P = dims
    grad_out = np.ascontiguousarray(grad_out, dtype=value.dtype).reshape(N, Lq, M * C)
    gv = np.zeros_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(aw)
    fn = getattr(lib(), f"msda3d_oracle_backward_{suf}")
    fn.restype = None
    fn(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(aw), *map(_c_int, dims),
       _ptr(gv), _ptr(gl), _ptr(ga), _c_int(int(contract)))
    return gv, gl, ga


def indices(shapes, loc, contract: bool = True):
    """Sampling-index arithmetic only -> (idx int32 [N,Lq,M,L,P,4] = in_range,d_low,h_low,w_low ; frac [...,3] = ld,lh,lw)."""
    loc = np.ascontiguousarray(loc)
    suf = "f32" if loc.dtype == np.float32 else "f64"
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    N, Lq, M, L, P, _ = loc.shape
    idx = np.empty((N, Lq, M, L, P, 4), dtype=np.int32)
    frac = np.empty((N, Lq, M, L, P, 3), dtype=loc.dtype)
    fn = getattr(lib(), f"msda3d_oracle_indices_{suf}")
    fn.restype = None
    fn(_ptr(shapes), _ptr(loc), _c_int(N), _c_int(M), _c_int(L), _c_int(Lq), _c_int(P), _ptr(idx), _ptr(frac),
       _c_int(int(contract)))
    return idx, frac


def gridsample_path(value, shapes, loc, aw):
    """Restatement of the reference's use_cuda=False route (ms_deform_attn_func.py:41-65) on torch tensors.

    value [N,S,M,C], shapes [(D,H,W)]*L, loc [N,Lq,M,L,P,3] in (x,y,z) = (W,H,D) order, aw [N,Lq,M,L,P]
    -> [N,Lq,M*C].  Differentiable (autograd) -- this is what the reference trains through on CPU.
    """
    import torch
    import torch.nn.functional as F

    N, S, M, C = value.shape
    _, Lq, _, L, P, _ = loc.shape
    sizes = [int(d) * int(h) * int(w) for d, h, w in shapes]
    grid_all = loc * 2 - 1                                            # func.py:51
    per_level = []
    begin = 0
    for lvl, (D, H, W) in enumerate(shapes):
        D, H, W = int(D), int(H), int(W)
        vol = value[:, begin:begin + sizes[lvl]]                      # func.py:49
        begin += sizes[lvl]
        vol = vol.permute(0, 2, 3, 1).reshape(N * M, C, D, H, W)      # func.py:54
        grid = grid_all[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, P, 3)   # func.py:55,59
        smp = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=False)  # func.py:58-60
        per_level.append(smp.reshape(N * M, C, Lq, P))
    stacked = torch.stack(per_level, dim=3).reshape(N * M, C, Lq, L * P)   # func.py:64 (stack on -2, flatten)
    w = aw.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)             # func.py:63
    out = (stacked * w).sum(-1).reshape(N, M * C, Lq)                      # func.py:64
    return out.transpose(1, 2).contiguous()                                # func.py:65


# --------------------------------------------------------------------------------------------------------------
# The reference's own compiled CUDA op (GPU box only).  Takes torch CUDA tensors.
# --------------------------------------------------------------------------------------------------------------

def refcuda_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libmsda3d_refcuda.so"))


def refcuda() -> ctypes.CDLL:
    global _REFCUDA
    if _REFCUDA is None:
        _REFCUDA = ctypes.CDLL(os.path.join(_HERE, "_ref", "libmsda3d_refcuda.so"))
    return _REFCUDA


def _tp(t):
    return _c_void_p(t.data_ptr())


def refcuda_forward(value, shapes, starts, loc, aw):
    """Reference kernels on the current CUDA stream; zero-inits like ms_deform_attn_cuda.cu:54."""
    import torch
    N, S, M, C = value.shape
    _, Lq, _, L, P, _ = loc.shape
    suf = {torch.float32: "f32", torch.float64: "f64"}[value.dtype]
    out = torch.zeros(N, Lq, M * C, dtype=value.dtype, device=value.device)
    fn = getattr(refcuda(), f"msda3d_refcuda_forward_{suf}")
    fn.restype = _c_int
    rc = fn(_c_void_p(torch.cuda.current_stream().cuda_stream), _tp(value), _tp(shapes), _tp(starts), _tp(loc), _tp(aw),
            *map(_c_int, (N, S, M, C, L, Lq, P)), _tp(out))
    if rc != 0:
        raise RuntimeError(f"reference CUDA forward launch failed: cudaError {rc}")
    return out


def refcuda_backward(grad_out, value, shapes, starts, loc, aw, out=None):
    """-> (grad_value, grad_loc, grad_aw); zero-inits like ms_deform_attn_cuda.cu:122-124 (or reuses ``out``, pre-zeroed)."""
    import torch
    N, S, M, C = value.shape
    _, Lq, _, L, P, _ = loc.shape
    suf = {torch.float32: "f32", torch.float64: "f64"}[value.dtype]
    if out is None:
        out = (torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(aw))
    gv, gl, ga = out
    fn = getattr(refcuda(), f"msda3d_refcuda_backward_{suf}")
    fn.restype = _c_int
    rc = fn(_c_void_p(torch.cuda.current_stream().cuda_stream), _tp(grad_out), _tp(value), _tp(shapes), _tp(starts),
            _tp(loc), _tp(aw), *map(_c_int, (N, S, M, C, L, Lq, P)), _tp(gv), _tp(gl), _tp(ga))
    if rc != 0:
        raise RuntimeError(f"reference CUDA backward launch failed: cudaError {rc}")
    return gv, gl, ga
