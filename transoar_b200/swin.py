"""Swin encoder stages -- mirrors of transoar/models/backbones/encoder_blocks.py: ``EncoderSwinBlock`` (:56-120), ``SwinBlock``
(:122-211), ``WindowAttention3D`` (:213-289), ``Mlp`` (:291-307), ``PatchMerging`` (:309-340), ``ConvPatchMerging`` (:342-364) and the
window helpers (:366-400).  Used by the AttnFPN encoder for stages >= 2 when ``use_encoder_attn`` is set (attn_fpn.py:172-190;
BASELINE.json configs[3]).

Module / parameter / buffer names equal the reference's (``blocks.<i>.{norm1,attn.{relative_position_bias_table,
relative_position_index,qkv,proj},norm2,mlp.{fc1,fc2}}``, ``downsample.{reduction,norm}``), so reference checkpoints load.
The four Linear layers of every block are ``TCLinear`` (TF32 tcgen05 GEMMs when TF32 is the requested precision); the 125-token
window attention itself (head dim 16) is ``scaled_dot_product_attention`` with the relative-position bias and the shift mask
as one additive term -- a library call: at 5x5x5 windows it is tiny next to the projections.  Device-agnostic torch code, so the
fixtures generated from the reference modules are checked on the CPU (tests/test_swin_cpu.py)."""
import ctypes
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

from .fused_ln import layer_norm
from .linear import TCLinear


def effective_window(size, window, shift=None):
    """An axis shorter than the window uses the axis length as window and no shift (encoder_blocks.py:376-389)."""
    win = tuple(s if s <= w else w for s, w in zip(size, window))
    if shift is None:
        return win
    return win, tuple(0 if s <= w else sh for s, w, sh in zip(size, window, shift))


def to_windows(x, win):
    """[B, D, H, W, C] (D, H, W multiples of the window) -> [B * nW, wd*wh*ww, C], windows in (d, h, w) raster order."""
    B, D, H, W, C = x.shape
    wd, wh, ww = win
    x = x.reshape(B, D // wd, wd, H // wh, wh, W // ww, ww, C).permute(0, 1, 3, 5, 2, 4, 6, 7)
    return x.reshape(-1, wd * wh * ww, C)


def from_windows(windows, win, B, D, H, W):
    wd, wh, ww = win
    x = windows.reshape(B, D // wd, H // wh, W // ww, wd, wh, ww, -1).permute(0, 1, 4, 2, 5, 3, 6, 7)
    return x.reshape(B, D, H, W, -1)


_MASKS = {}


def shift_mask(dims, win, shift, device):
    """Additive attention mask [nW, n, n] of the shifted-window scheme: 0 where two tokens of a window come from the same region of
    the (padded, rolled) volume, -100 otherwise (encoder_blocks.py:387-400).  Cached per geometry."""
    key = (tuple(dims), tuple(win), tuple(shift), str(device))
    if key not in _MASKS:
        labels = []
        for size, w, s in zip(dims, win, shift):
            idx = torch.arange(size, device=device)
            # three bands per axis: before the last window, the last window up to the shift, the wrapped-around tail;
            # with no shift on an axis the reference's last slice covers the whole axis (one band)
            labels.append(torch.full((size,), 2, device=device) if s == 0 else (idx >= size - w).long() + (idx >= size - s).long())
        region = (labels[0][:, None, None] * 9 + labels[1][None, :, None] * 3 + labels[2][None, None, :]).float()
        per_window = to_windows(region[None, ..., None], win).squeeze(-1)                   # [nW, n]
        differs = per_window[:, None, :] != per_window[:, :, None]
        _MASKS[key] = torch.zeros(differs.shape, device=device).masked_fill(differs, -100.0)
    return _MASKS[key]


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath, which the reference imports at encoder_blocks.py:10)."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1.0 - self.p
        gate = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * gate / keep


class WindowAttentionFunction(Function):
    """softmax((q * scale) k^T + bias[head] (+ mask[window])) v on the fused sm_100a kernel (include/win_attn.h).
    ``qkv`` [Bw, n, 3, H, 16] (the qkv Linear's output, viewed), ``bias`` [H, n, n], ``mask`` [nW, n, n] or None -> [Bw, n, H * 16]."""

    @staticmethod
    def forward(ctx, qkv, bias, mask, scale):
        if not qkv.is_cuda:
            raise RuntimeError("window attention: Not implemented on the CPU")
        qkv = qkv.float().contiguous()
        bias = bias.float().contiguous()
        bias_t = bias.transpose(1, 2).contiguous()
        mask = None if mask is None else mask.float().contiguous()
        Bw, n, _, H, hd = qkv.shape
        out = torch.empty(Bw, n, H * hd, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(Bw, H, n, dtype=torch.float32, device=qkv.device)
        with torch.cuda.device(qkv.device):
            rc = _lib.lib().win_attn_forward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _ptr(qkv), _ptr(bias_t), _ptr(mask), Bw, n, H, hd,
                                             0 if mask is None else mask.shape[0], float(scale), _ptr(out), _ptr(lse))
        _lib.check(rc, "win_attn_forward")
        ctx.save_for_backward(qkv, bias, bias_t, mask, out, lse)
        ctx.scale = float(scale)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        qkv, bias, bias_t, mask, out, lse = ctx.saved_tensors
        Bw, n, _, H, hd = qkv.shape
        dout = dout.float().contiguous()
        dqkv, dbias = torch.empty_like(qkv), torch.empty_like(bias)
        with torch.cuda.device(qkv.device):
            rc = _lib.lib().win_attn_backward(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), _ptr(qkv), _ptr(bias), _ptr(bias_t), _ptr(mask),
                                              _ptr(out), _ptr(dout), _ptr(lse), Bw, n, H, hd, 0 if mask is None else mask.shape[0], ctx.scale,
                                              _ptr(dqkv), _ptr(dbias))
        _lib.check(rc, "win_attn_backward")
        return dqkv, dbias, None, None


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class WindowAttention3D(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias, qk_scale, attn_drop, proj_drop):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wd, wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), num_heads))
        pos = torch.stack(torch.meshgrid(torch.arange(wd), torch.arange(wh), torch.arange(ww), indexing="ij")).flatten(1)    # [3, n]
        rel = pos[:, :, None] - pos[:, None, :] + torch.tensor([wd - 1, wh - 1, ww - 1])[:, None, None]
        index = (rel[0] * (2 * wh - 1) + rel[1]) * (2 * ww - 1) + rel[2]
        self.register_buffer("relative_position_index", index)
        self.qkv = TCLinear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = TCLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def forward(self, x, mask=None):
        """x [B * nW, n, C]; mask [nW, n, n] or None."""
        Bw, n, C = x.shape
        H = self.num_heads
        qkv = self.qkv(x)
        bias = self.relative_position_bias_table[self.relative_position_index[:n, :n].reshape(-1)].reshape(n, n, H).permute(2, 0, 1)
        if (x.is_cuda and not (self.training and self.attn_drop.p > 0) and (mask is None or Bw % mask.shape[0] == 0)
                and _lib.lib().win_attn_supported(n, C // H)):
            # fused kernel: reads q, k, v in place out of the Linear's output, keeps K / V in shared memory, never builds [Bw, H, n, n]
            out = WindowAttentionFunction.apply(qkv.reshape(Bw, n, 3, H, C // H), bias, mask, self.scale)
            return self.proj_drop(self.proj(out.to(qkv.dtype)))
        q, k, v = qkv.reshape(Bw, n, 3, H, C // H).permute(2, 0, 3, 1, 4)                          # each [Bw, H, n, hd]
        if mask is None:
            add = bias[None]                                                                        # [1, H, n, n]
        else:
            nW = mask.shape[0]
            add = bias[None, None] + mask[None, :, None]                                            # [1, nW, H, n, n]
            q, k, v = (t.reshape(Bw // nW, nW, H, n, C // H) for t in (q, k, v))
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=add.to(q.dtype), dropout_p=self.attn_drop.p if self.training else 0.0,
                                             scale=self.scale)
        out = out.reshape(Bw, H, n, C // H).transpose(1, 2).reshape(Bw, n, C)
        return self.proj_drop(self.proj(out))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = TCLinear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = TCLinear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class SwinBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim, self.num_heads, self.mlp_ratio = dim, num_heads, mlp_ratio
        self.window_size, self.shift_size = tuple(window_size), tuple(shift_size)
        assert all(0 <= s < w for s, w in zip(self.shift_size, self.window_size)), "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention3D(dim, self.window_size, num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _attend(self, x, mask):
        B, D, H, W, C = x.shape
        win, shift = effective_window((D, H, W), self.window_size, self.shift_size)
        x = layer_norm(x, self.norm1)
        pad = [(w - s % w) % w for s, w in zip((D, H, W), win)]
        x = F.pad(x, (0, 0, 0, pad[2], 0, pad[1], 0, pad[0]))                                      # pad the high side of W, H, D
        Dp, Hp, Wp = x.shape[1:4]
        shifted = any(shift)
        if shifted:
            x = torch.roll(x, shifts=tuple(-s for s in shift), dims=(1, 2, 3))
        y = self.attn(to_windows(x, win), mask if shifted else None)
        y = from_windows(y, win, B, Dp, Hp, Wp)
        if shifted:
            y = torch.roll(y, shifts=shift, dims=(1, 2, 3))
        return y[:, :D, :H, :W].contiguous() if any(pad) else y

    def forward(self, x, mask_matrix):
        x = x + self.drop_path(self._attend(x, mask_matrix))
        return x + self.drop_path(self.mlp(layer_norm(x, self.norm2)))


class PatchMerging(nn.Module):
    """2x2x2 neighbourhoods concatenated on the channel axis (order d, then h before w as in the reference :324-334) -> LN -> Linear."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = TCLinear(8 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(8 * dim)

    def forward(self, x):
        B, D, H, W, C = x.shape
        if H % 2 or W % 2:
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        parts = [x[:, d::2, h::2, w::2] for d in (0, 1) for w in (0, 1) for h in (0, 1)]            # x0..x7 of the reference
        return self.reduction(layer_norm(torch.cat(parts, -1), self.norm))


class ConvPatchMerging(nn.Module):
    def __init__(self, dim, bias=False, affine=True, eps=1e-05):
        super().__init__()
        self._reduction = nn.Sequential(nn.Conv3d(dim, dim * 2, kernel_size=2, stride=2, padding=0, bias=bias),
                                        nn.InstanceNorm3d(dim * 2, affine=affine, eps=eps), nn.ReLU(inplace=True))

    def forward(self, x):
        return self._reduction(x.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)


class EncoderSwinBlock(nn.Module):
    """``depth`` Swin blocks (even ones unshifted, odd ones shifted by half a window) + patch merging; [B,C,D,H,W] in and out."""

    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, downsample,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(w // 2 for w in window_size)
        self.blocks = nn.ModuleList(
            SwinBlock(dim, num_heads, self.window_size, (0, 0, 0) if i % 2 == 0 else self.shift_size, mlp_ratio, qkv_bias, qk_scale, drop,
                      attn_drop, drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer)
            for i in range(depth))
        self.downsample = None
        if downsample is not None:
            self.downsample = downsample(dim=dim) if downsample is ConvPatchMerging else downsample(dim=dim, norm_layer=norm_layer)

    def forward(self, x):
        B, C, D, H, W = x.shape
        win, shift = effective_window((D, H, W), self.window_size, self.shift_size)
        padded = tuple(math.ceil(s / w) * w for s, w in zip((D, H, W), win))
        mask = shift_mask(padded, win, shift, x.device)
        x = x.permute(0, 2, 3, 4, 1)                                   # a view when the backbone is channels-last
        for blk in self.blocks:
            x = blk(x, mask)
        if self.downsample is not None:
            x = self.downsample(x)
        return x.permute(0, 4, 1, 2, 3)
