"""AttnFPN backbone -- mirror of transoar/models/backbones/attn_fpn.py (``AttnFPN`` :18-32, ``Decoder`` :34-145,
``Encoder`` :148-213) and ``EncoderCnnBlock`` (transoar/models/backbones/encoder_blocks.py:14-54).

Module / parameter names equal the reference's (``_encoder._stages.<i>._block.<k>``, ``_decoder._lateral/_up/_out/_refine``)
so reference checkpoints load.  The encoder's first convolution (1 input channel) is a direct sm_100a stencil (include/stem_conv.h) when the model is
channels-last; every other 3x3x3 convolution (stride 1 / 2, with / without bias) runs on the general tcgen05 implicit-GEMM kernels
(include/conv3d_gen.h), the 1x1 lateral and the kernel-2 / stride-2 transposed convolutions as GEMMs over the NDHWC rows (transoar_b200/conv3d_gen.py) --
cuDNN is only the route of NCDHW / autocast / strict-fp32 models; every
InstanceNorm3d -> ReLU pair runs as a fused sm_100a kernel (include/instnorm.h); the deformable refinement (`use_decoder_attn`) runs on the sm_100a kernels through
``transoar_b200.refine.DecoderDefAttnBlock``; the Swin encoder variant (``use_encoder_attn``, encoder_blocks.py:56-334) through
``transoar_b200.swin``."""
import torch
from torch import nn

from .conv3d_gen import (conv3d_1x1, conv3d_k3_gen, conv_1x1_eligible, conv_gen_eligible, conv_transpose3d_k2s2,
                         conv_transpose_eligible)
from .conv3d_tc import conv3d_k3, conv_tc_eligible
from .instnorm import instance_norm_relu
from .position_encoding import PositionEmbeddingSine3D
from .refine import DecoderDefAttnBlock
from .stem_conv import stem_conv3d, stem_eligible
from .swin import ConvPatchMerging, EncoderSwinBlock, PatchMerging


def _tf32_backbone_under_autocast(conv, x):
    """Inside a bf16 autocast region (the reference's trainer runs the model under autocast, trainer.py:67-69) a channels-last CUDA model keeps
    its convolutional backbone on this library's fp32-storage / TF32-multiply kernels: autocast would hand the convolutions to cuDNN in bf16,
    which this library has no kernels for -- TF32 is the HIGHER precision (10 mantissa bits against 7), so the 1e-2 bf16 parity bound holds a
    fortiori.  The transformer parts (every Linear, the op, the attention kernels) stay on their bf16 routes."""
    return (torch.is_autocast_enabled() and x.is_cuda and torch.backends.cudnn.allow_tf32 and conv.weight.dtype == torch.float32
            and conv.weight.is_contiguous(memory_format=torch.channels_last_3d) and not conv.weight.is_contiguous())


class EncoderCnnBlock(nn.Module):
    """(Conv3d no-bias -> InstanceNorm3d(affine) -> ReLU) x 2; the first conv carries the stride (encoder_blocks.py:14-54)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding=1, bias=False, affine=True, eps=1e-05):
        super().__init__()
        layers = []
        for cin, s in ((in_channels, stride), (out_channels, 1)):
            layers += [nn.Conv3d(cin, out_channels, kernel_size=kernel_size, stride=s, padding=padding, bias=bias),
                       nn.InstanceNorm3d(num_features=out_channels, affine=affine, eps=eps), nn.ReLU(inplace=True)]
        self._block = nn.Sequential(*layers)

    def forward(self, x):
        if _tf32_backbone_under_autocast(self._block[3], x):
            with torch.autocast(device_type="cuda", enabled=False):
                return self._forward(x.float())
        return self._forward(x)

    def _forward(self, x):
        # same Sequential (so the reference's parameter names are kept), but each InstanceNorm3d -> ReLU pair runs as one
        # fused sm_100a kernel (transoar_b200/instnorm.py) instead of cuDNN batch-norm + an elementwise ReLU
        conv1, norm1, _, conv2, norm2, _ = self._block
        cl_model = conv2.weight.is_contiguous(memory_format=torch.channels_last_3d) and not conv2.weight.is_contiguous()
        if cl_model and stem_eligible(conv1, x):
            x = stem_conv3d(x, conv1.weight)                    # 1 -> C direct stencil, writes NDHWC (include/stem_conv.h)
        elif cl_model and conv_gen_eligible(conv1, x):
            x = conv3d_k3_gen(x, conv1.weight, conv1.bias, conv1.stride[0])      # stride-2 stage entry on the general tcgen05 kernel (include/conv3d_gen.h)
        else:
            x = conv1(x)
            if cl_model and conv1.in_channels == 1 and not x.is_contiguous(memory_format=torch.channels_last_3d):
                # a 1-channel input / weight is layout-ambiguous and cuDNN answers NCDHW: move to NDHWC once, here
                x = x.contiguous(memory_format=torch.channels_last_3d)
        x = instance_norm_relu(x, norm1.weight, norm1.bias, norm1.eps)
        # narrow full-resolution stage (24 -> 24): tcgen05 implicit-GEMM convolution, forward and input gradient (include/conv3d_tc.h)
        if conv_tc_eligible(conv2, x):
            x = conv3d_k3(x, conv2.weight)
        elif cl_model and conv_gen_eligible(conv2, x):
            x = conv3d_k3_gen(x, conv2.weight, conv2.bias, 1)                    # every wider stage: the general tcgen05 kernel
        else:
            x = conv2(x)
        return instance_norm_relu(x, norm2.weight, norm2.bias, norm2.eps)


class Encoder(nn.Module):
    """attn_fpn.py:148-213: one stage per entry of conv_kernels, channels start_channels * 2^stage."""

    def __init__(self, config, debug=False):
        super().__init__()
        cin, cout = config["in_channels"], config["start_channels"]
        depths = config["depths"]
        drop_path = [v.item() for v in torch.linspace(0, config["drop_path_rate"], sum(depths))]       # stochastic-depth schedule (:160-161)
        merging = ConvPatchMerging if config["conv_merging"] else PatchMerging
        self._stages = nn.ModuleList()
        for stage, (kernel, stride) in enumerate(zip(config["conv_kernels"], config["strides"])):
            if config["use_encoder_attn"] and stage > 1:                                               # stages 0-1 stay convolutional (:172)
                k = stage - 2
                self._stages.append(EncoderSwinBlock(dim=cin, depth=depths[k], num_heads=config["num_heads"][k],
                                                     window_size=config["window_size"], mlp_ratio=config["mlp_ratio"],
                                                     qkv_bias=config["qkv_bias"], qk_scale=config["qk_scale"], drop=config["drop_rate"],
                                                     attn_drop=config["attn_drop_rate"],
                                                     drop_path=drop_path[sum(depths[:k]):sum(depths[:k + 1])], downsample=merging))
            else:
                self._stages.append(EncoderCnnBlock(cin, cout, kernel, stride))
            cin, cout = cout, cout * 2

    def forward(self, x):
        outputs = {}
        if x.is_cuda and self._stages[0]._block[3].weight.is_contiguous(memory_format=torch.channels_last_3d):
            # channels-last model (TransoarNet.to(memory_format=torch.channels_last_3d)): hand the first convolution an NDHWC
            # input so that every activation of the backbone is produced and consumed in the tensor-core convolutions' own layout
            x = x.contiguous(memory_format=torch.channels_last_3d)
        for i, stage in enumerate(self._stages):
            x = stage(x)
            outputs["C" + str(i)] = x
        return outputs


class Decoder(nn.Module):
    """attn_fpn.py:34-145: 1x1 lateral convs, transposed-conv top-down path, 3x3x3 output convs, optional deformable refine."""

    def __init__(self, config, debug=False):
        super().__init__()
        n_stages = len(config["conv_kernels"])
        self._refine_fmaps = config["use_decoder_attn"]
        self._refine_feature_levels = config["feature_levels"]
        self._seg_proxy = config["use_seg_proxy_loss"]
        enc_ch = [config["start_channels"] * 2 ** s for s in range(n_stages)]
        wanted = config["out_fmaps"] + config["feature_levels"] if config["use_decoder_attn"] else config["out_fmaps"]
        required = set(int(name[-1]) for name in wanted)                                   # :50-53 (a set, iterated ascending)
        if self._seg_proxy:
            required.add(0)
        self._required_stages = required
        first = min(required)
        lat_in = enc_ch if self._seg_proxy else enc_ch[first:]
        lat_out = [min(c, config["fpn_channels"]) for c in lat_in]                         # :57-58
        self._lateral = nn.ModuleList(nn.Conv3d(i, o, kernel_size=1) for i, o in zip(lat_in, lat_out))
        self._lateral_levels = len(self._lateral)
        out_in = [lat_out[-n_stages + stage] for stage in required]                        # :67
        out_out = [int(config["fpn_channels"])] * len(out_in)
        out_out[0] = enc_ch[0] if self._seg_proxy else int(config["fpn_channels"])         # :69
        self._out = nn.ModuleList(nn.Conv3d(i, o, kernel_size=3, padding=1) for i, o in zip(out_in, out_out))
        rev_ch, rev_strides = lat_out[::-1], list(reversed(config["strides"]))
        self._up = nn.ModuleList(nn.ConvTranspose3d(rev_ch[l], rev_ch[l + 1], kernel_size=rev_strides[l], stride=rev_strides[l])
                                 for l in range(len(lat_out) - 1))                         # :76-83
        if self._refine_fmaps:
            if config["pos_encoding"] != "sine":
                raise NotImplementedError("only the sine positional encoding is mirrored (the shipped configs use it)")
            self._pos_enc = PositionEmbeddingSine3D(channels=config["hidden_dim"])
            self._refine = DecoderDefAttnBlock(d_model=config["hidden_dim"], nhead=config["nheads"], num_layers=config["layers"],
                                               dim_feedforward=config["dim_feedforward"], dropout=config["dropout"],
                                               feature_levels=config["feature_levels"], n_points=config["n_points"],
                                               use_cuda=config["use_cuda"])

    def forward(self, x):
        feats = list(x.values())[-self._lateral_levels:]
        if _tf32_backbone_under_autocast(self._out[0], feats[0]):
            with torch.autocast(device_type="cuda", enabled=False):
                outputs = self._fpn([f.float() for f in feats])
        else:
            outputs = self._fpn(feats)
        if self._refine_fmaps:
            fmaps = [outputs[name] for name in self._refine_feature_levels]
            refined = self._refine(fmaps, [self._pos_enc(f) for f in fmaps])
            outputs.update(zip(self._refine_feature_levels, refined))
        return outputs

    def _fpn(self, feats):
        # channels-last model: lateral 1x1 convolutions and the k = s = 2 transposed convolutions are GEMMs over the NDHWC rows (include/tc_gemm.h),
        # the 3x3x3 output convolutions run on the general tcgen05 kernel (include/conv3d_gen.h); otherwise the library modules
        own = self._out[0].weight.is_contiguous(memory_format=torch.channels_last_3d) and not self._out[0].weight.is_contiguous()
        lateral = [conv3d_1x1(f, conv.weight, conv.bias) if own and conv_1x1_eligible(conv, f) else conv(f) for conv, f in zip(self._lateral, feats)]
        top_down, cur = [], None
        for idx, lat in enumerate(reversed(lateral)):                                     # coarsest first (:110-118)
            if idx != 0:
                up = self._up[idx - 1]
                if own and conv_transpose_eligible(up, cur) and lat.is_contiguous(memory_format=torch.channels_last_3d):
                    cur = conv_transpose3d_k2s2(cur, up.weight, up.bias, lat)             # up(cur) + lat: the pixel shuffle rides in the addition
                else:
                    cur = lat + up(cur)
            else:
                cur = lat
            top_down.append(cur)
        fine_first = top_down[::-1]
        if self._seg_proxy:
            pairs = [(fine_first[stage], stage) for stage in self._required_stages]
        else:
            pairs = zip(fine_first, self._required_stages)                                # :124 (positional pairing, as the reference)
        outputs = {"P" + str(stage): conv3d_k3_gen(f, self._out[i].weight, self._out[i].bias, 1) if own and conv_gen_eligible(self._out[i], f)
                   else self._out[i](f) for i, (f, stage) in enumerate(pairs)}
        return outputs


class AttnFPN(nn.Module):
    def __init__(self, fpn_config, debug=False):
        super().__init__()
        self._encoder = Encoder(fpn_config, debug)
        self._decoder = Decoder(fpn_config, debug)

    def forward(self, src):
        return self._decoder(self._encoder(src))

    def init_weights(self):
        pass
