"""Parameter-holding module tree shared by the two denoiser networks.

The classes mirror the reference's module structure -- names, constructor arguments and
``state_dict`` keys -- so that checkpoints written for the reference load unchanged
(SURVEY.md section 8b "state_dict contract"):

  ccdm/ddpm/models/unet_openai/unet.py:70-311      TimestepEmbedSequential, Upsample, Downsample,
                                                   ResBlock, AttentionBlock
  ccdm/ddpm/models/unet_openai/nn.py:17-32,100     GroupNorm32 / conv_nd / linear
  latentdiffusion/ldm/modules/attention.py:37-261  GEGLU, FeedForward, CrossAttention,
                                                   BasicTransformerBlock, SpatialTransformer

They hold parameters only.  None of them has a ``forward``: the arithmetic is planned and
executed by ``unet_engine.UNetEngine`` as a sequence of calls into libguidegen_sm100.so.
"""
import math
from typing import Dict

import torch
from torch import nn


class ParamConv(nn.Module):
    """weight [Cout, Cin, k, ...(dims)] + bias, initialised like nn.ConvNd (conv_nd, nn.py:22-32)."""

    def __init__(self, dims, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.dims, self.in_channels, self.out_channels = dims, in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty((out_channels, in_channels) + (kernel_size,) * dims))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size ** self.dims
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)


class ParamLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(in_features)
            nn.init.uniform_(self.bias, -bound, bound)


class ParamNorm(nn.Module):
    """affine parameters of GroupNorm32(32, C) (nn.py:17-19,100) or nn.LayerNorm(C)."""

    def __init__(self, channels, eps=1e-5, groups=32):
        super().__init__()
        self.channels, self.eps, self.groups = channels, eps, groups
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class Slots(nn.Module):
    """Children registered under explicit integer names: reproduces the keys of an nn.Sequential
    whose parameter-free members (SiLU, Dropout, Identity, Softmax) are simply absent."""

    def __init__(self, children: Dict[int, nn.Module]):
        super().__init__()
        for k, m in children.items():
            self.add_module(str(k), m)

    def __getitem__(self, k):
        return self._modules[str(k)]

    def __contains__(self, k):
        return str(k) in self._modules


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        p.detach().zero_()
    return m


class TimestepEmbedSequential(nn.Module):
    """unet.py:70-84: an ordered list of layers; the engine dispatches on the layer type."""

    def __init__(self, *layers):
        super().__init__()
        for i, l in enumerate(layers):
            self.add_module(str(i), l)

    def __iter__(self):
        return iter(self._modules.values())


class Upsample(nn.Module):
    """unet.py:87-116: nearest x2 in every spatial dim, then a 3^d conv."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = ParamConv(dims, self.channels, self.out_channels, 3, padding=1)


class Downsample(nn.Module):
    """unet.py:119-146: 3^d conv with stride 2 (avg-pool variant is not on the shipped path)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if not use_conv:
            raise NotImplementedError("Downsample(use_conv=False) (avg_pool_nd) is not used by any shipped config")
        self.op = ParamConv(dims, self.channels, self.out_channels, 3, stride=2, padding=1)


class ResBlock(nn.Module):
    """unet.py:149-262 (use_scale_shift_norm / up / down are not used by the shipped configs)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down:
            raise NotImplementedError("ResBlock(use_scale_shift_norm/up/down) is not used by any shipped config")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.dims = dims
        self.in_layers = Slots({0: ParamNorm(channels), 2: ParamConv(dims, channels, self.out_channels, 3, padding=1)})
        self.emb_layers = Slots({1: ParamLinear(emb_channels, self.out_channels)})
        self.out_layers = Slots({0: ParamNorm(self.out_channels),
                                 3: zero_module(ParamConv(dims, self.out_channels, self.out_channels, 3, padding=1))})
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = ParamConv(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = ParamConv(dims, channels, self.out_channels, 1)


class AttentionBlock(nn.Module):
    """unet.py:265-311 with QKVAttentionLegacy (:334-360)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False, use_new_attention_order=False):
        super().__init__()
        if use_new_attention_order:
            raise NotImplementedError("use_new_attention_order is not used by any shipped config")
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0
            self.num_heads = channels // num_head_channels
        self.norm = ParamNorm(channels)
        self.qkv = ParamConv(1, channels, channels * 3, 1)
        self.proj_out = zero_module(ParamConv(1, channels, channels, 1))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = ParamLinear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """ldm/modules/attention.py:48-64 (glu=True is what BasicTransformerBlock passes)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=True, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("FeedForward(glu=False) is not on the sampler path")
        inner = int(dim * mult)
        self.net = Slots({0: GEGLU(dim, inner), 2: ParamLinear(inner, dim_out or dim)})


class CrossAttention(nn.Module):
    """ldm/modules/attention.py:152-193."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        self.to_q = ParamLinear(query_dim, inner, bias=False)
        self.to_k = ParamLinear(context_dim, inner, bias=False)
        self.to_v = ParamLinear(context_dim, inner, bias=False)
        self.to_out = Slots({0: ParamLinear(inner, query_dim)})


class BasicTransformerBlock(nn.Module):
    """ldm/modules/attention.py:196-215."""

    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1 = ParamNorm(dim, eps=1e-5, groups=0)
        self.norm2 = ParamNorm(dim, eps=1e-5, groups=0)
        self.norm3 = ParamNorm(dim, eps=1e-5, groups=0)


class SpatialTransformer(nn.Module):
    """ldm/modules/attention.py:218-261.  The reference is 2-D only; here all
    spatial axes flatten into the token axis, which is what 'b c h w -> b (h w) c' does in 2-D,
    so the same block serves the 3-D text-conditioned CCDM.  Extra keyword arguments of the
    CCDM call site (unet.py:585-588: disable_self_attn, use_linear, use_checkpoint) are accepted."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=False):
        super().__init__()
        if disable_self_attn or use_linear:
            raise NotImplementedError("disable_self_attn / use_linear have no reference semantics (SURVEY.md D2)")
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.n_heads, self.d_head = n_heads, d_head
        self.norm = ParamNorm(in_channels, eps=1e-6)     # Normalize(): GroupNorm(32, C, eps=1e-6)
        self.proj_in = ParamConv(2, in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim) for _ in range(depth or 1)])
        self.proj_out = zero_module(ParamConv(2, inner, in_channels, 1))


def build_unet_tree(model: nn.Module, *, dims, in_channels, model_channels, out_channels, num_res_blocks,
                    attention_resolutions, channel_mult, conv_resample, dropout, make_attn):
    """Populate ``model`` with time_embed / input_blocks / middle_block / output_blocks / out in the
    reference's construction order (unet.py:511-721 == openaimodel.py:507-693).
    ``make_attn(ch, upsample_side)`` returns the attention layer for ``ch`` channels."""
    time_embed_dim = model_channels * 4
    model.time_embed = Slots({0: ParamLinear(model_channels, time_embed_dim), 2: ParamLinear(time_embed_dim, time_embed_dim)})
    ch = input_ch = int(channel_mult[0] * model_channels)
    model.input_blocks = nn.ModuleList([TimestepEmbedSequential(ParamConv(dims, in_channels, ch, 3, padding=1))])
    chans, ds = [ch], 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=int(mult * model_channels), dims=dims)]
            ch = int(mult * model_channels)
            if ds in attention_resolutions:
                layers.append(make_attn(ch, False))
            model.input_blocks.append(TimestepEmbedSequential(*layers))
            chans.append(ch)
        if level != len(channel_mult) - 1:
            model.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
            chans.append(ch)
            ds *= 2
    model.middle_block = TimestepEmbedSequential(ResBlock(ch, time_embed_dim, dropout, dims=dims), make_attn(ch, False),
                                                 ResBlock(ch, time_embed_dim, dropout, dims=dims))
    model.output_blocks = nn.ModuleList([])
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            layers = [ResBlock(ch + ich, time_embed_dim, dropout, out_channels=int(model_channels * mult), dims=dims)]
            ch = int(model_channels * mult)
            if ds in attention_resolutions:
                layers.append(make_attn(ch, True))
            if level and i == num_res_blocks:
                layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                ds //= 2
            model.output_blocks.append(TimestepEmbedSequential(*layers))
    model.out = Slots({0: ParamNorm(ch), 2: zero_module(ParamConv(dims, input_ch, out_channels, 3, padding=1))})
