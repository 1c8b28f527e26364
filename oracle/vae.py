"""CPU restatement (torch fp32) of the decode side of the first-stage autoencoder -- TEST INFRASTRUCTURE ONLY.

Follows latentdiffusion/ldm/modules/diffusionmodules/model.py: ``nonlinearity`` :33-35 (x * sigmoid(x)), ``Normalize``
:38-39 (GroupNorm 32 groups, eps 1e-6), ``Upsample`` :42-58, ``ResnetBlock.forward`` :120-147, ``AttnBlock2d.forward``
:237-261, ``Decoder.forward`` :598-631, and ``AutoencoderKL.decode`` (ldm/models/autoencoder.py:355-359).
Pinned against the unmodified reference ``Decoder`` (tests/golden/vae_decoder.npz, oracle/make_golden.py).
"""
from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(sd, p, x):
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)        # dropout 0, no temb
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    h = _gn(sd, p + ".norm", x)
    q, k, v = _conv(sd, p + ".q", h, 0), _conv(sd, p + ".k", h, 0), _conv(sd, p + ".v", h, 0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h, 0)


@torch.no_grad()
def decoder_forward(sd: Dict[str, Tensor], z: Tensor, num_resolutions: int, num_res_blocks: int, prefix: str = "") -> Tensor:
    p = prefix
    h = _conv(sd, p + "conv_in", z, 1)
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    for i_level in reversed(range(num_resolutions)):
        for i_block in range(num_res_blocks + 1):
            h = resnet_block(sd, p + f"up.{i_level}.block.{i_block}", h)
            if (p + f"up.{i_level}.attn.{i_block}.norm.weight") in sd:
                h = attn_block(sd, p + f"up.{i_level}.attn.{i_block}", h)
        if i_level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, p + f"up.{i_level}.upsample.conv", h, 1)
    h = _swish(_gn(sd, p + "norm_out", h))
    return _conv(sd, p + "conv_out", h, 1)


@torch.no_grad()
def encoder_forward(sd: Dict[str, Tensor], x: Tensor, num_resolutions: int, num_res_blocks: int, prefix: str = "") -> Tensor:
    """Encoder.forward (model.py:493-520); Downsample :74-78 = zero pad (0, 1, 0, 1) then a pad-0 stride-2 conv."""
    p = prefix
    h = _conv(sd, p + "conv_in", x, 1)
    for i_level in range(num_resolutions):
        for i_block in range(num_res_blocks):
            h = resnet_block(sd, p + f"down.{i_level}.block.{i_block}", h)
            if (p + f"down.{i_level}.attn.{i_block}.norm.weight") in sd:
                h = attn_block(sd, p + f"down.{i_level}.attn.{i_block}", h)
        if i_level != num_resolutions - 1:
            q = p + f"down.{i_level}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[q + ".weight"], sd[q + ".bias"], stride=2)
    h = resnet_block(sd, p + "mid.block_1", h)
    h = attn_block(sd, p + "mid.attn_1", h)
    h = resnet_block(sd, p + "mid.block_2", h)
    h = _swish(_gn(sd, p + "norm_out", h))
    return _conv(sd, p + "conv_out", h, 1)


@torch.no_grad()
def autoencoder_encode_moments(sd: Dict[str, Tensor], x: Tensor, num_resolutions: int, num_res_blocks: int) -> Tensor:
    """AutoencoderKL.encode up to the posterior's parameters (autoencoder.py:350-352): quant_conv(encoder(x));
    mean = first half of the channels (DiagonalGaussianDistribution.mode, distributions.py:27,61-62)."""
    h = encoder_forward(sd, x, num_resolutions, num_res_blocks, prefix="encoder.")
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


@torch.no_grad()
def autoencoder_decode(sd: Dict[str, Tensor], z: Tensor, num_resolutions: int, num_res_blocks: int) -> Tensor:
    """AutoencoderKL.decode: post_quant_conv (1x1) then the decoder; keys as in the reference checkpoint."""
    z = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    return decoder_forward(sd, z, num_resolutions, num_res_blocks, prefix="decoder.")
