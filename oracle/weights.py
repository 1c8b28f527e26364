"""Deterministic synthetic weights (TEST INFRASTRUCTURE).

The reference zero-initialises the last conv of every ResBlock, every attention proj_out and
the output conv (unet.py:216-218,300,719), so a random-init network returns a
constant; parity on it would test nothing.  Instead every tensor of a ``state_dict`` is
filled from a numpy ``RandomState`` keyed on (seed, parameter name): stable across machines
and independent of module construction order, so the build container (which loads the tensors
into the *reference* modules to make tests/golden) and the GPU box (which loads them into
the CUDA modules and into oracle.nets) see identical weights without shipping them.
"""
import zlib
from typing import Dict, Sequence

import numpy as np
import torch


def synth_tensor(name: str, shape: Sequence[int], seed: int) -> torch.Tensor:
    rs = np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(int(s) for s in shape)
    if len(shape) <= 1:
        v = rs.standard_normal(shape).astype(np.float32)
        if name.endswith(".weight"):          # GroupNorm / LayerNorm scale
            v = 1.0 + 0.1 * v
        else:                                 # any bias
            v = 0.05 * v
    else:
        fan_in = int(np.prod(shape[1:]))
        v = (rs.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))


def synth_state_dict(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def shapes_of(module_or_sd) -> Dict[str, tuple]:
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, "state_dict") else module_or_sd
    return {k: tuple(v.shape) for k, v in sd.items() if v.dtype.is_floating_point}


def exp_noise(seed: int, shape) -> np.ndarray:
    """Exp(1) noise block for the categorical draw, float32, strictly positive."""
    rs = np.random.RandomState(seed)
    q = rs.standard_exponential(shape).astype(np.float32)
    return np.maximum(q, np.float32(1e-30))


def uniform_one_hot(seed: int, B: int, C: int, spatial) -> torch.Tensor:
    """x_T ~ Uniform over classes, one-hot float32 [B, C, *spatial] (evaluator.py:135-136)."""
    rs = np.random.RandomState(seed)
    idx = rs.randint(0, C, size=(B,) + tuple(spatial))
    return torch.from_numpy(np.moveaxis(np.eye(C, dtype=np.float32)[idx], -1, 1).copy())


def normal(seed: int, shape) -> torch.Tensor:
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def encoder_shapes(dim: int, heads: int, d_head: int, depth: int) -> Dict[str, tuple]:
    """Parameter names / shapes of PreloadedBERTEncoder (ccdm/ddpm/models/encoder.py:103-113: `depth`
    BasicTransformerBlock(dim, heads, d_head) without context, unet_openai/attention.py:127-137; GEGLU feed-forward
    with inner width 4 dim).  Checked against the reference module when it is importable."""
    inner, out = heads * d_head, {}
    for i in range(depth):
        p = f"transformer_blocks.{i}."
        for a in ("attn1", "attn2"):
            for n in ("to_q", "to_k", "to_v"):
                out[p + f"{a}.{n}.weight"] = (inner, dim)
            out[p + f"{a}.to_out.0.weight"] = (dim, inner)
            out[p + f"{a}.to_out.0.bias"] = (dim,)
        out[p + "ff.net.0.proj.weight"] = (8 * dim, dim)
        out[p + "ff.net.0.proj.bias"] = (8 * dim,)
        out[p + "ff.net.2.weight"] = (dim, 4 * dim)
        out[p + "ff.net.2.bias"] = (dim,)
        for n in ("norm1", "norm2", "norm3"):
            out[p + n + ".weight"] = (dim,)
            out[p + n + ".bias"] = (dim,)
    return out


def vae_decoder_shapes(ch: int, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 4, out_ch: int = 1, embed_dim: int = 4,
                       prefix: str = "decoder.") -> Dict[str, tuple]:
    """Parameter names / shapes of AutoencoderKL's decode half (model.py:524-596 Decoder with attn_resolutions = [], plus
    autoencoder.py:318 post_quant_conv).  Checked against the reference Decoder when it is importable."""
    out: Dict[str, tuple] = {}

    def conv(p, ci, co, k):
        out[p + ".weight"] = (co, ci, k, k)
        out[p + ".bias"] = (co,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def res(p, ci, co):
        norm(p + ".norm1", ci), conv(p + ".conv1", ci, co, 3), norm(p + ".norm2", co), conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".nin_shortcut", ci, co, 1)

    nres = len(ch_mult)
    block_in = ch * ch_mult[-1]
    conv(prefix + "conv_in", z_channels, block_in, 3)
    res(prefix + "mid.block_1", block_in, block_in)
    norm(prefix + "mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(prefix + "mid.attn_1." + n, block_in, block_in, 1)
    res(prefix + "mid.block_2", block_in, block_in)
    for i_level in reversed(range(nres)):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks + 1):
            res(prefix + f"up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            conv(prefix + f"up.{i_level}.upsample.conv", block_in, block_in, 3)
    norm(prefix + "norm_out", block_in)
    conv(prefix + "conv_out", block_in, out_ch, 3)
    if embed_dim:
        conv("post_quant_conv", embed_dim, z_channels, 1)
    return out


def vae_encoder_shapes(ch: int, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 4, in_channels: int = 1,
                       embed_dim: int = 4, prefix: str = "encoder.") -> Dict[str, tuple]:
    """Parameter names / shapes of AutoencoderKL's encode half (model.py:398-491 Encoder with attn_resolutions = [],
    double_z, plus autoencoder.py:317 quant_conv).  Checked against the reference Encoder when it is importable."""
    out: Dict[str, tuple] = {}

    def conv(p, ci, co, k):
        out[p + ".weight"] = (co, ci, k, k)
        out[p + ".bias"] = (co,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def res(p, ci, co):
        norm(p + ".norm1", ci), conv(p + ".conv1", ci, co, 3), norm(p + ".norm2", co), conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".nin_shortcut", ci, co, 1)

    conv(prefix + "conv_in", in_channels, ch, 3)
    in_mult = (1,) + tuple(ch_mult)
    block_in = ch
    for i_level in range(len(ch_mult)):
        block_in, block_out = ch * in_mult[i_level], ch * ch_mult[i_level]
        for i_block in range(num_res_blocks):
            res(prefix + f"down.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != len(ch_mult) - 1:
            conv(prefix + f"down.{i_level}.downsample.conv", block_in, block_in, 3)
    res(prefix + "mid.block_1", block_in, block_in)
    norm(prefix + "mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(prefix + "mid.attn_1." + n, block_in, block_in, 1)
    res(prefix + "mid.block_2", block_in, block_in)
    norm(prefix + "norm_out", block_in)
    conv(prefix + "conv_out", block_in, 2 * z_channels, 3)
    if embed_dim:
        conv("quant_conv", 2 * z_channels, 2 * embed_dim, 1)
    return out


def reference_shapes(name: str) -> Dict[str, tuple]:
    """Parameter shapes of a named reference network (oracle/param_shapes.json, written by
    make_golden.py from the reference modules themselves)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "param_shapes.json")) as f:
        return {k: tuple(v) for k, v in json.load(f)[name].items()}
