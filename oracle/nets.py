"""Functional fp32 restatement of the two denoiser networks (TEST INFRASTRUCTURE).

The network is *not* rebuilt from constructor arguments: it is walked from the
reference ``state_dict`` keys (``input_blocks.N.M.in_layers.2.weight`` ...), so the
restatement shares no structural logic with the product modules it checks.

Restated reference code (paths relative to /root/reference):
  ccdm/ddpm/models/unet_openai/unet.py:758-823   UNetModel.forward (CCDM, softmax head)
  latentdiffusion/ldm/modules/diffusionmodules/openaimodel.py:713-745  UNetModel.forward (LDM)
  unet.py:242-262 / openaimodel.py:258-278       ResBlock._forward
  unet.py:305-311,343-360                        AttentionBlock, QKVAttentionLegacy
  unet.py:87-146                                 Upsample / Downsample
  nn.py:17-19,103-121                            GroupNorm32, timestep_embedding
  ldm/modules/attention.py:37-64,152-261         GEGLU, FeedForward, CrossAttention,
                                                 BasicTransformerBlock, SpatialTransformer
"""
import math
import re
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def timestep_embedding(timesteps: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """nn.py:103-121 / util.py:151-171 -- [cos | sin] halves, zero pad if dim is odd."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _conv(sd, p, x, stride=1):
    w = sd[p + ".weight"]
    b = sd.get(p + ".bias")
    nd = w.ndim - 2
    k = w.shape[2]
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    return fn(x, w, b, stride=stride, padding=k // 2)


def _gn(sd, p, x, eps=1e-5):
    """GroupNorm32(32, C): statistics in fp32 (nn.py:17-19)."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps).type(x.dtype)


def _linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def resblock(sd, p, x, emb):
    """unet.py:242-262 (no scale-shift norm, no up/down -- the shipped configs)."""
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x)))
    e = _linear(sd, p + ".emb_layers.1", F.silu(emb)).type(h.dtype)
    while e.ndim < h.ndim:
        e = e[..., None]
    h = h + e
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h)))
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x)
    return x + h


def qkv_attention_legacy(qkv: Tensor, n_heads: int) -> Tensor:
    """unet.py:343-360: heads split first, then q|k|v; scale ch^-1/4 on q and on k; fp32 softmax."""
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v)
    return a.reshape(bs, -1, length)


def attention_block(sd, p, x, n_heads):
    """unet.py:305-311."""
    b, c, *spatial = x.shape
    x = x.reshape(b, c, -1)
    qkv = _conv(sd, p + ".qkv", _gn(sd, p + ".norm", x))
    h = qkv_attention_legacy(qkv, n_heads)
    h = _conv(sd, p + ".proj_out", h)
    return (x + h).reshape(b, c, *spatial)


def cross_attention(sd, p, x, context, heads):
    """ldm/modules/attention.py:170-193 (no mask)."""
    q = _linear(sd, p + ".to_q", x)
    ctx = x if context is None else context
    k = _linear(sd, p + ".to_k", ctx)
    v = _linear(sd, p + ".to_v", ctx)
    b, n, inner = q.shape
    d = inner // heads

    def split(t):
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _linear(sd, p + ".to_out.0", out)


def basic_transformer_block(sd, p, x, context, heads):
    """ldm/modules/attention.py:211-215; GEGLU feed-forward :37-64."""
    dim = x.shape[-1]

    def ln(q, t):
        return F.layer_norm(t, (dim,), sd[q + ".weight"], sd[q + ".bias"], 1e-5)

    x = cross_attention(sd, p + ".attn1", ln(p + ".norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", ln(p + ".norm2", x), context, heads) + x
    h = _linear(sd, p + ".ff.net.0.proj", ln(p + ".norm3", x))
    a, gate = h.chunk(2, dim=-1)
    h = _linear(sd, p + ".ff.net.2", a * F.gelu(gate))
    return h + x


def spatial_transformer(sd, p, x, context, heads):
    """ldm/modules/attention.py:248-261.  The reference is 2-D only; the N-d
    generalisation used for the text-conditioned CCDM flattens all spatial axes into the
    token axis, which is what `rearrange('b c h w -> b (h w) c')` does in 2-D."""
    b, c, *spatial = x.shape
    x_in = x
    h = _gn(sd, p + ".norm", x, eps=1e-6)
    w = sd[p + ".proj_in.weight"].reshape(sd[p + ".proj_in.weight"].shape[0], c)
    h = torch.einsum("bc...,oc->bo...", h, w) + sd[p + ".proj_in.bias"].reshape(1, -1, *([1] * len(spatial)))
    inner = h.shape[1]
    h = h.reshape(b, inner, -1).permute(0, 2, 1)
    i = 0
    while (p + f".transformer_blocks.{i}.norm1.weight") in sd:
        h = basic_transformer_block(sd, p + f".transformer_blocks.{i}", h, context, heads)
        i += 1
    h = h.permute(0, 2, 1).reshape(b, inner, *spatial)
    w = sd[p + ".proj_out.weight"].reshape(c, inner)
    h = torch.einsum("bc...,oc->bo...", h, w) + sd[p + ".proj_out.bias"].reshape(1, -1, *([1] * len(spatial)))
    return h + x_in


def preloaded_bert_encoder(sd, x: Tensor, heads: int) -> Tensor:
    """PreloadedBERTEncoder.forward (ccdm/ddpm/models/encoder.py:115-123): x [b, c, l]; depth BasicTransformerBlocks
    without context over the token axis, then inputs + outputs."""
    h = x.transpose(1, 2)
    i = 0
    while f"transformer_blocks.{i}.norm1.weight" in sd:
        h = basic_transformer_block(sd, f"transformer_blocks.{i}", h, None, heads)
        i += 1
    return x + h.transpose(1, 2)


def upsample(sd, p, x):
    """unet.py:105-116: nearest x2 in ALL spatial dims (3-D too), then 3^d conv."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    if (p + ".conv.weight") in sd:
        x = _conv(sd, p + ".conv", x)
    return x


def downsample(sd, p, x):
    """unet.py:135-146: 3^d conv stride 2 pad 1."""
    return _conv(sd, p + ".op", x, stride=2)


def _heads_for(ch, num_heads, num_head_channels):
    return num_heads if num_head_channels == -1 else ch // num_head_channels


def _run_sequential(sd, p, h, emb, context, num_heads, num_head_channels):
    """TimestepEmbedSequential.forward (unet.py:76-84): walk children p.0, p.1, ... by key shape."""
    j = 0
    while True:
        q = f"{p}.{j}"
        if (q + ".in_layers.0.weight") in sd:
            h = resblock(sd, q, h, emb)
        elif (q + ".qkv.weight") in sd:
            h = attention_block(sd, q, h, _heads_for(h.shape[1], num_heads, num_head_channels))
        elif (q + ".proj_in.weight") in sd:
            h = spatial_transformer(sd, q, h, context, _heads_for(h.shape[1], num_heads, num_head_channels))
        elif (q + ".op.weight") in sd:
            h = downsample(sd, q, h)
        elif (q + ".conv.weight") in sd:
            h = upsample(sd, q, h)
        elif (q + ".weight") in sd:
            h = _conv(sd, q, h)
        else:
            break
        j += 1
    if j == 0:
        raise KeyError(f"no layers under {p}")
    return h


def _count(sd, prefix):
    idx = set()
    pat = re.compile(re.escape(prefix) + r"\.(\d+)\.")
    for k in sd:
        m = pat.match(k)
        if m:
            idx.add(int(m.group(1)))
    return max(idx) + 1 if idx else 0


@torch.no_grad()
def unet_forward(sd: Dict[str, Tensor], x: Tensor, timesteps: Tensor, context: Optional[Tensor] = None,
                 input_condition: Optional[Tensor] = None, num_heads: int = 1, num_head_channels: int = 32,
                 softmax_output: bool = False) -> Tensor:
    """unet.py:758-823 (CCDM: input_condition concat + softmax head) and
    openaimodel.py:713-745 (LDM: no concat inside, no softmax)."""
    mc = sd["time_embed.0.weight"].shape[1]
    emb = timestep_embedding(timesteps, mc)
    emb = _linear(sd, "time_embed.2", F.silu(_linear(sd, "time_embed.0", emb)))
    if input_condition is not None:
        x = torch.cat([x, input_condition], dim=1)
    hs = []
    h = x.float()
    for i in range(_count(sd, "input_blocks")):
        h = _run_sequential(sd, f"input_blocks.{i}", h, emb, context, num_heads, num_head_channels)
        hs.append(h)
    h = _run_sequential(sd, "middle_block", h, emb, context, num_heads, num_head_channels)
    for i in range(_count(sd, "output_blocks")):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_sequential(sd, f"output_blocks.{i}", h, emb, context, num_heads, num_head_channels)
    h = _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h)))
    if softmax_output:
        h = torch.softmax(h, dim=1)
    return h
