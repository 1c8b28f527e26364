"""Helpers shared by the GPU parity tests: torch fp32 references of the float kernels
(the CPU oracle in oracle/ covers the integer / bit-exact paths and the whole networks)."""
import math

import torch
import torch.nn.functional as F


def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def cl_from_nchw(x):
    """fp32 [N, C, *sp] -> CL bf16 [N, D, H, W, C] (C must already be a multiple of 8)."""
    sp = tuple(x.shape[2:])
    sp3 = (1,) * (3 - len(sp)) + sp
    n, c = x.shape[:2]
    return x.reshape(n, c, *sp3).permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def nchw_from_cl(y, c=None):
    """CL [N, D, H, W, C] -> fp32 [N, C, D, H, W]."""
    y = y.float().permute(0, 4, 1, 2, 3).contiguous()
    return y if c is None else y[:, :c]


def conv_ref(xs_cl, w, bias, dims, stride=1, emb=None, residual_cl=None, extra=()):
    """fp32 reference of gg_conv_fwd: conv over the channel concat of the full-filter sources
    `xs_cl` (CL bf16) with torch weight `w` [Cout, sum C, *k] + 1x1 terms `extra` = [(x_cl, w1x1)]
    + bias + per-sample emb + residual.  bf16-rounded operands, fp32 math.  Returns [N, Cout, Do, Ho, Wo]."""
    no_tf32()
    x = torch.cat([nchw_from_cl(t) for t in xs_cl], 1)
    n = x.shape[0]
    wq = w.to(torch.bfloat16).float()
    k = w.shape[-1]
    if dims == 3:
        y = F.conv3d(x, wq, None, stride=stride, padding=k // 2)
    elif dims == 2:
        y = F.conv2d(x[:, :, 0], wq, None, stride=stride, padding=k // 2)[:, :, None]
    else:
        y = F.conv1d(x[:, :, 0, 0], wq, None, stride=stride, padding=k // 2)[:, :, None, None]
    for (xe, we) in extra:
        xe = nchw_from_cl(xe)
        weq = we.to(torch.bfloat16).float().reshape(we.shape[0], -1)
        if stride == 2:
            xe = xe[:, :, ::2 if dims >= 3 else 1, ::2 if dims >= 2 else 1, ::2]
        y = y + torch.einsum("ncdhw,oc->nodhw", xe, weq)
    if bias is not None:
        y = y + bias.float().reshape(1, -1, 1, 1, 1)
    if emb is not None:
        y = y + emb.float()[:, :y.shape[1]].reshape(n, -1, 1, 1, 1)
    if residual_cl is not None:
        y = y + nchw_from_cl(residual_cl)[:, :y.shape[1]]
    return y


def rel_err(got, want):
    got, want = got.float(), want.float()
    denom = want.abs().max().clamp_min(1e-6)
    return float((got - want).abs().max() / denom)


def attention_ref(q, k, v, scale):
    """q [B, H, Tq, d], k/v [B, H, Tk, d] fp32."""
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * scale
    return torch.einsum("bhqk,bhkd->bhqd", torch.softmax(s, -1), v)


def heads_of(t, H):
    """[B, T, H*d] -> [B, H, T, d] fp32."""
    B, T, W = t.shape
    return t.float().reshape(B, T, H, W // H).permute(0, 2, 1, 3)
