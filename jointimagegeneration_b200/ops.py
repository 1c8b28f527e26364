"""Tensor-level wrappers over the C ABI (one function per entry point of include/guidegen_sm100.h).

Conventions: "CL" = channels-last bf16 activation [N, D, H, W, C] (2-D data: D = 1, tokens:
D = H = 1), C a multiple of 8.  Every function enqueues on torch's current CUDA stream and
raises if the CUDA library is missing or a call fails -- there is no fallback path.
"""
import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _C

CAT_POSTERIOR, CAT_SAMPLE, CAT_ARGMAX, CAT_PROBS, CAT_SAMPLE_GIVEN, CAT_ARGMAX_GIVEN, CAT_PROBS_GIVEN = range(7)
BLOCK_K = 64


def _chk(t, dtype=None, contiguous=True):
    _C.require_cuda(t)
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return t


# ------------------------------------------------------------------------- per-voxel kernels
def cat_posterior_sample(x0, xt, coef, mode, q=None, clamp_min=1e-12, out=None, out_i64=None, labels=None,
                         seed=0, offset=0):
    """K11-K13 at the reference interface.  x0/xt fp32 [B, C, *spatial]; q fp32 [B*V, C] or None
    (in-kernel Philox); coef fp32 [B, 2] = (alpha_t, cumalpha_{t-1}).  Returns whichever of
    (out, out_i64, labels) were passed / allocated by mode."""
    _chk(x0, torch.float32)
    B, Cc = x0.shape[:2]
    V = x0[0, 0].numel()
    given = mode >= CAT_SAMPLE_GIVEN
    if not given:
        _chk(xt, torch.float32)
        _chk(coef, torch.float32)
        assert xt.shape == x0.shape and coef.numel() == 2 * B
    if q is not None:
        _chk(q, torch.float32)
        assert q.numel() == B * V * Cc
    if out is None and out_i64 is None and labels is None:
        out = torch.empty_like(x0)
    a = _C.CatArgs(_C.ptr(x0), _C.ptr(xt), _C.ptr(q), _C.ptr(coef), _C.ptr(out), _C.ptr(out_i64), _C.ptr(labels),
                   B, Cc, V, float(clamp_min if clamp_min else 0.0), mode, seed, offset)
    _C.check(_C.lib().gg_cat_posterior_sample(C.byref(a), _C.stream()), "gg_cat_posterior_sample")
    return out, out_i64, labels


def cat_step_cl(logits, labels_in, coef, labels_out, B, V, Cc, mode=CAT_SAMPLE, q=None, cond=None, n_cond=0,
                next_x=None, probs_out=None, clamp_min=1e-12, seed=0, offset=0, vox_base=0):
    """Device-resident sampler-loop form (channels-last)."""
    _chk(logits, torch.float32)
    Cpad = logits.shape[-1]
    Cin_pad = next_x.shape[-1] if next_x is not None else 0
    a = _C.CatStepCLArgs(_C.ptr(logits), _C.ptr(labels_in), _C.ptr(q), _C.ptr(coef), _C.ptr(cond), _C.ptr(labels_out),
                         _C.ptr(next_x), _C.ptr(probs_out), B, Cc, Cpad, n_cond, Cin_pad, V, float(clamp_min), mode,
                         seed, offset, vox_base)
    _C.check(_C.lib().gg_cat_step_cl(C.byref(a), _C.stream()), "gg_cat_step_cl")


def ddim_update(x, e_t, coef, noise=None, temperature=1.0, want_pred_x0=True, x_prev=None, pred_x0=None,
                e_uncond=None, guidance_scale=1.0):
    """K14.  coef: device fp32 [4] = (a_t, a_prev, sigma_t, sqrt(1 - a_t)).  With e_uncond the
    classifier-free-guidance combination e = e_u + s (e_t - e_u) (ddim.py:175-179) is fused in."""
    _chk(x, torch.float32), _chk(e_t, torch.float32), _chk(coef, torch.float32)
    if noise is not None:
        _chk(noise, torch.float32)
    if e_uncond is not None:
        _chk(e_uncond, torch.float32)
    if x_prev is None:
        x_prev = torch.empty_like(x)
    if pred_x0 is None and want_pred_x0:
        pred_x0 = torch.empty_like(x)
    a = _C.DdimArgs(_C.ptr(x), _C.ptr(e_t), _C.ptr(noise), _C.ptr(coef), _C.ptr(x_prev), _C.ptr(pred_x0), x.numel(),
                    float(temperature), _C.ptr(e_uncond), float(guidance_scale))
    _C.check(_C.lib().gg_ddim_update(C.byref(a), _C.stream()), "gg_ddim_update")
    return x_prev, pred_x0


def plms_eps(e_t, olds, order, e_uncond=None, guidance_scale=1.0, e_cur=None, e_prime=None):
    """PLMS combination of noise predictions (plms.py:219-230).  olds = [newest, ..., oldest] previous predictions
    (order 0: [e_t_next]).  Returns (e_cur, e_prime): the (guided) current prediction and the combination."""
    _chk(e_t, torch.float32)
    for o in olds:
        _chk(o, torch.float32)
    assert 0 <= order <= 3 and len(olds) >= max(order, 1)
    if e_uncond is not None:
        _chk(e_uncond, torch.float32)
        if e_cur is None:
            e_cur = torch.empty_like(e_t)
    if e_prime is None:
        e_prime = torch.empty_like(e_t)
    o = list(olds) + [None, None, None]
    a = _C.PlmsArgs(_C.ptr(e_t), _C.ptr(e_uncond), _C.ptr(o[0]), _C.ptr(o[1]), _C.ptr(o[2]), _C.ptr(e_cur), _C.ptr(e_prime),
                    e_t.numel(), int(order), float(guidance_scale))
    _C.check(_C.lib().gg_plms_eps(C.byref(a), _C.stream()), "gg_plms_eps")
    return (e_cur if e_cur is not None else e_t), e_prime


def ddpm_update(x, e_t, coef, noise=None, temperature=1.0, clip_denoised=False, want_x0=False):
    """Ancestral DDPM step (ddpm.py:1060-1120).  coef: device fp32 [B, 6]."""
    _chk(x, torch.float32), _chk(e_t, torch.float32), _chk(coef, torch.float32)
    if noise is not None:
        _chk(noise, torch.float32)
    B = x.shape[0]
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_x0 else None
    a = _C.DdpmArgs(_C.ptr(x), _C.ptr(e_t), _C.ptr(noise), _C.ptr(coef), _C.ptr(x_prev), _C.ptr(x0), B, x[0].numel(),
                    float(temperature), int(clip_denoised))
    _C.check(_C.lib().gg_ddpm_update(C.byref(a), _C.stream()), "gg_ddpm_update")
    return x_prev, x0


def labels_to_mask(labels, fh, fw, divisor=255.0, out=None):
    """uint8 labels [D, H, W] -> fp32 mask [D, H*fh, W*fw] (nearest zoom, / divisor)."""
    _chk(labels, torch.uint8)
    D, H, W = labels.shape
    if out is None:
        out = torch.empty((D, H * fh, W * fw), dtype=torch.float32, device=labels.device)
    _C.check(_C.lib().gg_labels_to_mask(_C.ptr(labels), _C.ptr(out), D, H, W, fh, fw, float(divisor), _C.stream()), "gg_labels_to_mask")
    return out


def zoom_index(n_in: int, n_out: int):
    """Input index of every output index under scipy.ndimage.zoom(order=0, mode='constant', grid_mode=False) -- the call of
    sample_diffusion.py:200: output o samples input coordinate o * (n_in - 1) / (n_out - 1) (corner-aligned) and order 0
    takes floor(c + 0.5).  Double arithmetic as in scipy's NI_ZoomShift; int32 numpy array [n_out]."""
    import numpy as np
    if n_out == 1 or n_in == 1:
        return np.zeros(n_out, dtype=np.int32)
    c = np.arange(n_out, dtype=np.float64) * (np.float64(n_in - 1) / np.float64(n_out - 1))
    return np.clip(np.floor(c + 0.5), 0, n_in - 1).astype(np.int32)


def labels_zoom(labels, out_shape, divisor=255.0, out=None):
    """uint8 labels [D, H, W] -> fp32 [Do, Ho, Wo] = scipy.ndimage.zoom(labels, out_shape / shape, order=0) / divisor."""
    _chk(labels, torch.uint8)
    D, H, W = labels.shape
    Do, Ho, Wo = (int(v) for v in out_shape)
    idx = [torch.from_numpy(zoom_index(n, m)).to(labels.device) for n, m in ((D, Do), (H, Ho), (W, Wo))]
    if out is None:
        out = torch.empty((Do, Ho, Wo), dtype=torch.float32, device=labels.device)
    _C.check(_C.lib().gg_labels_gather(_C.ptr(labels), _C.ptr(out), D, H, W, Do, Ho, Wo, _C.ptr(idx[0]), _C.ptr(idx[1]), _C.ptr(idx[2]),
                                       float(divisor), _C.stream()), "gg_labels_gather")
    return out


def minmax_normalize(x, y_view, scratch):
    """y_view[b] = (x[b] - x.min()) / (x.max() - x.min()); y_view: [B, ...] whose rows are contiguous blocks."""
    _chk(x, torch.float32)
    B = x.shape[0]
    per = x[0].numel()
    assert y_view.dtype == torch.float32 and y_view[0].is_contiguous() and y_view[0].numel() == per
    _C.check(_C.lib().gg_minmax_normalize(_C.ptr(x), _C.ptr(y_view), _C.ptr(scratch), B, per, y_view.stride(0), _C.stream()),
             "gg_minmax_normalize")
    return y_view


def nchw_to_cl(x1, x2=None, c_pad=None, out=None):
    """fp32 [N, C1, *sp] (+ [N, C2, *sp]) -> CL bf16 [N, *sp3, Cpad] with zero-filled padding."""
    _chk(x1, torch.float32)
    N, C1 = x1.shape[:2]
    sp = tuple(x1.shape[2:])
    V = int(math.prod(sp))
    C2 = 0
    if x2 is not None:
        _chk(x2, torch.float32)
        C2 = x2.shape[1]
        assert tuple(x2.shape[2:]) == sp
    Cp = c_pad if c_pad is not None else (C1 + C2 + 7) // 8 * 8
    if out is None:
        out = torch.empty((N,) + _sp3(sp) + (Cp,), dtype=torch.bfloat16, device=x1.device)
    _C.check(_C.lib().gg_nchw_to_cl(_C.ptr(x1), C1, _C.ptr(x2), C2, _C.ptr(out), Cp, N, V, _C.stream()), "gg_nchw_to_cl")
    return out


def cl_to_nchw(x_cl, Cc, spatial, softmax=False, out=None):
    """CL (bf16 or fp32) [N, ..., Cstride] -> fp32 [N, C, *spatial] (optional softmax over C)."""
    _C.require_cuda(x_cl)
    N = x_cl.shape[0]
    Cs = x_cl.shape[-1]
    V = int(math.prod(spatial))
    if out is None:
        out = torch.empty((N, Cc) + tuple(spatial), dtype=torch.float32, device=x_cl.device)
    _C.check(_C.lib().gg_cl_to_nchw(_C.ptr(x_cl), Cs, int(x_cl.dtype == torch.float32), _C.ptr(out), Cc, N, V,
                                    int(softmax), _C.stream()), "gg_cl_to_nchw")
    return out


def _sp3(sp: Sequence[int]) -> Tuple[int, int, int]:
    sp = tuple(int(s) for s in sp)
    return (1,) * (3 - len(sp)) + sp


# --------------------------------------------------------------------------------- GroupNorm
def gn_num_chunks(S, Cc):
    return int(_C.lib().gg_gn_num_chunks(S, Cc))


def gn_partial(x_cl, partial=None):
    _chk(x_cl, torch.bfloat16)
    N, Cc = x_cl.shape[0], x_cl.shape[-1]
    S = x_cl[0].numel() // Cc
    nch = gn_num_chunks(S, Cc)
    if partial is None:
        partial = torch.empty((N, nch, Cc, 2), dtype=torch.float32, device=x_cl.device)
    _C.check(_C.lib().gg_gn_partial(_C.ptr(x_cl), N, S, Cc, _C.ptr(partial), _C.stream()), "gg_gn_partial")
    return partial


def gn_finalize(partial1, partial2, gamma, beta, S, eps, groups=32, scale_shift=None):
    N, nch1, C1 = partial1.shape[:3]
    nch2, C2 = (partial2.shape[1], partial2.shape[2]) if partial2 is not None else (0, 0)
    if scale_shift is None:
        scale_shift = torch.empty((N, C1 + C2, 2), dtype=torch.float32, device=partial1.device)
    a = _C.GnFinalizeArgs(_C.ptr(partial1), C1, nch1, _C.ptr(partial2), C2, nch2, _C.ptr(gamma), _C.ptr(beta),
                          _C.ptr(scale_shift), N, groups, S, float(eps))
    _C.check(_C.lib().gg_gn_finalize(C.byref(a), _C.stream()), "gg_gn_finalize")
    return scale_shift


def gn_apply(x1, x2, scale_shift, silu, out=None):
    _chk(x1, torch.bfloat16)
    N, C1 = x1.shape[0], x1.shape[-1]
    S = x1[0].numel() // C1
    C2 = 0
    if x2 is not None:
        _chk(x2, torch.bfloat16)
        C2 = x2.shape[-1]
    if out is None:
        out = torch.empty(tuple(x1.shape[:-1]) + (C1 + C2,), dtype=torch.bfloat16, device=x1.device)
    _C.check(_C.lib().gg_gn_apply(_C.ptr(x1), C1, _C.ptr(x2), C2, _C.ptr(scale_shift), _C.ptr(out), N, S, int(silu),
                                  _C.stream()), "gg_gn_apply")
    return out


def gn_fused(x1, x2, gamma, beta, eps=1e-5, silu=False, groups=32, out=None):
    """GroupNorm32 over th.cat([x1, x2], channel) (+SiLU) in one launch (small, L2-resident tensors)."""
    _chk(x1, torch.bfloat16)
    N, C1 = x1.shape[0], x1.shape[-1]
    S = x1[0].numel() // C1
    C2 = 0
    if x2 is not None:
        _chk(x2, torch.bfloat16)
        C2 = x2.shape[-1]
    if out is None:
        out = torch.empty(tuple(x1.shape[:-1]) + (C1 + C2,), dtype=torch.bfloat16, device=x1.device)
    _C.check(_C.lib().gg_gn_fused(_C.ptr(x1), C1, _C.ptr(x2), C2, _C.ptr(gamma), _C.ptr(beta), _C.ptr(out), N, S, groups, float(eps),
                                  int(silu), _C.stream()), "gg_gn_fused")
    return out


def group_norm_cl(x1, x2, gamma, beta, eps=1e-5, silu=False, groups=32, out=None):
    """GroupNorm32 over th.cat([x1, x2], channel) (+SiLU), statistics in fp32/fp64."""
    p1 = gn_partial(x1)
    p2 = gn_partial(x2) if x2 is not None else None
    S = x1[0].numel() // x1.shape[-1]
    ss = gn_finalize(p1, p2, gamma, beta, S, eps, groups)
    return gn_apply(x1, x2, ss, silu, out)


# ------------------------------------------------------------------------------ convolution
def conv_block_n(cout):
    return int(_C.lib().gg_conv_pick_block_n(cout))


def pack_conv_weight(w: torch.Tensor, splits: Sequence[int], centre_only: Sequence[bool] = None,
                     extra: Sequence[torch.Tensor] = (), chunk_major: bool = False) -> torch.Tensor:
    """torch conv / linear weight [Cout, Cin, *k] -> bf16 [Cout, Ktot] in the library's K order
    (source -> tap -> 64-channel chunk -> channel; include/guidegen_sm100.h).  `splits` = channels
    of each full-filter source (sum = Cin); `extra` = 1x1 weights [Cout, Ci(, 1...)] of
    centre-only sources appended after them."""
    Cout = w.shape[0]
    w = w.detach().float().reshape(Cout, w.shape[1], -1)          # [Cout, Cin, taps]
    cols = []
    c0 = 0
    for cs in splits:
        ws = w[:, c0:c0 + cs]                                     # [Cout, cs, taps]
        c0 += cs
        nch = (cs + BLOCK_K - 1) // BLOCK_K
        ws = ws.permute(0, 2, 1)                                  # [Cout, taps, cs]
        ws = torch.nn.functional.pad(ws, (0, nch * BLOCK_K - cs))
        if chunk_major:                                           # algo 1: source -> chunk -> tap -> channel
            ws = ws.reshape(Cout, ws.shape[1], nch, BLOCK_K).permute(0, 2, 1, 3)
        cols.append(ws.reshape(Cout, -1))
    assert c0 == w.shape[1]
    for e in extra:
        e = e.detach().float().reshape(Cout, -1)
        cs = e.shape[1]
        nch = (cs + BLOCK_K - 1) // BLOCK_K
        cols.append(torch.nn.functional.pad(e, (0, nch * BLOCK_K - cs)))
    return torch.cat(cols, dim=1).to(torch.bfloat16).contiguous()


def pad_vec(v: Optional[torch.Tensor], n: int) -> Optional[torch.Tensor]:
    """fp32 vector zero-padded to a multiple of 8 entries (bias / emb rows read 8 at a time)."""
    if v is None:
        return None
    v = v.detach().float()
    n8 = (n + 7) // 8 * 8
    if v.shape[-1] != n8:
        v = torch.nn.functional.pad(v, (0, n8 - v.shape[-1]))
    return v.contiguous()


def make_conv_args(srcs, w_packed, cout, y, *, dims, ksize=3, stride=1, bias=None, emb=None, residual=None,
                   taps=None, offsets=None, out_spatial=None, y_strides=None, block_n=0, brick=None, d_shift=0,
                   y_f32=False, algo=0, split_k=0, workspace=None, src_ss=None, ss_stride=0, xf_silu=True,
                   xf_z=None) -> _C.ConvArgs:
    """srcs: list of (CL tensor [N, D, H, W, C], centre_only).  y: CL tensor [N, Do, Ho, Wo, >= Cout8]
    (bf16 or fp32).  src_ss (algo 4): per source, None or the device address of its first (scale, shift) pair in
    gg_gn_finalize's output -- the kernel then normalises (+ SiLU) that source itself; ss_stride = floats per
    sample; xf_z = (lo, hi) local depth planes holding real data (default 0..D).
    Returns the filled gg_conv_args (keeps nothing alive: the caller owns the tensors)."""
    a = _C.ConvArgs()
    x0 = srcs[0][0]
    N, D, H, W = x0.shape[:4]
    a.nsrc = len(srcs)
    for i, (t, centre) in enumerate(srcs):
        _chk(t, torch.bfloat16)
        assert tuple(t.shape[:4]) == (N, D, H, W)
        a.src[i].x = _C.ptr(t)
        a.src[i].C = t.shape[-1]
        a.src[i].centre_only = int(bool(centre))
        a.src[i].d_shift = d_shift
    a.N, a.D, a.H, a.W, a.dims = N, D, H, W, dims
    if taps is None:
        k = ksize
        taps = (k if dims >= 3 else 1, k if dims >= 2 else 1, k)
    if offsets is None:
        offsets = tuple(-(t // 2) for t in taps)
    a.kd, a.kh, a.kw = taps
    a.od, a.oh, a.ow = offsets
    a.stride = stride
    if out_spatial is None:
        if stride == 1:
            out_spatial = (D, H, W)
        else:
            f = lambda n, on: (n - 1) // 2 + 1 if on else n
            out_spatial = (f(D, dims >= 3), f(H, dims >= 2), f(W, True))
    a.Do, a.Ho, a.Wo = out_spatial
    a.w_packed = _C.ptr(_chk(w_packed, torch.bfloat16))
    a.bias = bias if isinstance(bias, int) else _C.ptr(bias)
    a.emb = _C.ptr(emb)
    a.emb_stride = emb.shape[-1] if emb is not None else 0
    a.residual = residual if isinstance(residual, int) else _C.ptr(residual)
    a.res_stride = 0 if residual is None else (cout + 7) // 8 * 8 if isinstance(residual, int) else residual.shape[-1]
    a.y = y if isinstance(y, int) else _C.ptr(y)
    if y_strides is None:
        cs = (cout + 7) // 8 * 8 if isinstance(y, int) else y.shape[-1]
        Do, Ho, Wo = out_spatial
        y_strides = (Do * Ho * Wo * cs, Ho * Wo * cs, Wo * cs, cs)
    a.y_sn, a.y_sd, a.y_sh, a.y_sw = y_strides
    a.y_is_f32 = int(y_f32) if isinstance(y, int) else int(y.dtype == torch.float32)
    a.Cout = cout
    a.algo = algo
    a.split_k = split_k
    a.workspace = _C.ptr(workspace)
    if src_ss is not None:
        for i, v in enumerate(src_ss):
            a.src_ss[i] = v if (v is None or isinstance(v, int)) else _C.ptr(v)
        a.ss_stride, a.xf_silu = int(ss_stride), int(bool(xf_silu))
    a.xf_z_lo, a.xf_z_hi = xf_z if xf_z is not None else (0, D)
    a.block_n = block_n
    if brick is not None:
        for i in range(4):
            a.brick[i] = brick[i]
    return a


def conv_packed_k(a: _C.ConvArgs) -> int:
    return int(_C.lib().gg_conv_packed_k(C.byref(a)))


def conv_fwd(a: _C.ConvArgs):
    _C.check(_C.lib().gg_conv_fwd(C.byref(a), _C.stream()), "gg_conv_fwd")


def upsample2x(x_cl, dims, out=None):
    _chk(x_cl, torch.bfloat16)
    N, D, H, W, Cc = x_cl.shape
    fd, fh = (2 if dims >= 3 else 1), (2 if dims >= 2 else 1)
    if out is None:
        out = torch.empty((N, D * fd, H * fh, W * 2, Cc), dtype=torch.bfloat16, device=x_cl.device)
    _C.check(_C.lib().gg_upsample2x(_C.ptr(x_cl), _C.ptr(out), N, D, H, W, Cc, dims, _C.stream()), "gg_upsample2x")
    return out


# -------------------------------------------------------------------------------- attention
def make_attn_args(q, k, v, o, B, H, Tq, Tk, d, scale, q_str, k_str, v_str, o_str) -> _C.AttnArgs:
    """*_str = (batch stride, row stride, head stride) in elements."""
    return _C.AttnArgs(_C.ptr(q), _C.ptr(k), _C.ptr(v), _C.ptr(o), q_str[0], k_str[0], v_str[0], o_str[0],
                       q_str[1], k_str[1], v_str[1], o_str[1], q_str[2], k_str[2], v_str[2], o_str[2],
                       B, H, Tq, Tk, d, float(scale), None, 0)


def attention_workspace(a: _C.AttnArgs, device) -> Optional[torch.Tensor]:
    """Scratch the tensor-core attention kernel wants for this shape (V^T staging), attached to `a`; None if it does not
    take the shape.  The caller keeps the tensor alive until the launch has run."""
    n = int(_C.lib().gg_attention_workspace_bytes(C.byref(a)))
    if n <= 0:
        return None
    ws = torch.empty(n, dtype=torch.uint8, device=device)
    a.workspace, a.workspace_bytes = ws.data_ptr(), n
    return ws


def attention_fwd(a: _C.AttnArgs, device=None):
    ws = None
    if device is not None and not a.workspace:
        ws = attention_workspace(a, device)
    _C.check(_C.lib().gg_attention_fwd(C.byref(a), _C.stream()), "gg_attention_fwd")
    return ws


def attention_legacy(qkv_cl, n_heads, out=None):
    """QKVAttentionLegacy on the 1x1 qkv conv output in CL form [B, T, 3*C]; channel order is
    head-major then (q | k | v) (unet.py:353).  Returns CL bf16 [B, T, C]."""
    _chk(qkv_cl, torch.bfloat16)
    B, T, W3 = qkv_cl.shape
    ch = W3 // (3 * n_heads)
    if out is None:
        out = torch.empty((B, T, W3 // 3), dtype=torch.bfloat16, device=qkv_cl.device)
    base = qkv_cl.data_ptr()
    es = 2
    a = _C.AttnArgs(base, base + ch * es, base + 2 * ch * es, _C.ptr(out),
                    T * W3, T * W3, T * W3, T * (W3 // 3), W3, W3, W3, W3 // 3, 3 * ch, 3 * ch, 3 * ch, ch,
                    B, n_heads, T, T, ch, 1.0 / math.sqrt(ch))
    attention_fwd(a, qkv_cl.device)
    return out


# ------------------------------------------------------------------------------ small pieces
def timestep_embedding(t, dim, max_period=10000.0, out=None):
    _chk(t, torch.float32)
    B = t.numel()
    if out is None:
        out = torch.empty((B, dim), dtype=torch.float32, device=t.device)
    _C.check(_C.lib().gg_timestep_embedding(_C.ptr(t), _C.ptr(out), B, dim, float(max_period), _C.stream()),
             "gg_timestep_embedding")
    return out


def small_linear(x, w, b, act_in=False, act_out=False, out=None):
    _chk(x, torch.float32), _chk(w, torch.float32)
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    _C.check(_C.lib().gg_small_linear(_C.ptr(x), _C.ptr(w), _C.ptr(b), _C.ptr(out), M, N, K, int(act_in), int(act_out),
                                      _C.stream()), "gg_small_linear")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _chk(x, torch.bfloat16)
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    if out is None:
        out = torch.empty_like(x)
    _C.check(_C.lib().gg_layernorm(_C.ptr(x), _C.ptr(gamma), _C.ptr(beta), _C.ptr(out), rows, Cc, float(eps), _C.stream()),
             "gg_layernorm")
    return out


def geglu(x, out=None):
    _chk(x, torch.bfloat16)
    inner = x.shape[-1] // 2
    rows = x.numel() // x.shape[-1]
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (inner,), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().gg_geglu(_C.ptr(x), _C.ptr(out), rows, inner, _C.stream()), "gg_geglu")
    return out


def softmax_rows(x, scale=1.0, out=None):
    """softmax(scale * x) over the last axis: fp32 [rows, n] -> bf16 (AttnBlock2d, model.py:250-251)."""
    _chk(x, torch.float32)
    rows, n = x.numel() // x.shape[-1], x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().gg_softmax_rows(_C.ptr(x), _C.ptr(out), rows, n, float(scale), _C.stream()), "gg_softmax_rows")
    return out


def transpose_bf16(x, out=None):
    """bf16 [R, C] -> [C, R]."""
    _chk(x, torch.bfloat16)
    R, Cc = x.shape
    if out is None:
        out = torch.empty((Cc, R), dtype=torch.bfloat16, device=x.device)
    _C.check(_C.lib().gg_transpose_bf16(_C.ptr(x), _C.ptr(out), R, Cc, _C.stream()), "gg_transpose_bf16")
    return out
