"""First-stage autoencoder at inference: drop-in for
``latentdiffusion/ldm/modules/diffusionmodules/model.py::Decoder`` (:524-631) and for
``AutoencoderKL.decode`` (``ldm/models/autoencoder.py:355-359``: ``post_quant_conv`` then the decoder), i.e. what
``decode_first_stage`` (``ddpm.py:717-776``) runs on every generated slice of the ``_ae`` configuration.

Parameter names and shapes are the reference's (``decoder.mid.block_1.norm1.weight``, ``decoder.up.3.block.0.conv1.weight``,
``decoder.up.1.upsample.conv.weight``, ``post_quant_conv.weight`` ...), so the decoder half of an AutoencoderKL checkpoint
loads with ``load_state_dict(strict=False)``.  The arithmetic is planned once per input shape and executed by the same
sm_100a kernels as the denoiser:

* ``ResnetBlock`` (:82-147, no time embedding): GroupNorm(32, eps 1e-6) + swish fused into / in front of the 3x3 conv,
  ``nin_shortcut`` (1x1) accumulated into the second conv's tile, identity shortcut as the epilogue residual;
* ``Upsample`` (:42-58): nearest x2 folded into the conv (four parity-class 2x2 convs with pre-summed weights);
* ``AttnBlock2d`` (:209-261), single head over ALL channels (d = C = 512 at the shipped size): scores = a tcgen05 GEMM
  with the keys as the weight operand (fp32 out), row softmax (``gg_softmax_rows``), values transposed
  (``gg_transpose_bf16``) so that the second product is the same GEMM kernel;
* ``norm_out`` + swish + ``conv_out`` (:621-626).

``Encoder`` (:398-520) mirrors it: ``Downsample`` (:61-80, zero pad (0, 1, 0, 1) + pad-0 stride-2 conv) is the stride-2 conv
kernel with tap offset 0; ``AutoencoderKL.encode`` returns the reference's ``DiagonalGaussianDistribution`` (mean / logvar /
``mode()`` / ``sample()``) of ``quant_conv(encoder(x))``.  There is no CPU / PyTorch fallback.
"""
from typing import Dict, Optional

import torch
from torch import nn

from .. import _C, ops
from ..unet_engine import Act, Plan, UNetEngine, _Arena
from ..unet_modules import ParamConv, ParamNorm, Upsample as _UpsampleParams


class ResnetBlock(nn.Module):
    """model.py:82-147 (temb_channels = 0, conv_shortcut = False: the only form the autoencoder builds)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0, dims=2):
        super().__init__()
        if conv_shortcut or temb_channels:
            raise NotImplementedError("conv_shortcut / temb projections are not used by the autoencoder")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.dims = in_channels, out_channels, dims
        self.norm1 = ParamNorm(in_channels, eps=1e-6, groups=32)
        self.conv1 = ParamConv(dims, in_channels, out_channels, 3, padding=1)
        self.norm2 = ParamNorm(out_channels, eps=1e-6, groups=32)
        self.conv2 = ParamConv(dims, out_channels, out_channels, 3, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = ParamConv(dims, in_channels, out_channels, 1)


class AttnBlock2d(nn.Module):
    """model.py:209-235."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = ParamNorm(in_channels, eps=1e-6, groups=32)
        self.q = ParamConv(2, in_channels, in_channels, 1)
        self.k = ParamConv(2, in_channels, in_channels, 1)
        self.v = ParamConv(2, in_channels, in_channels, 1)
        self.proj_out = ParamConv(2, in_channels, in_channels, 1)


class Upsample(nn.Module):
    """model.py:42-58."""

    def __init__(self, in_channels, with_conv, dims=2):
        super().__init__()
        self.with_conv, self.dims, self.in_channels = with_conv, dims, in_channels
        if with_conv:
            self.conv = ParamConv(dims, in_channels, in_channels, 3, padding=1)


class Downsample(nn.Module):
    """model.py:61-80: zero pad (0, 1, 0, 1), then a 3x3 stride-2 conv without padding."""

    def __init__(self, in_channels, with_conv, dims=2):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("avg_pool2d downsampling (resamp_with_conv=False) is not used by the shipped config")
        self.with_conv, self.dims, self.in_channels = with_conv, dims, in_channels
        self.conv = ParamConv(dims, in_channels, in_channels, 3, stride=2, padding=0)


class _Planned(nn.Module):
    """Shared plumbing of Encoder / Decoder: one engine, one plan per input shape, weights re-packed after a load / move."""

    def _init_plumbing(self):
        self._engine: Optional[UNetEngine] = None
        self._plans: Dict[tuple, Plan] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        self._engine = None
        self._plans.clear()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate()
        return r

    def _eng(self) -> UNetEngine:
        if self._engine is None:
            self._engine = UNetEngine(self, 2, 1, -1)
        self._engine.lib = _C.lib()
        return self._engine


class Encoder(_Planned):
    """model.py:398-520.  ``post`` (AutoencoderKL.quant_conv, 1x1) is planned behind conv_out when set."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True,
                 in_channels, resolution, z_channels, double_z=True, use_linear_attn=False, attn_type="vanilla", dims=2,
                 **ignore_kwargs):
        super().__init__()
        if dims != 2 or use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("only the shipped 2-D vanilla-attention encoder is implemented")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = ParamConv(2, in_channels, ch, 3, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock2d(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock2d(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.norm_out = ParamNorm(block_in, eps=1e-6, groups=32)
        self.out_channels = 2 * z_channels if double_z else z_channels
        self.conv_out = ParamConv(2, block_in, self.out_channels, 3, padding=1)
        self.post: Optional[ParamConv] = None
        self._init_plumbing()

    def _build_plan(self, N: int, hw) -> Plan:
        eng = self._eng()
        dev = next(self.parameters()).device
        plan, ar = Plan(), _Arena(dev)
        cpad = (self.in_channels + 7) // 8 * 8
        x_in = torch.zeros((N, 1, hw[0], hw[1], cpad), dtype=torch.bfloat16, device=dev)
        plan.inputs["x"] = x_in
        plan.keep.append(x_in)
        ci = self.conv_in
        bi = eng._vec8((id(ci.bias), "b"), lambda: ci.bias, ci.out_channels)
        h = eng._conv(plan, ar, [(Act(x_in), False)], eng._packer(ci, [cpad]), ci.out_channels, dims=2, bias=_C.ptr(bi), stats=True)
        for i_level in range(self.num_resolutions):
            dn = self.down[i_level]
            for i_block in range(self.num_res_blocks):
                nxt = Decoder._resnet(self, eng, plan, ar, dn.block[i_block], h)
                eng._free(ar, h)
                h = nxt
                if len(dn.attn) > 0:
                    nxt = Decoder._attn(self, eng, plan, ar, dn.attn[i_block], h)
                    eng._free(ar, h)
                    h = nxt
            if i_level != self.num_resolutions - 1:
                dc = dn.downsample.conv
                bd = eng._vec8((id(dc.bias), "b"), lambda dc=dc: dc.bias, dc.out_channels)
                Ho, Wo = (h.sp[1] + 1 - 3) // 2 + 1, (h.sp[2] + 1 - 3) // 2 + 1
                # pad (0, 1, 0, 1) + pad-0 stride-2 conv == stride-2 conv whose taps start at offset 0 (right / bottom
                # zeros come from the TMA out-of-bounds fill)
                nxt = eng._conv(plan, ar, [(h, False)], eng._pack(dc, [h.C]), dc.out_channels, dims=2, stride=2, bias=_C.ptr(bd),
                                offsets=(0, 0, 0), out_spatial=(1, Ho, Wo), stats=True)
                eng._free(ar, h)
                h = nxt
        for blk in (self.mid.block_1, self.mid.attn_1, self.mid.block_2):
            nxt = (Decoder._resnet(self, eng, plan, ar, blk, h) if isinstance(blk, ResnetBlock)
                   else Decoder._attn(self, eng, plan, ar, blk, h))
            eng._free(ar, h)
            h = nxt
        co = self.conv_out
        bo = eng._vec8((id(co.bias), "b"), lambda: co.bias, co.out_channels)
        last = self.post is None
        y = eng._gn_conv(plan, ar, h, None, self.norm_out, True, lambda splits: eng._packer(co, splits), co.out_channels, dims=2,
                         bias=_C.ptr(bo), f32_out=last)
        eng._free(ar, h)
        if not last:       # AutoencoderKL.encode: moments = quant_conv(h)
            pq = self.post
            w = pq.weight.detach().reshape(pq.out_channels, pq.in_channels)
            wpad = torch.zeros((pq.out_channels, y.C), dtype=w.dtype, device=dev)
            wpad[:, :pq.in_channels] = w
            plan.keep.append(wpad)
            bq = eng._vec8((id(pq.bias), "b"), lambda: pq.bias, pq.out_channels)
            wp = eng._cached((id(pq.weight), "pq"), lambda: ops.pack_conv_weight(wpad.reshape(pq.out_channels, -1, 1), [y.C]))
            m = eng._conv(plan, ar, [(y, False)], wp, pq.out_channels, dims=3, ksize=1, bias=_C.ptr(bq), f32_out=True)
            eng._free(ar, y)
            y = m
        plan.outputs["y"] = y.interior
        plan.keep.append(ar.stores)
        plan.arena_bytes = ar.total
        return plan

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:493-520.  x fp32 [N, in_channels, H, W] -> fp32 [N, 2 z_channels, H / 2^(levels-1), W / 2^(levels-1)]."""
        if not x.is_cuda:
            raise RuntimeError("the encoder runs on the sm_100a library only (no CPU fallback)")
        N, hw = x.shape[0], tuple(x.shape[2:])
        key = (N, hw)
        if key not in self._plans:
            self._eng()
            self._plans[key] = self._build_plan(N, hw)
        plan = self._plans[key]
        ops.nchw_to_cl(x.float().contiguous(), None, c_pad=plan.inputs["x"].shape[-1], out=plan.inputs["x"])
        plan.run()
        f = 2 ** (self.num_resolutions - 1)
        cout = self.post.out_channels if self.post is not None else self.out_channels
        return ops.cl_to_nchw(plan.outputs["y"], cout, (hw[0] // f, hw[1] // f), softmax=False)


class DiagonalGaussianDistribution(object):
    """ldm/modules/distributions/distributions.py:24-62 (the members sampling uses; tensors of a few thousand elements)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.parameters.device)

    def mode(self):
        return self.mean


class Decoder(_Planned):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True,
                 in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False, use_linear_attn=False,
                 attn_type="vanilla", dims=2, **ignorekwargs):
        super().__init__()
        if dims != 2 or use_linear_attn or attn_type != "vanilla" or give_pre_end or tanh_out:
            raise NotImplementedError("only the shipped 2-D vanilla-attention decoder is implemented")
        self.ch, self.out_ch, self.num_resolutions, self.num_res_blocks = ch, out_ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.z_channels = resolution, in_channels, z_channels
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = ParamConv(2, z_channels, block_in, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock2d(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock2d(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)        # prepend to get consistent order (:592)
        self.norm_out = ParamNorm(block_in, eps=1e-6, groups=32)
        self.conv_out = ParamConv(2, block_in, out_ch, 3, padding=1)
        self.pre: Optional[ParamConv] = None          # AutoencoderKL.post_quant_conv, planned in front of conv_in
        self._init_plumbing()

    # ------------------------------------------------------------------------------------ planning
    def _resnet(self, eng: UNetEngine, plan, ar, rb: ResnetBlock, x: Act) -> Act:
        cout = rb.out_channels
        b1 = eng._vec8((id(rb.conv1.bias), "b"), lambda: rb.conv1.bias, cout)
        h1 = eng._gn_conv(plan, ar, x, None, rb.norm1, True, lambda splits: eng._packer(rb.conv1, splits), cout, dims=2,
                          bias=_C.ptr(b1), stats=True)
        if rb.in_channels == cout:
            b2 = eng._vec8((id(rb.conv2.bias), "b"), lambda: rb.conv2.bias, cout)
            out = eng._gn_conv(plan, ar, h1, None, rb.norm2, True, lambda splits: eng._packer(rb.conv2, splits), cout, dims=2,
                               bias=_C.ptr(b2), residual=x, stats=True)
        else:
            sk = rb.nin_shortcut
            skw = sk.weight.detach().reshape(cout, -1)
            key = (id(rb.conv2.weight), id(sk.weight))

            def packer(splits):
                return lambda cm: eng._cached(key + (tuple(splits), cm),
                                              lambda: ops.pack_conv_weight(rb.conv2.weight, list(splits), extra=[skw], chunk_major=cm))

            b2 = eng._vec8((id(rb.conv2.bias), id(sk.bias), "b"), lambda: rb.conv2.bias.detach() + sk.bias.detach(), cout)
            out = eng._gn_conv(plan, ar, h1, None, rb.norm2, True, packer, cout, dims=2, extra_srcs=[(x, True)], bias=_C.ptr(b2),
                               stats=True)
        eng._free(ar, h1)
        return out

    def _attn(self, eng: UNetEngine, plan, ar, ab: AttnBlock2d, x: Act) -> Act:
        N, T, Cc = x.N, x.S, x.C
        xn = eng._gn(plan, ar, x, None, ab.norm, False)
        lin = lambda conv: eng._linear(plan, ar, xn, conv.weight.reshape(Cc, Cc), (id(conv.weight), "p"), Cc, conv.bias)   # noqa: E731
        q, k, v = lin(ab.q), lin(ab.k), lin(ab.v)
        ar.release(xn.t)
        o = eng._new_act(ar, N, x.sp, Cc)
        scale = float(int(Cc) ** (-0.5))
        for b in range(N):       # one score matrix at a time (T x T fp32): weights = this sample's keys / values
            qb = Act(q.t[b:b + 1].reshape(1, 1, 1, T, Cc))
            s = eng._new_act(ar, 1, (1, 1, T), T, torch.float32)
            eng._conv(plan, ar, [(qb, False)], k.t[b].reshape(T, Cc), T, dims=3, ksize=1, f32_out=True, out=s)
            pr = eng._new_act(ar, 1, (1, 1, T), T)
            plan.add(eng.lib.gg_softmax_rows, s.ip, pr.ip, T, T, scale)
            ar.release(s.t)
            vt = ar.alloc((Cc, T), torch.bfloat16)
            plan.add(eng.lib.gg_transpose_bf16, _C.ptr(v.t[b]), _C.ptr(vt), T, Cc)
            ob = Act(o.t[b:b + 1].reshape(1, 1, 1, T, Cc))
            eng._conv(plan, ar, [(pr, False)], vt, Cc, dims=3, ksize=1, out=ob)
            ar.release(pr.t)
            plan.keep.append(vt)
            ar.release(vt)
        for t_ in (q, k, v):
            ar.release(t_.t)
        po = ab.proj_out
        out = eng._linear(plan, ar, o, po.weight.reshape(Cc, Cc), (id(po.weight), "p"), Cc, po.bias, x)
        ar.release(o.t)
        return out

    def _build_plan(self, N: int, hw) -> Plan:
        eng = self._eng()
        dev = next(self.parameters()).device
        plan, ar = Plan(), _Arena(dev)
        zc = self.pre.in_channels if self.pre is not None else self.z_channels
        zpad = (zc + 7) // 8 * 8
        z_in = torch.zeros((N, 1, hw[0], hw[1], zpad), dtype=torch.bfloat16, device=dev)
        plan.inputs["z"] = z_in
        plan.keep.append(z_in)
        h = Act(z_in)
        if self.pre is not None:           # AutoencoderKL.decode: z = post_quant_conv(z)
            w = self.pre.weight.detach().reshape(self.pre.out_channels, zc)
            wpad = torch.zeros((self.pre.out_channels, zpad), dtype=w.dtype, device=dev)
            wpad[:, :zc] = w
            plan.keep.append(wpad)
            h = eng._linear(plan, ar, h, wpad, (id(self.pre.weight), "pq"), self.pre.out_channels, self.pre.bias)
        ci = self.conv_in
        bi = eng._vec8((id(ci.bias), "b"), lambda: ci.bias, ci.out_channels)
        h0 = eng._conv(plan, ar, [(h, False)], eng._packer(ci, [h.C]), ci.out_channels, dims=2, bias=_C.ptr(bi), stats=True)
        ar.release(h.t)
        h = h0
        for blk in (self.mid.block_1, self.mid.attn_1, self.mid.block_2):
            nxt = self._resnet(eng, plan, ar, blk, h) if isinstance(blk, ResnetBlock) else self._attn(eng, plan, ar, blk, h)
            eng._free(ar, h)
            h = nxt
        for i_level in reversed(range(self.num_resolutions)):
            up = self.up[i_level]
            for i_block in range(self.num_res_blocks + 1):
                nxt = self._resnet(eng, plan, ar, up.block[i_block], h)
                eng._free(ar, h)
                h = nxt
                if len(up.attn) > 0:
                    nxt = self._attn(eng, plan, ar, up.attn[i_block], h)
                    eng._free(ar, h)
                    h = nxt
            if i_level != 0:
                us = up.upsample
                shim = _UpsampleParams.__new__(_UpsampleParams)       # the engine's view of an upsample: fields only
                nn.Module.__init__(shim)
                shim.channels = shim.out_channels = us.in_channels
                shim.use_conv, shim.dims = us.with_conv, 2
                if us.with_conv:
                    shim.conv = us.conv
                plan.keep.append(shim)
                nxt = eng._upsample(plan, ar, shim, h)
                eng._free(ar, h)
                h = nxt
        co = self.conv_out
        bo = eng._vec8((id(co.bias), "b"), lambda: co.bias, co.out_channels)
        y = eng._gn_conv(plan, ar, h, None, self.norm_out, True, lambda splits: eng._packer(co, splits), co.out_channels, dims=2,
                         bias=_C.ptr(bo), f32_out=True)
        eng._free(ar, h)
        plan.outputs["y"] = y.interior
        plan.keep.append(ar.stores)
        plan.arena_bytes = ar.total
        return plan

    @torch.no_grad()
    def forward(self, z: torch.Tensor) -> torch.Tensor:
        """model.py:598-631.  z fp32 [N, z_channels, h, w] -> fp32 [N, out_ch, 2^(levels-1) h, 2^(levels-1) w]."""
        if not z.is_cuda:
            raise RuntimeError("the decoder runs on the sm_100a library only (no CPU fallback)")
        self.last_z_shape = z.shape
        N, hw = z.shape[0], tuple(z.shape[2:])
        key = (N, hw)
        if key not in self._plans:
            self._plans[key] = self._build_plan(N, hw)
        plan = self._plans[key]
        ops.nchw_to_cl(z.float().contiguous(), None, c_pad=plan.inputs["z"].shape[-1], out=plan.inputs["z"])
        plan.run()
        f = 2 ** (self.num_resolutions - 1)
        return ops.cl_to_nchw(plan.outputs["y"], self.out_ch, (hw[0] * f, hw[1] * f), softmax=False)


class AutoencoderKL(nn.Module):
    """Inference surface of ldm/models/autoencoder.py::AutoencoderKL (:304-371): ``encode(x)`` = posterior of
    ``quant_conv(encoder(x))`` (:350-354), ``decode(z)`` = ``decoder(post_quant_conv(z))`` (:355-359).
    Constructor arguments other than ``ddconfig`` / ``embed_dim`` are accepted and ignored (losses, checkpoints, keys)."""

    def __init__(self, ddconfig, embed_dim, lossconfig=None, ckpt_path=None, ignore_keys=(), image_key="image", colorize_nlabels=None,
                 monitor=None, dims=2, **kwargs):
        super().__init__()
        assert ddconfig["double_z"]
        self.embed_dim = embed_dim
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = ParamConv(2, 2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = ParamConv(2, embed_dim, ddconfig["z_channels"], 1)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.encoder.invalidate(), self.decoder.invalidate()
        return r

    @torch.no_grad()
    def encode(self, x):
        # the 1x1 quant_conv is planned behind conv_out (object.__setattr__: not a second registration)
        if self.encoder.post is None:
            object.__setattr__(self.encoder, "post", self.quant_conv)
            self.encoder.invalidate()
        return DiagonalGaussianDistribution(self.encoder(x))

    @torch.no_grad()
    def decode(self, z):
        # the 1x1 post_quant_conv is planned in front of conv_in (object.__setattr__: not a second registration)
        if self.decoder.pre is None:
            object.__setattr__(self.decoder, "pre", self.post_quant_conv)
            self.decoder.invalidate()
        return self.decoder(z)
