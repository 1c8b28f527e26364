"""CCDM denoiser network: drop-in for ccdm/ddpm/models/unet_openai/unet.py::UNetModel (:402-823)
and unet_openai/__init__.py::create_unet_openai (:4-65).

Same constructor arguments, same ``state_dict`` keys and the same ``forward`` signature / return
dict; the forward pass is executed by ``UNetEngine`` on the sm_100a kernels (no torch ops).
Differences from the reference, all in its favour: ``use_spatial_transformer=True`` works
(the reference raises TypeError) and is N-d; ``context`` may be ``[B, L, ctx]``
as CrossAttention expects (D3).  Not implemented because no shipped config reaches them:
feature_cond_encoder concat, class-conditional label_emb, ce_head, scale-shift norm,
resblock up/down, new attention order.
"""
from typing import Optional

import torch
from torch import nn

from .. import ops
from .. import unet_modules as M
from ..unet_engine import UNetEngine


class UNetModel(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, cond_encoded_shape, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, softmax_output=True, ce_head=False,
                 feature_cond_encoder=None, use_spatial_transformer=False, transformer_depth=None, context_dim=None,
                 disabled_sa=False, use_linear_in_transformer=False):
        super().__init__()
        if num_classes is not None or ce_head or feature_cond_encoder is not None or resblock_updown or use_scale_shift_norm:
            raise NotImplementedError("option not reachable from any shipped CCDM config (see module docstring)")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_classes, self.use_checkpoint = num_classes, use_checkpoint
        self.dtype = torch.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.cond_encoded_shape = cond_encoded_shape
        self.sofmtax_output = softmax_output          # (sic) attribute name of the reference, unet.py:486
        self.use_ce_head, self.out_ce = ce_head, None
        self.feature_cond_encoder, self.feature_condition_idx = None, []
        self.dims = dims
        self.context_dim = context_dim

        def attn(ch, upsample_side):
            heads = num_heads_upsample if upsample_side else num_heads
            if num_head_channels == -1:
                nh, dh = heads, ch // heads
            else:
                nh, dh = ch // num_head_channels, num_head_channels
            if use_spatial_transformer:
                return M.SpatialTransformer(ch, nh, dh, depth=transformer_depth, context_dim=context_dim,
                                            disable_self_attn=disabled_sa, use_linear=use_linear_in_transformer)
            return M.AttentionBlock(ch, num_heads=heads, num_head_channels=num_head_channels)

        M.build_unet_tree(self, dims=dims, in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                          num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions,
                          channel_mult=channel_mult, conv_resample=conv_resample, dropout=dropout, make_attn=attn)
        self._engine: Optional[UNetEngine] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # ------------------------------------------------------------------------------------
    @property
    def engine(self) -> UNetEngine:
        if self._engine is None:
            self._engine = UNetEngine(self, self.dims, self.num_heads, self.num_head_channels)
        return self._engine

    def enable_slab(self, comm):
        """Depth-slab decomposition of ONE volume over the ranks of ``comm`` (sharding.SlabComm): every
        forward then takes this rank's slab [1, C, D/world, H, W] and exchanges halos / GroupNorm
        partial sums / attention keys+values with the other ranks (BASELINE config 5).  None disables."""
        self.engine.slab = comm
        self.engine.plans.clear()

    def invalidate(self):
        """Re-pack weights on the next forward (call after changing parameters in place)."""
        if self._engine is not None:
            self._engine.invalidate()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate()
        return r

    @property
    def in_channels_padded(self):
        return (self.in_channels + 7) // 8 * 8

    def _ctx_cl(self, context, N):
        if context is None:
            return None
        if context.ndim != 3:
            raise ValueError("context must be [B, L, context_dim]")
        if context.shape[-1] != self.context_dim and context.shape[1] == self.context_dim:
            context = context.transpose(1, 2)          # dataset layout 'c l' 
        return context.to(torch.bfloat16).contiguous().reshape(N, 1, 1, context.shape[1], context.shape[2])

    def plan_for(self, N, spatial, context=None):
        uses_ctx = any(isinstance(mm, M.SpatialTransformer) for mm in self.modules())
        if uses_ctx and context is not None and context.shape[-1] != self.context_dim and context.shape[1] == self.context_dim:
            context = context.transpose(1, 2)          # dataset layout 'c l' 
        ctx_shape = (context.shape[-2], context.shape[-1]) if (uses_ctx and context is not None) else None
        return self.engine.get_plan(N, tuple(spatial), self.in_channels_padded, ctx_shape)

    @torch.no_grad()
    def forward(self, x, input_condition, feature_condition, timesteps, context=None, y=None):
        """unet.py:758-823.  x fp32 [N, C, *spatial] (+ input_condition [N, 1, *spatial]) ->
        {"diffusion_out": fp32 [N, out_channels, *spatial] (softmax over dim 1), "logits": None}."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        if feature_condition is not None:
            raise NotImplementedError("feature_condition is not used by the shipped ruijin config (params.yml:49)")
        N = x.shape[0]
        spatial = tuple(x.shape[2:])
        plan = self.plan_for(N, spatial, context)
        xin = x.float().contiguous()
        cond = input_condition.float().contiguous() if input_condition is not None else None
        ops.nchw_to_cl(xin, cond, c_pad=self.in_channels_padded, out=plan.inputs["x"])
        plan.inputs["t"].copy_(timesteps.to(plan.inputs["t"].device, torch.float32))
        if "context" in plan.inputs:
            plan.inputs["context"].copy_(self._ctx_cl(context.to(x.device), N))
        plan.run()
        out = ops.cl_to_nchw(plan.outputs["head"], self.out_channels, spatial, softmax=self.sofmtax_output)
        return {"diffusion_out": out.to(x.dtype), "logits": None}


def create_unet_openai(image_size, base_channels, in_channels, out_channels, num_res_blocks, cond_encoded_shape,
                       channel_mult=None, use_checkpoint=False, attention_resolutions=[32, 16, 8], num_heads=1,
                       num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0, resblock_updown=False,
                       use_fp16=False, use_new_attention_order=False, softmax_output=True, ce_head=False,
                       feature_cond_encoder=None, dims=None, use_spatial_transformer=False, transformer_depth=None,
                       context_dim=None):
    """unet_openai/__init__.py:4-65 (+ the cross-attention arguments the reference forgets to forward, D1)."""
    if channel_mult is None:
        defaults = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}
        if image_size not in defaults:
            raise ValueError(f"unsupported image size: {image_size}")
        channel_mult = defaults[image_size]
    if dims not in [1, 2, 3]:
        raise NotImplementedError(f"got convnd dims={dims}")
    return UNetModel(in_channels=in_channels, model_channels=base_channels, out_channels=out_channels,
                     num_res_blocks=num_res_blocks, cond_encoded_shape=cond_encoded_shape,
                     attention_resolutions=attention_resolutions, dropout=dropout, channel_mult=channel_mult, num_classes=None,
                     use_checkpoint=use_checkpoint, use_fp16=use_fp16, num_heads=num_heads, num_head_channels=num_head_channels,
                     num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                     resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order,
                     softmax_output=softmax_output, ce_head=ce_head, feature_cond_encoder=feature_cond_encoder, dims=dims,
                     use_spatial_transformer=use_spatial_transformer, transformer_depth=transformer_depth,
                     context_dim=context_dim)
