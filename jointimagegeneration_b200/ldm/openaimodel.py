"""LDM denoiser network: drop-in for
latentdiffusion/ldm/modules/diffusionmodules/openaimodel.py::UNetModel (:416-745).

Same constructor arguments, ``state_dict`` keys and ``forward(x, timesteps, context, y)``; the
forward pass is executed by ``UNetEngine`` on the sm_100a kernels.  Both attention flavours are
supported: ``AttentionBlock`` self-attention (the two shipped configs) and
``use_spatial_transformer=True`` cross-attention to ``context`` [B, L, context_dim]
(the 'hybrid' / 'crossattn' conditioning of ddpm.py:1421-1427).
"""
from typing import Optional

import torch
from torch import nn

from .. import ops
from .. import unet_modules as M
from ..unet_engine import UNetEngine


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=3, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        if use_spatial_transformer:
            assert context_dim is not None, "use_spatial_transformer needs context_dim"
        if context_dim is not None:
            assert use_spatial_transformer, "context_dim needs use_spatial_transformer"
            context_dim = list(context_dim) if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig" else context_dim
            if isinstance(context_dim, list):
                raise NotImplementedError("per-block context_dim lists have no consumer in the reference UNet")
        if num_classes is not None or n_embed is not None or resblock_updown or use_scale_shift_norm or use_new_attention_order:
            raise NotImplementedError("option not reachable from any shipped LDM config")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        if num_head_channels == -1:
            assert num_heads != -1, "Either num_heads or num_head_channels has to be set"
        self.image_size, self.in_channels, self.model_channels, self.out_channels = image_size, in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_classes, self.use_checkpoint, self.dtype = num_classes, use_checkpoint, torch.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.predict_codebook_ids = False
        self.dims, self.context_dim = dims, context_dim

        def attn(ch, upsample_side):
            heads = num_heads_upsample if upsample_side else num_heads
            if num_head_channels == -1:
                nh, dh = num_heads, ch // num_heads
            else:
                nh, dh = ch // num_head_channels, num_head_channels
            if legacy:
                dh = ch // nh if use_spatial_transformer else num_head_channels
            if use_spatial_transformer:
                return M.SpatialTransformer(ch, nh, dh, depth=transformer_depth, context_dim=context_dim)
            return M.AttentionBlock(ch, num_heads=heads, num_head_channels=dh)

        M.build_unet_tree(self, dims=dims, in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                          num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions,
                          channel_mult=channel_mult, conv_resample=conv_resample, dropout=dropout, make_attn=attn)
        self._engine: Optional[UNetEngine] = None
        self.use_cuda_graph = False
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    @property
    def engine(self) -> UNetEngine:
        if self._engine is None:
            self._engine = UNetEngine(self, self.dims, self.num_heads, self.num_head_channels)
        return self._engine

    def invalidate(self):
        if self._engine is not None:
            self._engine.invalidate()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate()
        return r

    def plan_for(self, N, spatial, context=None, in_ch=None):
        in_pad = ((in_ch or self.in_channels) + 7) // 8 * 8
        ctx_shape = (context.shape[-2], context.shape[-1]) if (self.context_dim is not None and context is not None) else None
        return self.engine.get_plan(N, tuple(spatial), in_pad, ctx_shape)

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, concat=None, **kwargs):
        """openaimodel.py:713-745.  x fp32 [N, C, *spatial] -> fp32 [N, out_channels, *spatial].
        ``concat`` (extension): a second tensor whose channels follow x's -- the c_concat of
        DiffusionWrapper.forward -- so that torch.cat([x] + c_concat, 1) never materialises."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        N, spatial = x.shape[0], tuple(x.shape[2:])
        cin = x.shape[1] + (concat.shape[1] if concat is not None else 0)
        assert cin == self.in_channels, f"expected {self.in_channels} input channels, got {cin}"
        if self.context_dim is not None and context is None:
            raise ValueError("this UNet was built with use_spatial_transformer: context is required")
        plan = self.plan_for(N, spatial, context)
        if self.use_cuda_graph and plan.graph is None:
            plan.capture()
        ops.nchw_to_cl(x.float().contiguous(), concat.float().contiguous() if concat is not None else None,
                       c_pad=plan.inputs["x"].shape[-1], out=plan.inputs["x"])
        plan.inputs["t"].copy_(timesteps.to(x.device, torch.float32))
        if "context" in plan.inputs:
            c = context.to(x.device, torch.bfloat16).contiguous()
            plan.inputs["context"].copy_(c.reshape(N, 1, 1, c.shape[1], c.shape[2]))
        plan.run()
        return ops.cl_to_nchw(plan.outputs["head"], self.out_channels, spatial, softmax=False)
