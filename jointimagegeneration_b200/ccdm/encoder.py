"""Drop-in for ccdm/ddpm/models/encoder.py::PreloadedBERTEncoder (:103-123): the text-context encoder of the
text-conditioned CCDM (built by condition_encoder.py:83-98 as ``feature_cond_encoder``).

Input: cached BERT features [b, embed_dim, length] (the reference's 'b c l' layout); output the same shape,
``inputs + blocks(inputs)`` with ``depth`` BasicTransformerBlocks (self-attention twice -- attn2 has no context --
and a GEGLU feed-forward, unet_openai/attention.py:127-146) over the token axis.  ``state_dict`` keys match the
reference module (``transformer_blocks.N.attn1.to_q.weight`` ...).  Runs once per volume, before the denoising
loop; the blocks execute on the same sm_100a kernels as the UNet's SpatialTransformer (layernorm, tcgen05 GEMMs,
attention, GEGLU) through a planned, CUDA-graph-capturable launch list.  The frozen BERT model itself
(``FrozenBERTEmbedder``, a HuggingFace checkpoint) stays outside: its output is this module's input.
"""
from typing import Dict, Optional

import torch
from torch import nn

from ..unet_engine import Plan, UNetEngine
from ..unet_modules import BasicTransformerBlock


class PreloadedBERTEncoder(nn.Module):
    def __init__(self, embed_dim=768, n_heads=8, depth=4, d_head=64, dropout=.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(embed_dim, n_heads, d_head, dropout=dropout)
                                                 for _ in range(depth)])
        self._engine: Optional[UNetEngine] = None
        self._plans: Dict[tuple, Plan] = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        """Re-pack weights on the next forward (call after changing parameters in place)."""
        self._engine = None
        self._plans.clear()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate()
        return r

    @torch.no_grad()
    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        # inputs: encoded text features of size [b embed_dim length]   (encoder.py:116)
        assert inputs.dim() == 3 and inputs.shape[1] == self.embed_dim, "expected [b, embed_dim, length]"
        if not inputs.is_cuda:
            raise RuntimeError("PreloadedBERTEncoder runs on the sm_100a library only (no CPU fallback)")
        if self._engine is None:
            self._engine = UNetEngine(self, 1, self.transformer_blocks[0].attn1.heads, -1)
        B, Cc, L = inputs.shape
        key = (B, L)
        if key not in self._plans:
            self._plans[key] = self._engine.build_token_plan(B, L, Cc, self.transformer_blocks)
        plan = self._plans[key]
        plan.inputs["tokens"].copy_(inputs.transpose(1, 2).reshape(B, 1, 1, L, Cc))        # 'b c l -> b l c', bf16
        plan.run()
        out = plan.outputs["tokens"].reshape(B, L, Cc).transpose(1, 2).to(inputs.dtype)    # 'b l c -> b c l'
        return inputs + out                                                                  # encoder.py:123
