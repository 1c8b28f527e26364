"""Drop-in for ccdm/ddpm/models/one_hot_categorical.py::OneHotCategoricalBCHW (:10-54).

The reference subclasses torch.distributions.OneHotCategorical: ``sample()`` permutes to
channels-last, normalises ``probs / probs.sum(-1)`` (Categorical.__init__), draws with
``torch.multinomial(probs_2d, 1, True)`` -- which is ``argmax(p / q)`` with ``q ~ Exp(1)`` drawn as
one ``[rows, C]`` block -- one-hot encodes and permutes back.  Here one fused sm_100a kernel does
normalise + argmax(p / q) + one-hot directly in the [B, C, *spatial] layout.

Noise: by default ``q`` is drawn with ``torch.empty(rows, C).exponential_()`` from torch's
generator on the tensor's device, i.e. the same draw, in the same order, the reference's
``torch.multinomial`` makes -- a seeded reference run and a seeded run of this class consume the
RNG stream identically.  Pass ``q=`` to inject noise, or ``rng="philox"`` for the in-kernel
counter-based generator (no noise tensor in HBM).
"""
from typing import Optional

import torch

from .. import ops


class OneHotCategoricalBCHW:
    def __init__(self, probs: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None, validate_args=None):
        if (probs is None) == (logits is None):
            raise ValueError("Either `probs` or `logits` must be specified, but not both.")
        if probs is not None and probs.ndim < 2:
            raise ValueError("`probs.ndim` should be at least 2")
        if logits is not None and logits.ndim < 2:
            raise ValueError("`logits.ndim` should be at least 2")
        if logits is not None:
            # only used to draw x_T once per volume (evaluator.py:135-136), outside the denoising loop
            probs = torch.softmax(logits.float(), dim=1)
        self._p = probs.float().contiguous()       # un-normalised, class axis at dim 1

    @property
    def probs(self):
        """Normalised probabilities, channels-last (what Categorical.probs holds in the reference)."""
        return self.prob_sample().permute((0,) + tuple(range(2, self._p.ndim)) + (1,))

    def sample(self, sample_shape=torch.Size(), q: Optional[torch.Tensor] = None, rng: str = "torch", seed: int = 0,
               offset: int = 0) -> torch.Tensor:
        if len(sample_shape) != 0:
            raise NotImplementedError("sample_shape != () is not used by the reference sampler")
        p = self._p
        B, C = p.shape[:2]
        V = p[0, 0].numel()
        if q is None and rng == "torch":
            q = torch.empty((B * V, C), dtype=torch.float32, device=p.device).exponential_(1)
        out, _, _ = ops.cat_posterior_sample(p, None, None, ops.CAT_SAMPLE_GIVEN, q=q, clamp_min=0.0, seed=seed, offset=offset)
        return out

    def max_prob_sample(self) -> torch.Tensor:
        """one_hot(argmax) as int64, like F.one_hot in the reference (:42-46)."""
        o64 = torch.empty(self._p.shape, dtype=torch.int64, device=self._p.device)
        ops.cat_posterior_sample(self._p, None, None, ops.CAT_ARGMAX_GIVEN, clamp_min=0.0, out_i64=o64)
        return o64

    def prob_sample(self) -> torch.Tensor:
        out = torch.empty_like(self._p)
        ops.cat_posterior_sample(self._p, None, None, ops.CAT_PROBS_GIVEN, clamp_min=0.0, out=out)
        return out
