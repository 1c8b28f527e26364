"""CCDM categorical-diffusion math, restated on CPU (TEST INFRASTRUCTURE).

Restated reference code (paths relative to /root/reference/ccdm/ddpm/models):
  diffusion_denoising.py:18-39     linear_schedule / cosine_schedule
  diffusion_denoising.py:105-139   DiffusionModel.theta_post_prob
  diffusion_denoising.py:176-227   DenoisingModel.forward_denoising
  one_hot_categorical.py:25-54     OneHotCategoricalBCHW.sample / max_prob_sample / prob_sample
  (torch) Categorical.__init__: probs / probs.sum(-1); torch.multinomial(p, 1, True) on one
  row == argmax(p / q), q ~ Exp(1) drawn as a [rows, C] block (probe-verified)

Two statements of the posterior are kept on purpose:
  * ``theta_post_prob_literal`` follows the reference op for op (O(C^2), same torch calls) and
    is bit-identical to the reference on the same machine;
  * ``theta_post_prob_closed`` is the O(C) closed form the CUDA kernel implements, written
    in numpy float32 with a FIXED sequential operation order (no FMA, left-to-right sums)
    so the kernel can match it bit for bit.  tests/ pins closed against literal (<= 1e-6
    abs; the einsum/sum order inside torch is platform-defined, so bit equality between
    those two is not a meaningful requirement).
"""
import math
from typing import Callable, Optional

import numpy as np
import torch

Tensor = torch.Tensor
f32 = np.float32


def linear_schedule(time_steps: int, start=1e-2, end=0.2):
    """diffusion_denoising.py:18-22."""
    betas = torch.linspace(start, end, time_steps)
    alphas = 1 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def cosine_schedule(time_steps: int, s: float = 8e-3):
    """diffusion_denoising.py:25-39.  Quirks kept: ``s`` is overwritten with 0.008; cumalphas is
    the fp32 tensor cos^2 curve (NOT the cumprod of alphas); betas come from Python doubles."""
    t = torch.arange(0, time_steps)
    s = 0.008
    cumalphas = torch.cos(((t / time_steps + s) / (1 + s)) * (math.pi / 2)) ** 2

    def f(u):
        return math.cos((u + s) / (1.0 + s) * math.pi / 2) ** 2

    betas = torch.tensor([min(1 - f((i + 1) / time_steps) / f(i / time_steps), 0.999) for i in range(time_steps)])
    return betas, 1 - betas, cumalphas


def schedule(name: str, time_steps: int, params: Optional[dict] = None):
    fn = {"linear": linear_schedule, "cosine": cosine_schedule}[name]
    return fn(time_steps, **(params or {}))


def step_coefficients(alphas: Tensor, cumalphas: Tensor, t: int):
    """(alpha_t, cumalpha_{t-1}) for 1-based step t, as fp32 scalars
    (diffusion_denoising.py:114-122: t==1 -> alpha=0, cumalpha_tm1=1)."""
    if t == 1:
        return f32(0.0), f32(1.0)
    return f32(alphas[t - 1].item()), f32(cumalphas[t - 2].item())


@torch.no_grad()
def theta_post_prob_literal(alphas: Tensor, cumalphas: Tensor, num_classes: int, xt: Tensor, theta_x0: Tensor,
                            t: Tensor) -> Tensor:
    """diffusion_denoising.py:105-139, same torch ops in the same order (any spatial rank)."""
    t = t - 1
    nsp = xt.ndim - 2
    alphas_t = alphas[t][(...,) + (None,) * (nsp + 1)].clone()
    cumalphas_tm1 = cumalphas[t - 1][(...,) + (None,) * (nsp + 2)].clone()
    alphas_t[t == 0] = 0.0
    cumalphas_tm1[t == 0] = 1.0
    x0 = torch.eye(num_classes, device=xt.device)[(None, slice(None), slice(None)) + (None,) * nsp]
    theta_xt_xtm1 = alphas_t * xt + (1 - alphas_t) / num_classes
    theta_xtm1_x0 = cumalphas_tm1 * x0 + (1 - cumalphas_tm1) / num_classes
    aux = theta_xt_xtm1[:, :, None] * theta_xtm1_x0
    theta_xtm1_xtx0 = aux / aux.sum(dim=1, keepdim=True)
    sp = "lhw"[3 - nsp:]
    return torch.einsum(f"bcd{sp},bd{sp}->bc{sp}", theta_xtm1_xtx0, theta_x0)


def theta_post_prob_closed(a: np.float32, g: np.float32, xt: np.ndarray, x0: np.ndarray) -> np.ndarray:
    """O(C) closed form of the same posterior (derivation: SURVEY.md section 7).

    xt, x0: float32 [C, V] (class-major, V voxels of ONE sample); a = alpha_t, g = cumalpha_{t-1}.
    Operation order is part of the contract with the CUDA kernel:
      k   = (1 - a) / C ;  h = (1 - g) / C
      u_c = a * xt_c + k                         (mul, then add)
      U   = ((u_0 + u_1) + u_2) + ...            (left to right)
      hU  = h * U
      S_d = g * u_d + hU                         (mul, then add)
      r_d = x0_d / S_d
      R   = ((r_0 + r_1) + ...)                  (left to right)
      hR  = h * R
      out_c = u_c * (g * r_c + hR)               (mul, add, mul)
    """
    C = xt.shape[0]
    a, g = f32(a), f32(g)
    k = f32(f32(1) - a) / f32(C)
    h = f32(f32(1) - g) / f32(C)
    u = (a * xt.astype(f32) + k).astype(f32)
    U = u[0].copy()
    for c in range(1, C):
        U = (U + u[c]).astype(f32)
    hU = (h * U).astype(f32)
    S = ((g * u).astype(f32) + hU).astype(f32)
    r = (x0.astype(f32) / S).astype(f32)
    R = r[0].copy()
    for c in range(1, C):
        R = (R + r[c]).astype(f32)
    hR = (h * R).astype(f32)
    return (u * ((g * r).astype(f32) + hR).astype(f32)).astype(f32)


def categorical_sample(probs: np.ndarray, q: np.ndarray, clamp: Optional[float] = None) -> np.ndarray:
    """Label draw for one block of voxels.  probs [C, V] float32, q [V, C] float32 Exp(1) noise
    (channels-last rows, exactly the block torch.multinomial draws).  Returns int64 [V].

      p_c = max(p_c, clamp)                       diffusion_denoising.py:216
      S   = ((p_0 + p_1) + ...)                   Categorical.__init__ normalisation
      idx = first argmax_c (p_c / S) / q_c        torch.multinomial(.., 1, True)
    """
    p = probs.astype(f32)
    if clamp is not None:
        p = np.maximum(p, f32(clamp))
    S = p[0].copy()
    for c in range(1, p.shape[0]):
        S = (S + p[c]).astype(f32)
    ratio = ((p / S).astype(f32) / q.T.astype(f32)).astype(f32)
    return np.argmax(ratio, axis=0)


def near_tie_mask(probs: np.ndarray, q: np.ndarray, rel: float = 1e-5, clamp: Optional[float] = 1e-12) -> np.ndarray:
    """Voxels whose two largest p/q ratios are within `rel` of each other: the only places where
    a label may legitimately differ between two fp32 evaluation orders of the same formula."""
    p = np.maximum(probs.astype(np.float64), clamp or 0.0)
    ratio = p / p.sum(0) / q.T.astype(np.float64)
    top2 = np.sort(ratio, axis=0)[-2:]
    return (top2[1] - top2[0]) <= rel * top2[1]


def one_hot(idx: np.ndarray, C: int) -> np.ndarray:
    """[V] -> float32 [C, V] (F.one_hot(...).to(probs) + channels_second, one_hot_categorical.py:30-50)."""
    out = np.zeros((C, idx.shape[0]), dtype=f32)
    out[idx, np.arange(idx.shape[0])] = 1
    return out


@torch.no_grad()
def forward_denoising(unet_fn: Callable[[Tensor, Tensor], Tensor], alphas: Tensor, cumalphas: Tensor, x_T: Tensor,
                      q_noise: Tensor, time_steps: Optional[int] = None, step_T_sample: str = "majority",
                      t_values=None, record=None):
    """diffusion_denoising.py:176-227 with INJECTED noise.

    unet_fn(xt, t_float[B]) -> x0 probabilities [B, C, *spatial]  (the network is a parameter so
    this loop can drive the restated UNet, the live reference UNet, or the CUDA UNet).
    x_T: one-hot float32 [B, C, *spatial].  q_noise: [n_steps, B*V, C] Exp(1), row order = the
    channels-last flattening torch.multinomial sees.  Returns final xt
    (t>1: float one-hot; last step 'majority': int64 one-hot; 'confidence': probs).
    """
    T = time_steps if time_steps is not None else len(alphas)
    if t_values is None:
        t_values = range(T, 0, -1)
    B, C = x_T.shape[:2]
    spatial = tuple(x_T.shape[2:])
    V = int(np.prod(spatial))
    xt = x_T.clone()
    for i, t in enumerate(t_values):
        t_ = torch.full((B,), t)
        x0pred = unet_fn(xt, t_.float())
        a, g = step_coefficients(alphas, cumalphas, t)
        xt_np = xt.reshape(B, C, V).numpy().astype(f32)
        x0_np = x0pred.reshape(B, C, V).float().numpy()
        probs = np.stack([theta_post_prob_closed(a, g, xt_np[b], x0_np[b]) for b in range(B)])
        probs = np.maximum(probs, f32(1e-12))
        if t > 1:
            q = q_noise[i].numpy().reshape(B, V, C)
            idx = np.stack([categorical_sample(probs[b], q[b]) for b in range(B)])
            xt = torch.from_numpy(np.stack([one_hot(idx[b], C) for b in range(B)])).reshape(B, C, *spatial)
        elif step_T_sample in (None, "majority"):
            # max_prob_sample takes the argmax of the NORMALISED probs (Categorical.__init__)
            S = probs[:, 0].copy()
            for c in range(1, C):
                S = (S + probs[:, c]).astype(f32)
            idx = (probs / S[:, None]).astype(f32).argmax(axis=1)
            xt = torch.from_numpy(np.stack([one_hot(idx[b], C) for b in range(B)])).reshape(B, C, *spatial).long()
        elif step_T_sample == "confidence":
            # OneHotCategoricalBCHW.prob_sample returns the NORMALISED probs (Categorical.__init__)
            S = probs[:, 0].copy()
            for c in range(1, C):
                S = (S + probs[:, c]).astype(f32)
            xt = torch.from_numpy((probs / S[:, None]).astype(f32)).reshape(B, C, *spatial)
        else:
            raise ValueError(step_T_sample)
        if record is not None:
            record.append(xt.clone())
    return xt
