"""LDM DDIM sampler math, restated on CPU (TEST INFRASTRUCTURE).

Restated reference code (paths relative to /root/reference/latentdiffusion):
  ldm/modules/diffusionmodules/util.py:21-43   make_beta_schedule ("linear": f64 linspace of sqrt)
  ldm/models/diffusion/ddpm.py:118-127         register_schedule: alphas_cumprod (f64 cumprod -> f32)
  ldm/modules/diffusionmodules/util.py:46-74   make_ddim_timesteps / make_ddim_sampling_parameters
  ldm/models/diffusion/ddim.py:24-53           DDIMSampler.make_schedule
  ldm/models/diffusion/ddim.py:115-164         ddim_sampling loop
  ldm/models/diffusion/ddim.py:167-205         p_sample_ddim update

dtype notes that matter for bit parity: alphas_cumprod is rounded to f32 when registered;
ddim_alphas is that f32 tensor indexed; ddim_alphas_prev and ddim_sigmas are float64 numpy
arrays built FROM the f32 values; every per-step coefficient then passes through
``torch.full(..., fill)`` and is rounded to f32 again before the fp32 tensor arithmetic.
"""
from typing import Callable, List, Optional

import numpy as np
import torch

Tensor = torch.Tensor
f32 = np.float32


def make_beta_schedule_linear(n_timestep: int, linear_start: float, linear_end: float) -> np.ndarray:
    """util.py:22-25."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def alphas_cumprod_f32(betas: np.ndarray) -> np.ndarray:
    """ddpm.py:125-138: f64 cumprod, stored as an f32 buffer."""
    return np.cumprod(1.0 - betas, axis=0).astype(f32)


def make_ddim_timesteps(num_ddim: int, num_ddpm: int) -> np.ndarray:
    """util.py:46-60, 'uniform': range(0, T, T // S) + 1."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def ddim_tables(acp_f32: np.ndarray, ddim_timesteps: np.ndarray, eta: float):
    """util.py:63-74 + ddim.py:44-47 -> dict of float64 arrays whose f32 rounding is what
    p_sample_ddim feeds to torch.full."""
    a = acp_f32[ddim_timesteps].astype(f32)                           # torch f32 tensor in the reference
    a_prev = np.asarray([float(acp_f32[0])] + [float(v) for v in acp_f32[ddim_timesteps[:-1]]])  # f64 of f32 values
    a64 = a.astype(np.float64)
    # `(1 - alphas_prev) / (1 - alphas)` is ndarray(f64) / Tensor(f32): numpy defers to
    # Tensor.__rtruediv__, which evaluates `(1 - alphas).reciprocal() * other` -- the subtraction and
    # the reciprocal are f32 tensor ops, only the product is f64 (probe-verified, bit-exact).
    recip = (f32(1.0) / (f32(1.0) - a).astype(f32)).astype(f32).astype(np.float64)
    sig = eta * np.sqrt(recip * (1 - a_prev) * (1 - a64 / a_prev))
    # 1 - ddim_alphas is an f32 tensor op, np.sqrt of it stays f32
    s1m = np.sqrt((f32(1.0) - a).astype(f32)).astype(f32)
    return {"alphas": a.astype(np.float64), "alphas_prev": a_prev, "sigmas": sig,
            "sqrt_one_minus_alphas": s1m.astype(np.float64)}


def ddim_update(x: np.ndarray, e_t: np.ndarray, a_t, a_prev, sigma_t, sqrt_one_minus_at, noise: Optional[np.ndarray],
                temperature: float = 1.0):
    """ddim.py:190-205 in float32, one op per line (no FMA):
        pred_x0 = (x - s1m * e) / sqrt(a_t)
        dir_xt  = sqrt(1 - a_prev - sigma^2) * e
        x_prev  = sqrt(a_prev) * pred_x0 + dir_xt + sigma * noise * temperature
    """
    a_t, a_prev, sigma_t, s1m = f32(a_t), f32(a_prev), f32(sigma_t), f32(sqrt_one_minus_at)
    x = x.astype(f32)
    e = e_t.astype(f32)
    pred_x0 = ((x - (s1m * e).astype(f32)).astype(f32) / np.sqrt(a_t)).astype(f32)
    c2 = np.sqrt(f32(f32(f32(1.0) - a_prev) - f32(sigma_t * sigma_t)))
    dir_xt = (c2 * e).astype(f32)
    n = np.zeros_like(x) if noise is None else noise.astype(f32)
    nz = ((sigma_t * n).astype(f32) * f32(temperature)).astype(f32)
    x_prev = (((np.sqrt(a_prev) * pred_x0).astype(f32) + dir_xt).astype(f32) + nz).astype(f32)
    return x_prev, pred_x0


@torch.no_grad()
def ddim_sample(eps_fn: Callable[[Tensor, Tensor], Tensor], acp_f32: np.ndarray, x_T: Tensor, S: int, eta: float = 0.0,
                noises: Optional[List[Tensor]] = None, temperature: float = 1.0, record=None) -> Tensor:
    """ddim.py:115-164 with injected per-step noise.  eps_fn(x, t_long[B]) -> e_t."""
    T = acp_f32.shape[0]
    ts = make_ddim_timesteps(S, T)
    tab = ddim_tables(acp_f32, ts, eta)
    img = x_T.clone()
    b = img.shape[0]
    total = ts.shape[0]
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        e_t = eps_fn(img, t).float()
        nz = None if noises is None else noises[i].numpy()
        x_prev, pred_x0 = ddim_update(img.numpy(), e_t.numpy(), tab["alphas"][index], tab["alphas_prev"][index],
                                      tab["sigmas"][index], tab["sqrt_one_minus_alphas"][index], nz, temperature)
        img = torch.from_numpy(x_prev)
        if record is not None:
            record.append((img.clone(), torch.from_numpy(pred_x0)))
    return img


def plms_combine(e_t: np.ndarray, old_eps) -> np.ndarray:
    """Adams-Bashforth combination of plms.py:224-230 (old_eps = chronological list, newest last), fp32,
    torch's left-to-right evaluation with one rounding per operation."""
    f = np.float32
    if len(old_eps) == 1:
        return ((f(3) * e_t - old_eps[-1]) / f(2)).astype(np.float32)
    if len(old_eps) == 2:
        return ((f(23) * e_t - f(16) * old_eps[-1] + f(5) * old_eps[-2]) / f(12)).astype(np.float32)
    return ((f(55) * e_t - f(59) * old_eps[-1] + f(37) * old_eps[-2] - f(9) * old_eps[-3]) / f(24)).astype(np.float32)


def plms_sample(eps_fn: Callable[[Tensor, Tensor], Tensor], acp_f32: np.ndarray, x_T: Tensor, S: int, record=None) -> Tensor:
    """PLMSSampler.plms_sampling + p_sample_plms (plms.py:119-236) at eta = 0 without guidance: the DDIM update
    (get_x_prev_and_pred_x0, :198-213) fed with the multistep combination; the first step evaluates the model a
    second time at (x_prev, t_next) and averages (:219-223).  `record(i, pred_x0, e_t)` is called per step."""
    ts = make_ddim_timesteps(S, acp_f32.shape[0])
    tab = ddim_tables(acp_f32, ts, 0.0)
    sig, a, ap, s1m = tab["sigmas"], tab["alphas"], tab["alphas_prev"], tab["sqrt_one_minus_alphas"]
    time_range = np.flip(ts)
    x = x_T.clone()
    B = x.shape[0]
    old = []
    for i, step in enumerate(time_range):
        index = len(ts) - i - 1
        t = torch.full((B,), int(step), dtype=torch.long)
        t_next = torch.full((B,), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        e_t = eps_fn(x, t).numpy().astype(np.float32)
        xn = x.numpy()
        if len(old) == 0:
            x_prev0, _ = ddim_update(xn, e_t, a[index], ap[index], sig[index], s1m[index], None)
            e_next = eps_fn(torch.from_numpy(x_prev0), t_next).numpy().astype(np.float32)
            e_prime = ((e_t + e_next) / np.float32(2)).astype(np.float32)
        else:
            e_prime = plms_combine(e_t, old)
        x_prev, pred_x0 = ddim_update(xn, e_prime, a[index], ap[index], sig[index], s1m[index], None)
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)
        if record is not None:
            record(i, pred_x0, e_t)
        x = torch.from_numpy(x_prev)
    return x


def ddpm_tables(betas: np.ndarray):
    """ddpm.py:118-163 -> dict of f32 arrays used by the ancestral sampler."""
    alphas = 1.0 - betas
    acp = np.cumprod(alphas, axis=0)
    acp_prev = np.append(1.0, acp[:-1])
    post_var = betas * (1.0 - acp_prev) / (1.0 - acp)
    return {"sqrt_recip": np.sqrt(1.0 / acp).astype(f32), "sqrt_recipm1": np.sqrt(1.0 / acp - 1).astype(f32),
            "coef1": (betas * np.sqrt(acp_prev) / (1.0 - acp)).astype(f32),
            "coef2": ((1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp)).astype(f32),
            "logvar": np.log(np.maximum(post_var, 1e-20)).astype(f32)}


def ddpm_update(x, e_t, tab, t: int, noise, temperature=1.0, clip=False):
    """ddpm.py:215-230,1060-1120 for one timestep index t (0-based), float32 op by op."""
    x, e = x.astype(f32), e_t.astype(f32)
    x0 = ((tab["sqrt_recip"][t] * x).astype(f32) - (tab["sqrt_recipm1"][t] * e).astype(f32)).astype(f32)
    if clip:
        x0 = np.clip(x0, f32(-1), f32(1))
    mean = ((tab["coef1"][t] * x0).astype(f32) + (tab["coef2"][t] * x).astype(f32)).astype(f32)
    nz = f32(0.0 if t == 0 else 1.0)
    sd = f32(nz * np.exp(f32(0.5) * tab["logvar"][t], dtype=f32))
    n = (noise.astype(f32) * f32(temperature)).astype(f32)
    return (mean + (sd * n).astype(f32)).astype(f32), x0


def sample_cond(eps_fn, acp_f32, wholemask: np.ndarray, n_samples: int, S: int, x_T_fn, eta: float = 0.0) -> np.ndarray:
    """latentdiffusion/sample_diffusion.py:196-224 restated (numpy + ddim_sample above).
    eps_fn(x [n,1,H,W], t, cond [n,2,H,W]) -> eps; wholemask [1,1,D,H,W]; returns [n,2,D,H,W]."""
    _, _, D, H, W = wholemask.shape
    nz = np.where(wholemask.sum((0, 1, 3, 4)))[0]
    start, end = int(nz[0]), int(nz[-1])
    samples = np.zeros((n_samples, 1, D, H, W), dtype=f32)
    gen_mask = np.repeat(wholemask.astype(f32), n_samples, axis=0)
    for m in range(start - 1, end + 1):
        cond = np.concatenate([samples[:, :, max(0, m - 1)], gen_mask[:, :, m]], axis=1)
        ct = torch.from_numpy(cond)
        s = ddim_sample(lambda x, t: eps_fn(x, t, ct), acp_f32, x_T_fn(m), S, eta).numpy()
        samples[:, 0, m] = ((s - s.min()) / (s.max() - s.min())).astype(f32)[:, 0]
    return np.concatenate([samples, gen_mask], axis=1)
