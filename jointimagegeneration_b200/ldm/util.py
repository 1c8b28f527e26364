"""Host-side schedule tables of the LDM sampler (float64 numpy, once per ``sample()`` call).
Drop-in for the functions of latentdiffusion/ldm/modules/diffusionmodules/util.py:21-74,264-267
that ``DDIMSampler`` uses.  The dtype of every intermediate matters for bit parity of the
fp32 coefficients that reach the kernels; see the notes on each function."""
import numpy as np
import torch


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3) -> np.ndarray:
    """util.py:21-43 -> float64 numpy betas."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        ab = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        ab = ab / ab[0]
        betas = torch.clamp(1 - ab[1:] / ab[:-1], min=0, max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True) -> np.ndarray:
    """util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ts = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ts = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    out = ts + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {out}")
    return out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:63-74.  ``alphacums`` is the fp32 alphas_cumprod buffer.  In the reference
    ``alphas`` stays an fp32 tensor while ``alphas_prev`` becomes a float64 array, and
    ``ndarray / Tensor`` dispatches to Tensor.__rtruediv__ = ``reciprocal(1 - alphas)`` in fp32
    times the float64 numerator; ``alphas / alphas_prev`` is evaluated in float64.  The same
    rounding points are reproduced here so that the fp32 value of every sigma is identical."""
    ac = np.asarray(alphacums.detach().cpu().numpy() if torch.is_tensor(alphacums) else alphacums, dtype=np.float32)
    alphas = ac[ddim_timesteps]                                                    # fp32
    alphas_prev = np.asarray([float(ac[0])] + [float(v) for v in ac[ddim_timesteps[:-1]]])   # float64 of fp32 values
    one = np.float32(1.0)
    recip = (one / (one - alphas).astype(np.float32)).astype(np.float32).astype(np.float64)
    sigmas = eta * np.sqrt(recip * (1 - alphas_prev) * (1 - alphas.astype(np.float64) / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule {sigmas}")
    return sigmas, alphas, alphas_prev


def noise_like(shape, device, repeat=False):
    """util.py:264-267 -- drawn from torch's generator exactly as the reference does (also when
    sigma == 0, so a seeded run consumes the RNG stream identically)."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)
