"""Drop-in for latentdiffusion/ldm/models/diffusion/plms.py::PLMSSampler (SAMPLING ONLY).

``make_schedule`` (:25-57), ``sample`` (:60-116), ``plms_sampling`` (:119-175) and ``p_sample_plms`` (:178-236)
keep their signatures and return values.  PLMS is the DDIM update at eta = 0 fed with an Adams-Bashforth
combination of the last noise predictions; here the combination (incl. the classifier-free-guidance mix) is one
sm_100a kernel (``gg_plms_eps``) and the update is the DDIM kernel (``gg_ddim_update``), both reading device
buffers, so a step is  UNet forward -> 2 launches  instead of ~20 elementwise torch launches and four
``torch.full`` fills.  Branches ``sample_diffusion.py`` never reaches raise NotImplementedError: inpainting
``mask``, ``quantize_denoised``, ``score_corrector``, ``noise_dropout``, ``ddim_use_original_steps``.
The schedule and its rounding points are DDIMSampler's (same reference helpers, util.py).
"""
import numpy as np
import torch

from .. import ops
from .ddim import DDIMSampler


class PLMSSampler(DDIMSampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")          # plms.py:26-27
        super().make_schedule(ddim_num_steps, ddim_discretize=ddim_discretize, ddim_eta=ddim_eta, verbose=verbose)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, **kwargs):
        if conditioning is not None:
            first = conditioning[list(conditioning.keys())[0]] if isinstance(conditioning, dict) else conditioning
            first = first[0] if isinstance(first, (list, tuple)) else first
            if first.shape[0] != batch_size:
                print(f"Warning: Got {first.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        size = (batch_size,) + tuple(shape)
        if verbose:
            print(f"Data shape for PLMS sampling is {size}")
        return self.plms_sampling(conditioning, size, dims=kwargs.get("dims", 2), callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature, score_corrector=score_corrector,
                                  corrector_kwargs=corrector_kwargs, x_T=x_T, log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, verbose=verbose)

    @torch.no_grad()
    def plms_sampling(self, cond, shape, dims=2, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, verbose=False):
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is not used by the shipped samplers")
        if mask is not None:
            raise NotImplementedError("inpainting mask blend is not used by the shipped samplers")
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device, torch.float32).contiguous()
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {"x_inter": [img], "pred_x0": [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        if verbose:
            print(f"Running PLMS Sampling with {total_steps} timesteps")
        ts_all = torch.from_numpy(np.ascontiguousarray(time_range)).to(device=device, dtype=torch.long)
        old_eps = []
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = ts_all[i].expand(b)
            ts_next = ts_all[min(i + 1, len(time_range) - 1)].expand(b)
            img, pred_x0, e_t = self.p_sample_plms(img, cond, ts, index=index, quantize_denoised=quantize_denoised,
                                                   temperature=temperature, noise_dropout=noise_dropout,
                                                   score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                                   unconditional_guidance_scale=unconditional_guidance_scale,
                                                   unconditional_conditioning=unconditional_conditioning, old_eps=old_eps,
                                                   t_next=ts_next)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates["x_inter"].append(img)
                intermediates["pred_x0"].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_plms(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, old_eps=None, t_next=None):
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("branch not taken by the shipped samplers (see module docstring)")
        old_eps = [] if old_eps is None else old_eps
        guided = not (unconditional_conditioning is None or unconditional_guidance_scale == 1.)

        def model_output(xx, tt):
            # two passes instead of one doubled batch (plms.py:183-188): same arithmetic, half the activation memory
            e_u = self.model.apply_model(xx, tt, unconditional_conditioning) if guided else None
            return self.model.apply_model(xx, tt, c), e_u

        e_t, e_u = model_output(x, t)
        coef = self._coef[index]
        # sigma = 0 (eta = 0): the noise term sigma * noise_like(...) * temperature of :208 vanishes exactly
        if len(old_eps) == 0:
            # first step (:219-223): a DDIM step with e_t, a second prediction at (x_prev, t_next), then the average
            self.noise_fn(x.shape, x.device, repeat_noise)       # drawn (and multiplied by sigma = 0) by the reference: :208
            x_prev0, _ = ops.ddim_update(x, e_t, coef, None, temperature, e_uncond=e_u, guidance_scale=unconditional_guidance_scale)
            e_next, e_un = model_output(x_prev0, t_next)
            if guided:
                e_next, _ = ops.plms_eps(e_next, [e_next], 0, e_uncond=e_un, guidance_scale=unconditional_guidance_scale)
            e_cur, e_prime = ops.plms_eps(e_t, [e_next], 0, e_uncond=e_u, guidance_scale=unconditional_guidance_scale)
        else:
            order = min(len(old_eps), 3)
            e_cur, e_prime = ops.plms_eps(e_t, old_eps[::-1][:order], order, e_uncond=e_u,
                                          guidance_scale=unconditional_guidance_scale)
        # the reference draws noise_like(x.shape) in every get_x_prev_and_pred_x0 call (:208) although sigma = 0: keep
        # torch's RNG stream in the same position so that seeded runs interleave with other draws as the reference's do
        self.noise_fn(x.shape, x.device, repeat_noise)
        x_prev, pred_x0 = ops.ddim_update(x, e_prime, coef, None, temperature)
        return x_prev, pred_x0, e_cur
