"""GuideGen two-stage pipeline (BASELINE config 4): CCDM mask volume -> stage bridge -> autoregressive,
slice-by-slice conditional CT generation with the LDM DDIM sampler.

Drop-in for the inference driver ``sample_cond`` in latentdiffusion/sample_diffusion.py:166-273 (the part
that is arithmetic: :196-224; PNG / NIfTI dumps and metrics are I/O and stay out).  Everything stays on
the device between the 64 x 50 UNet forwards: the label volume never goes through a NIfTI file
(evaluator.py:147-148 -> sample_diffusion.py:199-201), the per-slice min-max normalisation and the
write-back are one fused pair of kernels, and the conditioning of slice m (previous generated slice |
mask slice) is assembled by views.  Slices are strictly sequential (slice m conditions on slice m-1);
parallelism over GPUs comes from independent volumes / samples (``sharding``).
"""
import torch

from . import ops
from .ldm.ddim import DDIMSampler


class GuideGenPipeline:
    def __init__(self, ldm_model, mask_model=None, ddim_steps: int = 50, ddim_eta: float = 0.0):
        self.ldm = ldm_model
        self.mask_model = mask_model
        self.ddim_steps, self.ddim_eta = ddim_steps, ddim_eta
        self.sampler = DDIMSampler(ldm_model)

    # ---- stage 1 -> stage 2 bridge ------------------------------------------------------------
    @torch.no_grad()
    def generate_mask(self, x_T, condition, context=None, init_t=None) -> torch.Tensor:
        """Runs the CCDM chain; returns uint8 labels [B, D, H, W] (argmax of the final one-hot)."""
        t = None if init_t is None else torch.tensor(init_t)
        out = self.mask_model(x_T, condition, t=t, context=context)["diffusion_out"]
        B, C = out.shape[:2]
        labels = torch.empty((B, out[0, 0].numel()), dtype=torch.uint8, device=out.device)
        ops.cat_posterior_sample(out.float().contiguous(), None, None, ops.CAT_ARGMAX_GIVEN, clamp_min=0.0, labels=labels)
        return labels.view((B,) + tuple(out.shape[2:]))

    @torch.no_grad()
    def mask_to_ct_grid(self, labels: torch.Tensor, size=(512, 512), depth=None, rot90_k: int = 0, mode: str = "scipy") -> torch.Tensor:
        """uint8 [D, h, w] -> fp32 whole-mask [1, 1, D', H, W]: the stage bridge of sample_diffusion.py:199-201,
        ``rot90(zoom(labels, out / in, order=0), k=3, dims=(1, 2)) / 255``.  mode "scipy" reproduces scipy.ndimage.zoom's
        corner-aligned nearest rule exactly (ops.zoom_index; pinned against scipy in tests/test_gpu_models.py), "block" is
        plain integer-factor replication.  ``depth`` resamples the slice axis too (the reference zooms to 96 slices)."""
        D, h, w = labels.shape
        Do = D if depth is None else int(depth)
        if mode == "block":
            assert size[0] % h == 0 and size[1] % w == 0 and Do == D, "block replication needs integer in-plane factors"
            m = ops.labels_to_mask(labels.contiguous(), size[0] // h, size[1] // w, 255.0)
        else:
            m = ops.labels_zoom(labels.contiguous(), (Do, size[0], size[1]), 255.0)
        if rot90_k % 4:
            m = torch.rot90(m, k=rot90_k, dims=(1, 2)).contiguous()      # a view permutation of the finished mask (once per volume)
        return m.view(1, 1, Do, m.shape[1], m.shape[2])

    # ---- stage 2 ----------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_cond(self, wholemask: torch.Tensor, n_samples: int = 1, x_T_fn=None) -> torch.Tensor:
        """sample_diffusion.py:196-224.  wholemask fp32 [1, 1, D, H, W] on the device.
        Returns pred = cat([samples, gen_mask], 1): [n, 2, D, H, W].  ``x_T_fn(m)`` optionally supplies the
        initial noise of slice m (tests); default = torch.randn as in DDIMSampler."""
        assert wholemask.shape[0] == 1, "batch size should be 1"
        model, dev = self.ldm, wholemask.device
        _, _, D, H, W = wholemask.shape
        nz = torch.where(wholemask.sum((0, 1, 3, 4)))[0]          # occupied slices (host sync once per volume)
        start_layer, end_layer = int(nz[0]), int(nz[-1])
        if model.no_first_stage:
            shape = (1, H, W)                                      # pixel-space LDM (sample_diffusion.py:203)
        else:                                                      # `_ae` configuration: the sampler works on latents
            dec = model.first_stage_model.decoder
            f = 2 ** (dec.num_resolutions - 1)
            shape = (dec.z_channels, H // f, W // f)
        samples = torch.zeros((n_samples, 1, D, H, W), dtype=torch.float32, device=dev)
        gen_mask = wholemask.repeat(n_samples, 1, 1, 1, 1)
        scratch = torch.empty(1024, dtype=torch.float32, device=dev)
        with model.ema_scope():
            for m_ in range(start_layer - 1, end_layer + 1):
                concat_cond = torch.cat([samples[:, :, max(0, m_ - 1)], gen_mask[:, :, m_]], dim=1)   # [n, 2, H, W]
                c = model.get_learned_conditioning(concat_cond)
                s, _ = self.sampler.sample(S=self.ddim_steps, dims=2, conditioning=c, batch_size=n_samples, shape=shape,
                                           verbose=False, eta=self.ddim_eta, x_T=None if x_T_fn is None else x_T_fn(m_))
                ds = model.decode_first_stage(s)                   # identity for the pixel config (:219)
                ops.minmax_normalize(ds.contiguous(), samples[:, 0, m_], scratch)
        return torch.cat([samples, gen_mask], dim=1)
