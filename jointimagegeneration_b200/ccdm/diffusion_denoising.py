"""Drop-in for ccdm/ddpm/models/diffusion_denoising.py: schedules (:18-39), DiffusionModel
(:42-139) and DenoisingModel (:142-227), with the per-step math on sm_100a kernels.

The reverse chain (forward_denoising, :176-227) runs entirely on the device.  Two execution
forms of the same step are provided:

* ``interface`` (default): exactly the reference's tensor interface per step --
  ``unet(...)["diffusion_out"]`` fp32 [B, C, *sp] -> fused posterior + clamp + draw kernel ->
  fp32 one-hot ``xt``.  Noise ``q ~ Exp(1)`` is drawn per step from torch's generator exactly as
  the reference's ``torch.multinomial`` does (same stream, same order), or injected (``q_noise``).
* ``resident``: the sampler loop never leaves the channels-last device layout -- head-conv
  logits -> fused softmax + posterior + draw -> next UNet input written in place (uint8 labels,
  Philox in-kernel noise).  Only the last step goes through the interface form to produce the
  reference's output tensor.

The posterior uses the O(C) closed form of theta_post_prob -- no
[B, C, C, D, H, W] intermediates.
"""
import logging
import math
from typing import Optional, Union

import numpy as np
import torch
from torch import Tensor, nn

from .. import ops
from .one_hot_categorical import OneHotCategoricalBCHW

LOGGER = logging.getLogger(__name__)


def linear_schedule(time_steps: int, start=1e-2, end=0.2):
    betas = torch.linspace(start, end, time_steps)
    alphas = 1 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def cosine_schedule(time_steps: int, s: float = 8e-3):
    """:25-39.  The argument ``s`` is ignored in the reference as well (overwritten by 0.008)."""
    s = 0.008
    steps = torch.arange(0, time_steps)
    cumalphas = torch.cos(((steps / time_steps + s) / (1 + s)) * (math.pi / 2)) ** 2
    curve = lambda u: math.cos((u + s) / (1.0 + s) * math.pi / 2) ** 2  # noqa: E731
    betas = torch.tensor([min(1 - curve((i + 1) / time_steps) / curve(i / time_steps), 0.999) for i in range(time_steps)])
    return betas, 1 - betas, cumalphas


class DiffusionModel(nn.Module):
    betas: Tensor
    alphas: Tensor
    cumalphas: Tensor

    def __init__(self, schedule: str, time_steps: int, num_classes: int, schedule_params=None, dims=3):
        super().__init__()
        fn = {"linear": linear_schedule, "cosine": cosine_schedule}[schedule]
        betas, alphas, cumalphas = fn(time_steps, **schedule_params) if schedule_params is not None else fn(time_steps)
        self.dims = dims
        self.register_buffer("betas", betas)
        self.register_buffer("alphas", alphas)
        self.register_buffer("cumalphas", cumalphas)
        self.num_classes = num_classes

    @property
    def time_steps(self):
        return len(self.betas)

    # -- forward (noising) process: training-side helpers, not on the sampler hot path ------
    def _bc(self, v: Tensor, extra: int = 0) -> Tensor:
        return v[(...,) + (None,) * (self.dims + 1 + extra)]

    def q_xt_given_xtm1(self, xtm1: Tensor, t: Tensor) -> OneHotCategoricalBCHW:
        betas = self._bc(self.betas[t - 1])
        return OneHotCategoricalBCHW((1 - betas) * xtm1 + betas / self.num_classes)

    def q_xt_given_x0(self, x0: Tensor, t: Tensor) -> OneHotCategoricalBCHW:
        ca = self._bc(self.cumalphas[t - 1])
        return OneHotCategoricalBCHW(ca * x0 + (1 - ca) / self.num_classes)

    def theta_post(self, xt: Tensor, x0: Tensor, t: Tensor) -> Tensor:
        """:92-103 (x0 one-hot; training loss target -- plain tensor ops, not on the sampler path)."""
        a, g = self._step_coef(t)
        a, g = self._bc(a), self._bc(g)
        theta = (a * xt + (1 - a) / self.num_classes) * (g * x0 + (1 - g) / self.num_classes)
        return theta / theta.sum(dim=1, keepdim=True)

    # -- reverse process -------------------------------------------------------------------
    def _step_coef(self, t: Tensor):
        """(alpha_t, cumalpha_{t-1}) per sample for 1-based t; t == 1 -> (0, 1)  (:114-122)."""
        tt = t.long() - 1
        a = self.alphas[tt].clone().float()
        g = self.cumalphas[tt - 1].clone().float()
        a[tt == 0] = 0.0
        g[tt == 0] = 1.0
        return a, g

    def step_coef_tensor(self, t: Tensor) -> Tensor:
        a, g = self._step_coef(t)
        return torch.stack([a, g], dim=1).contiguous()

    @torch.no_grad()
    def theta_post_prob(self, xt: Tensor, theta_x0: Tensor, t: Tensor) -> Tensor:
        """:105-139  theta_post(x_{t-1} | x_t, p(x0)) -- one fused kernel, O(C) per voxel."""
        coef = self.step_coef_tensor(t.to(self.alphas.device)).to(xt.device)
        out, _, _ = ops.cat_posterior_sample(theta_x0.float().contiguous(), xt.float().contiguous(), coef, ops.CAT_POSTERIOR)
        return out


class DenoisingModel(nn.Module):
    def __init__(self, diffusion: DiffusionModel, unet: nn.Module, dataset_file: str, step_T_sample: str = "majority", dims=3,
                 loop: str = "interface"):
        super().__init__()
        self.diffusion = diffusion
        self.unet = unet
        self.dims = dims
        self.dataset_file = dataset_file
        self.step_T_sample = step_T_sample
        self.loop = loop                 # "interface" | "resident"
        self.use_cuda_graph = False      # capture the UNet forward of the loop in a CUDA graph
        self.q_noise = None              # optional injected noise: indexable [step] -> fp32 [B*V, C]
        # Key of the in-kernel Philox noise of the resident loop.  None (default): every forward_denoising call draws a
        # fresh key from torch's default generator -- so torch.manual_seed governs the result and successive calls
        # (an evaluator looping over volumes, repeated samples) get independent noise, as the reference's
        # torch.multinomial does by advancing the global RNG.  An int pins the key (reproducible runs, tests).
        self.philox_seed: Optional[int] = None
        # global index of this process' first chain when a batch is sharded over ranks (sharding.shard_range): the
        # Philox counter is the GLOBAL voxel index, so results do not depend on the number of ranks
        self.chain_base = 0
        # resident loop: draw inside the head conv's epilogue (no logits tensor, no separate per-voxel launch) whenever the
        # plan offers it (depth-rolling head conv) and the noise is the in-kernel Philox stream
        self.fuse_head = True
        self.record = None               # optional list: receives the uint8 label volume after every step

    @property
    def time_steps(self):
        return self.diffusion.time_steps

    def forward(self, x: Tensor, condition: Tensor, feature_condition: Tensor = None, t: Optional[Tensor] = None,
                label_ref_logits: Optional[Tensor] = None, validation: bool = False, context=None) -> Union[Tensor, dict]:
        if self.training:
            if not isinstance(t, Tensor):
                raise ValueError("'t' needs to be a Tensor at training time")
            if not isinstance(x, Tensor):
                raise ValueError("'x' needs to be a Tensor at training time")
            return self.forward_step(x, condition, feature_condition, t, context=context)
        if validation:
            return self.forward_step(x, condition, feature_condition, t, context=context)
        if t is None:
            return self.forward_denoising(x, condition, feature_condition, label_ref_logits=label_ref_logits, context=context)
        return self.forward_denoising(x, condition, feature_condition, int(t.item()), label_ref_logits, context=context)

    def forward_step(self, x, condition, feature_condition, t, context=None):
        return self.unet(x, condition, feature_condition=feature_condition, timesteps=t, context=context)

    def _t_values(self, init_t):
        if init_t is None:
            init_t = self.time_steps
        if init_t > 10000:                      # step skipping (:190-197)
            K = init_t % 10000
            assert 0 < K <= self.time_steps
            if K == self.time_steps:
                return list(range(K, 0, -1))
            LOGGER.warning(f"Override default {self.time_steps} time steps with {K}.")
            return [round(v) for v in np.linspace(self.time_steps, 1, K)]
        return list(range(init_t, 0, -1))

    @torch.no_grad()
    def forward_denoising(self, x: Optional[Tensor], condition: Tensor, feature_condition: Tensor, init_t: Optional[int] = None,
                          label_ref_logits: Optional[Tensor] = None, context: Tensor = None) -> dict:
        if label_ref_logits is not None:
            raise NotImplementedError("guidance branch is broken in the reference (undefined attributes, SURVEY.md D4)")
        if feature_condition is not None:
            raise NotImplementedError("feature_condition is not used by the shipped ruijin config")
        dev = next(self.unet.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DenoisingModel needs its parameters on a CUDA device (no CPU path exists)")
        t_values = self._t_values(init_t)
        xt = x.to(dev, torch.float32).contiguous()
        cond = condition.to(dev, torch.float32).contiguous() if condition is not None else None
        if context is not None:
            context = context.to(dev)
        B, Cc = xt.shape[:2]
        spatial = tuple(xt.shape[2:])
        V = int(math.prod(spatial))
        coefs = self.diffusion.step_coef_tensor(torch.tensor(t_values)).to(dev)          # [steps, 2]
        coefs = coefs[:, None, :].expand(-1, B, -1).contiguous()                          # [steps, B, 2]
        slab = getattr(getattr(self.unet, "_engine", None), "slab", None)
        if self.loop == "resident" and len(t_values) > 1:
            xt = self._resident_steps(xt, cond, context, t_values[:-1], coefs, spatial)
            t_values, coefs, step0 = t_values[-1:], coefs[-1:], len(t_values) - 1
        else:
            step0 = 0
        unet = self.unet
        plan = unet.plan_for(B, spatial, context)
        if self.use_cuda_graph and plan.graph is None:
            plan.capture()
        if "context" in plan.inputs:
            plan.inputs["context"].copy_(unet._ctx_cl(context, B))
        probs_x0 = torch.empty_like(xt)
        labels = torch.empty((B, V), dtype=torch.uint8, device=dev) if self.record is not None else None
        for i, t in enumerate(t_values):
            # x0pred = unet(xt, condition, None, t_.float(), context)["diffusion_out"]      (:203-207)
            ops.nchw_to_cl(xt, cond, c_pad=unet.in_channels_padded, out=plan.inputs["x"])
            plan.inputs["t"].fill_(float(t))
            plan.run()
            ops.cl_to_nchw(plan.outputs["head"], Cc, spatial, softmax=bool(unet.sofmtax_output), out=probs_x0)
            # probs = theta_post_prob(xt, x0pred, t_); clamp(1e-12); sample / argmax / probs   (:209-224)
            if t > 1:
                q = None
                if self.q_noise is not None:
                    q = self.q_noise[step0 + i].to(dev, torch.float32).contiguous()
                elif slab is not None and slab.world > 1:
                    # depth slabs: every rank draws the noise field of the WHOLE volume from its (identically seeded)
                    # generator and keeps its own rows -- the slabs then see different noise, and the same noise as an
                    # unsplit run from the same generator state
                    q = torch.empty((slab.world, B * V, Cc), dtype=torch.float32, device=dev).exponential_(1)[slab.rank].contiguous()
                else:
                    q = torch.empty((B * V, Cc), dtype=torch.float32, device=dev).exponential_(1)
                nxt = torch.empty_like(xt)
                ops.cat_posterior_sample(probs_x0, xt, coefs[i], ops.CAT_SAMPLE, q=q, out=nxt, labels=labels)
                xt = nxt
            elif self.step_T_sample is None or self.step_T_sample == "majority":
                o64 = torch.empty(xt.shape, dtype=torch.int64, device=dev)
                ops.cat_posterior_sample(probs_x0, xt, coefs[i], ops.CAT_ARGMAX, out_i64=o64, labels=labels)
                xt = o64
            elif self.step_T_sample == "confidence":
                out = torch.empty_like(xt)
                ops.cat_posterior_sample(probs_x0, xt, coefs[i], ops.CAT_PROBS, out=out)
                xt = out
            if self.record is not None and (t > 1 or self.step_T_sample in (None, "majority")):
                self.record.append(labels.clone())
        return {"diffusion_out": xt}

    def _resident_steps(self, xt, cond, context, t_values, coefs, spatial):
        """Steps t > 1 without leaving the channels-last device layout; returns fp32 one-hot xt."""
        st = self.resident_begin(xt, cond, context)
        for i, t in enumerate(t_values):
            q = None
            if self.q_noise is not None:
                q = self.q_noise[i].to(xt.device, torch.float32).contiguous()
            self.resident_step(st, t, coefs[i], q=q, offset=i)
            if self.record is not None:
                self.record.append(st["lab_a"].view(st["B"], st["V"]).clone())
        return self.resident_end(st)

    # -- device-resident sampler state (also used by bench.py to time exactly K steps) -----------
    def resident_begin(self, xt, cond, context=None) -> dict:
        """xt fp32 one-hot [B, C, *sp] and cond [B, 1, *sp] on the device -> loop state: the UNet
        plan (input buffer filled), uint8 label ping-pong buffers, channels-last condition."""
        unet = self.unet
        dev = xt.device
        B, Cc = xt.shape[:2]
        spatial = tuple(xt.shape[2:])
        V = int(math.prod(spatial))
        if not unet.sofmtax_output:
            raise NotImplementedError("the resident loop applies the head softmax in its fused per-voxel kernel; a network built with "
                                      "softmax_output=False (raw head output fed to theta_post_prob) must use loop='interface'")
        if not bool(((xt.amax(1) == 1) & (xt.sum(1) == 1)).all()):
            raise ValueError("the resident loop keeps x_t as uint8 labels: x must be one-hot over dim 1 (use loop='interface' for soft x_t)")
        plan = unet.plan_for(B, spatial, context)
        if "context" in plan.inputs:
            plan.inputs["context"].copy_(unet._ctx_cl(context, B))
        xin = plan.inputs["x"]
        if self.use_cuda_graph and plan.graph is None:
            plan.capture()
        ops.nchw_to_cl(xt, cond, c_pad=unet.in_channels_padded, out=xin)
        lab_a = torch.empty((B * V,), dtype=torch.uint8, device=dev)
        lab_b = torch.empty_like(lab_a)
        ops.cat_posterior_sample(xt, None, None, ops.CAT_ARGMAX_GIVEN, clamp_min=0.0, labels=lab_a.view(B, V))
        n_cond = cond.shape[1] if cond is not None else 0
        cond_cl = ops.nchw_to_cl(cond, None, c_pad=8)[..., :n_cond].contiguous() if cond is not None else None
        slab = unet.engine.slab
        # global voxel index of this process' first voxel: slabs are equal-sized and contiguous in depth (one volume);
        # sharded batches start at chain `chain_base` of the global batch
        vox_base = slab.rank * V if slab is not None else self.chain_base * V
        seed = self.philox_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            if slab is not None and slab.world > 1:
                seed = slab.broadcast_int(seed)          # one volume, one key
        fused = bool(self.fuse_head and plan.fused_head is not None and self.q_noise is None)
        if fused and self.use_cuda_graph:
            plan.capture_body()
        return dict(plan=plan, xin=xin, lab_a=lab_a, lab_b=lab_b, cond_cl=cond_cl, n_cond=n_cond, B=B, C=Cc, V=V,
                    spatial=spatial, vox_base=vox_base, seed=seed, fused_head=fused)

    def resident_step(self, st: dict, t: int, coef: Tensor, q: Optional[Tensor] = None, offset: int = 0):
        """One reverse step t -> t-1 (t > 1): UNet forward, then ONE fused kernel: softmax over the
        head logits + posterior + clamp + draw + next UNet input (one-hot | condition) in place."""
        plan = st["plan"]
        plan.inputs["t"].fill_(float(t))
        if st.get("fused_head") and q is None:
            plan.run_body()
            self._launch_fused_head(st, coef, offset)
            st["lab_a"], st["lab_b"] = st["lab_b"], st["lab_a"]
            return
        plan.run()
        ops.cat_step_cl(plan.outputs["head"], st["lab_a"], coef, st["lab_b"], st["B"], st["V"], st["C"], mode=ops.CAT_SAMPLE,
                        q=q, cond=st["cond_cl"], n_cond=st["n_cond"], next_x=st["xin"], seed=st["seed"], offset=offset,
                        vox_base=st["vox_base"])
        st["lab_a"], st["lab_b"] = st["lab_b"], st["lab_a"]

    def _launch_fused_head(self, st: dict, coef: Tensor, offset: int):
        """The head conv with softmax + posterior + clamp + draw + next-input write in its epilogue (gg_conv_args.cat); one
        launch per lane of the plan, each over its own sample range."""
        import ctypes as C
        from .. import _C
        plan = st["plan"]
        V, xin = st["V"], st["xin"]
        refs = []
        for fa, cat, n0 in plan.fused_heads:
            cat.labels_in, cat.labels_out = st["lab_a"].data_ptr() + n0 * V, st["lab_b"].data_ptr() + n0 * V
            cat.next_x = xin.data_ptr() + n0 * V * xin.shape[-1] * xin.element_size()
            cat.cond = (st["cond_cl"].data_ptr() + n0 * V * st["n_cond"] * st["cond_cl"].element_size()) if st["cond_cl"] is not None else None
            cat.coef = coef.data_ptr() + n0 * 2 * coef.element_size()
            cat.C, cat.n_cond, cat.Cin_pad, cat.mode = st["C"], st["n_cond"], xin.shape[-1], ops.CAT_SAMPLE
            cat.clamp_min, cat.seed, cat.offset, cat.vox_base = 1e-12, int(st["seed"]), int(offset), int(st["vox_base"]) + n0 * V
            ref = C.byref(fa)
            _C.check(_C.lib().gg_conv_fwd(ref, _C.stream()), "gg_conv_fwd (sampler epilogue)")
            refs.append(ref)
        return tuple(refs)

    def launches_per_step(self, plan) -> int:
        """libguidegen_sm100 launches of one resident step."""
        return plan.num_launches if (self.fuse_head and plan.fused_head is not None and self.q_noise is None) else plan.num_launches + 1

    def resident_tail_launcher(self, st: dict, coef: Tensor, offset: int = 0):
        """For per-launch timing (bench.py): (plan steps to run, [(name, launch)]) that together make one resident step."""
        plan = st["plan"]
        if st.get("fused_head"):
            return plan.body_steps, [("gg_conv_fwd", lambda: self._launch_fused_head(st, coef, offset))]

        def cat_step():
            ops.cat_step_cl(plan.outputs["head"], st["lab_a"], coef, st["lab_b"], st["B"], st["V"], st["C"], mode=ops.CAT_SAMPLE,
                            cond=st["cond_cl"], n_cond=st["n_cond"], next_x=st["xin"], seed=st["seed"], offset=offset,
                            vox_base=st["vox_base"])
            return ()
        return plan.steps, [("gg_cat_step_cl", cat_step)]

    def resident_end(self, st: dict) -> Tensor:
        return ops.cl_to_nchw(st["xin"], st["C"], st["spatial"])
