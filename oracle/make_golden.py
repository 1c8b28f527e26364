"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference modules are imported read-only through oracle.refshim, loaded with
oracle.weights.synth_state_dict (zero-initialised modules re-randomised),
driven with seeded inputs and injected noise, and their outputs are stored.  Inputs that
are cheap to regenerate from a numpy seed are NOT stored; the seed is.
"""
import os
import sys

import numpy as np
import torch

from . import configs, refshim, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _patched_multinomial(q_list):
    """torch.multinomial(p, 1, True) == argmax(p / q) with q ~ Exp(1) (ATen fast path);
    here q is popped from a pre-drawn list so the draw can be replayed on the GPU."""
    state = {"i": 0}

    def fn(probs_2d, num_samples, replacement=False, *, generator=None, out=None):
        assert num_samples == 1
        q = q_list[state["i"]]
        state["i"] += 1
        return torch.argmax(probs_2d / q, dim=-1, keepdim=True)

    return fn


def _ccdm_model(ref, params, T, C, spatial, seed_w):
    m = ref.build_model(time_steps=T, schedule="cosine", schedule_params={"s": 0.008},
                        input_shapes=[(1,) + spatial, (C,) + spatial], cond_encoded_shape=None,
                        backbone="unet_openai", backbone_params=dict(params), dataset_file="datasets.ruijin",
                        step_T_sample="majority", dims=3)
    m.eval()
    shapes = {k: v for k, v in weights.shapes_of(m).items() if k.startswith("unet.")}
    sd = weights.synth_state_dict({k[5:]: s for k, s in shapes.items()}, seed_w)
    m.unet.load_state_dict(sd)
    return m


@torch.no_grad()
def ccdm_chain(name, params, T, B, C, spatial, seed_w=1, seed_x=2, seed_q=3, sub=1):
    ref = refshim.ccdm()
    m = _ccdm_model(ref, params, T, C, spatial, seed_w)
    V = int(np.prod(spatial))
    x_T = weights.uniform_one_hot(seed_x, B, C, spatial)
    cond = torch.zeros(B, 1, *spatial)
    q = torch.from_numpy(weights.exp_noise(seed_q, (T, B * V, C)))
    # first-step network output and posterior, through the reference's own methods
    t_ = torch.full((B,), T)
    probs0 = m.unet(x_T, cond, None, t_.float())["diffusion_out"]
    post0 = m.diffusion.theta_post_prob(x_T, probs0, t_)
    # full chain with injected noise; record every intermediate label volume
    steps = []
    real_multinomial = torch.multinomial
    torch.multinomial = _patched_multinomial(list(q))
    try:
        cls = ref.OneHotCategoricalBCHW
        real_sample = cls.sample

        def rec_sample(self, sample_shape=torch.Size()):
            r = real_sample(self, sample_shape)
            steps.append(r.argmax(dim=1).to(torch.uint8).numpy())
            return r

        cls.sample = rec_sample
        out = m(x_T, cond, feature_condition=None, context=None)["diffusion_out"]
        cls.sample = real_sample
    finally:
        torch.multinomial = real_multinomial
    final = out.argmax(dim=1).to(torch.uint8).numpy()
    assert out.dtype == torch.int64
    sl = (slice(None), slice(None)) + tuple(slice(None, None, sub) for _ in spatial)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        T=T, B=B, C=C, spatial=np.asarray(spatial), seed_w=seed_w, seed_x=seed_x, seed_q=seed_q,
                        sub=sub, probs0=probs0[sl].numpy(), post0=post0[sl].numpy(),
                        step_labels=np.stack(steps), final_labels=final,
                        betas=m.diffusion.betas.numpy(), alphas=m.diffusion.alphas.numpy(),
                        cumalphas=m.diffusion.cumalphas.numpy())
    print(name, "steps", len(steps), "label hist", np.bincount(final.ravel(), minlength=C))


@torch.no_grad()
def posterior_cases(name="posterior_cases"):
    ref = refshim.ccdm()
    C, spatial, B = 12, (4, 8, 8), 2
    dm = ref.DiffusionModel("cosine", 1000, C, schedule_params={"s": 0.008}, dims=3)
    rs = np.random.RandomState(7)
    out = {}
    for t in (1, 2, 17, 500, 999, 1000):
        xt = weights.uniform_one_hot(100 + t, B, C, spatial)
        x0 = torch.softmax(torch.from_numpy(rs.standard_normal((B, C) + spatial).astype(np.float32)) * 3, dim=1)
        out[f"x0_{t}"] = x0.numpy()
        out[f"post_{t}"] = dm.theta_post_prob(xt, x0, torch.full((B,), t)).numpy()
    # soft (non one-hot) x_t as well: theta_post_prob is a public method
    xt = torch.softmax(torch.from_numpy(rs.standard_normal((B, C) + spatial).astype(np.float32)), dim=1)
    x0 = torch.softmax(torch.from_numpy(rs.standard_normal((B, C) + spatial).astype(np.float32)), dim=1)
    out["soft_xt"], out["soft_x0"] = xt.numpy(), x0.numpy()
    out["soft_post_300"] = dm.theta_post_prob(xt, x0, torch.full((B,), 300)).numpy()
    dm10 = ref.DiffusionModel("cosine", 10, C, schedule_params={"s": 0.008}, dims=3)
    dml = ref.DiffusionModel("linear", 50, C, schedule_params=None, dims=3)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), C=C, B=B, spatial=np.asarray(spatial),
                        cos1000_betas=dm.betas.numpy(), cos1000_alphas=dm.alphas.numpy(),
                        cos1000_cumalphas=dm.cumalphas.numpy(), cos10_cumalphas=dm10.cumalphas.numpy(),
                        cos10_alphas=dm10.alphas.numpy(), lin50_alphas=dml.alphas.numpy(),
                        lin50_cumalphas=dml.cumalphas.numpy(), **out)
    print(name, "ok")


class _DuckLDM:
    """The six attributes DDIMSampler needs from `model`; LatentDiffusion
    itself cannot be imported here (pytorch_lightning / taming absent).  The schedule buffers
    follow ddpm.py:118-138 via the reference's own make_beta_schedule."""

    def __init__(self, ref, unet, conditioning_key="concat"):
        sch = configs.LDM_SCHEDULE
        betas = ref.util.make_beta_schedule("linear", sch["timesteps"], linear_start=sch["linear_start"],
                                            linear_end=sch["linear_end"])
        acp = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = int(betas.shape[0])
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(acp, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, acp[:-1]), dtype=torch.float32)
        self.device = torch.device("cpu")
        self.unet = unet
        self.conditioning_key = conditioning_key
        self.parameterization = "eps"

    def apply_model(self, x, t, c):
        # DiffusionWrapper.forward, ddpm.py:1415-1434
        if self.conditioning_key == "concat":
            return self.unet(torch.cat([x, c], dim=1), t)
        if self.conditioning_key == "hybrid":
            # the reference's DDIMSampler.sample reads `.shape` off the first dict value (ddim.py:82),
            # so dict-of-lists conditionings cannot pass through it; tensors are wrapped here instead
            cc, ca = c["c_concat"], c["c_crossattn"]
            cc = cc if isinstance(cc, list) else [cc]
            ca = ca if isinstance(ca, list) else [ca]
            return self.unet(torch.cat([x] + cc, dim=1), t, context=torch.cat(ca, 1))
        raise NotImplementedError


@torch.no_grad()
def ldm_ddim(name, params, B, hw, S, eta, hybrid=False, seed_w=11, store_full=True):
    ref = refshim.ldm()
    unet = ref.UNetModel(**params).eval()
    unet.load_state_dict(weights.synth_state_dict(weights.shapes_of(unet), seed_w))
    model = _DuckLDM(ref, unet, "hybrid" if hybrid else "concat")
    x_T = weights.normal(21, (B, 4) + hw)
    cc = weights.normal(22, (B, 4) + hw)
    cond = cc
    ctx = None
    if hybrid:
        ctx = weights.normal(23, (B, 7, params["context_dim"]))
        cond = {"c_concat": cc, "c_crossattn": ctx}
    noises = [weights.normal(1000 + i, (B, 4) + hw) for i in range(S)]
    sampler = ref.DDIMSamplerCPU(model)
    # inject the per-step Gaussian noise (noise_like always draws, ddim.py:201 / util.py:264-267)
    it = iter(noises)
    real = ref.ddim_module.noise_like
    ref.ddim_module.noise_like = lambda shape, device, repeat=False: next(it)
    inter = []
    try:
        out, _ = sampler.sample(S=S, batch_size=B, shape=(4,) + hw, conditioning=cond, eta=eta, x_T=x_T,
                                verbose=False, dims=2, img_callback=lambda p, i: inter.append(p.clone()))
    finally:
        ref.ddim_module.noise_like = real
    t0 = torch.full((B,), int(sampler.ddim_timesteps[-1]), dtype=torch.long)
    eps0 = model.apply_model(x_T, t0, cond)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), B=B, hw=np.asarray(hw), S=S, eta=eta, seed_w=seed_w,
                        hybrid=hybrid, eps0=eps0.numpy(), final=out.numpy(), pred_x0_first=inter[0].numpy(),
                        pred_x0_last=inter[-1].numpy(),
                        ddim_timesteps=np.asarray(sampler.ddim_timesteps),
                        ddim_alphas=np.asarray(sampler.ddim_alphas, dtype=np.float64),
                        ddim_alphas_prev=np.asarray(sampler.ddim_alphas_prev, dtype=np.float64),
                        ddim_sigmas=np.asarray(sampler.ddim_sigmas, dtype=np.float64),
                        ddim_sqrt_one_minus_alphas=np.asarray(sampler.ddim_sqrt_one_minus_alphas, dtype=np.float64),
                        alphas_cumprod=model.alphas_cumprod.numpy())
    print(name, "final std", float(out.std()))


@torch.no_grad()
def ldm_plms(name, params, B, hw, S, seed_w=11):
    """PLMSSampler.sample of the unmodified reference (plms.py) on the tiny concat-conditioned LDM network."""
    ref = refshim.ldm()
    unet = ref.UNetModel(**params).eval()
    unet.load_state_dict(weights.synth_state_dict(weights.shapes_of(unet), seed_w))
    model = _DuckLDM(ref, unet, "concat")
    x_T = weights.normal(21, (B, 4) + hw)
    cond = weights.normal(22, (B, 4) + hw)
    sampler = ref.PLMSSamplerCPU(model)
    inter, eps = [], []
    real = sampler.p_sample_plms

    def spy(*a, **k):
        out = real(*a, **k)
        eps.append(out[2].clone())
        return out

    sampler.p_sample_plms = spy
    out, _ = sampler.sample(S=S, batch_size=B, shape=(4,) + hw, conditioning=cond, eta=0.0, x_T=x_T, verbose=False,
                            img_callback=lambda p, i: inter.append(p.clone()))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), B=B, hw=np.asarray(hw), S=S, seed_w=seed_w, final=out.numpy(),
                        pred_x0=np.stack([p.numpy() for p in inter]), e_t=np.stack([e.numpy() for e in eps]),
                        alphas_cumprod=model.alphas_cumprod.numpy())
    print(name, "final std", float(out.std()))


@torch.no_grad()
def vae_decoder(name="vae_decoder", seed_w=19):
    """Decoder of the unmodified reference (ldm/modules/diffusionmodules/model.py:524-631) behind a 1x1
    post_quant_conv (AutoencoderKL.decode, autoencoder.py:355-359; AutoencoderKL itself needs taming / lightning and
    cannot be imported here): a narrow instance in full and the shipped widths (ch 128, mult 1-2-4-4) sub-sampled."""
    import importlib
    refshim.ldm()
    mdl = importlib.import_module("ldm.modules.diffusionmodules.model")
    out = {}
    for tag, (ch, zhw, sub) in (("small", (32, 16, 1)), ("wide", (128, 16, 4))):
        dd = dict(ch=ch, out_ch=1, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0, in_channels=1,
                  resolution=zhw * 8, z_channels=4, double_z=True, dims=2)
        dec = mdl.Decoder(**dd).eval()
        sd = weights.synth_state_dict({"decoder." + k: v for k, v in weights.shapes_of(dec).items()}, seed_w)
        sd["post_quant_conv.weight"] = weights.synth_tensor("post_quant_conv.weight", (4, 4, 1, 1), seed_w)
        sd["post_quant_conv.bias"] = weights.synth_tensor("post_quant_conv.bias", (4,), seed_w)
        dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
        z = weights.normal(51, (2, 4, zhw, zhw))
        y = dec(torch.nn.functional.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"]))
        out[tag + "_cfg"] = np.asarray([ch, zhw, sub])
        out[tag + "_out"] = y[:, :, ::sub, ::sub].numpy()
        out[tag + "_std"] = float(y.std())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed_w=seed_w, **out)
    print(name, {k: v for k, v in out.items() if k.endswith("_std")})


@torch.no_grad()
def vae_encoder(name="vae_encoder", seed_w=23):
    """Encoder of the unmodified reference (model.py:398-520) followed by a 1x1 quant_conv (AutoencoderKL.encode,
    autoencoder.py:350-352): the posterior's moments; narrow instance and the shipped widths (ch 128) on a 128 x 128 slice."""
    import importlib
    refshim.ldm()
    mdl = importlib.import_module("ldm.modules.diffusionmodules.model")
    out = {}
    for tag, (ch, hw) in (("small", (32, 128)), ("wide", (128, 128))):
        dd = dict(ch=ch, out_ch=1, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0, in_channels=1,
                  resolution=hw, z_channels=4, double_z=True, dims=2)
        enc = mdl.Encoder(**dd).eval()
        sd = weights.synth_state_dict({"encoder." + k: v for k, v in weights.shapes_of(enc).items()}, seed_w)
        sd["quant_conv.weight"] = weights.synth_tensor("quant_conv.weight", (8, 8, 1, 1), seed_w)
        sd["quant_conv.bias"] = weights.synth_tensor("quant_conv.bias", (8,), seed_w)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
        x = weights.normal(61, (2, 1, hw, hw))
        y = torch.nn.functional.conv2d(enc(x), sd["quant_conv.weight"], sd["quant_conv.bias"])
        out[tag + "_cfg"] = np.asarray([ch, hw])
        out[tag + "_out"] = y.numpy()
        out[tag + "_std"] = float(y.std())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed_w=seed_w, **out)
    print(name, {k: v for k, v in out.items() if k.endswith("_std")})


@torch.no_grad()
def text_encoder(name="ccdm_text_encoder", seed_w=17):
    """PreloadedBERTEncoder of the unmodified reference (ccdm/ddpm/models/encoder.py:103-123): a small instance in full
    and the shipped size (768 wide, 8 heads x 64, depth 4, 512 tokens) sub-sampled."""
    import importlib
    refshim.ccdm()
    enc = importlib.import_module("ddpm.models.encoder")
    out = {}
    for tag, (dim, heads, d_head, depth, B, L, sub) in (("small", (128, 2, 64, 2, 2, 24, 1)), ("full", (768, 8, 64, 4, 1, 512, 8))):
        m = enc.PreloadedBERTEncoder(embed_dim=dim, n_heads=heads, depth=depth, d_head=d_head).eval()
        m.load_state_dict(weights.synth_state_dict(weights.shapes_of(m), seed_w))
        x = weights.normal(41, (B, dim, L))
        y = m(x)
        out[tag + "_cfg"] = np.asarray([dim, heads, d_head, depth, B, L, sub])
        out[tag + "_out"] = y[:, ::sub, ::sub].numpy()
        out[tag + "_std"] = float((y - x).std())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed_w=seed_w, **out)
    print(name, {k: v for k, v in out.items() if k.endswith("_std")})


@torch.no_grad()
def ldm_forward(name, params, B, hw, sub, seed_w=12):
    ref = refshim.ldm()
    unet = ref.UNetModel(**params).eval()
    unet.load_state_dict(weights.synth_state_dict(weights.shapes_of(unet), seed_w))
    x = weights.normal(31, (B, params["in_channels"]) + hw)
    t = torch.tensor([981] * B, dtype=torch.long)
    y = unet(x, t)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), B=B, hw=np.asarray(hw), sub=sub, seed_w=seed_w,
                        out=y[:, :, ::sub, ::sub].numpy(), out_mean=float(y.mean()), out_std=float(y.std()))
    print(name, "out std", float(y.std()))


def ddim_tables(name="ddim_tables"):
    ref = refshim.ldm()
    model = _DuckLDM(ref, None)
    out = {}
    for S, eta in ((50, 0.0), (50, 1.0), (20, 0.5), (250, 0.0)):
        s = ref.DDIMSamplerCPU(model)
        s.make_schedule(S, ddim_eta=eta, verbose=False)
        k = f"S{S}_eta{eta}"
        out[k + "_timesteps"] = np.asarray(s.ddim_timesteps)
        out[k + "_alphas"] = np.asarray(s.ddim_alphas, dtype=np.float64)
        out[k + "_alphas_prev"] = np.asarray(s.ddim_alphas_prev, dtype=np.float64)
        out[k + "_sigmas"] = np.asarray(s.ddim_sigmas, dtype=np.float64)
        out[k + "_sqrt_one_minus_alphas"] = np.asarray(s.ddim_sqrt_one_minus_alphas, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), alphas_cumprod=model.alphas_cumprod.numpy(),
                        betas=model.betas.numpy(), **out)
    print(name, "ok")


def param_shapes(name="param_shapes"):
    """Parameter names/shapes of the reference networks (taken from the reference modules): lets the
    oracle build synthetic state_dicts on machines without /root/reference."""
    import json
    out = {}
    rc = refshim.ccdm()
    for key, params, C, sp in (("CCDM_PARAMS_YML", configs.CCDM_PARAMS_YML, 12, (32, 32, 32)), ("CCDM_TINY", configs.CCDM_TINY, 4, (8, 8, 8)),
                               ("CCDM_TINY_C12", configs.CCDM_TINY, 12, (8, 8, 8))):
        m = rc.build_model(time_steps=10, schedule="cosine", schedule_params={"s": 0.008}, input_shapes=[(1,) + sp, (C,) + sp],
                           cond_encoded_shape=None, backbone="unet_openai", backbone_params=dict(params),
                           dataset_file="x", step_T_sample="majority", dims=3)
        out[key] = {k[5:]: list(v) for k, v in weights.shapes_of(m).items() if k.startswith("unet.")}
    rl = refshim.ldm()
    for key in ("LDM_AE", "LDM_PIXEL", "LDM_TINY", "LDM_TINY_XATTN"):
        out[key] = {k: list(v) for k, v in weights.shapes_of(rl.UNetModel(**getattr(configs, key))).items()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".json"), "w") as f:
        json.dump(out, f)
    print(name, {k: len(v) for k, v in out.items()})


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    which = set(argv) or {"all"}

    def want(k):
        return "all" in which or k in which

    if want("shapes"):
        param_shapes()
    if want("posterior"):
        posterior_cases()
    if want("ddim_tables"):
        ddim_tables()
    if want("ccdm_tiny"):
        ccdm_chain("ccdm_tiny", configs.CCDM_TINY, T=6, B=2, C=4, spatial=(8, 8, 8))
    if want("ldm_tiny"):
        ldm_ddim("ldm_tiny_eta0", configs.LDM_TINY, B=2, hw=(16, 16), S=5, eta=0.0)
        ldm_ddim("ldm_tiny_eta05", configs.LDM_TINY, B=2, hw=(16, 16), S=5, eta=0.5)
        ldm_ddim("ldm_tiny_hybrid", configs.LDM_TINY_XATTN, B=2, hw=(16, 16), S=4, eta=0.0, hybrid=True)
    if want("vae_decoder"):
        vae_decoder()
    if want("vae_encoder"):
        vae_encoder()
    if want("text_encoder"):
        text_encoder()
    if want("ldm_plms"):
        ldm_plms("ldm_tiny_plms", configs.LDM_TINY, B=2, hw=(16, 16), S=7)
    if want("ccdm_cfg1"):
        ccdm_chain("ccdm_cfg1", configs.CCDM_PARAMS_YML, T=10, B=1, C=12, spatial=(32, 32, 32), sub=4)
    if want("ldm_ae"):
        ldm_forward("ldm_ae_fwd", configs.LDM_AE, B=1, hw=(64, 64), sub=4)


if __name__ == "__main__":
    main(sys.argv[1:])
