#!/usr/bin/env python
"""bench.py -- denoising steps/sec of the GuideGen CCDM mask sampler on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Workload (config.workload): BASELINE.json configs[1] -- CCDM 3-D UNet (ccdm/params.yml:69-75) mask
sampler, volume 128x128x64 (tensor [B, 12, 64, 128, 128]), 12 classes, batch 8 per GPU, bf16.
One "step" = one reverse-diffusion step of the whole batch: UNet forward + categorical posterior
+ categorical draw + next-input assembly.  N > 1: independent chains are sharded over the ranks
(batch 8 per GPU, no data-path collective) -> weak scaling; value = N * K / max-over-ranks time.

Keys beyond the base contract: "roofline" (dominant kernel = the tcgen05 implicit-GEMM conv, tensor
bound, timed live with CUDA events per launch in a separate instrumented pass), "roofline_hbm"
(the fused per-voxel posterior/sampling kernel), "cpu_baseline", "kernel_ms" (per-kernel share of
one step), "volumes_per_sec" (= value * batch / 1000 steps).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CCDM_NET = dict(base_channels=64, channel_mult=[1, 2, 2, 4, 5], attention_resolutions=[32, 16, 8], num_heads=1,
                num_head_channels=32, softmax_output=True)                       # ccdm/params.yml:69-75
WORKLOADS = {
    # name: (spatial D,H,W, classes, batch, chain length, FLOP per sample-forward from BASELINE.md section 3)
    "ccdm_cfg2": dict(spatial=(64, 128, 128), C=12, batch=8, T=1000, flop_per_sample=6.324e12,
                      desc="CCDM mask sampler 128x128x64 (tensor [8,12,64,128,128]), 12 classes, 1000-step chain, batch 8, bf16"),
    "ccdm_cfg1": dict(spatial=(32, 32, 32), C=12, batch=1, T=10, flop_per_sample=1.97e11,
                      desc="CCDM mask sampler 32x32x32, 12 classes, 10 steps, batch 1"),
    "ccdm_cfg2_text": dict(spatial=(64, 128, 128), C=12, batch=8, T=1000, flop_per_sample=6.324e12, text=True,
                           desc="CCDM TEXT-CONDITIONED mask sampler 128x128x64, 12 classes, batch 8, bf16: attention sites are "
                                "SpatialTransformer blocks (self-attn + cross-attn to a [8,512,768] BERT-like context + GEGLU FF); the "
                                "reference declares but cannot construct this network (SURVEY.md D1/D2), so there is no reference arm"),
    "ccdm_cfg5": dict(spatial=(128, 256, 256), C=12, batch=1, T=1000, flop_per_sample=5.182e13, attn_flop=1.41e12, slab=True,
                      desc="CCDM mask sampler, ONE 256x256x128 volume (tensor [1,12,128,256,256]) split into depth slabs over the "
                           "GPUs: halo exchange per 3x3x3 conv, GroupNorm partial-sum gather, attention K/V gather (NCCL)"),
    "ldm_cfg3": dict(spatial=(64, 64), C=4, batch=16, T=50, flop_per_sample=1.24e11, kind="ldm",
                     desc="LDM conditional CT slice generator (ruijin-ldm_from_controlnet_ae.yaml UNet), latent 4x64x64, "
                          "concat mask/prev-slice context, DDIM 50 steps eta 0, batch 16, bf16"),
}
LDM_PIXEL_NET = dict(dims=2, image_size=512, in_channels=3, out_channels=1, model_channels=128, attention_resolutions=[32, 16, 8],
                     num_res_blocks=2, channel_mult=[1, 2, 4, 4, 5], num_head_channels=32)      # ruijin-ldm_from_controlnet.yaml:17-40
LDM_AE_NET = dict(dims=2, image_size=512, in_channels=8, out_channels=4, model_channels=160, attention_resolutions=[8, 4, 2],
                  num_res_blocks=2, channel_mult=[1, 2, 4, 4, 5], num_head_channels=32)      # ruijin-ldm_from_controlnet_ae.yaml:17-40
LDM_SCHEDULE = dict(timesteps=1000, linear_start=0.0015, linear_end=0.0195)
WORKLOADS["ldm_cfg3"]["net"] = LDM_AE_NET
WORKLOADS["ldm_cfg4"] = dict(spatial=(512, 512), C=1, batch=2, T=50, flop_per_sample=4.63e12, kind="ldm", net=LDM_PIXEL_NET,
                             desc="stage 2 of the full GuideGen pipeline: pixel-space LDM (ruijin-ldm_from_controlnet.yaml UNet), one CT slice "
                                  "512x512 conditioned on (previous slice | mask slice), DDIM 50 steps, n_samples 2; slices are sequential")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def build_and_seed_like(module, seed):
    """Deterministic state_dict (same on every rank) for replicated-weight multi-rank runs."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in module.state_dict().items():
        if v.ndim >= 2:
            out[k] = (torch.randn(v.shape, generator=g) / v[0].numel() ** 0.5).to(v.device)
        elif k.endswith("weight"):
            out[k] = (1.0 + 0.1 * torch.randn(v.shape, generator=g)).to(v.device)
        else:
            out[k] = (0.05 * torch.randn(v.shape, generator=g)).to(v.device)
    return out


def randomize_zero_modules(model, seed):
    """The reference zero-initialises the last conv of every ResBlock, every attention proj_out and the
    output conv: a random-init network would return a constant.  Re-draw those."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim >= 2 and float(p.abs().max()) == 0.0:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
            elif p.ndim == 1 and name.endswith("bias") and float(p.abs().max()) == 0.0 and "norm" not in name and ".0.bias" not in name:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


# =============================================================================== this repo's arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from jointimagegeneration_b200 import _C, ops
    from jointimagegeneration_b200.ccdm import build_model

    wl = WORKLOADS[args.workload]
    if wl.get("kind") == "ldm":
        return run_ours_ldm(args)
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N > 1 with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _C.check(_C.lib().gg_device_check(), "gg_device_check")
    K, W = args.steps, max(args.warmup, 3)
    B, Cc, sp, T = args.batch or wl["batch"], wl["C"], wl["spatial"], wl["T"]
    slab = bool(wl.get("slab"))
    full_sp = sp
    if slab:
        assert sp[0] % (16 * world) == 0, "depth must split into slabs that are multiples of 16 planes"
        sp = (sp[0] // world, sp[1], sp[2])
    V = sp[0] * sp[1] * sp[2]

    torch.manual_seed(1234 + rank)
    net = dict(CCDM_NET)
    context = None
    if wl.get("text"):
        net.update(use_spatial_transformer=True, transformer_depth=1, context_dim=768)
    model = build_model(T, "cosine", {"s": 0.008}, [(1,) + sp, (Cc,) + sp], None, "unet_openai", net, "synthetic",
                        "majority", dims=3)
    randomize_zero_modules(model.unet, 7)
    model = model.to(dev).eval()
    model.loop, model.use_cuda_graph, model.philox_seed = "resident", True, 99 + rank
    comm = None
    if slab:
        torch.manual_seed(1234)                       # every rank holds the same weights
        model.unet.load_state_dict({k: v for k, v in build_and_seed_like(model.unet, 1234).items()})
        model.philox_seed = 99
        if world > 1:
            from jointimagegeneration_b200.sharding import SlabComm
            comm = SlabComm()
            model.unet.enable_slab(comm)
            model.use_cuda_graph = False              # collectives between kernels: eager launches

    # ---- synthetic inputs, resident in HBM before the timed region
    lab0 = torch.randint(0, Cc, (B,) + sp, device=dev)
    x_T = torch.zeros((B, Cc) + sp, dtype=torch.float32, device=dev).scatter_(1, lab0[:, None], 1.0)
    cond = torch.zeros((B, 1) + sp, dtype=torch.float32, device=dev)          # ruijin.py:181-182: zeros
    t_values = list(range(T, 0, -1))
    coefs = model.diffusion.step_coef_tensor(torch.tensor(t_values)).to(dev)[:, None, :].expand(-1, B, -1).contiguous()
    if wl.get("text"):
        context = torch.randn((B, 512, 768), device=dev)
    st = model.resident_begin(x_T, cond, context)
    plan = st["plan"]

    def step(i):
        model.resident_step(st, t_values[i % (T - 1)], coefs[i % (T - 1)], offset=i)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    value = (1 if slab else world) * K / (ms / 1e3)
    launches_per_step = plan.num_launches + 1                                   # UNet plan + fused per-voxel kernel

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    x_host = x_T.cpu().pin_memory()
    c_host = cond.cpu().pin_memory()
    out_host = torch.empty((B, Cc) + sp, dtype=torch.int64).pin_memory()
    Ke = max(2, min(K, 10))
    # one untimed short call first: the public path's own one-time work (its plan / graph for this call signature,
    # pinned staging buffers) is not part of a steady-state step
    model(x_host, c_host, t=torch.tensor(10000 + 2), context=context)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = model(x_host, c_host, t=torch.tensor(10000 + Ke), context=context)["diffusion_out"]   # reference's own K-step knob (:190-197)
    out_host.copy_(res, non_blocking=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tm = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e = {"value": (1 if slab else world) * Ke / e2e_s, "unit": "steps/s", "steps": Ke,
           "h2d_bytes_per_step": (x_host.numel() * 4 + c_host.numel() * 4) // Ke,
           "d2h_bytes_per_step": out_host.numel() * 8 // Ke,
           "call": "DenoisingModel.forward(x_host, condition_host, t=10000+K) -> int64 one-hot on host (loop='resident', CUDA graph)"}

    line = {"metric": "denoising steps/sec", "value": value, "unit": "steps/s (1 step = UNet forward + categorical posterior/draw for a batch of %d volumes)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if slab else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "global_batch": B if slab else B * world, "batch_per_gpu": B,
                       "volume": list(full_sp), "local_volume": list(sp), "classes": Cc, "network": "ccdm/params.yml unet_openai (95.4 M params), random init, zero-init modules re-randomised",
                       "text_conditioning": "off -- the reference cannot construct its text-conditioned CCDM (SURVEY.md D1/D2); --text enables ours",
                       "parallelism": ("one volume in %d depth slabs: halo exchange / GN gather / KV gather over NCCL" % world) if slab
                       else "independent chains sharded over ranks (dp%d), no per-step collective" % world,
                       "l2": "inputs larger than L2 (activations are GBs per step); no explicit flush",
                       "rng": "in-kernel Philox", "cuda_graph": bool(model.use_cuda_graph)},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K,
            "volumes_per_sec": value * B / T}
    if comm is not None:
        nf = W + K + Ke + 1
        line["comm"] = {"halo_exchanges_per_forward": comm.n_exchanges / nf, "gathers_per_forward": comm.n_gathers / nf,
                        "halo_bytes_sent_per_forward": comm.bytes_sent / nf}

    if rank == 0 or comm is not None:      # slab mode: the instrumented pass contains collectives -> every rank runs it
        peaks = load_peaks()
        # ---- instrumented pass: CUDA-event time of every launch of one step (eager, same stream)
        kinds = {}
        torch.cuda.synchronize()
        reps = 2
        for _ in range(reps):
            evs = []
            s = _C.stream()
            for fn, fargs in plan.steps:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if hasattr(fn, "fn"):
                    fn.fn(*fargs)                                   # collective / halo exchange (host-side step)
                else:
                    _C.check(fn(*fargs, s), fn.__name__)
                b.record()
                evs.append((fn.__name__, a, b))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.cat_step_cl(plan.outputs["head"], st["lab_a"], coefs[5], st["lab_b"], B, V, Cc, mode=ops.CAT_SAMPLE,
                            cond=st["cond_cl"], n_cond=st["n_cond"], next_x=st["xin"], seed=1, offset=12345)
            b.record()
            evs.append(("gg_cat_step_cl", a, b))
            torch.cuda.synchronize()
            for name, a, b in evs:
                d = kinds.setdefault(name, [0.0, 0])
                d[0] += a.elapsed_time(b) / reps
                d[1] += 1
        if args.detail:
            import ctypes
            rows = []
            for (fn, fargs), (name, a, b) in zip(plan.steps, evs):
                d = a.elapsed_time(b)
                if name == "gg_conv_fwd":
                    ca = ctypes.cast(fargs[0], ctypes.POINTER(_C.ConvArgs)).contents if not hasattr(fargs[0], "_obj") else fargs[0]._obj
                    cin = sum(ca.src[i].C for i in range(ca.nsrc))
                    kk = _C.lib().gg_conv_packed_k(ctypes.byref(ca))
                    fl = 2.0 * ca.N * ca.Do * ca.Ho * ca.Wo * ca.Cout * kk
                    rows.append("conv N%d in %dx%dx%d out %dx%dx%d Cin %d(nsrc %d) Cout %d taps %dx%dx%d s%d K %d : %.3f ms %.0f TF/s"
                                % (ca.N, ca.D, ca.H, ca.W, ca.Do, ca.Ho, ca.Wo, cin, ca.nsrc, ca.Cout, ca.kd, ca.kh, ca.kw, ca.stride, kk, d,
                                   fl / d / 1e9))
                else:
                    rows.append("%s : %.3f ms" % (name, d))
            with open(os.path.join(ROOT, "gpurun_out", "bench_detail.txt"), "w") as f:
                f.write("\n".join(rows) + "\n")
        kernel_ms = {k: round(v[0], 4) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}
        # conv launches by kernel (gg_conv_args.algo): launches, ms and issued TFLOP/s (issued = incl. channel padding)
        import ctypes as _ct
        by_kernel = {}
        for (fn, fargs), (name, a, b) in zip(plan.steps, evs):
            if name != "gg_conv_fwd":
                continue
            ca = fargs[0]._obj if hasattr(fargs[0], "_obj") else _ct.cast(fargs[0], _ct.POINTER(_C.ConvArgs)).contents
            kname = {0: "conv_tcgen05_kernel", 4: "conv_roll_kernel"}.get(int(ca.algo), "conv_halo_kernel")
            kk = _C.lib().gg_conv_packed_k(_ct.byref(ca))
            e = by_kernel.setdefault(kname, {"launches": 0, "ms": 0.0, "issued_flop": 0.0})
            e["launches"] += 1
            e["ms"] += a.elapsed_time(b)
            e["issued_flop"] += 2.0 * ca.N * ca.Do * ca.Ho * ca.Wo * ca.Cout * kk
        for e in by_kernel.values():
            e["tflops_issued"] = round(e["issued_flop"] / max(e["ms"], 1e-9) / 1e9, 1)
            e["frac_of_peak"] = round(e["tflops_issued"] / peaks["tf_sustained"], 3)
            e["ms"] = round(e["ms"], 3)
        n_conv = kinds["gg_conv_fwd"][1] // reps
        conv_ms = kinds["gg_conv_fwd"][0]
        share = (1.0 / world) if slab else 1.0                          # FLOPs this rank executes
        attn_flops = wl.get("attn_flop", 0.022e12 / 6.324e12 * wl["flop_per_sample"])
        conv_alg = (wl["flop_per_sample"] - attn_flops) * B * share
        ach = conv_alg / (conv_ms / 1e3) / 1e12
        # DRAM bytes per conv launch from the committed ncu pass (profiles/r1_conv_dram_ccdm_cfg2.md): only valid for the
        # exact workload it was captured on (config 2, 8 volumes per GPU, no text conditioning, one rank's full volume)
        traffic, traffic_src = None, None
        if args.workload == "ccdm_cfg2" and B == 8 and not slab:
            traffic = 55.87e9 / 114
            traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 114 conv launches of one forward = 55.87 GB " \
                          "(profiles/r1_conv_dram_ccdm_cfg2.md); algorithmic conv input + output bytes: 56 GB (SURVEY.md 8d)"
        line["roofline"] = {"bound": "tensor", "kernel": "conv_roll_kernel + conv_halo_kernel + conv_tcgen05_kernel (all %d conv launches of one step)" % n_conv,
                            "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                            "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_flop": conv_alg, "issued_flop": plan.flops, "avg_launch_ms": conv_ms / n_conv,
                            "by_kernel": by_kernel,
                            "peak_source": peaks["source"] + " (bf16 sustained: kernel timed inside a long step)",
                            "whole_step_frac": wl["flop_per_sample"] * B * share / (ms / K / 1e3) / 1e12 / peaks["tf_sustained"]}
        cat_ms = kinds["gg_cat_step_cl"][0]
        alg_b = 50.0 * B * V
        act_b = (64 + 1 + 1 + 2 * plan.inputs["x"].shape[-1] + 2) * B * V
        line["roofline_hbm_resident"] = {"bound": "hbm", "kernel": "cat_step_cl_fast_kernel (softmax + posterior + clamp + Philox inverse-CDF draw + next input)",
                                         "achieved": alg_b / (cat_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                         "frac": alg_b / (cat_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                                         # ncu capture of this kernel's I/O at B = 8 (profiles/r1_ncu_full_summaries.md)
                                         "traffic": 786.8e6 if (args.workload == "ccdm_cfg2" and B == 8 and not slab) else None,
                                         "algorithmic_bytes": alg_b, "moved_bytes": act_b,
                                         "moved_frac": act_b / (cat_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "launch_ms": cat_ms}
        # the per-voxel kernel at the reference's tensor interface (fp32 [B,C,V] in/out, injected Exp(1) noise):
        # 16*C = 192 B/voxel algorithmic; tensors (4 x 403 MB) exceed L2
        x0p = torch.softmax(torch.randn((B, Cc) + sp, device=dev), 1)
        qn = torch.empty((B * V, Cc), device=dev).exponential_(1)
        outp = torch.empty_like(x0p)
        for _ in range(2):
            ops.cat_posterior_sample(x0p, x_T, coefs[5], ops.CAT_SAMPLE, q=qn, out=outp)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            ops.cat_posterior_sample(x0p, x_T, coefs[5], ops.CAT_SAMPLE, q=qn, out=outp)
        eb.record()
        torch.cuda.synchronize()
        pm = ea.elapsed_time(eb) / 5
        ib = 16.0 * Cc * B * V
        line["roofline_hbm"] = {"bound": "hbm", "kernel": "cat_posterior_kernel<12> (theta_post_prob + clamp + categorical draw, reference interface)",
                                "achieved": ib / (pm / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": ib / (pm / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": ib,
                                "launch_ms": pm, "peak_source": peaks["source"] + " (copy bandwidth)"}
        del x0p, qn, outp
        line["kernel_ms"] = kernel_ms
        line["arena_bytes"] = plan.arena_bytes
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args, wl, budget_s=20.0)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        model.unet.invalidate()           # drop plans (and any CUDA graph holding NCCL work) before the communicator
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def describe_launch(name, fargs, ms):
    import ctypes
    from jointimagegeneration_b200 import _C
    if name != "gg_conv_fwd":
        return "%s : %.3f ms" % (name, ms)
    ca = fargs[0]._obj
    cin = sum(ca.src[i].C for i in range(ca.nsrc))
    kk = _C.lib().gg_conv_packed_k(ctypes.byref(ca))
    fl = 2.0 * ca.N * ca.Do * ca.Ho * ca.Wo * ca.Cout * kk
    return ("conv N%d in %dx%dx%d out %dx%dx%d Cin %d(nsrc %d) Cout %d taps %dx%dx%d s%d algo%d K %d : %.3f ms %.0f TF/s"
            % (ca.N, ca.D, ca.H, ca.W, ca.Do, ca.Ho, ca.Wo, cin, ca.nsrc, ca.Cout, ca.kd, ca.kh, ca.kw, ca.stride, ca.algo, kk, ms,
               fl / ms / 1e9))


def instrument_plan(plan, extra=None, reps=2, detail_path=None):
    """CUDA-event time of every launch of one planned forward (eager, current stream) -> {name: [ms, count]}."""
    import torch
    from jointimagegeneration_b200 import _C
    kinds = {}
    torch.cuda.synchronize()
    for _ in range(reps):
        evs = []
        s = _C.stream()
        for fn, fargs in plan.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _C.check(fn(*fargs, s), fn.__name__)
            b.record()
            evs.append((fn.__name__, a, b))
        if extra is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            name = extra()
            b.record()
            evs.append((name, a, b))
        torch.cuda.synchronize()
        for name, a, b in evs:
            d = kinds.setdefault(name, [0.0, 0])
            d[0] += a.elapsed_time(b) / reps
            d[1] += 1
    if detail_path:
        with open(detail_path, "w") as f:
            for (fn, fargs), (name, a, b) in zip(plan.steps, evs):
                f.write(describe_launch(name, fargs, a.elapsed_time(b)) + "\n")
    return {k: (v[0], v[1] // reps) for k, v in kinds.items()}


def run_ours_ldm(args):
    """BASELINE config 3: one step = UNet eps-prediction for the batch + fused DDIM update."""
    import torch
    import torch.distributed as dist
    from jointimagegeneration_b200 import _C, ops
    from jointimagegeneration_b200.ldm import DDIMSampler, LatentDiffusion, UNetModel

    wl = WORKLOADS[args.workload]
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    assert world == args.gpus
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    B, S = args.batch or wl["batch"], wl["T"]
    hw = wl["spatial"]
    torch.manual_seed(4321 + rank)
    net = wl["net"]
    xc, cc = net["out_channels"], net["in_channels"] - net["out_channels"]
    unet = UNetModel(**net)
    randomize_zero_modules(unet, 7)
    ld = LatentDiffusion(unet, conditioning_key="concat", **LDM_SCHEDULE).to(dev).eval()
    unet.use_cuda_graph = True
    sampler = DDIMSampler(ld)
    sampler.make_schedule(S, ddim_eta=0.0, verbose=False)
    x = torch.randn((B, xc) + hw, device=dev)
    c = torch.randn((B, cc) + hw, device=dev)
    steps_t = [int(v) for v in sampler.ddim_timesteps[::-1]]
    ts = [torch.full((B,), v, device=dev, dtype=torch.long) for v in steps_t]
    state = {"x": x}

    def step(i):
        j = i % S
        state["x"], _ = sampler.p_sample_ddim(state["x"], c, ts[j], 2, index=S - 1 - j)
        if j == S - 1:
            state["x"] = x
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    if world > 1:
        tm = torch.tensor([ms], device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = world * K / (ms / 1e3)
    plan = unet.plan_for(B, hw)
    # e2e: the public call sample_cond makes (sample_diffusion.py:212), host conditioning in, host samples out
    c_host = c.cpu().pin_memory()
    out_host = torch.empty((B, xc) + hw).pin_memory()
    # one untimed short call first (as in the CCDM branch): one-time work of the public path is not a steady-state step
    sampler.sample(S=2, batch_size=B, shape=(xc,) + hw, conditioning=c_host.to(dev, non_blocking=True), eta=0.0, verbose=False, dims=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp, _ = sampler.sample(S=S, batch_size=B, shape=(xc,) + hw, conditioning=c_host.to(dev, non_blocking=True), eta=0.0, verbose=False, dims=2)
    out_host.copy_(smp)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    line = {"metric": "denoising steps/sec", "value": value, "unit": "steps/s (1 step = UNet eps forward + DDIM update for a batch of %d latents)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "global_batch": B * world, "batch_per_gpu": B,
                       "l2": "working set (267.5 M params bf16 + activations) exceeds L2; no explicit flush", "cuda_graph": True,
                       "parallelism": "independent samples sharded over ranks (dp%d)" % world},
            "clocks": clk, "gpu_launches": (plan.num_launches + 3) * K,
            "e2e": {"value": world * S / e2e_s, "unit": "steps/s", "steps": S, "h2d_bytes_per_step": c_host.numel() * 4 // S,
                    "d2h_bytes_per_step": out_host.numel() * 4 // S, "call": "DDIMSampler.sample(S=50, conditioning=host tensor) -> host"},
            "slices_per_sec": value * B / S}
    if rank == 0:
        peaks = load_peaks()
        kinds = instrument_plan(plan, detail_path=os.path.join(ROOT, "gpurun_out", "bench_detail_ldm.txt") if args.detail else None)
        conv_ms, n_conv = kinds["gg_conv_fwd"]
        ach = wl["flop_per_sample"] * B / (conv_ms / 1e3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv_roll_kernel + conv_halo_kernel + conv_tcgen05_kernel (all %d conv launches of one step)" % n_conv, "achieved": ach,
                            "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"], "traffic": None,
                            "algorithmic_flop": wl["flop_per_sample"] * B, "issued_flop": plan.flops,
                            "peak_source": peaks["source"] + " (bf16 sustained)",
                            "whole_step_frac": wl["flop_per_sample"] * B / (ms / K / 1e3) / 1e12 / peaks["tf_sustained"]}
        line["kernel_ms"] = {k: round(v[0], 4) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ============================================================================ CPU reference arm
def _oracle_step_fn(wl, B, sp):
    """One denoising step of the CPU oracle (restated reference, fp32 torch on all host threads) on a
    [B, C, *sp] sample.  bench.py may execute oracle/ only here (cpu_baseline / --impl reference)."""
    import numpy as np
    import torch
    from oracle import diffusion, nets, weights
    if wl.get("kind") == "ldm":
        from oracle import ddim
        sd = weights.synth_state_dict(weights.reference_shapes("LDM_AE"), 1)
        acp = ddim.alphas_cumprod_f32(ddim.make_beta_schedule_linear(1000, LDM_SCHEDULE["linear_start"], LDM_SCHEDULE["linear_end"]))
        x = weights.normal(2, (B, 4) + tuple(sp))
        c = weights.normal(3, (B, 4) + tuple(sp))

        def step_ldm():
            return ddim.ddim_sample(lambda xx, tt: nets.unet_forward(sd, torch.cat([xx, c], 1), tt, num_head_channels=32), acp, x, 1, 0.0)
        return step_ldm
    Cc, T = wl["C"], wl["T"]
    sd = weights.synth_state_dict(weights.reference_shapes("CCDM_PARAMS_YML"), 1)
    _, alphas, cumalphas = diffusion.cosine_schedule(T)
    V = int(np.prod(sp))
    xt = weights.uniform_one_hot(2, B, Cc, sp)
    cond = torch.zeros(B, 1, *sp)
    q = torch.from_numpy(weights.exp_noise(3, (1, B * V, Cc)))

    def step():
        def unet_fn(x, t):
            return nets.unet_forward(sd, x, t, input_condition=cond, softmax_output=True, num_head_channels=32)
        return diffusion.forward_denoising(unet_fn, alphas, cumalphas, xt, q, time_steps=T, t_values=[T // 2])
    return step


def cpu_reference(args, wl, budget_s=20.0, steps=1, warmup=0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if wl.get("kind") == "ldm":
        B = args.batch or wl["batch"]
        f = _oracle_step_fn(wl, 1, wl["spatial"])
        f()
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            f()
        dt = (time.perf_counter() - t0) / max(1, steps)
        return {"value": 1.0 / (dt * B), "unit": "steps/s", "cores": cores, "kind": "port",
                "sample": "oracle port on 1 of the %d latents of a batch step; %.2f s per sample step, extrapolated x%d" % (B, dt, B),
                "sample_seconds_per_step": dt}
    full_vox = wl["spatial"][0] * wl["spatial"][1] * wl["spatial"][2] * (args.batch or wl["batch"])
    # probe a small crop to size the sample for ~budget seconds of CPU work
    probe_sp = tuple(max(16, s // 8) for s in wl["spatial"])   # 4 stride-2 levels need multiples of 16
    f = _oracle_step_fn(wl, 1, probe_sp)
    f()
    t0 = time.perf_counter()
    f()
    probe_s = time.perf_counter() - t0
    per_vox = probe_s / (probe_sp[0] * probe_sp[1] * probe_sp[2])
    target_vox = budget_s / max(1, steps + warmup) / per_vox
    sp = list(probe_sp)
    full = list(wl["spatial"])
    # grow the crop by doubling dims (last first) while it stays within budget and within the volume
    for ax in (2, 1, 0, 2, 1, 0, 2, 1, 0):
        if sp[ax] * 2 <= full[ax] and sp[0] * sp[1] * sp[2] * 2 <= target_vox:
            sp[ax] *= 2
    sp = tuple(sp)
    f = _oracle_step_fn(wl, 1, sp)
    for _ in range(warmup):
        f()
    t0 = time.perf_counter()
    for _ in range(steps):
        f()
    dt = (time.perf_counter() - t0) / steps
    scale = full_vox / (sp[0] * sp[1] * sp[2])
    return {"value": 1.0 / (dt * scale), "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": "oracle port (fp32 torch CPU restatement of the reference) on 1 volume crop %dx%dx%d = 1/%.0f of a batch step; "
                      "%.2f s per sample step, extrapolated linearly in voxels" % (sp[0], sp[1], sp[2], scale, dt),
            "sample_seconds_per_step": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    # bounded: the whole run (W + K sample steps) stays within ~2.5 minutes of CPU work
    cb = cpu_reference(args, wl, budget_s=float(os.environ.get("BENCH_CPU_BUDGET_S", 150.0)), steps=max(1, K), warmup=max(0, min(W, 1)))
    B = args.batch or wl["batch"]
    line = {"impl": "reference", "metric": "denoising steps/sec", "value": cb["value"],
            "unit": "steps/s (1 step = UNet forward + categorical posterior/draw for a batch of %d volumes)" % B,
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "global_batch": B, "note": "CPU arm runs on rank 0 only; "
                       "the reference is pure Python/PyTorch and /root/reference does not travel to the GPU box, so the oracle port "
                       "(pinned against the unmodified reference in tests/test_oracle_pinning.py) stands in for it"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    # NCCL prints its version banner on STDOUT when NCCL_DEBUG asks for it; stdout must carry ONE JSON line
    if not os.environ.get("BENCH_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ccdm_cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override batch per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--detail", action="store_true", help="write per-launch times of one step to gpurun_out/bench_detail.txt")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
