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
_DIST = {}


def dist_setup(args):
    """(rank, local_rank, world, device); initialises NCCL once per process."""
    import torch
    import torch.distributed as dist
    if not _DIST:
        rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N > 1 with torch.distributed.run)"
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        _DIST.update(rank=rank, local=local, world=world, dev=dev)
    return _DIST["rank"], _DIST["local"], _DIST["world"], _DIST["dev"]


def _max_over_ranks(v, dev, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(v)
    t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _ccdm_model(T, sp, Cc, dev, text=False, same_weights_seed=None):
    import torch
    from jointimagegeneration_b200.ccdm import build_model
    net = dict(CCDM_NET)
    if text:
        net.update(use_spatial_transformer=True, transformer_depth=1, context_dim=768)
    model = build_model(T, "cosine", {"s": 0.008}, [(1,) + tuple(sp), (Cc,) + tuple(sp)], None, "unet_openai", net, "synthetic",
                        "majority", dims=3)
    if same_weights_seed is None:
        randomize_zero_modules(model.unet, 7)
    else:
        model.unet.load_state_dict(build_and_seed_like(model.unet, same_weights_seed))     # identical on every rank
    return model.to(dev).eval()


def run_ours(args, workload=None, K=None, W=None, sub=False):
    """One workload on this repo's sm_100a path -> its JSON line as a dict on rank 0 (None elsewhere).
    sub=True: a short secondary measurement (no CPU / eager baseline legs)."""
    import torch
    import torch.distributed as dist
    from jointimagegeneration_b200 import _C, ops

    workload = workload or args.workload
    wl = WORKLOADS[workload]
    if wl.get("kind") == "ldm":
        return run_ours_ldm(args, workload, K, W)
    rank, local, world, dev = dist_setup(args)
    _C.check(_C.lib().gg_device_check(), "gg_device_check")
    K, W = K or args.steps, max(W or args.warmup, 3)
    B, Cc, sp, T = (args.batch if not sub else 0) or wl["batch"], wl["C"], wl["spatial"], wl["T"]
    slab = bool(wl.get("slab"))
    full_sp = sp
    if slab:
        assert sp[0] % (16 * world) == 0, "depth must split into slabs that are multiples of 16 planes"
        sp = (sp[0] // world, sp[1], sp[2])
    V = sp[0] * sp[1] * sp[2]

    torch.manual_seed(1234 + rank)
    context = None
    model = _ccdm_model(T, sp, Cc, dev, text=bool(wl.get("text")), same_weights_seed=1234 if slab else None)
    # one Philox key for the whole job; the counter is the GLOBAL (chain, voxel) index, so results do not depend on N
    model.loop, model.use_cuda_graph, model.philox_seed, model.chain_base = "resident", True, 99, (0 if slab else rank * B)
    comm = None
    if slab and world > 1:
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
    n0 = _C.launch_count()
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    ms = _max_over_ranks(ms, dev, world)
    value = (1 if slab else world) * K / (ms / 1e3)
    launches_per_step = model.launches_per_step(plan)                           # kernels of libguidegen_sm100 per resident step

    # ---- end to end through the public API, as the reference's own caller drives it (Trainer.test_step, ccdm/ddpm/trainer.py:
    # 412-421; evaluator.py:135-148): the batch's condition volume comes from HOST memory, x_T is drawn ON THE DEVICE with
    # OneHotCategoricalBCHW(logits=zeros).sample() (trainer.py:418), DenoisingModel.forward runs the chain, and the result
    # the caller keeps -- the arg-max label volume it writes to NIfTI -- goes back to the host.  Copies are inside the timed region.
    from jointimagegeneration_b200.ccdm import OneHotCategoricalBCHW
    c_host = cond.cpu().pin_memory()
    lab_host = torch.empty((B,) + sp, dtype=torch.uint8).pin_memory()
    Ke = max(2, min(K, 10))

    def test_step(k):
        image = c_host.to(dev, non_blocking=True)
        x = OneHotCategoricalBCHW(logits=torch.zeros((B, Cc) + sp, device=dev)).sample()
        pred = model(x, image, t=torch.tensor(10000 + k), context=context)["diffusion_out"]     # the reference's own K-step knob (:190-197)
        lab_host.copy_(pred.argmax(dim=1).to(torch.uint8), non_blocking=False)

    # one untimed short call first: the public path's own one-time work (its plan / graph for this call signature,
    # pinned staging buffers) is not part of a steady-state step
    test_step(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    test_step(Ke)
    torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev, world)
    e2e = {"value": (1 if slab else world) * Ke / e2e_s, "unit": "steps/s", "steps": Ke,
           "h2d_bytes_per_step": c_host.numel() * 4 // Ke, "d2h_bytes_per_step": lab_host.numel() // Ke,
           "call": "Trainer.test_step flow (trainer.py:418-421): condition from pinned host memory -> x_T = OneHotCategoricalBCHW(logits=0).sample() on "
                   "the device -> DenoisingModel.forward(x_T, condition, t=10000+K) -> arg-max label volume (uint8) to the host "
                   "(loop='resident', CUDA graph)"}

    line = {"metric": "denoising steps/sec", "value": value, "unit": "steps/s (1 step = UNet forward + categorical posterior/draw for a batch of %d volumes)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if slab else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(workload, B, world),
            "detail": {"batch_per_gpu": B, "volume": list(full_sp), "local_volume": list(sp), "classes": Cc,
                       "network": "ccdm/params.yml unet_openai (95.4 M params), random init, zero-init modules re-randomised",
                       "text_conditioning": "off -- the reference cannot construct its text-conditioned CCDM (SURVEY.md D1/D2); --workload ccdm_cfg2_text enables ours",
                       "parallelism": ("one volume in %d depth slabs: halo exchange / GN gather / KV gather over NCCL" % world) if slab
                       else "independent chains sharded over ranks (dp%d), no per-step collective" % world,
                       "l2": "inputs larger than L2 (activations are GBs per step); no explicit flush",
                       "rng": "in-kernel Philox keyed on the global (chain, voxel) index", "cuda_graph": bool(model.use_cuda_graph),
                       "head": "sampler fused into the head conv" if st.get("fused_head") else "head logits + per-voxel kernel"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K,
            "volumes_per_sec": value * B / T}
    if comm is not None:
        nf = max(1, comm.n_forwards) if hasattr(comm, "n_forwards") else (W + K + Ke + 1)
        line["comm"] = {"halo_exchanges_per_forward": comm.n_exchanges / nf, "gathers_per_forward": comm.n_gathers / nf,
                        "halo_bytes_sent_per_forward": comm.bytes_sent / nf}

    if rank == 0 or comm is not None:      # slab mode: the instrumented pass contains collectives -> every rank runs it
        peaks = load_peaks()
        # ---- instrumented pass: CUDA-event time of every launch of one step (eager, same stream)
        body, tail = model.resident_tail_launcher(st, coefs[5], offset=12345)
        kinds, evs = instrument_plan(plan, extra=tail, reps=2, steps=body,
                                     detail_path=os.path.join(ROOT, "gpurun_out", "bench_detail_%s.txt" % workload) if args.detail else None)
        kernel_ms = {k: round(v[0], 4) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}
        # conv launches by kernel (gg_conv_args.algo): launches, ms and issued TFLOP/s (issued = the MACs really executed)
        import ctypes as _ct
        by_kernel = {}
        conv_ms, n_conv, issued = 0.0, 0, 0.0
        for (name, fargs, d) in evs:
            if name != "gg_conv_fwd":
                continue
            ca = fargs[0]._obj
            kname = {0: "conv_tcgen05_kernel", 4: "conv_roll_kernel"}.get(int(ca.algo), "conv_halo_kernel")
            kk = _C.lib().gg_conv_packed_k(_ct.byref(ca))
            fl = 2.0 * ca.N * ca.Do * ca.Ho * ca.Wo * ca.Cout * kk
            e = by_kernel.setdefault(kname, {"launches": 0, "ms": 0.0, "issued_flop": 0.0})
            e["launches"] += 1
            e["ms"] += d
            e["issued_flop"] += fl
            conv_ms, n_conv, issued = conv_ms + d, n_conv + 1, issued + fl
        for e in by_kernel.values():
            e["tflops_issued"] = round(e["issued_flop"] / max(e["ms"], 1e-9) / 1e9, 1)
            e["frac_of_peak"] = round(e["tflops_issued"] / peaks["tf_sustained"], 3)
            e["ms"] = round(e["ms"], 3)
        share = (1.0 / world) if slab else 1.0                          # FLOPs this rank executes
        attn_flops = wl.get("attn_flop", 0.022e12 / 6.324e12 * wl["flop_per_sample"])
        conv_alg = (wl["flop_per_sample"] - attn_flops) * B * share
        ach = conv_alg / (conv_ms / 1e3) / 1e12
        # DRAM bytes per conv launch: a CONSTANT from a committed ncu pass, valid only for the exact workload it was
        # captured on (config 2, 8 volumes per GPU, no text conditioning, one rank's full volume); not measured in-run
        traffic, traffic_src = None, None
        if workload == "ccdm_cfg2" and B == 8 and not slab:
            traffic = 55.36e9 / 114
            traffic_src = {"constant_from": "profiles/r2_conv_dram_ccdm_cfg2.md",
                           "what": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 114 conv launches of one forward = 55.36 GB (final code of round 2); "
                                   "algorithmic conv input + output bytes: 56 GB (SURVEY.md 8d)"}
        line["roofline"] = {"bound": "tensor", "kernel": "conv_roll_kernel + conv_halo_kernel + conv_tcgen05_kernel (all %d conv launches of one step)" % n_conv,
                            "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
                            "achieved_issued": issued / (conv_ms / 1e3) / 1e12,
                            "frac_issued": issued / (conv_ms / 1e3) / 1e12 / peaks["tf_sustained"],
                            "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_flop": conv_alg, "issued_flop": issued, "avg_launch_ms": conv_ms / n_conv,
                            "by_kernel": by_kernel,
                            "peak_source": peaks["source"] + " (bf16 sustained: kernel timed inside a long step)",
                            "note": "frac counts the reference's FLOPs (the folded upsample is credited 27 taps for 8 issued); frac_issued counts "
                                    "the MACs the tensor core executes (incl. channel padding)",
                            "whole_step_frac": wl["flop_per_sample"] * B * share / (ms / K / 1e3) / 1e12 / peaks["tf_sustained"]}
        alg_b = 50.0 * B * V
        if st.get("fused_head"):
            # the sampler lives in the head conv's epilogue: its HBM traffic is the conv's input planes (2 * 64 B/voxel, read once)
            # + 1 B label in + 1 B label out + 32 B next-input row; there is no logits tensor and no separate per-voxel kernel
            head_ms = [d for (name, fargs, d) in evs if name == "gg_conv_fwd"][-1]
            moved = (128.0 + 1 + 1 + 2 * plan.inputs["x"].shape[-1]) * B * V
            line["roofline_hbm_resident"] = {"bound": "hbm", "kernel": "conv_roll_kernel<.., 16, .., SAMPLER> (64->12 head conv with softmax + posterior + clamp + "
                                             "Philox inverse-CDF draw + next-input write in its epilogue; the separate per-voxel kernel and the fp32 logits are gone)",
                                             "achieved": moved / (head_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                             "frac": moved / (head_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                             "algorithmic_bytes": moved, "per_voxel_algorithmic_bytes_sampler_only": 50, "launch_ms": head_ms,
                                             "note": "algorithmic bytes of the FUSED launch = conv input (64 ch bf16) + labels + next input; "
                                                     "the sampler adds 34 B/voxel to a launch that already streams 128 B/voxel"}
        else:
            cat_ms = kinds["gg_cat_step_cl"][0]
            act_b = (64 + 1 + 1 + 2 * plan.inputs["x"].shape[-1] + 2) * B * V
            line["roofline_hbm_resident"] = {"bound": "hbm", "kernel": "cat_step_cl_fast_kernel (softmax + posterior + clamp + Philox inverse-CDF draw + next input)",
                                             "achieved": alg_b / (cat_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                             "frac": alg_b / (cat_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                                             "traffic": None, "algorithmic_bytes": alg_b, "moved_bytes": act_b,
                                             "moved_frac": act_b / (cat_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "launch_ms": cat_ms}
        if not sub:
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
    del st, plan
    model.unet.invalidate()           # drop plans (and any CUDA graph) before anything else is built / torn down
    del model
    torch.cuda.empty_cache()
    return line if rank == 0 else None


def bench_config(workload, B, world):
    """The `config` object: identical in this repo's arm and in the reference arm for the same workload."""
    wl = WORKLOADS[workload]
    slab = bool(wl.get("slab"))
    return {"workload": wl["desc"], "name": workload, "global_batch": B if slab else B * world}


def describe_launch(name, fargs, ms):
    import ctypes
    from jointimagegeneration_b200 import _C
    if name == "gg_peer_exchange":
        xa = fargs[0]._obj
        return "gg_peer_exchange sends %d (%s B) waits %d copies %d : %.3f ms" % (
            xa.nsend, "+".join(str(int(xa.bytes[i])) for i in range(xa.nsend)), xa.nflag_in, xa.ncopy, ms)
    if name != "gg_conv_fwd":
        return "%s : %.3f ms" % (name, ms)
    ca = fargs[0]._obj
    cin = sum(ca.src[i].C for i in range(ca.nsrc))
    kk = _C.lib().gg_conv_packed_k(ctypes.byref(ca))
    fl = 2.0 * ca.N * ca.Do * ca.Ho * ca.Wo * ca.Cout * kk
    return ("conv N%d in %dx%dx%d out %dx%dx%d Cin %d(nsrc %d) Cout %d taps %dx%dx%d s%d algo%d K %d : %.3f ms %.0f TF/s"
            % (ca.N, ca.D, ca.H, ca.W, ca.Do, ca.Ho, ca.Wo, cin, ca.nsrc, ca.Cout, ca.kd, ca.kh, ca.kw, ca.stride, ca.algo, kk, ms,
               fl / ms / 1e9))


def instrument_plan(plan, extra=None, reps=2, detail_path=None, steps=None):
    """CUDA-event time of every launch of one planned forward (eager, current stream).
    Returns ({name: (ms, count)}, [(name, args, ms) per launch of the last repetition])."""
    import torch
    from jointimagegeneration_b200 import _C
    kinds, evs = {}, []
    torch.cuda.synchronize()
    steps = plan.steps if steps is None else steps
    for _ in range(reps):
        evs = []
        s = _C.stream()
        for fn, fargs in steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if hasattr(fn, "fn"):
                fn.fn(*fargs)                                   # collective / halo exchange (host-side step)
            else:
                _C.check(fn(*fargs, s), fn.__name__)
            b.record()
            evs.append((fn.__name__, fargs, a, b))
        if extra is not None:
            for name, launch in extra:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fargs = launch()
                b.record()
                evs.append((name, fargs, a, b))
        torch.cuda.synchronize()
        for name, fargs, a, b in evs:
            d = kinds.setdefault(name, [0.0, 0])
            d[0] += a.elapsed_time(b) / reps
            d[1] += 1
    out = [(name, fargs, a.elapsed_time(b)) for name, fargs, a, b in evs]
    if detail_path:
        os.makedirs(os.path.dirname(detail_path), exist_ok=True)
        with open(detail_path, "w") as f:
            for name, fargs, d in out:
                f.write(describe_launch(name, fargs, d) + "\n")
    return {k: (v[0], v[1] // reps) for k, v in kinds.items()}, out


def run_ours_ldm(args, workload, K=None, W=None):
    """BASELINE configs 3 / 4 (stage 2): one step = UNet eps-prediction for the batch + fused DDIM update."""
    import torch
    import torch.distributed as dist
    from jointimagegeneration_b200 import _C
    from jointimagegeneration_b200.ldm import DDIMSampler, LatentDiffusion, UNetModel

    wl = WORKLOADS[workload]
    rank, local, world, dev = dist_setup(args)
    K, W = K or args.steps, max(W or args.warmup, 3)
    B, S = wl["batch"], wl["T"]
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
    ms = _max_over_ranks(e0.elapsed_time(e1), dev, world)
    clk = clocks.stop()
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
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev, world)
    line = {"metric": "denoising steps/sec", "value": value, "unit": "steps/s (1 step = UNet eps forward + DDIM update for a batch of %d)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(workload, B, world),
            "detail": {"batch_per_gpu": B, "l2": "working set (weights bf16 + activations) exceeds L2; no explicit flush", "cuda_graph": True,
                       "parallelism": "independent samples sharded over ranks (dp%d)" % world},
            "clocks": clk, "gpu_launches": (plan.num_launches + 3) * K,
            "e2e": {"value": world * S / e2e_s, "unit": "steps/s", "steps": S, "h2d_bytes_per_step": c_host.numel() * 4 // S,
                    "d2h_bytes_per_step": out_host.numel() * 4 // S, "call": "DDIMSampler.sample(S=%d, conditioning=host tensor) -> host" % S},
            "slices_per_sec": value * B / S}
    if rank == 0:
        peaks = load_peaks()
        line["detail"]["lanes"] = len(plan.lanes)
        iso = plan
        if len(plan.lanes) > 1:
            # the timed step runs the batch as concurrent sample lanes (Plan.lanes); per-launch times are taken from the
            # SINGLE-lane plan of the same batch, every kernel alone on the GPU
            line["detail"]["lanes_note"] = ("the step runs %d sample lanes concurrently (parallel branches of one CUDA graph); roofline / kernel_ms "
                                            "time every launch of the single-lane plan alone on the GPU" % len(plan.lanes))
            prev = os.environ.get("GG_LANES")
            os.environ["GG_LANES"] = "1"
            iso = unet.plan_for(B, hw)
            iso.run()
            if prev is None:
                del os.environ["GG_LANES"]
            else:
                os.environ["GG_LANES"] = prev
        kinds, _ = instrument_plan(iso, detail_path=os.path.join(ROOT, "gpurun_out", "bench_detail_%s.txt" % workload) if args.detail else None)
        del iso
        conv_ms, n_conv = kinds["gg_conv_fwd"]
        ach = wl["flop_per_sample"] * B / (conv_ms / 1e3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv_roll_kernel + conv_halo_kernel + conv_tcgen05_kernel (all %d conv launches of one step)" % n_conv, "achieved": ach,
                            "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"], "traffic": None,
                            "algorithmic_flop": wl["flop_per_sample"] * B, "issued_flop": plan.flops,
                            "peak_source": peaks["source"] + " (bf16 sustained)",
                            "whole_step_frac": wl["flop_per_sample"] * B / (ms / K / 1e3) / 1e12 / peaks["tf_sustained"]}
        line["kernel_ms"] = {k: round(v[0], 4) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}
    del plan, sampler
    unet.invalidate()
    del ld, unet
    torch.cuda.empty_cache()
    return line if rank == 0 else None


# ===================================================================== BASELINE config 4: one whole volume, both stages
def run_pipeline_cfg4(args, chain_steps=1000, n_samples=1):
    """Full GuideGen pipeline for ONE volume through jointimagegeneration_b200.pipeline.GuideGenPipeline:
    stage 1 (CCDM mask sampler, 128x128x64, `chain_steps` reverse steps) -> bridge (argmax labels -> scipy-rule zoom to
    64 x 512 x 512, / 255) -> stage 2 (pixel-space LDM, every slice: 50 DDIM steps at 512^2 conditioned on the previous
    generated slice and the mask slice, min-max normalise).  sample_diffusion.py:196-224.  Synthetic weights."""
    import torch
    from jointimagegeneration_b200.ldm import LatentDiffusion, UNetModel
    from jointimagegeneration_b200.pipeline import GuideGenPipeline
    rank, local, world, dev = dist_setup(args)
    wl2 = WORKLOADS["ccdm_cfg2"]
    sp, Cc = wl2["spatial"], wl2["C"]
    torch.manual_seed(777 + rank)
    mask_model = _ccdm_model(1000, sp, Cc, dev)
    mask_model.loop, mask_model.use_cuda_graph = "resident", True
    unet = UNetModel(**LDM_PIXEL_NET)
    randomize_zero_modules(unet, 7)
    ld = LatentDiffusion(unet, conditioning_key="concat", **LDM_SCHEDULE).to(dev).eval()
    unet.use_cuda_graph = True
    pipe = GuideGenPipeline(ld, mask_model, ddim_steps=50, ddim_eta=0.0)
    lab0 = torch.randint(0, Cc, (1,) + sp, device=dev)
    x_T = torch.zeros((1, Cc) + sp, device=dev).scatter_(1, lab0[:, None], 1.0)
    cond = torch.zeros((1, 1) + sp, device=dev)
    # warm both stages once (plans, graphs) -- a service generates many volumes with the same plans
    pipe.generate_mask(x_T, cond, init_t=10000 + 2)
    wm = torch.zeros((1, 1, 4, 512, 512), device=dev)
    wm[:, :, 1:3] = 0.01
    pipe.sample_cond(wm, n_samples=n_samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    labels = pipe.generate_mask(x_T, cond, init_t=None if chain_steps >= 1000 else 10000 + chain_steps)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    wholemask = pipe.mask_to_ct_grid(labels[0], size=(512, 512))
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    pred = pipe.sample_cond(wholemask, n_samples=n_samples)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    D = wholemask.shape[2]
    nz = torch.where(wholemask.sum((0, 1, 3, 4)))[0]
    n_slices = int(nz[-1]) - int(nz[0]) + 2
    # device time of one steady-state DDIM step at this batch (graph replay + fused update), for the host-gap figure
    sampler = pipe.sampler
    x = torch.randn((n_samples, 1, 512, 512), device=dev)
    c = torch.randn((n_samples, 2, 512, 512), device=dev)
    ts = torch.full((n_samples,), 501, device=dev, dtype=torch.long)
    for _ in range(3):
        sampler.p_sample_ddim(x, c, ts, 2, index=25)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        sampler.p_sample_ddim(x, c, ts, 2, index=25)
    b.record()
    torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / 10
    total = t3 - t0
    out = {"metric": "volumes/sec (full pipeline, one 512x512x64 CT volume)", "value": 1.0 / total, "unit": "volumes/s", "seconds_per_volume": total,
           "stage1_s": t1 - t0, "stage1_steps": chain_steps, "stage1_ms_per_step": (t1 - t0) / chain_steps * 1e3,
           "bridge_ms": (t2 - t1) * 1e3, "stage2_s": t3 - t2, "slices": n_slices, "ddim_steps": 50, "n_samples": n_samples,
           "ms_per_slice": (t3 - t2) / n_slices * 1e3, "device_ms_per_ddim_step": step_ms,
           "host_gap_ms_per_slice": (t3 - t2) / n_slices * 1e3 - 50 * step_ms,
           "output_shape": list(pred.shape), "finite": bool(torch.isfinite(pred).all()),
           "config": {"workload": "Full GuideGen pipeline: mask sampler (128x128x64, 12 classes, %d steps) -> zoom/slice bridge -> autoregressive "
                                  "pixel-space LDM CT 512x512x64 (50 DDIM steps per slice, n_samples %d)" % (chain_steps, n_samples), "name": "pipeline_cfg4"}}
    del pipe, ld, unet, mask_model, pred, wholemask
    torch.cuda.empty_cache()
    return out


# ======================================================= the same torch modules on the SAME GPU
def gpu_eager_baseline(workload, dev, budget_s=25.0):
    """The oracle port's torch functions (F.conv3d / group_norm / softmax / the reference's O(C^2) theta_post_prob einsum /
    argmax(p/q) draw) executed on the B200 under cuDNN / cuBLAS: fp32 (TF32 off), TF32, bf16 autocast.  One sample of
    the batch (the reference's [B, C, C, ...] posterior intermediates do not fit at B = 8), extrapolated x B.
    A bench/test leg only: nothing in the product path can reach it."""
    import torch
    from oracle import diffusion, nets, weights
    wl = WORKLOADS[workload]
    Cc, sp, T, B = wl["C"], wl["spatial"], wl["T"], wl["batch"]
    V = sp[0] * sp[1] * sp[2]
    sd = {k: v.to(dev) for k, v in weights.synth_state_dict(weights.reference_shapes("CCDM_PARAMS_YML"), 1).items()}
    _, alphas, cumalphas = diffusion.cosine_schedule(T)
    alphas, cumalphas = alphas.to(dev), cumalphas.to(dev)
    xt = weights.uniform_one_hot(2, 1, Cc, sp).to(dev)
    cond = torch.zeros((1, 1) + tuple(sp), device=dev)
    q = torch.empty((V, Cc), device=dev).exponential_(1)
    t = torch.full((1,), T // 2, device=dev, dtype=torch.long)

    def step():
        x0 = nets.unet_forward(sd, xt, t.float(), input_condition=cond, softmax_output=True, num_head_channels=32).float()
        post = diffusion.theta_post_prob_literal(alphas, cumalphas, Cc, xt, x0, t).clamp(min=1e-12)
        p2 = post.permute(0, 2, 3, 4, 1).reshape(-1, Cc)
        return (p2 / p2.sum(-1, keepdim=True) / q).argmax(-1)

    out = {"batch": 1, "extrapolated": True, "extrapolation": "x%d samples (linear; B = %d of the reference's posterior does not fit)" % (B, B),
           "what": "oracle port (restated reference modules) on cuda: cuDNN conv3d, torch group_norm/softmax/einsum posterior, same B200", "modes": {}}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    t_begin = time.perf_counter()
    try:
        for name, tf32, amp in (("bf16_autocast", True, True), ("tf32", True, False), ("fp32", False, False)):
            if time.perf_counter() - t_begin > budget_s:
                out["modes"][name] = {"skipped": "time budget"}
                continue
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    step()                                             # cuDNN autotune + warm-up
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    n = 2
                    for _ in range(n):
                        step()
                    b.record()
                    torch.cuda.synchronize()
                ms1 = a.elapsed_time(b) / n
                out["modes"][name] = {"ms_per_sample_step": ms1, "ms_per_step": ms1 * B, "steps_per_sec": 1e3 / (ms1 * B)}
            except Exception as e:  # noqa: BLE001  (e.g. out of memory in one precision: report, keep the others)
                out["modes"][name] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    best = [m for m in out["modes"].values() if "ms_per_step" in m]
    if best:
        f = min(best, key=lambda m: m["ms_per_step"])
        out["ms_per_step"], out["dtype"] = f["ms_per_step"], [k for k, v in out["modes"].items() if v is f][0]
    del sd
    torch.cuda.empty_cache()
    return out


# ================================================================ depth slabs over the N GPUs of this run (config 5)
def slab_section(args):
    """N > 1: (1) slab parity vs the unsplit computation on the same GPU (jointimagegeneration_b200.slab_check, the
    check tests/test_gpu_multi.py runs), (2) BASELINE config 5 -- ONE 256x256x128 volume in N depth slabs -- timed against
    the same volume unsplit on one GPU of this run.  Every rank calls this."""
    import torch
    import torch.distributed as dist
    from jointimagegeneration_b200 import slab_check
    from jointimagegeneration_b200.sharding import SlabComm, slab_ranges
    rank, local, world, dev = dist_setup(args)
    out = {}
    Cc = 12
    # ---- (1) parity
    spatial = (16 * world, 32, 32)
    m = _ccdm_model(20, spatial, Cc, dev, same_weights_seed=1234)
    g = torch.Generator().manual_seed(77)
    lab = torch.randint(0, Cc, (1,) + spatial, generator=g).to(dev)
    x = torch.zeros((1, Cc) + spatial, device=dev).scatter_(1, lab[:, None], 1.0)
    cond = torch.zeros((1, 1) + spatial, device=dev)
    rec = slab_check.unsplit_chain(m, x, cond, [13, 12, 11], seed=5)
    # transport: gg_peer_exchange kernels over NVLink peer memory (graph-capturable, no NCCL on the data path); if the
    # cudaIpc mapping is not available on this box, the host-enqueued NCCL transport of round 1
    comm, why = None, None
    if os.environ.get("GG_SLAB_TRANSPORT", "peer") == "peer":
        try:
            from jointimagegeneration_b200.sharding import PeerSlabComm
            comm = PeerSlabComm(arena_bytes=int(os.environ.get("GG_SLAB_ARENA_MB", "3072")) << 20)
        except Exception as e:  # noqa: BLE001
            comm, why = None, repr(e)[:200]
    ok = torch.tensor([1.0 if comm is not None else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok) == 0.0:
        if comm is not None:
            comm.close()
        comm = SlabComm()
        out["transport_fallback_reason"] = why or "a peer rank could not map the arenas"
    res = slab_check.reduce_over_ranks(slab_check.slab_vs_unsplit(m, rec, world, comm), dev)
    out.update(parity_volume=list(spatial), parity_max_abs=res["parity_max_abs"], bit_equal=res["bit_equal"],
               label_agreement_teacher_forced=res["agree"],
               parity_note="slab probabilities vs the unsplit plan on the same GPU (worst rank); labels of 3 teacher-forced sampler steps")
    m.unet.invalidate()
    del m, rec
    torch.cuda.empty_cache()
    # ---- (2) config 5 timing
    wl = WORKLOADS["ccdm_cfg5"]
    full, T = wl["spatial"], wl["T"]
    if full[0] % (16 * world) != 0:
        out["timing_skipped"] = "depth %d does not split into %d slabs of a multiple of 16 planes" % (full[0], world)
        return out
    model = _ccdm_model(T, full, Cc, dev, same_weights_seed=1234)
    model.loop, model.philox_seed, model.use_cuda_graph = "resident", 99, True
    g = torch.Generator().manual_seed(78)
    lab = torch.randint(0, Cc, (1,) + full, generator=g, dtype=torch.uint8).to(dev).long()
    x_T = torch.zeros((1, Cc) + full, device=dev).scatter_(1, lab[:, None], 1.0)
    cond = torch.zeros((1, 1) + full, device=dev)
    t_values = list(range(T, 0, -1))
    coefs = model.diffusion.step_coef_tensor(torch.tensor(t_values)).to(dev)[:, None, :].contiguous()

    def timed(st, n_warm, n):
        for i in range(n_warm):
            model.resident_step(st, t_values[i], coefs[i], offset=i)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            model.resident_step(st, t_values[n_warm + i], coefs[n_warm + i], offset=n_warm + i)
        b.record()
        torch.cuda.synchronize()
        return _max_over_ranks(a.elapsed_time(b) / n, dev, world)

    st = model.resident_begin(x_T, cond)
    ms1 = timed(st, 2, 3)                               # the whole volume on ONE GPU (every rank measures it; max reported)
    del st
    model.unet.invalidate()
    torch.cuda.empty_cache()
    lo, hi = slab_ranges(full[0], world)[rank]
    model.unet.enable_slab(comm)
    model.use_cuda_graph = bool(getattr(comm, "peer", False))     # kernels only -> the slab forward replays as a CUDA graph
    e0, g0, b0 = comm.n_exchanges, comm.n_gathers, comm.bytes_sent
    st = model.resident_begin(x_T[:, :, lo:hi].contiguous(), cond[:, :, lo:hi].contiguous())
    n_warm, n = 2, 5
    msN = timed(st, n_warm, n)
    nf = n_warm + n
    peer = bool(getattr(comm, "peer", False))
    nfp = 1 if peer else nf           # the peer transport counts at PLAN time (one plan), the NCCL one at every call
    out.update(volume=list(full), slabs=world, ms_per_step=msN, ms_per_step_n1=ms1, steps_per_sec=1e3 / msN,
               speedup_vs_n1=ms1 / msN, efficiency=ms1 / msN / world,
               halo_exchanges_per_forward=(comm.n_exchanges - e0) / nfp, gathers_per_forward=(comm.n_gathers - g0) / nfp,
               nvlink_bytes_sent_per_forward=(comm.bytes_sent - b0) / nfp, transport=getattr(comm, "transport", "nccl"),
               cuda_graph=bool(model.use_cuda_graph))
    # communication share: CUDA-event time around the collective steps of one forward (eager; includes waiting for the peers)
    kinds, _ = instrument_plan(st["plan"], reps=2, detail_path=os.path.join(ROOT, "gpurun_out", "bench_detail_slab_n%d_r%d.txt" % (world, rank))
                               if args.detail else None)
    out["kernel_ms_rank0"] = {k: round(v[0], 3) for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}
    cm = {k: round(_max_over_ranks(v[0], dev, world), 3) for k, v in kinds.items()
          if k in ("all_gather", "exchange_halo", "gg_peer_exchange", "gg_peer_epoch_inc")}
    out["comm_ms"] = cm
    out["comm_frac_of_step"] = sum(cm.values()) / msN if cm else None
    del st
    model.unet.invalidate()
    model.unet.enable_slab(None)
    del model
    torch.cuda.synchronize()
    dist.barrier()
    if hasattr(comm, "close"):
        comm.close()
    torch.cuda.empty_cache()
    return out


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
    """Times the CPU oracle (kind "port") on a bounded sample of one batch step: ONE full sample of the batch when
    (steps + warmup) of them fit the budget (they do on the GPU box's 16+ cores: ~5 s per config-2 volume), else the largest
    power-of-two crop of it that does.  Both CPU legs (cpu_baseline inside the default run, --impl reference) use this rule."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.batch or wl["batch"]
    if wl.get("kind") == "ldm":
        f = _oracle_step_fn(wl, 1, wl["spatial"])
        f()
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            f()
        dt = (time.perf_counter() - t0) / max(1, steps)
        return {"value": 1.0 / (dt * B), "unit": "steps/s", "cores": cores, "kind": "port", "extrapolated": True, "scale": float(B),
                "sample": "oracle port on 1 of the %d samples of a batch step; %.2f s per sample step, extrapolated x%d" % (B, dt, B),
                "sample_seconds_per_step": dt, "steps_run": max(1, steps), "warmup_run": 1}
    full = list(wl["spatial"])
    full_vox = full[0] * full[1] * full[2] * B
    # probe a crop to size the sample (4 stride-2 levels need multiples of 16)
    probe_sp = tuple(max(16, min(32, s)) for s in full)
    f = _oracle_step_fn(wl, 1, probe_sp)
    f()
    t0 = time.perf_counter()
    f()
    probe_s = time.perf_counter() - t0
    per_vox = probe_s / (probe_sp[0] * probe_sp[1] * probe_sp[2])
    target_vox = budget_s / max(1, steps + warmup) / per_vox
    sp = list(probe_sp)
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
    whole = list(sp) == full
    return {"value": 1.0 / (dt * scale), "unit": "steps/s", "cores": cores, "kind": "port", "extrapolated": True, "scale": scale,
            "sample": "oracle port (fp32 torch CPU restatement of the reference) on %s %dx%dx%d = 1/%.0f of a batch step; "
                      "%.2f s per sample step, extrapolated linearly" % ("ONE full volume" if whole else "a crop of one volume,", sp[0], sp[1], sp[2], scale, dt),
            "sample_seconds_per_step": dt, "steps_run": steps, "warmup_run": warmup}


def run_reference(args):
    """CPU reference arm: the oracle port (restated reference, fp32 torch on all host threads) on ONE full sample of the
    batch per step; value extrapolated linearly to the batch.  Same config / metric / unit as this repo's arm."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    K, W = max(1, args.steps), max(0, args.warmup)
    B = args.batch or wl["batch"]
    # bounded: the whole run (W + K sample steps) stays within BENCH_CPU_BUDGET_S (default 150 s) of CPU work
    cb = cpu_reference(args, wl, budget_s=float(os.environ.get("BENCH_CPU_BUDGET_S", 150.0)), steps=K, warmup=W)
    unit = ("steps/s (1 step = UNet forward + categorical posterior/draw for a batch of %d volumes)" % B) if wl.get("kind") != "ldm" \
        else ("steps/s (1 step = UNet eps forward + DDIM update for a batch of %d)" % B)
    line = {"impl": "reference", "metric": "denoising steps/sec", "value": cb["value"], "unit": unit,
            "n_gpus": args.gpus, "steps": cb["steps_run"], "warmup": cb["warmup_run"], "ms_per_step": cb["sample_seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.workload, B, 1),
            "extrapolated": True,
            "note": "CPU arm (rank 0 only): the reference is pure Python/PyTorch and /root/reference does not travel to the GPU box, so the "
                    "oracle port (pinned against the unmodified reference in tests/test_oracle_pinning.py) stands in for it.  Each timed step "
                    "is the oracle on a BOUNDED SAMPLE of a batch step (cpu_baseline.sample); ms_per_step is that measured sample time, "
                    "value = 1 / (sample time x cpu_baseline.scale) is the whole-batch figure, extrapolated linearly",
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ccdm_cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override batch per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / eager baseline / slab section")
    ap.add_argument("--detail", action="store_true", help="write per-launch times of one step to gpurun_out/bench_detail_<workload>.txt")
    args = ap.parse_args()
    # stdout carries ONE JSON line.  NCCL's INFO lines (the driver counts ranks in them) go to stderr instead of being
    # silenced: NCCL_DEBUG is left as the caller set it (INIT-level INFO by default at N > 1)
    if args.impl == "reference":
        run_reference(args)
        return
    json_fd = 1
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", 1)) > 1:
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        # NCCL writes its INFO lines to fd 1: point fd 1 at stderr for the whole process and keep a private duplicate of
        # the real stdout for the one JSON line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank, local, world, dev = dist_setup(args)
    state = {"line": None, "printed": False}
    lock = threading.Lock()

    def emit():
        with lock:
            if rank == 0 and state["line"] is not None and not state["printed"]:
                sys.stdout.flush()
                os.write(json_fd, (json.dumps(state["line"]) + "\n").encode())
                state["printed"] = True

    line = run_ours(args)
    state["line"] = line
    extras = args.workload == "ccdm_cfg2" and not args.no_extras and not args.batch
    if extras and world == 1:
        # ---- BASELINE configs 3, 4 (stage 2) and 5 (one large volume, here on one GPU): short runs, same contract keys
        line["workloads"] = {}
        for name, k in (("ldm_cfg3", 20), ("ldm_cfg4", 6), ("ccdm_cfg5", 3)):
            try:
                sub = run_ours(args, workload=name, K=k, W=3, sub=True)
                line["workloads"][name] = {kk: sub[kk] for kk in ("value", "unit", "ms_per_step", "steps", "e2e", "roofline", "kernel_ms",
                                                                   "gpu_launches", "config", "roofline_hbm_resident") if kk in sub}
            except Exception as e:  # noqa: BLE001
                line["workloads"][name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        try:
            line["workloads"]["pipeline_cfg4"] = run_pipeline_cfg4(args)
        except Exception as e:  # noqa: BLE001
            line["workloads"]["pipeline_cfg4"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(args.workload, dev)
            if "ms_per_step" in line["gpu_eager_baseline"]:
                line["gpu_eager_baseline"]["ours_ms_per_step"] = line["ms_per_step"]
                line["gpu_eager_baseline"]["speedup_vs_best_eager"] = line["gpu_eager_baseline"]["ms_per_step"] / line["ms_per_step"]
        except Exception as e:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": repr(e)[:300]}
    if world == 1 and rank == 0 and not args.no_cpu_baseline and line is not None:
        line["cpu_baseline"] = cpu_reference(args, WORKLOADS[args.workload], budget_s=20.0)
    if extras and world > 1:
        # a hang in the (never before seen at this N) slab section must not cost the weak-scaling line: after the
        # limit rank 0 prints what it has and every rank leaves
        limit = float(os.environ.get("BENCH_SLAB_LIMIT_S", 240.0))

        def bail():
            if state["line"] is not None:
                state["line"]["slab"] = {"error": "slab section exceeded %.0f s" % limit}
            emit()
            os._exit(0)
        timer = threading.Timer(limit, bail)
        timer.daemon = True
        timer.start()
        try:
            slab = slab_section(args)
        except Exception as e:  # noqa: BLE001
            slab = {"error": repr(e)[:300]}
        timer.cancel()
        if line is not None:
            line["slab"] = slab
    emit()
    if world > 1:
        try:
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
        sys.stdout.flush()
        os._exit(0)          # communicator teardown has dead-locked before (round 1): the line is out, leave


if __name__ == "__main__":
    main()
