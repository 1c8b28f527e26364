"""CPU-only tests: the C-ABI library loads and exports every symbol include/guidegen_sm100.h
declares (no compute calls), the drop-in modules keep the reference's constructor / state_dict
surface, host-side tables are bit-identical to the reference's, weight packing / upsample folding
are algebraically right, the product path fails loudly off-GPU, and the multi-rank host logic
works over gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden


def test_library_exports_every_declared_symbol():
    from jointimagegeneration_b200 import _C
    hdr = open(os.path.join(ROOT, "include", "guidegen_sm100.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _C.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_C.SYMBOLS), declared ^ set(_C.SYMBOLS)
    assert lib.gg_version() >= 100
    assert b"unsupported" in lib.gg_status_string(-2)
    # argument validation happens before any CUDA call: no GPU needed
    assert lib.gg_conv_pick_block_n(64) == 64 and lib.gg_conv_pick_block_n(12) == 16 and lib.gg_conv_pick_block_n(320) == 160
    assert lib.gg_gn_num_chunks(1 << 20, 192) >= 1
    import ctypes
    a = _C.ConvArgs()
    a.nsrc = 2
    a.src[0].C, a.src[1].C, a.src[1].centre_only = 192, 64, 1
    a.kd = a.kh = a.kw = 3
    assert lib.gg_conv_packed_k(ctypes.byref(a)) == 27 * 192 + 64
    assert lib.gg_cat_posterior_sample(None, None) == -1
    # the ctypes mirrors of the argument structs have the C structs' sizes (ABI drift shows up here, without a GPU)
    n = lib.gg_abi_sizes(None, 0)
    assert n == len(_C.ABI_STRUCTS)
    sizes = (ctypes.c_int32 * n)()
    lib.gg_abi_sizes(sizes, n)
    for cls, want in zip(_C.ABI_STRUCTS, list(sizes)):
        assert ctypes.sizeof(cls) == want, (cls.__name__, ctypes.sizeof(cls), want)


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """cuobjdump evidence that the conv kernel is tcgen05 + TMA (UTC*MMA / UTMALDG / LDTM)."""
    from jointimagegeneration_b200 import _C
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2gg19conv_tcgen05_kernelENS_10ConvParamsE", _C.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass, "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"


def test_product_path_fails_loudly_without_gpu_or_library(monkeypatch):
    from jointimagegeneration_b200 import _C, ops
    with pytest.raises(_C.GuideGenLibraryError):
        ops.cat_posterior_sample(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8), torch.zeros(1, 2), ops.CAT_POSTERIOR)
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libguidegen_sm100.so")
    with pytest.raises(_C.GuideGenLibraryError):
        _C.lib()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jointimagegeneration_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
                assert "/root/reference" not in src, f"{f} reads the reference tree"


def test_state_dict_surface_matches_reference_shapes():
    from jointimagegeneration_b200.ccdm import build_model
    from jointimagegeneration_b200.ldm import UNetModel
    from oracle import configs, weights
    m = build_model(10, "cosine", {"s": 0.008}, [(1, 32, 32, 32), (12, 32, 32, 32)], None, "unet_openai",
                    dict(configs.CCDM_PARAMS_YML), "x", "majority", dims=3)
    mine = {k[5:]: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith("unet.")}
    assert mine == weights.reference_shapes("CCDM_PARAMS_YML")
    assert {"diffusion.betas", "diffusion.alphas", "diffusion.cumalphas"} <= set(m.state_dict())
    for name in ("LDM_AE", "LDM_PIXEL", "LDM_TINY", "LDM_TINY_XATTN"):
        u = UNetModel(**getattr(configs, name))
        assert {k: tuple(v.shape) for k, v in u.state_dict().items()} == weights.reference_shapes(name), name
    # zero-initialised modules as in the reference (unet.py:216-218,300,719)
    sd = m.unet.state_dict()
    for k in ("out.2.weight", "input_blocks.1.0.out_layers.3.weight", "middle_block.1.proj_out.weight"):
        assert float(sd[k].abs().max()) == 0.0


def test_ccdm_schedules_and_ddim_tables_bit_identical_to_reference():
    from jointimagegeneration_b200.ccdm import DiffusionModel
    from jointimagegeneration_b200.ldm import DDIMSampler, LatentDiffusion, UNetModel
    from oracle import configs
    g = golden("posterior_cases")
    dm = DiffusionModel("cosine", 1000, 12, schedule_params={"s": 0.008}, dims=3)
    assert np.array_equal(dm.betas.numpy(), g["cos1000_betas"]) and np.array_equal(dm.cumalphas.numpy(), g["cos1000_cumalphas"])
    dl = DiffusionModel("linear", 50, 12, schedule_params=None, dims=3)
    assert np.array_equal(dl.alphas.numpy(), g["lin50_alphas"]) and np.array_equal(dl.cumalphas.numpy(), g["lin50_cumalphas"])
    # step coefficients: t == 1 -> (0, 1)
    co = dm.step_coef_tensor(torch.tensor([1, 2, 1000]))
    assert co[0].tolist() == [0.0, 1.0]
    assert float(co[1, 0]) == float(dm.alphas[1]) and float(co[1, 1]) == float(dm.cumalphas[0])
    assert float(co[2, 0]) == float(dm.alphas[999]) and float(co[2, 1]) == float(dm.cumalphas[998])
    gt = golden("ddim_tables")
    ld = LatentDiffusion(UNetModel(**configs.LDM_TINY), **configs.LDM_SCHEDULE)
    assert np.array_equal(ld.alphas_cumprod.numpy(), gt["alphas_cumprod"]) and np.array_equal(ld.betas.numpy(), gt["betas"])
    for S, eta in ((50, 0.0), (50, 1.0), (20, 0.5), (250, 0.0)):
        s = DDIMSampler(ld)
        s.make_schedule(S, ddim_eta=eta, verbose=False)
        k = f"S{S}_eta{eta}"
        assert np.array_equal(s.ddim_timesteps, gt[k + "_timesteps"])
        tab = s._coef.numpy()
        for j, name in enumerate(("alphas", "alphas_prev", "sigmas", "sqrt_one_minus_alphas")):
            assert np.array_equal(tab[:, j], gt[k + "_" + name].astype(np.float32)), (k, name)


def test_weight_packing_and_upsample_fold():
    from jointimagegeneration_b200 import ops
    from jointimagegeneration_b200.unet_engine import _fold_upsample_weight
    rs = np.random.RandomState(0)
    w = torch.from_numpy(rs.standard_normal((5, 70 + 24, 3, 3, 3)).astype(np.float32))
    e = torch.from_numpy(rs.standard_normal((5, 24)).astype(np.float32))
    p = ops.pack_conv_weight(w, [70, 24], extra=[e]).float()
    assert p.shape == (5, 27 * 128 + 27 * 64 + 64)
    wq = w.to(torch.bfloat16).float()
    # source 0, tap (1, 2, 0) = index 15, channel 69 -> column 15 * 128 + 69; channel 70.. padded with zeros
    assert torch.equal(p[:, 15 * 128 + 69], wq[:, 69, 1, 2, 0]) and float(p[:, 15 * 128 + 70:16 * 128].abs().max()) == 0
    base = 27 * 128
    assert torch.equal(p[:, base + 3 * 64 + 5], wq[:, 70 + 5, 0, 1, 0])
    assert torch.equal(p[:, base + 27 * 64 + 23], e.to(torch.bfloat16).float()[:, 23])
    # nearest-x2 upsample + 3^d conv == per-parity 2^d conv over the coarse grid with folded weights
    for dims in (2, 3):
        x = torch.from_numpy(rs.standard_normal((1, 3) + (4,) * dims).astype(np.float32))
        w = torch.from_numpy(rs.standard_normal((2, 3) + (3,) * dims).astype(np.float32))
        conv = torch.nn.functional.conv3d if dims == 3 else torch.nn.functional.conv2d
        want = conv(torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
        got = torch.zeros_like(want)
        for code in range(2 ** dims):
            par = tuple((code >> (dims - 1 - i)) & 1 for i in range(dims))
            wf = _fold_upsample_weight(w, dims, (0,) * (3 - dims) + par)
            # taps {pi - 1, pi}: pad the coarse grid by one on each side, then crop
            xp = torch.nn.functional.pad(x, (1, 1) * dims)
            y = conv(xp, wf)
            sl = tuple(slice(p_, p_ + 4) for p_ in par)
            idx = (slice(None), slice(None)) + tuple(slice(p_, None, 2) for p_ in par)
            got[idx] = y[(slice(None), slice(None)) + sl]
        assert float((got - want).abs().max()) <= 1e-5


def test_one_hot_categorical_and_loop_surface():
    from jointimagegeneration_b200.ccdm import DenoisingModel, OneHotCategoricalBCHW, build_model
    from oracle import configs
    with pytest.raises(ValueError):
        OneHotCategoricalBCHW(probs=torch.ones(3))
    with pytest.raises(ValueError):
        OneHotCategoricalBCHW()
    m = build_model(1000, "cosine", {"s": 0.008}, [(1, 8, 8, 8), (4, 8, 8, 8)], None, "unet_openai", dict(configs.CCDM_TINY), "x",
                    "majority", dims=3)
    assert isinstance(m, DenoisingModel) and m.time_steps == 1000
    assert m._t_values(None)[:3] == [1000, 999, 998] and m._t_values(None)[-1] == 1
    assert m._t_values(7) == [7, 6, 5, 4, 3, 2, 1]
    tv = m._t_values(10000 + 5)      # step skipping (diffusion_denoising.py:190-197)
    assert tv == [round(v) for v in np.linspace(1000, 1, 5)]
    m.eval()
    with pytest.raises(RuntimeError):   # parameters on CPU: no CPU path
        m(torch.zeros(1, 4, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(NotImplementedError):
        m.forward_denoising(torch.zeros(1, 4, 8, 8, 8), None, None, label_ref_logits=torch.zeros(1))


def test_sharding_host_logic():
    from jointimagegeneration_b200 import sharding
    for n, w in ((8, 8), (8, 3), (5, 8), (16, 4), (1, 2)):
        rs = [sharding.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert sharding.slab_ranges(128, 8) == [(16 * i, 16 * (i + 1)) for i in range(8)]
    assert sharding.slab_ranges(128, 2) == [(0, 64), (64, 128)]
    with pytest.raises(ValueError):          # unequal slabs would break the GroupNorm combine / voxel indexing: refused
        sharding.slab_ranges(48, 2)
    assert len(set(sharding.chain_seeds(3, 16))) == 16


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from jointimagegeneration_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
full = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 3)
mine = sharding.shard_batch(full, r, 2)
assert mine.shape[0] == (3 if r == 0 else 2)
back = sharding.gather_batch(mine * 2, 5)
assert torch.equal(back, full * 2)
mn, mx = sharding.global_minmax(mine)
assert (mn, mx) == (0.0, 14.0)
# depth-slab communication (config 5): halo planes from the neighbours, zeros at the ends of the volume
comm = sharding.SlabComm()
lead, D = 2, 3
t = torch.full((1, lead + D + 1, 2, 2, 4), -7.0)
t[0, lead:lead + D] = torch.arange(D, dtype=torch.float32).reshape(D, 1, 1, 1) + 10 * (r + 1)
comm.exchange_halo(t, lead, D)
lo, hi = t[0, lead - 1], t[0, lead + D]
if r == 0:
    assert torch.all(lo == 0) and torch.all(hi == 20.0)          # rank 1's first interior plane
else:
    assert torch.all(lo == 12.0) and torch.all(hi == 0)          # rank 0's last interior plane
assert torch.all(t[0, 0] == -7.0)                                # spare plane untouched
t2 = torch.full_like(t, -7.0)
t2[0, lead:lead + D] = t[0, lead:lead + D]
comm.exchange_halo(t2, lead, D, need_lo=True, need_hi=False)     # what a stride-2 conv asks for
assert torch.all(t2[0, lead + D] == -7.0) and torch.all(t2[0, lead - 1] == (0.0 if r == 0 else 12.0))
part = torch.full((1, 2, 4, 2), float(r + 1))
g = torch.empty((1, 4, 4, 2))
comm.all_gather(g, part)
assert torch.all(g[0, :2] == 1.0) and torch.all(g[0, 2:] == 2.0)
assert comm.n_exchanges == 2 and comm.n_gathers == 1
dist.barrier()
dist.destroy_process_group()
print("ok", r)
'''


def test_sample_sharding_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_next_row_modules_keep_the_reference_parameter_names():
    """state_dict contract of the drop-ins added for SURVEY 8f (N1, N3): parameter names / shapes equal the reference
    modules' (oracle.weights.*_shapes are checked against the live reference in tests/test_oracle_pinning.py); no
    forward is run here, and a CPU forward must refuse loudly (no fallback)."""
    import pytest
    import torch
    from jointimagegeneration_b200.ccdm import PreloadedBERTEncoder
    from jointimagegeneration_b200.ldm import PLMSSampler  # noqa: F401  (importable without a GPU)
    from jointimagegeneration_b200.ldm.autoencoder import AutoencoderKL
    from oracle import weights
    dd = dict(ch=128, out_ch=1, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0, in_channels=1,
              resolution=512, z_channels=4, double_z=True, dims=2)       # ruijin-ldm_from_controlnet_ae.yaml:48-63
    ae = AutoencoderKL(dd, 4)
    want = dict(weights.vae_encoder_shapes(128))
    want.update(weights.vae_decoder_shapes(128))
    assert weights.shapes_of(ae) == want
    with pytest.raises(RuntimeError):
        ae.decode(torch.zeros(1, 4, 8, 8))
    with pytest.raises(RuntimeError):
        ae.encode(torch.zeros(1, 1, 64, 64))
    enc = PreloadedBERTEncoder()                                          # encoder.py:104-105 defaults: 768 / 8 heads / depth 4 / 64
    assert weights.shapes_of(enc) == weights.encoder_shapes(768, 8, 64, 4)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 768, 16))


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU oracle port, rank 0 only): one JSON line with the contract's keys; other ranks
    print nothing.  BENCH_CPU_BUDGET_S bounds the sample so that the test takes seconds."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, BENCH_CPU_BUDGET_S="1", RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoising steps/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["name"] == "ccdm_cfg2"
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=600, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_round2_host_planning_logic(monkeypatch):
    """Host-side decisions added in round 2, none of which needs a GPU: the cluster plan of the one-launch GroupNorm
    (gg_gn_fused_resident), the sample lanes of a plan (Plan.lanes / UNetEngine.pick_lanes) and the SASS of the remote
    mbarrier arrive (no GPU-scope fence in front of the accumulator hand-backs of the pair kernels)."""
    from jointimagegeneration_b200 import _C
    from jointimagegeneration_b200.unet_engine import Plan, UNetEngine
    lib = _C.lib()
    # ---- gg_gn_fused_resident: smallest power-of-two cluster with >= one CTA per 48 KB that holds the sample; 0 = does not fit
    assert lib.gg_gn_fused_resident(16, 800) == 1              # 25 KB
    assert lib.gg_gn_fused_resident(64, 640) == 2              # 80 KB
    assert lib.gg_gn_fused_resident(64, 1440) == 4             # 180 KB
    assert lib.gg_gn_fused_resident(4096, 160) == 8            # 1.3 MB: 164 KB per CTA
    assert lib.gg_gn_fused_resident(4096, 320) == 0            # 2.6 MB: more than eight CTAs hold
    assert lib.gg_gn_fused_resident(256, 12) == 0              # channels not a multiple of 8
    # ---- Plan.lanes bookkeeping
    p = Plan()
    f = lambda *a: 0
    p.add(f, 1), p.add(f, 2)
    assert len(p.lanes) == 1 and p.num_launches == 2 and p.steps is p.lanes[0]
    p.begin_lane(1)
    p.add(f, 3), p.add(f, 4), p.add(f, 5)
    assert len(p.lanes) == 2 and p.num_launches == 5
    assert [a for _, a in p.steps] == [(1,), (2,), (3,), (4,), (5,)]
    assert [a for _, a in p.body_steps] == [(1,), (3,), (4,)]          # everything but each lane's head conv
    # ---- pick_lanes: pinned by GG_LANES (reduced until it divides the batch), one lane in depth-slab mode, two for mid-sized 2-D batches
    import torch
    eng = UNetEngine(torch.nn.Linear(1, 1), dims=2, num_heads=1, num_head_channels=-1)
    monkeypatch.delenv("GG_LANES", raising=False)
    assert eng.pick_lanes(16, (64, 64)) == 1 and eng.pick_lanes(2, (512, 512)) == 2 and eng.pick_lanes(8, (512, 512)) == 1
    eng3 = UNetEngine(torch.nn.Linear(1, 1), dims=3, num_heads=1, num_head_channels=-1)
    assert eng3.pick_lanes(8, (64, 128, 128)) == 1 and eng3.pick_lanes(2, (32, 64, 64)) == 1
    monkeypatch.setenv("GG_LANES", "4")
    assert eng.pick_lanes(16, (64, 64)) == 4 and eng.pick_lanes(6, (64, 64)) == 3 and eng.pick_lanes(1, (64, 64)) == 1
    eng.slab = object()
    assert eng.pick_lanes(16, (64, 64)) == 1
    # ---- SASS: the pair kernels keep exactly the two cluster-wide barriers of their prologue / exit (MEMBAR.ALL.GPU twice),
    # none in front of the per-tile / per-plane remote arrives
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        for fn in ("_ZN2gg16conv_halo_kernelILi3ELb1ELb1ELi0EEEvNS_10HaloParamsE", "_ZN2gg16conv_roll_kernelILi3ELb1ELi64ELi4EEEvNS_10RollParamsE"):
            sass = subprocess.run([cuobjdump, "-sass", "-fun", fn, _C.LIB_PATH], capture_output=True, text=True).stdout
            assert "UTCHMMA.2CTA" in sass, fn
            assert sass.count("MEMBAR.ALL.GPU") == 2, (fn, sass.count("MEMBAR.ALL.GPU"))
        sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN2gg16conv_halo_kernelILi2ELb1ELb1ELi0EEEvNS_10HaloParamsE", _C.LIB_PATH],
                              capture_output=True, text=True).stdout
        assert "STG.E.ENL2.256" in sass, "no 256-bit epilogue stores in the halo-brick conv"
