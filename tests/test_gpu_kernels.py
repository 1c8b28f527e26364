"""GPU parity tests of the individual sm_100a kernels, called through the C ABI.

Integer / label outputs: bit-exact against the CPU oracle (oracle/).  Float kernels: against
a plain torch fp32 reference of the same op on bf16-rounded operands; tolerances are stated
at each assert (bf16 output rounding = 2^-9 relative; fp32 accumulation order differs)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import golden  # noqa: E402
from gpu_util import attention_ref, cl_from_nchw, conv_ref, heads_of, nchw_from_cl, no_tf32, rel_err  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from jointimagegeneration_b200 import _C, ops as o
    assert torch.cuda.is_available()
    _C.check(_C.lib().gg_device_check(), "gg_device_check")
    return o


def dev(x):
    return torch.as_tensor(x).cuda()


# ------------------------------------------------------------------------------ posterior
def _posterior_case(ops, B, C, spatial, t, seed, T=1000):
    from oracle import diffusion, weights
    _, alphas, cumalphas = diffusion.cosine_schedule(T)
    V = int(np.prod(spatial))
    xt = weights.uniform_one_hot(seed, B, C, spatial)
    rs = np.random.RandomState(seed + 1)
    x0 = torch.softmax(torch.from_numpy(rs.standard_normal((B, C) + tuple(spatial)).astype(np.float32)) * 3, 1)
    q = weights.exp_noise(seed + 2, (B * V, C))
    a, g = diffusion.step_coefficients(alphas, cumalphas, t)
    coef = torch.tensor([[a, g]] * B, dtype=torch.float32)
    return xt, x0, q, coef, a, g, V


@pytest.mark.parametrize("B,C,spatial,t", [(2, 12, (4, 8, 8), 1000), (2, 12, (4, 8, 8), 500), (2, 12, (4, 8, 8), 2),
                                           (2, 12, (4, 8, 8), 1), (1, 12, (32, 32, 32), 777), (3, 4, (5, 7, 3), 17),
                                           (2, 19, (6, 10), 300), (1, 2, (3, 3, 3), 9), (2, 12, (3, 5, 7), 640)])
def test_posterior_and_sampling_bit_exact(ops, B, C, spatial, t):
    """posterior probs, sampled labels, argmax labels and normalised probs: bit-exact vs oracle
    (covers ragged V % 4 != 0 shapes -> scalar path, and C from 2 to 19)."""
    from oracle import diffusion
    xt, x0, q, coef, a, g, V = _posterior_case(ops, B, C, spatial, t, seed=100 + t)
    xt_n, x0_n = xt.reshape(B, C, V).numpy(), x0.reshape(B, C, V).numpy()
    want = np.stack([diffusion.theta_post_prob_closed(a, g, xt_n[b], x0_n[b]) for b in range(B)])
    got, _, _ = ops.cat_posterior_sample(dev(x0), dev(xt), dev(coef), ops.CAT_POSTERIOR)
    assert np.array_equal(got.reshape(B, C, V).cpu().numpy(), want), "posterior not bit-exact"
    # sample
    clamped = np.maximum(want, np.float32(1e-12))
    qb = q.reshape(B, V, C)
    idx = np.stack([diffusion.categorical_sample(clamped[b], qb[b]) for b in range(B)])
    labels = torch.empty((B, V), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(dev(x0))
    ops.cat_posterior_sample(dev(x0), dev(xt), dev(coef), ops.CAT_SAMPLE, q=dev(q), out=out, labels=labels)
    assert np.array_equal(labels.cpu().numpy(), idx.astype(np.uint8)), "sampled labels differ"
    oh = np.stack([diffusion.one_hot(idx[b], C) for b in range(B)])
    assert np.array_equal(out.reshape(B, C, V).cpu().numpy(), oh)
    # argmax (last step, 'majority') with int64 one-hot like F.one_hot
    S = clamped[:, 0].copy()
    for c in range(1, C):
        S = (S + clamped[:, c]).astype(np.float32)
    pn = (clamped / S[:, None]).astype(np.float32)
    o64 = torch.empty(x0.shape, dtype=torch.int64, device="cuda")
    ops.cat_posterior_sample(dev(x0), dev(xt), dev(coef), ops.CAT_ARGMAX, out=None, out_i64=o64, labels=labels)
    assert np.array_equal(labels.cpu().numpy(), pn.argmax(1).astype(np.uint8))
    assert np.array_equal(o64.reshape(B, C, V).argmax(1).cpu().numpy(), pn.argmax(1))
    assert int(o64.sum()) == B * V
    # normalised probabilities ('confidence')
    pr, _, _ = ops.cat_posterior_sample(dev(x0), dev(xt), dev(coef), ops.CAT_PROBS)
    assert np.array_equal(pr.reshape(B, C, V).cpu().numpy(), pn)


def test_posterior_matches_reference_golden(ops):
    """Against vectors produced by the UNMODIFIED reference's theta_post_prob (tests/golden)."""
    from oracle import diffusion, weights
    g = golden("posterior_cases")
    C, B, spatial = int(g["C"]), int(g["B"]), tuple(g["spatial"])
    _, alphas, cumalphas = diffusion.cosine_schedule(1000)
    for t in (1, 2, 17, 500, 999, 1000):
        xt = weights.uniform_one_hot(100 + t, B, C, spatial)
        a, gg = diffusion.step_coefficients(alphas, cumalphas, t)
        coef = torch.tensor([[a, gg]] * B, dtype=torch.float32)
        got, _, _ = ops.cat_posterior_sample(dev(g[f"x0_{t}"]), dev(xt), dev(coef), ops.CAT_POSTERIOR)
        # closed form vs the reference's O(C^2) einsum: <= 1e-6 abs (different fp32 evaluation order)
        assert np.abs(got.cpu().numpy() - g[f"post_{t}"]).max() <= 1e-6
    coef = torch.tensor([list(diffusion.step_coefficients(alphas, cumalphas, 300))] * B, dtype=torch.float32)
    got, _, _ = ops.cat_posterior_sample(dev(g["soft_x0"]), dev(g["soft_xt"]), dev(coef), ops.CAT_POSTERIOR)
    assert np.abs(got.cpu().numpy() - g["soft_post_300"]).max() <= 1e-6


def test_sample_given_and_philox(ops):
    from oracle import diffusion, weights
    B, C, spatial = 2, 12, (4, 4, 8)
    V = int(np.prod(spatial))
    p = torch.softmax(torch.from_numpy(np.random.RandomState(5).standard_normal((B, C) + spatial).astype(np.float32)), 1)
    q = weights.exp_noise(6, (B * V, C))
    labels = torch.empty((B, V), dtype=torch.uint8, device="cuda")
    ops.cat_posterior_sample(dev(p), None, None, ops.CAT_SAMPLE_GIVEN, q=dev(q), clamp_min=0.0, labels=labels)
    want = np.stack([diffusion.categorical_sample(p[b].reshape(C, V).numpy(), q.reshape(B, V, C)[b]) for b in range(B)])
    assert np.array_equal(labels.cpu().numpy(), want.astype(np.uint8))
    # in-kernel Philox: deterministic per (seed, offset), different across offsets, uniform-ish over classes
    big = torch.full((1, C, 64, 64, 16), 1.0 / C, device="cuda")
    l1 = torch.empty((1, 65536), dtype=torch.uint8, device="cuda")
    l2 = torch.empty_like(l1)
    l3 = torch.empty_like(l1)
    ops.cat_posterior_sample(big, None, None, ops.CAT_SAMPLE_GIVEN, clamp_min=0.0, labels=l1, seed=7, offset=1)
    ops.cat_posterior_sample(big, None, None, ops.CAT_SAMPLE_GIVEN, clamp_min=0.0, labels=l2, seed=7, offset=1)
    ops.cat_posterior_sample(big, None, None, ops.CAT_SAMPLE_GIVEN, clamp_min=0.0, labels=l3, seed=7, offset=2)
    assert torch.equal(l1, l2) and not torch.equal(l1, l3)
    hist = torch.bincount(l1.flatten().long(), minlength=C).float() / 65536
    assert float((hist - 1.0 / C).abs().max()) < 0.01


def test_cat_step_cl_matches_oracle_outside_near_ties(ops):
    from oracle import diffusion, weights
    B, C, spatial, t = 2, 12, (4, 8, 8), 400
    V = int(np.prod(spatial))
    _, alphas, cumalphas = diffusion.cosine_schedule(1000)
    a, g = diffusion.step_coefficients(alphas, cumalphas, t)
    rs = np.random.RandomState(3)
    logits = rs.standard_normal((B * V, 16)).astype(np.float32) * 2
    lab = rs.randint(0, C, size=(B * V,)).astype(np.uint8)
    q = weights.exp_noise(4, (B * V, C))
    x0 = torch.softmax(torch.from_numpy(logits[:, :C]), -1).numpy().reshape(B, V, C).transpose(0, 2, 1)
    xt = np.eye(C, dtype=np.float32)[lab].reshape(B, V, C).transpose(0, 2, 1)
    probs = np.stack([diffusion.theta_post_prob_closed(a, g, xt[b], x0[b]) for b in range(B)])
    probs = np.maximum(probs, np.float32(1e-12))
    want = np.stack([diffusion.categorical_sample(probs[b], q.reshape(B, V, C)[b]) for b in range(B)]).reshape(-1)
    tie = np.stack([diffusion.near_tie_mask(probs[b], q.reshape(B, V, C)[b], rel=1e-4) for b in range(B)]).reshape(-1)
    coef = torch.tensor([[a, g]] * B, dtype=torch.float32)
    lout = torch.empty(B * V, dtype=torch.uint8, device="cuda")
    nx = torch.empty((B * V, 16), dtype=torch.bfloat16, device="cuda")
    pout = torch.empty((B, C, V), dtype=torch.float32, device="cuda")
    ops.cat_step_cl(dev(logits), dev(lab), dev(coef), lout, B, V, C, q=dev(q), next_x=nx, probs_out=pout)
    got = lout.cpu().numpy()
    assert np.array_equal(got[~tie], want[~tie].astype(np.uint8))
    assert tie.mean() < 0.01
    # fused softmax + posterior (shuffle-order sums, expf): 1e-5 relative to the largest prob
    assert np.abs(pout.cpu().numpy() - probs).max() <= 1e-5
    nxr = nx.float().cpu().numpy()
    assert np.array_equal(nxr[:, :C].argmax(1), got) and np.all(nxr[:, C:] == 0) and np.all(nxr.sum(1) == 1)


# ------------------------------------------------------------------------------------ DDIM
@pytest.mark.parametrize("shape,eta", [((2, 4, 16, 16), 0.0), ((2, 4, 16, 16), 0.5), ((3, 1, 5, 7), 1.0), ((16, 4, 64, 64), 0.0)])
def test_ddim_update_bit_exact(ops, shape, eta):
    from oracle import configs, ddim
    betas = ddim.make_beta_schedule_linear(1000, configs.LDM_SCHEDULE["linear_start"], configs.LDM_SCHEDULE["linear_end"])
    acp = ddim.alphas_cumprod_f32(betas)
    ts = ddim.make_ddim_timesteps(50, 1000)
    tab = ddim.ddim_tables(acp, ts, eta)
    rs = np.random.RandomState(11)
    x = rs.standard_normal(shape).astype(np.float32)
    e = rs.standard_normal(shape).astype(np.float32)
    nz = rs.standard_normal(shape).astype(np.float32)
    for index in (49, 20, 0):
        co = [tab["alphas"][index], tab["alphas_prev"][index], tab["sigmas"][index], tab["sqrt_one_minus_alphas"][index]]
        want_prev, want_x0 = ddim.ddim_update(x, e, *co, nz, 1.0)
        coef = torch.tensor(co, dtype=torch.float64).float().cuda()
        got_prev, got_x0 = ops.ddim_update(dev(x), dev(e), coef, dev(nz))
        assert np.array_equal(got_prev.cpu().numpy(), want_prev) and np.array_equal(got_x0.cpu().numpy(), want_x0)
        if eta == 0.0:  # noise pointer omitted == reference value (sigma = 0)
            got2, _ = ops.ddim_update(dev(x), dev(e), coef, None)
            assert np.array_equal(got2.cpu().numpy(), want_prev)


def test_ddpm_ancestral_update(ops):
    """x0 and the mean are bit-exact; the exp(0.5 logvar) factor differs from numpy's by <= 1 ulp."""
    from oracle import configs, ddim
    betas = ddim.make_beta_schedule_linear(1000, configs.LDM_SCHEDULE["linear_start"], configs.LDM_SCHEDULE["linear_end"])
    tab = ddim.ddpm_tables(betas)
    rs = np.random.RandomState(12)
    shape = (3, 4, 8, 8)
    x, e, nz = (rs.standard_normal(shape).astype(np.float32) for _ in range(3))
    for t in (999, 500, 1, 0):
        coef = torch.tensor([[tab["sqrt_recip"][t], tab["sqrt_recipm1"][t], tab["coef1"][t], tab["coef2"][t], tab["logvar"][t],
                              0.0 if t == 0 else 1.0]] * 3, dtype=torch.float32).cuda()
        for clip in (False, True):
            want, want0 = ddim.ddpm_update(x, e, tab, t, nz, 1.0, clip)
            got, got0 = ops.ddpm_update(dev(x), dev(e), coef, dev(nz), 1.0, clip, want_x0=True)
            assert np.array_equal(got0.cpu().numpy(), want0)
            assert np.abs(got.cpu().numpy() - want).max() <= 1e-6 * max(1.0, np.abs(want).max())


# --------------------------------------------------------------------------- layout bridges
def test_layout_bridges(ops):
    rs = np.random.RandomState(0)
    x1 = torch.from_numpy(rs.standard_normal((2, 12, 3, 5, 7)).astype(np.float32))
    x2 = torch.from_numpy(rs.standard_normal((2, 1, 3, 5, 7)).astype(np.float32))
    y = ops.nchw_to_cl(dev(x1), dev(x2), c_pad=16)
    assert y.shape == (2, 3, 5, 7, 16)
    want = torch.cat([x1, x2, torch.zeros(2, 3, 3, 5, 7)], 1).to(torch.bfloat16).permute(0, 2, 3, 4, 1)
    assert torch.equal(y.cpu(), want)
    back = ops.cl_to_nchw(y, 13, (3, 5, 7))
    assert torch.equal(back.cpu(), torch.cat([x1, x2], 1).to(torch.bfloat16).float())
    lg = torch.from_numpy(rs.standard_normal((2, 3, 5, 7, 16)).astype(np.float32))
    sm = ops.cl_to_nchw(dev(lg), 12, (3, 5, 7), softmax=True)
    want = torch.softmax(lg[..., :12], -1).permute(0, 4, 1, 2, 3)
    assert float((sm.cpu() - want).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------- GroupNorm
@pytest.mark.parametrize("N,sp,C1,C2,silu,eps", [(2, (4, 8, 8), 64, 0, True, 1e-5), (2, (1, 16, 16), 160, 0, False, 1e-6),
                                                  (1, (8, 16, 16), 128, 64, True, 1e-5), (2, (1, 1, 37), 320, 0, True, 1e-5),
                                                  (1, (16, 32, 32), 32, 0, True, 1e-5), (2, (2, 4, 4), 640, 320, True, 1e-5)])
def test_group_norm_silu(ops, N, sp, C1, C2, silu, eps):
    no_tf32()
    rs = np.random.RandomState(1)
    C = C1 + C2
    x = torch.from_numpy(rs.standard_normal((N, C) + sp).astype(np.float32)) * 2 + 0.5
    gamma = torch.from_numpy(rs.standard_normal(C).astype(np.float32))
    beta = torch.from_numpy(rs.standard_normal(C).astype(np.float32))
    xc = cl_from_nchw(x.cuda())
    x1 = xc[..., :C1].contiguous()
    x2 = xc[..., C1:].contiguous() if C2 else None
    y = ops.group_norm_cl(x1, x2, gamma.cuda(), beta.cuda(), eps=eps, silu=silu)
    xr = nchw_from_cl(xc)
    want = torch.nn.functional.group_norm(xr, 32, gamma.cuda(), beta.cuda(), eps)
    if silu:
        want = torch.nn.functional.silu(want)
    # bf16 output rounding (2^-9 relative) dominates
    assert rel_err(nchw_from_cl(y), want) <= 6e-3


# ------------------------------------------------------------------------------ convolution
def _conv_case(ops, N, sp, Cs, Cout, dims, k=3, stride=1, bias=True, emb=False, residual=False, extra_C=0, f32_out=False,
               seed=0, block_n=0, brick=None, algo=0, split_k=0, stats=False):
    no_tf32()
    rs = np.random.RandomState(seed)
    sp3 = (1,) * (3 - len(sp)) + tuple(sp)
    xs = [torch.from_numpy(rs.standard_normal((N,) + sp3 + (c,)).astype(np.float32)).cuda().to(torch.bfloat16) for c in Cs]
    Cin = sum(Cs)
    w = torch.from_numpy((rs.standard_normal((Cout, Cin) + (k,) * dims) / math.sqrt(Cin * k ** dims)).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda() if bias else None
    extra, extra_w, srcs = [], [], [(x, False) for x in xs]
    if extra_C:
        xe = torch.from_numpy(rs.standard_normal((N,) + sp3 + (extra_C,)).astype(np.float32)).cuda().to(torch.bfloat16)
        we = torch.from_numpy((rs.standard_normal((Cout, extra_C)) / math.sqrt(extra_C)).astype(np.float32)).cuda()
        extra.append((xe, we))
        extra_w.append(we)
        srcs.append((xe, True))
    wp = ops.pack_conv_weight(w, Cs, extra=extra_w, chunk_major=(algo >= 1))
    Cout8 = (Cout + 7) // 8 * 8
    if stride == 1:
        osp = sp3
    else:
        f = lambda n, on: (n - 1) // 2 + 1 if on else n
        osp = (f(sp3[0], dims >= 3), f(sp3[1], dims >= 2), f(sp3[2], True))
    e = torch.from_numpy(rs.standard_normal((N, Cout8)).astype(np.float32)).cuda() if emb else None
    r = torch.from_numpy(rs.standard_normal((N,) + osp + (Cout8,)).astype(np.float32)).cuda().to(torch.bfloat16) if residual else None
    y = torch.full((N,) + osp + (Cout8,), float("nan"), dtype=torch.float32 if f32_out else torch.bfloat16, device="cuda")
    ws = torch.full((max(split_k, 1), N * osp[0] * osp[1] * osp[2], Cout8), float("nan"), device="cuda") if split_k > 1 else None
    b_pad = ops.pad_vec(b, Cout)       # the args struct holds raw pointers: keep the padded copy alive
    a = ops.make_conv_args(srcs, wp, Cout, y, dims=dims, ksize=k, stride=stride, bias=b_pad, emb=e,
                           residual=r, block_n=block_n, brick=brick, algo=algo, split_k=split_k, workspace=ws)
    assert ops.conv_packed_k(a) == wp.shape[1]
    part = None
    if stats:       # GroupNorm column sums from the conv epilogue; poisoned first: the launch must define every row
        import ctypes as C
        from jointimagegeneration_b200 import _C
        per = int(_C.lib().gg_conv_stats_chunks(C.byref(a)))
        assert per > 0
        part = torch.full((N, per + 3, Cout8, 2), float("nan"), device="cuda")
        a.gn_partial, a.gn_chunk_base, a.gn_nchunks_total = _C.ptr(part), 2, per + 3
    ops.conv_fwd(a)
    torch.cuda.synchronize()
    if split_k > 1:
        # the in-kernel fix-up (gg_conv_args.split_counters: the last split of a tile to finish sums the partials in split order
        # and writes the output in the same launch) must give exactly the two-launch result, launch after launch
        import ctypes as C
        from jointimagegeneration_b200 import _C
        cnt = torch.zeros((int(_C.lib().gg_conv_num_tiles(C.byref(a))),), dtype=torch.int32, device="cuda")
        y2 = torch.full_like(y, float("nan"))
        a.y, a.split_counters = y2.data_ptr(), cnt.data_ptr()
        n0 = _C.launch_count()
        for _ in range(3):
            y2.fill_(float("nan"))
            ops.conv_fwd(a)
            torch.cuda.synchronize()
            assert torch.equal(y2.view(torch.uint8), y.view(torch.uint8)), "split-K fix-up differs from the reduce launch"
            assert int(cnt.abs().sum()) == 0, "tile counters must wrap back to zero"
        assert _C.launch_count() - n0 == 3, "one launch per convolution"
        a.y, a.split_counters = y.data_ptr(), None
    if stats:
        assert torch.isnan(part[:, :2]).all() and torch.isnan(part[:, -1:]).all()      # stays inside its chunk range
        got_s = part[:, 2:-1].double().sum(1)
        yd = y.double().reshape(N, -1, Cout8)
        want_s = torch.stack([yd.sum(1), (yd * yd).sum(1)], -1)
        assert torch.allclose(got_s, want_s, rtol=1e-4, atol=1e-2), float((got_s - want_s).abs().max())
    want = conv_ref(xs, w, b, dims, stride, e, r, extra)
    got = nchw_from_cl(y, Cout)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert not torch.isnan(got).any(), f"{int(torch.isnan(got).sum())} outputs never written"
    return rel_err(got, want), got, want


CONV_CASES = {
    "linear_64": dict(N=2, sp=(256,), Cs=[64], Cout=64, dims=1, k=1),
    "linear_128_160": dict(N=2, sp=(100,), Cs=[128], Cout=160, dims=1, k=1),
    "linear_ragged": dict(N=3, sp=(37,), Cs=[320], Cout=960, dims=1, k=1, bias=False),
    "conv2d_64": dict(N=2, sp=(16, 16), Cs=[64], Cout=64, dims=2),
    "conv3d_64_128_all": dict(N=1, sp=(8, 8, 8), Cs=[64], Cout=128, dims=3, emb=True, residual=True),
    "conv3d_c16": dict(N=1, sp=(8, 8, 8), Cs=[16], Cout=64, dims=3),
    "conv3d_concat_skip": dict(N=2, sp=(4, 8, 8), Cs=[128, 64], Cout=64, dims=3, extra_C=192),
    "conv3d_stride2": dict(N=2, sp=(8, 8, 8), Cs=[64], Cout=64, dims=3, stride=2),
    "conv2d_stride2_odd": dict(N=1, sp=(15, 17), Cs=[32], Cout=32, dims=2, stride=2),
    "conv3d_head_f32": dict(N=1, sp=(8, 8, 8), Cs=[64], Cout=12, dims=3, f32_out=True),
    "conv2d_160_320": dict(N=2, sp=(16, 16), Cs=[160], Cout=320, dims=2, emb=True),
    "conv3d_many_tiles": dict(N=2, sp=(16, 32, 32), Cs=[128], Cout=128, dims=3, residual=True),
    "conv3d_tiny_spatial": dict(N=3, sp=(2, 2, 2), Cs=[320, 320], Cout=320, dims=3, emb=True),
    "conv2d_bn64_brick": dict(N=1, sp=(32, 32), Cs=[64], Cout=64, dims=2, block_n=64, brick=(1, 1, 8, 16)),
    # split-K (general kernel)
    "splitk_tiny_spatial": dict(N=3, sp=(2, 2, 2), Cs=[320, 320], Cout=320, dims=3, emb=True, residual=True, split_k=5),
    "splitk_2d_4x4": dict(N=16, sp=(4, 4), Cs=[800], Cout=800, dims=2, emb=True, split_k=8),
    "splitk_stride2": dict(N=1, sp=(8, 8, 8), Cs=[128], Cout=128, dims=3, stride=2, split_k=3),
    "splitk_f32": dict(N=1, sp=(8, 8), Cs=[64], Cout=12, dims=2, f32_out=True, split_k=2),
    # halo-brick kernel (algo 1)
    "halo3d_64": dict(N=1, sp=(4, 16, 16), Cs=[64], Cout=64, dims=3, algo=1),
    "halo3d_all": dict(N=2, sp=(3, 32, 24), Cs=[128, 64], Cout=64, dims=3, extra_C=192, emb=True, algo=1),
    "halo3d_res_128": dict(N=2, sp=(8, 32, 32), Cs=[128], Cout=128, dims=3, residual=True, algo=1),
    "halo3d_ragged": dict(N=1, sp=(5, 20, 13), Cs=[16], Cout=64, dims=3, algo=1),
    "halo2d_160_320": dict(N=2, sp=(32, 32), Cs=[160], Cout=320, dims=2, emb=True, algo=1),
    "halo3d_head_f32": dict(N=1, sp=(4, 16, 16), Cs=[64], Cout=12, dims=3, f32_out=True, algo=1),
    "halo3d_many_tiles": dict(N=2, sp=(16, 64, 64), Cs=[64], Cout=64, dims=3, residual=True, algo=1),
    # halo-brick kernel with GroupNorm statistics accumulated in registers (64-channel outputs)
    "halo3d_stats": dict(N=2, sp=(16, 64, 64), Cs=[64], Cout=64, dims=3, residual=True, algo=1, stats=True),
    "halo3d_stats_ragged": dict(N=3, sp=(5, 20, 13), Cs=[16], Cout=64, dims=3, emb=True, algo=1, stats=True),
    "halo3d_stats_few_tiles": dict(N=5, sp=(1, 16, 8), Cs=[128, 64], Cout=60, dims=3, algo=1, stats=True),
    # halo-brick kernel, single CTAs forced (algo 2) / CTA pairs forced (algo 3; odd brick counts along w included)
    "halo_single_64": dict(N=2, sp=(4, 32, 32), Cs=[64], Cout=64, dims=3, residual=True, algo=2),
    "halo_pair_64": dict(N=1, sp=(4, 16, 16), Cs=[64], Cout=64, dims=3, algo=3),
    "halo_pair_all": dict(N=2, sp=(3, 32, 24), Cs=[128, 64], Cout=64, dims=3, extra_C=192, emb=True, algo=3),
    "halo_pair_res_128": dict(N=2, sp=(8, 32, 32), Cs=[128], Cout=128, dims=3, residual=True, algo=3),
    "halo_pair_ragged_odd": dict(N=3, sp=(5, 20, 21), Cs=[16], Cout=64, dims=3, emb=True, algo=3, stats=True),
    "halo_pair_one_brick": dict(N=2, sp=(2, 16, 8), Cs=[64], Cout=128, dims=3, algo=3),
    "halo_pair_2d_160_320": dict(N=2, sp=(32, 32), Cs=[160], Cout=320, dims=2, emb=True, algo=3),
    "halo_pair_head_f32": dict(N=1, sp=(4, 16, 16), Cs=[64], Cout=12, dims=3, f32_out=True, algo=3),
    "halo_pair_many_tiles": dict(N=2, sp=(16, 64, 64), Cs=[64], Cout=64, dims=3, residual=True, algo=3, stats=True),
    "halo_pair_cout_192": dict(N=1, sp=(2, 32, 16), Cs=[128], Cout=192, dims=3, algo=3),
    "halo_pair_stats_128": dict(N=2, sp=(4, 32, 32), Cs=[128], Cout=128, dims=3, residual=True, emb=True, algo=3, stats=True),
    "halo_stats_2d_320": dict(N=2, sp=(32, 24), Cs=[160], Cout=320, dims=2, emb=True, algo=1, stats=True),
    "halo_single_stats_72": dict(N=1, sp=(3, 16, 8), Cs=[64], Cout=72, dims=3, algo=2, stats=True),
    # depth-rolling kernel (algo 4): three depth taps stacked along N, two interleaved bricks per CTA, CTA pairs
    "roll_64": dict(N=1, sp=(4, 32, 16), Cs=[64], Cout=64, dims=3, algo=4),
    "roll_one_plane": dict(N=2, sp=(1, 32, 16), Cs=[64], Cout=64, dims=3, algo=4),
    "roll_all": dict(N=2, sp=(5, 32, 32), Cs=[128, 64], Cout=64, dims=3, extra_C=192, emb=True, residual=True, algo=4, stats=True),
    "roll_ragged_odd": dict(N=3, sp=(7, 20, 21), Cs=[16], Cout=64, dims=3, emb=True, algo=4, stats=True),
    "roll_head_f32": dict(N=1, sp=(6, 32, 16), Cs=[64], Cout=12, dims=3, f32_out=True, algo=4),
    "roll_cout_60": dict(N=1, sp=(3, 16, 16), Cs=[192], Cout=60, dims=3, residual=True, algo=4),
    "roll_many_items": dict(N=2, sp=(16, 64, 64), Cs=[64], Cout=64, dims=3, residual=True, algo=4, stats=True),
    "roll_deep": dict(N=1, sp=(40, 32, 16), Cs=[64], Cout=64, dims=3, algo=4),
}


@pytest.mark.parametrize("name", list(CONV_CASES))
def test_conv_tcgen05(ops, name):
    err, _, _ = _conv_case(ops, **CONV_CASES[name])
    # bf16 operands are identical on both sides; fp32 accumulation order + bf16 output rounding
    tol = 2e-5 if CONV_CASES[name].get("f32_out") else 6e-3
    assert err <= tol, f"{name}: rel err {err}"


@pytest.mark.parametrize("N,sp,C1,C2,Cskip,Cout,silu", [(2, (5, 32, 32), 64, 0, 0, 64, True), (1, (3, 40, 24), 128, 64, 192, 64, True),
                                                       (2, (4, 32, 16), 64, 0, 0, 12, True), (1, (2, 32, 16), 64, 64, 0, 64, False)])
def test_conv_roll_fused_groupnorm(ops, N, sp, C1, C2, Cskip, Cout, silu):
    """algo 4 with src_ss: GroupNorm (+SiLU) applied to the raw input planes inside the conv kernel must give exactly
    what gg_gn_apply followed by the same conv gives (same bf16 operands into the same MMA sequence) -- including
    the zero padding of the NORMALISED tensor at the h / w / d borders, concatenated norms and an un-normalised
    1x1x1 skip source."""
    no_tf32()
    rs = np.random.RandomState(5)
    mk = lambda c: (torch.from_numpy(rs.standard_normal((N,) + sp + (c,)).astype(np.float32) * 1.7 + 0.3)).cuda().to(torch.bfloat16)
    x1, x2 = mk(C1), (mk(C2) if C2 else None)
    xs = mk(Cskip) if Cskip else None
    C = C1 + C2
    gamma = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    beta = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    S = sp[0] * sp[1] * sp[2]
    ss = ops.gn_finalize(ops.gn_partial(x1), ops.gn_partial(x2) if C2 else None, gamma, beta, S, 1e-5)
    a_norm = ops.gn_apply(x1, x2, ss, silu)
    w = torch.from_numpy((rs.standard_normal((Cout, C, 3, 3, 3)) / math.sqrt(C * 27)).astype(np.float32)).cuda()
    extra = []
    if Cskip:
        extra = [torch.from_numpy((rs.standard_normal((Cout, Cskip)) / math.sqrt(Cskip)).astype(np.float32)).cuda()]
    b_pad = ops.pad_vec(torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda(), Cout)
    Cout8 = (Cout + 7) // 8 * 8

    def run(srcs, splits, src_ss):
        wp = ops.pack_conv_weight(w, splits, extra=extra, chunk_major=True)
        y = torch.full((N,) + sp + (Cout8,), float("nan"), dtype=torch.bfloat16, device="cuda")
        a = ops.make_conv_args(srcs, wp, Cout, y, dims=3, ksize=3, stride=1, bias=b_pad, algo=4, src_ss=src_ss,
                               ss_stride=2 * C, xf_silu=silu)
        ops.conv_fwd(a)
        torch.cuda.synchronize()
        return y

    skip = [(xs, True)] if Cskip else []
    want = run([(a_norm, False)] + skip, [C], None)
    fused_srcs = [(x1, False)] + ([(x2, False)] if C2 else []) + skip
    ss_ptrs = [ss.data_ptr()] + ([ss.data_ptr() + 8 * C1] if C2 else []) + ([None] if Cskip else [])
    got = run(fused_srcs, [C1] + ([C2] if C2 else []), ss_ptrs)
    assert not torch.isnan(got.float()).any()
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())


@pytest.mark.parametrize("N,sp,C1,C2,Cskip,Cout,silu,dims", [(2, (1, 64, 64), 160, 0, 0, 160, True, 2), (3, (1, 32, 24), 320, 160, 0, 320, True, 2),
                                                            (1, (4, 16, 16), 128, 64, 192, 128, True, 3), (2, (1, 20, 40), 96, 0, 0, 4, False, 2),
                                                            (1, (3, 18, 8), 256, 0, 0, 256, True, 3)])
def test_conv_halo_fused_groupnorm(ops, N, sp, C1, C2, Cskip, Cout, silu, dims):
    """algo 1 with src_ss (round 2): GroupNorm (+SiLU) applied to the landed input windows inside the halo-brick conv
    (openaimodel.py:207,233 / unet.py:191,217: GroupNorm32 -> SiLU -> conv) must give exactly what gg_gn_apply followed by
    the same conv gives -- single CTAs and CTA pairs, 2-D and 3-D, channel counts that are not multiples of 64 (LDM: 160),
    concatenated norms, an un-normalised 1x1 skip source, zero padding of the NORMALISED tensor at every border."""
    no_tf32()
    rs = np.random.RandomState(6)
    mk = lambda c: (torch.from_numpy(rs.standard_normal((N,) + sp + (c,)).astype(np.float32) * 1.7 + 0.3)).cuda().to(torch.bfloat16)
    x1, x2 = mk(C1), (mk(C2) if C2 else None)
    xs = mk(Cskip) if Cskip else None
    C = C1 + C2
    gamma = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    beta = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    S = sp[0] * sp[1] * sp[2]
    ss = ops.gn_finalize(ops.gn_partial(x1), ops.gn_partial(x2) if C2 else None, gamma, beta, S, 1e-5)
    a_norm = ops.gn_apply(x1, x2, ss, silu)
    w = torch.from_numpy((rs.standard_normal((Cout, C) + (3,) * dims) / math.sqrt(C * 3 ** dims)).astype(np.float32)).cuda()
    extra = []
    if Cskip:
        extra = [torch.from_numpy((rs.standard_normal((Cout, Cskip)) / math.sqrt(Cskip)).astype(np.float32)).cuda()]
    b_pad = ops.pad_vec(torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda(), Cout)
    Cout8 = (Cout + 7) // 8 * 8

    def run(srcs, splits, src_ss):
        wp = ops.pack_conv_weight(w, splits, extra=extra, chunk_major=True)
        y = torch.full((N,) + sp + (Cout8,), float("nan"), dtype=torch.bfloat16, device="cuda")
        a = ops.make_conv_args(srcs, wp, Cout, y, dims=dims, ksize=3, stride=1, bias=b_pad, algo=1, src_ss=src_ss,
                               ss_stride=2 * C, xf_silu=silu)
        ops.conv_fwd(a)
        torch.cuda.synchronize()
        return y

    skip = [(xs, True)] if Cskip else []
    want = run([(a_norm, False)] + skip, [C], None)
    fused_srcs = [(x1, False)] + ([(x2, False)] if C2 else []) + skip
    ss_ptrs = [ss.data_ptr()] + ([ss.data_ptr() + 8 * C1] if C2 else []) + ([None] if Cskip else [])
    got = run(fused_srcs, [C1] + ([C2] if C2 else []), ss_ptrs)
    assert not torch.isnan(got.float()).any()
    if C2 == 0 or C1 % 64 == 0:
        # same bf16 operands into the same MMA sequence (a concat whose first part is not a multiple of 64 channels is chunked
        # differently as two sources than as one tensor: other summation order, compared with a tolerance instead)
        assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    else:
        assert rel_err(got.float(), want.float()) <= 6e-3


@pytest.mark.parametrize("N,sp,silu", [(2, (6, 32, 16), True), (1, (5, 40, 24), True), (3, (4, 48, 20), False)])
def test_conv_roll_sampler_epilogue_equals_logits_plus_per_voxel_kernel(ops, N, sp, silu):
    """gg_conv_args.cat: the 64 -> 12 head conv with softmax + posterior + clamp + Philox draw + next-input row in its
    epilogue must draw EXACTLY the labels (and write exactly the next-input rows) that the same conv writing fp32 logits
    followed by gg_cat_step_cl's production kernel gives -- same accumulators, same arithmetic, same random words.
    Shapes cover full bricks, bricks cut in h and w (partial rows take the byte-store path) and whole CTAs out of range."""
    import ctypes as C_
    from jointimagegeneration_b200 import _C
    no_tf32()
    rs = np.random.RandomState(11)
    Cin, Cout, Cin_pad = 64, 12, 16
    V = sp[0] * sp[1] * sp[2]
    assert V % 4 == 0
    x = torch.from_numpy(rs.standard_normal((N,) + sp + (Cin,)).astype(np.float32) * 1.3 + 0.2).cuda().to(torch.bfloat16)
    gamma = torch.from_numpy(rs.standard_normal(Cin).astype(np.float32)).cuda()
    beta = torch.from_numpy(rs.standard_normal(Cin).astype(np.float32)).cuda()
    ss = ops.gn_finalize(ops.gn_partial(x), None, gamma, beta, V, 1e-5)
    w = torch.from_numpy((rs.standard_normal((Cout, Cin, 3, 3, 3)) * 4.0 / math.sqrt(Cin * 27)).astype(np.float32)).cuda()
    b_pad = ops.pad_vec(torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda(), Cout)
    wp = ops.pack_conv_weight(w, [Cin], chunk_major=True)
    logits = torch.full((N,) + sp + (16,), float("nan"), dtype=torch.float32, device="cuda")

    def args():
        return ops.make_conv_args([(x, False)], wp, Cout, logits, dims=3, ksize=3, stride=1, bias=b_pad, algo=4, src_ss=[ss.data_ptr()],
                                  ss_stride=2 * Cin, xf_silu=silu)
    ops.conv_fwd(args())
    lab_in = torch.from_numpy(rs.randint(0, Cout, size=N * V).astype(np.uint8)).cuda()
    coef = torch.tensor([[0.93, 0.41], [0.5, 0.9], [0.999, 0.02]][:N], dtype=torch.float32).cuda()
    cond = torch.from_numpy(rs.standard_normal((N * V, 1)).astype(np.float32)).cuda().to(torch.bfloat16)
    want_lab = torch.empty(N * V, dtype=torch.uint8, device="cuda")
    want_nx = torch.full((N * V, Cin_pad), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.cat_step_cl(logits, lab_in, coef, want_lab, N, V, Cout, cond=cond, n_cond=1, next_x=want_nx, seed=21, offset=5, vox_base=4 * V)
    got_lab = torch.full((N * V,), 255, dtype=torch.uint8, device="cuda")
    got_nx = torch.full((N * V, Cin_pad), float("nan"), dtype=torch.bfloat16, device="cuda")
    cat = _C.CatEpilogue(lab_in.data_ptr(), got_lab.data_ptr(), got_nx.data_ptr(), cond.data_ptr(), coef.data_ptr(), Cout, 1, Cin_pad,
                         ops.CAT_SAMPLE, 1e-12, 21, 5, 4 * V)
    a = args()
    a.cat = C_.pointer(cat)
    logits.fill_(float("nan"))
    ops.conv_fwd(a)
    torch.cuda.synchronize()
    assert torch.isnan(logits).all(), "the sampler epilogue must not write the logits tensor"
    assert int((got_lab == 255).sum()) == 0
    assert torch.equal(got_lab, want_lab), f"{int((got_lab != want_lab).sum())} of {N * V} labels differ"
    assert torch.equal(got_nx.view(torch.int16), want_nx.view(torch.int16))
    hist = torch.bincount(got_lab.long(), minlength=Cout)
    assert int((hist > 0).sum()) >= Cout - 2          # a real draw, not a constant


# -------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,H,T,d", [(2, 4, 64, 32), (1, 8, 2048, 32), (2, 10, 256, 32), (3, 2, 16, 32), (1, 5, 100, 32),
                                     (1, 2, 130, 64), (4, 10, 1024, 32)])
def test_attention_legacy(ops, B, H, T, d):
    no_tf32()
    rs = np.random.RandomState(2)
    qkv = torch.from_numpy(rs.standard_normal((B, T, 3 * H * d)).astype(np.float32)).cuda().to(torch.bfloat16)
    out = ops.attention_legacy(qkv, H)
    f = qkv.float().reshape(B, T, H, 3, d)
    q, k, v = (f[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    want = attention_ref(q, k, v, 1.0 / math.sqrt(d))
    got = heads_of(out, H)
    assert rel_err(got, want) <= 1e-2   # bf16 P and bf16 output


def test_attention_cross(ops):
    no_tf32()
    rs = np.random.RandomState(3)
    B, H, Tq, Tk, d = 2, 5, 64, 7, 32
    q = torch.from_numpy(rs.standard_normal((B, Tq, H * d)).astype(np.float32)).cuda().to(torch.bfloat16)
    kv = torch.from_numpy(rs.standard_normal((B, Tk, 2 * H * d)).astype(np.float32)).cuda().to(torch.bfloat16)
    o = torch.empty_like(q)
    W = H * d
    a = ops.make_attn_args(q, kv, kv[..., W:], o, B, H, Tq, Tk, d, d ** -0.5, (Tq * W, W, d), (Tk * 2 * W, 2 * W, d),
                           (Tk * 2 * W, 2 * W, d), (Tq * W, W, d))
    a.v = kv.data_ptr() + W * 2
    ops.attention_fwd(a)
    want = attention_ref(heads_of(q, H), heads_of(kv[..., :W], H), heads_of(kv[..., W:], H), d ** -0.5)
    assert rel_err(heads_of(o, H), want) <= 1e-2


@pytest.mark.parametrize("B,H,Tq,Tk,d", [(2, 5, 300, 77, 64), (1, 8, 1024, 512, 64), (2, 3, 200, 333, 32), (1, 8, 4096, 4096, 32),
                                         (1, 2, 64, 2000, 32), (2, 16, 2500, 1100, 64), (8, 8, 2048, 2048, 32)])
def test_attention_tensor_core_kernel(ops, B, H, Tq, Tk, d):
    """attention_tc.cu (tcgen05 Q K^T and P V, TMEM accumulators, TMA operands, lazy online-softmax rescaling) through
    gg_attention_fwd with a workspace: cross-attention shapes (separate q / kv tensors, ragged Tq / Tk, key-tail masking,
    d = 32 and 64) and a long self-attention, against fp32 softmax attention.  Large score ranges exercise the rescaling; the
    last two shapes fill the GPU and run two query tiles per CTA, the others one."""
    from jointimagegeneration_b200 import _C
    import ctypes as C_
    no_tf32()
    rs = np.random.RandomState(7)
    W = H * d
    q = torch.from_numpy((rs.standard_normal((B, Tq, W)) * 2.0).astype(np.float32)).cuda().to(torch.bfloat16)
    kv = torch.from_numpy((rs.standard_normal((B, Tk, 2 * W)) * 2.0).astype(np.float32)).cuda().to(torch.bfloat16)
    o = torch.full((B, Tq, W), float("nan"), dtype=torch.bfloat16, device="cuda")
    a = ops.make_attn_args(q, kv, kv[..., W:], o, B, H, Tq, Tk, d, d ** -0.5, (Tq * W, W, d), (Tk * 2 * W, 2 * W, d),
                           (Tk * 2 * W, 2 * W, d), (Tq * W, W, d))
    a.v = kv.data_ptr() + W * 2
    assert int(_C.lib().gg_attention_workspace_bytes(C_.byref(a))) > 0, "the tensor-core kernel should take this shape"
    ws = ops.attention_fwd(a, q.device)
    assert ws is not None
    torch.cuda.synchronize()
    assert not torch.isnan(o.float()).any()
    want = attention_ref(heads_of(q, H), heads_of(kv[..., :W], H), heads_of(kv[..., W:], H), d ** -0.5)
    assert rel_err(heads_of(o, H), want) <= 1e-2


# ------------------------------------------------------------------------------ small pieces
def test_timestep_embedding_and_small_linear(ops):
    from oracle import nets
    t = torch.tensor([1000.0, 981.0, 1.0, 21.0, 500.0])
    for dim in (64, 160, 128):
        got = ops.timestep_embedding(t.cuda(), dim)
        want = nets.timestep_embedding(t, dim)
        assert float((got.cpu() - want).abs().max()) <= 2e-4   # |arg| up to 1e3 rad: sinf/cosf/expf ulps
    rs = np.random.RandomState(4)
    for M, K, N in ((1, 64, 256), (8, 256, 1000), (16, 640, 333)):
        x = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32))
        w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32) / math.sqrt(K))
        b = torch.from_numpy(rs.standard_normal(N).astype(np.float32))
        got = ops.small_linear(x.cuda(), w.cuda(), b.cuda(), act_in=True, act_out=True)
        want = torch.nn.functional.silu(torch.nn.functional.linear(torch.nn.functional.silu(x), w, b))
        assert float((got.cpu() - want).abs().max()) <= 1e-4


def test_layernorm_geglu_upsample(ops):
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.standard_normal((3, 50, 320)).astype(np.float32)).cuda().to(torch.bfloat16)
    g = torch.from_numpy(rs.standard_normal(320).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.standard_normal(320).astype(np.float32)).cuda()
    got = ops.layernorm(x, g, b)
    want = torch.nn.functional.layer_norm(x.float(), (320,), g, b, 1e-5)
    assert rel_err(got, want) <= 6e-3
    h = torch.from_numpy(rs.standard_normal((3, 50, 2 * 1280)).astype(np.float32)).cuda().to(torch.bfloat16)
    got = ops.geglu(h)
    a, gate = h.float().chunk(2, -1)
    assert rel_err(got, a * torch.nn.functional.gelu(gate)) <= 6e-3
    u = torch.from_numpy(rs.standard_normal((2, 3, 4, 5, 16)).astype(np.float32)).cuda().to(torch.bfloat16)
    got = ops.upsample2x(u, 3)
    want = torch.nn.functional.interpolate(nchw_from_cl(u), scale_factor=2, mode="nearest")
    assert torch.equal(nchw_from_cl(got), want)
    got2 = ops.upsample2x(u[:, :1].contiguous(), 2)
    want2 = torch.nn.functional.interpolate(nchw_from_cl(u[:, :1])[:, :, 0], scale_factor=2, mode="nearest")
    assert torch.equal(nchw_from_cl(got2)[:, :, 0], want2)


def test_cat_step_cl_fast_path_distribution_and_slab_invariance(ops):
    """Production per-voxel kernel (in-kernel Philox, inverse-CDF draw): (i) the empirical label
    distribution over 2^20 identical voxels matches the oracle posterior, (ii) labels depend only on
    (seed, offset, GLOBAL voxel index): a half-volume call with vox_base reproduces the full call,
    (iii) the next-input row is the one-hot of the drawn label plus the condition channel."""
    from oracle import diffusion
    C, V, t = 12, 1 << 20, 300
    _, alphas, cumalphas = diffusion.cosine_schedule(1000)
    a, g = diffusion.step_coefficients(alphas, cumalphas, t)
    rs = np.random.RandomState(8)
    row = (rs.standard_normal(16) * 1.5).astype(np.float32)
    logits = torch.from_numpy(np.tile(row, (V, 1))).cuda()
    lab_in = torch.full((V,), 5, dtype=torch.uint8, device="cuda")
    coef = torch.tensor([[a, g]], dtype=torch.float32).cuda()
    cond = torch.full((V, 1), 0.25, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(V, dtype=torch.uint8, device="cuda")
    nx = torch.empty((V, 16), dtype=torch.bfloat16, device="cuda")
    ops.cat_step_cl(logits, lab_in, coef, out, 1, V, C, cond=cond, n_cond=1, next_x=nx, seed=3, offset=17)
    x0 = torch.softmax(torch.from_numpy(row[:C]), 0).numpy().reshape(C, 1)
    xt = np.eye(C, dtype=np.float32)[5].reshape(C, 1)
    p = np.maximum(diffusion.theta_post_prob_closed(a, g, xt, x0)[:, 0], 1e-12)
    p = p / p.sum()
    hist = torch.bincount(out.long(), minlength=C).cpu().numpy() / V
    # binomial standard error at 2^20 draws is <= 5e-4; allow 5 sigma
    assert np.abs(hist - p).max() <= 2.5e-3, (hist, p)
    nxf = nx.float().cpu().numpy()
    assert np.array_equal(nxf[:, :C].argmax(1), out.cpu().numpy()) and np.all(nxf[:, :C].sum(1) == 1)
    assert np.all(nxf[:, C] == 0.25) and np.all(nxf[:, C + 1:] == 0)
    # determinism and dependence on the global voxel index only
    out2 = torch.empty_like(out)
    ops.cat_step_cl(logits, lab_in, coef, out2, 1, V, C, cond=cond, n_cond=1, seed=3, offset=17)
    assert torch.equal(out, out2)
    half = V // 2
    out3 = torch.empty(half, dtype=torch.uint8, device="cuda")
    ops.cat_step_cl(logits[half:], lab_in[half:], coef, out3, 1, half, C, cond=cond[half:], n_cond=1, seed=3, offset=17, vox_base=half)
    assert torch.equal(out3, out[half:])
    ops.cat_step_cl(logits, lab_in, coef, out2, 1, V, C, seed=3, offset=18)
    assert not torch.equal(out, out2)


def test_softmax_rows_and_transpose(ops):
    """gg_softmax_rows / gg_transpose_bf16 (the VAE's single-head attention, model.py:250-256) vs torch."""
    no_tf32()
    g = torch.Generator(device="cuda").manual_seed(0)
    for rows, n in ((5, 64), (300, 4096), (7, 8192), (33, 260)):
        x = torch.randn((rows, n), device="cuda", generator=g) * 6.0
        y = ops.softmax_rows(x, scale=0.37)
        want = torch.softmax(x * 0.37, -1)
        assert y.dtype == torch.bfloat16 and y.shape == x.shape
        assert float((y.float() - want).abs().max()) <= 2.0 ** -8 * float(want.max()) + 1e-6       # bf16 output rounding
        assert float((y.float().sum(-1) - 1).abs().max()) <= 2e-2
    for R, Cc in ((4096, 512), (100, 72), (64, 64), (1, 8)):
        a = torch.randn((R, Cc), device="cuda", generator=g).to(torch.bfloat16)
        assert torch.equal(ops.transpose_bf16(a), a.t().contiguous())


@pytest.mark.parametrize("sp,C,Cout", [((32, 32), 64, 64), ((18, 22), 32, 40), ((64, 64), 128, 128)])
def test_conv_stride2_asymmetric_pad(ops, sp, C, Cout):
    """The VAE Downsample (model.py:61-80): F.pad(x, (0, 1, 0, 1)) then a stride-2, padding-0 conv == tap offset 0."""
    no_tf32()
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.standard_normal((2, 1) + sp + (C,)).astype(np.float32)).cuda().to(torch.bfloat16)
    w = torch.from_numpy((rs.standard_normal((Cout, C, 3, 3)) / math.sqrt(9 * C)).astype(np.float32)).cuda()
    b = torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).cuda()
    osp = ((sp[0] + 1 - 3) // 2 + 1, (sp[1] + 1 - 3) // 2 + 1)
    Cout8 = (Cout + 7) // 8 * 8
    y = torch.full((2, 1) + osp + (Cout8,), float("nan"), dtype=torch.bfloat16, device="cuda")
    wp = ops.pack_conv_weight(w, [C])
    b_pad = ops.pad_vec(b, Cout)
    a = ops.make_conv_args([(x, False)], wp, Cout, y, dims=2, ksize=3, stride=2, bias=b_pad, offsets=(0, 0, 0), out_spatial=(1,) + osp)
    ops.conv_fwd(a)
    torch.cuda.synchronize()
    xn = nchw_from_cl(x)[:, :, 0]
    want = torch.nn.functional.conv2d(torch.nn.functional.pad(xn, (0, 1, 0, 1)), w.to(torch.bfloat16).float(), b, stride=2)
    got = nchw_from_cl(y, Cout)[:, :, 0]
    assert got.shape == want.shape and not torch.isnan(got).any()
    assert rel_err(got, want) <= 6e-3


@pytest.mark.parametrize("N,sp,C1,C2,silu", [(16, (64, 64), 320, 0, True), (3, (5, 7, 9), 64, 128, False), (2, (8, 8), 800, 800, True),
                                             (1, (4, 4, 4), 320, 0, True), (4, (16, 16), 640, 320, True), (16, (64, 64), 160, 0, True),
                                             (5, (32, 32), 320, 320, True), (8, (8, 16, 16), 256, 0, False), (2, (4, 4), 1600, 448, True),
                                             (3, (3, 3), 32, 0, True), (2, (17, 32), 160, 320, True)])
def test_group_norm_one_launch(ops, N, sp, C1, C2, silu):
    """gg_gn_fused vs the three-kernel GroupNorm and vs torch: the shared-memory-resident form (cluster of 1..8 CTAs per
    sample, bulk async loads, DSMEM reduction) where a sample fits, the L2-streaming cluster form where it does not
    (the first case)."""
    from jointimagegeneration_b200 import _C
    S = int(np.prod(sp))
    cl = int(_C.lib().gg_gn_fused_resident(S, C1 + C2))
    assert (cl == 0) == (N == 16 and C1 == 320), cl
    assert cl in (0, 1, 2, 4, 8)
    no_tf32()
    rs = np.random.RandomState(7)
    sp3 = (1,) * (3 - len(sp)) + tuple(sp)
    mk = lambda c: torch.from_numpy((rs.standard_normal((N,) + sp3 + (c,)) * 1.5 + 0.4).astype(np.float32)).cuda().to(torch.bfloat16)
    x1, x2 = mk(C1), (mk(C2) if C2 else None)
    C = C1 + C2
    gamma = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    beta = torch.from_numpy(rs.standard_normal(C).astype(np.float32)).cuda()
    got = ops.gn_fused(x1, x2, gamma, beta, 1e-6, silu)
    three = ops.group_norm_cl(x1, x2, gamma, beta, 1e-6, silu)
    xr = nchw_from_cl(torch.cat([x1, x2], -1) if C2 else x1)
    want = torch.nn.functional.group_norm(xr, 32, gamma, beta, 1e-6)
    if silu:
        want = torch.nn.functional.silu(want)
    assert rel_err(nchw_from_cl(got), want) <= 6e-3
    # same formula as the three-kernel path; only the summation order of the statistics differs (last-bit flips)
    d = (got.float() - three.float()).abs()
    assert float((d > 0).float().mean()) <= 5e-3 and float(d.max()) <= 2.0 ** -7 * float(three.float().abs().max())
    assert torch.equal(got, ops.gn_fused(x1, x2, gamma, beta, 1e-6, silu))
