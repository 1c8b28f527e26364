"""Depth-slab decomposition (BASELINE config 5) on ONE GPU: the ranks are emulated by sharding.LocalSlabGroup, which
runs the R slab plans -- the same plans, kernels, halo layout, GroupNorm combine, key/value gather and global voxel
indexing as the multi-GPU run -- in lock step and carries out the collectives on the R buffer sets directly.  Checked
against (a) the unsplit plan on the same GPU and (b) the fp32 CPU oracle of the reference forward
(ccdm/ddpm/models/unet_openai/unet.py:758-823): the slab result must be as close to the oracle as the unsplit one,
also on the planes next to the slab boundaries, and teacher-forced sampler steps must draw the same labels.
tests/test_gpu_multi.py runs the same check over NCCL when the box has more than one GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

C = 12


def _model(spatial, T=20, seed_w=9):
    from jointimagegeneration_b200.ccdm import build_model
    from oracle import configs, weights
    m = build_model(T, "cosine", {"s": 0.008}, [(1,) + spatial, (C,) + spatial], None, "unet_openai", dict(configs.CCDM_PARAMS_YML), "x",
                    "majority", dims=3)
    sd = weights.synth_state_dict(weights.shapes_of(m.unet), seed_w)
    m.unet.load_state_dict(sd)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("world", [2, 4])
def test_virtual_slab_ranks_match_unsplit_and_oracle(world):
    from jointimagegeneration_b200 import slab_check
    from oracle import nets, weights
    spatial = (16 * world, 32, 32)
    m, sd = _model(spatial)
    x = weights.uniform_one_hot(4, 1, C, spatial)
    cond = torch.zeros(1, 1, *spatial)
    t_values = [13, 12, 11]
    rec = slab_check.unsplit_chain(m, x.cuda(), cond.cuda(), t_values, seed=5)
    res = slab_check.slab_vs_unsplit(m, rec, world)
    # fp32 oracle of the reference forward on the same weights
    orc = nets.unet_forward(sd, x, torch.full((1,), float(t_values[0])), input_condition=cond, softmax_output=True,
                            num_head_channels=32).cuda()
    ref = rec["probs0"]
    e_ref = (ref - orc).abs().amax((0, 1, 3, 4))                       # per depth plane
    got = torch.empty_like(ref)
    Dl = res["planes_per_rank"]
    for r, prof in res["plane_max_abs"]:
        assert len(prof) == Dl
    ses = slab_check.SlabSession(m.unet, spatial, world)
    ses.load_input(rec["xins"][0], t_values[0])
    ses.run()
    for p, r in zip(ses.probs(C), ses.ranks):
        got[:, :, r * Dl:(r + 1) * Dl] = p
    ses.close()
    e_slab = (got - orc).abs().amax((0, 1, 3, 4))
    bnd = sorted({d for r in range(1, world) for d in (r * Dl - 1, r * Dl)})
    print(f"world {world}: slab vs unsplit max-abs {res['parity_max_abs']:.3e} (bit-equal {res['bit_equal']}); vs fp32 oracle: unsplit "
          f"{float(e_ref.max()):.3e}, slab {float(e_slab.max()):.3e}, slab at boundary planes {float(e_slab[bnd].max()):.3e}; "
          f"teacher-forced label agreement {res['agree']}")
    assert res["parity_max_abs"] <= 1e-2
    assert float(e_slab.max()) <= 1.25 * float(e_ref.max()) + 1e-3          # equal distance from the oracle
    assert float(e_slab[bnd].max()) <= 1.25 * float(e_ref.max()) + 1e-3     # a wrong halo plane would show here
    assert min(res["agree"]) >= 0.995, res["agree"]


def test_virtual_slab_world8_planes_per_rank_16():
    """Eight slabs of 16 planes (config 5's split at 8 GPUs: the coarsest level holds ONE plane per rank)."""
    from jointimagegeneration_b200 import slab_check
    from oracle import weights
    spatial = (128, 32, 16)
    m, _ = _model(spatial)
    x = weights.uniform_one_hot(6, 1, C, spatial).cuda()
    cond = torch.zeros(1, 1, *spatial).cuda()
    rec = slab_check.unsplit_chain(m, x, cond, [9, 8], seed=2)
    res = slab_check.slab_vs_unsplit(m, rec, 8)
    print(f"world 8: slab vs unsplit max-abs {res['parity_max_abs']:.3e}; label agreement {res['agree']}")
    assert res["parity_max_abs"] <= 1e-2 and min(res["agree"]) >= 0.995


@pytest.mark.parametrize("world", [2, 4])
def test_virtual_slab_ranks_over_peer_memory_kernels(world):
    """The NVLink peer-memory transport (csrc/peer_comm.cu: one gg_peer_exchange kernel per halo exchange / GroupNorm
    combine / key-value gather, flags + epochs instead of NCCL) with every rank's arena in this process: the R plans run
    in lock step, phase 1 (push + raise flags) of all ranks before phase 2 (wait + unpack) of any.  Same parity bar as the
    copy transport; two forwards in a row exercise the epoch / flag reuse."""
    from jointimagegeneration_b200 import slab_check
    from oracle import weights
    spatial = (16 * world, 32, 32)
    m, _ = _model(spatial)
    x = weights.uniform_one_hot(4, 1, C, spatial).cuda()
    cond = torch.zeros(1, 1, *spatial).cuda()
    rec = slab_check.unsplit_chain(m, x, cond, [13, 12, 11], seed=5)
    copy = slab_check.slab_vs_unsplit(m, rec, world, transport="copy")
    peer = slab_check.slab_vs_unsplit(m, rec, world, transport="peer")
    print(f"world {world} peer transport: slab vs unsplit max-abs {peer['parity_max_abs']:.3e}; label agreement {peer['agree']}")
    assert peer["parity_max_abs"] == copy["parity_max_abs"]          # the transport moves bytes: identical results
    assert peer["agree"] == copy["agree"]
    assert peer["parity_max_abs"] <= 1e-2 and min(peer["agree"]) >= 0.995
