"""Worker of tests/test_gpu_multi.py: depth-slab CCDM forward + sampler steps on WORLD_SIZE GPUs,
checked on every rank against the unsplit computation on the same GPU.  Launch with torchrun."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jointimagegeneration_b200.ccdm import build_model
    from jointimagegeneration_b200.sharding import SlabComm, slab_ranges
    from oracle import configs, weights
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    C, spatial, T = 12, (16 * world, 32, 32), 20
    m = build_model(T, "cosine", {"s": 0.008}, [(1,) + spatial, (C,) + spatial], None, "unet_openai",
                    dict(configs.CCDM_PARAMS_YML), "x", "majority", dims=3)
    m.unet.load_state_dict(weights.synth_state_dict(weights.shapes_of(m.unet), 9))
    m = m.cuda().eval()
    x = weights.uniform_one_hot(4, 1, C, spatial).cuda()
    cond = torch.zeros(1, 1, *spatial).cuda()
    t = torch.full((1,), 13.0).cuda()
    lo, hi = slab_ranges(spatial[0], world)[rank]
    # ---- unsplit reference on this GPU
    ref = m.unet(x, cond, None, t)["diffusion_out"].clone()
    m.loop, m.philox_seed = "resident", 5
    m.record = []
    ref_final = m(x, cond, t=torch.tensor(10000 + 3))["diffusion_out"].clone()
    ref_labels = [r.clone() for r in m.record]
    # ---- slab mode
    comm = SlabComm()
    m.unet.enable_slab(comm)
    xs, cs = x[:, :, lo:hi].contiguous(), cond[:, :, lo:hi].contiguous()
    got = m.unet(xs, cs, None, t)["diffusion_out"]
    want = ref[:, :, lo:hi]
    err = float((got - want).abs().max())
    m.record = []
    fin = m(xs, cs, t=torch.tensor(10000 + 3))["diffusion_out"]
    V = spatial[1] * spatial[2]
    agree = []
    for a, b in zip(m.record, ref_labels):
        agree.append(float((a.view(-1) == b.view(1, spatial[0], V)[:, lo:hi].reshape(-1)).float().mean()))
    fin_agree = float((fin == ref_final[:, :, lo:hi]).float().mean())
    print(f"rank {rank}/{world}: slab planes [{lo},{hi}) probs max-abs diff vs unsplit {err:.3e}; label agreement per step {agree}; "
          f"final one-hot agreement {fin_agree:.5f}; halo exchanges/forward {comm.n_exchanges}, gathers {comm.n_gathers}, "
          f"bytes sent {comm.bytes_sent}", flush=True)
    # step 1 starts from identical inputs: labels agree except at near-ties.  Later steps start from the previous step's
    # (slightly different) labels and the synthetic network's probabilities are nearly flat, so the chains drift apart
    # (0.997 -> 0.89 by step 3 with 5e-3 differences in the probabilities); the bound only catches gross errors.
    ok = err <= 2e-2 and agree[0] >= 0.995 and min(agree) >= 0.8
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
