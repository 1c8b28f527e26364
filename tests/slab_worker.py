"""Worker of tests/test_gpu_multi.py: depth-slab CCDM forward + teacher-forced sampler steps on WORLD_SIZE GPUs over
NCCL, checked on every rank against the unsplit computation on the same GPU (jointimagegeneration_b200.slab_check).
Launch with torchrun."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jointimagegeneration_b200 import slab_check
    from jointimagegeneration_b200.ccdm import build_model
    from jointimagegeneration_b200.sharding import SlabComm
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
    rec = slab_check.unsplit_chain(m, x, cond, [13, 12, 11], seed=5)          # unsplit reference on this GPU
    ok = True
    results = {}
    for transport in ("nccl", "peer"):
        # this rank's slab, teacher-forced: over host-enqueued NCCL calls, then over gg_peer_exchange kernels (NVLink peer memory)
        from jointimagegeneration_b200.sharding import PeerSlabComm
        comm = SlabComm() if transport == "nccl" else PeerSlabComm(arena_bytes=256 << 20)
        res = slab_check.slab_vs_unsplit(m, rec, world, comm)
        print(f"rank {rank}/{world} [{transport}]: probs max-abs diff vs unsplit {res['parity_max_abs']:.3e} (bit-equal {res['bit_equal']}); "
              f"teacher-forced label agreement per step {res['agree']}", flush=True)
        worst = slab_check.reduce_over_ranks(res, x.device)
        results[transport] = worst
        ok = ok and worst["parity_max_abs"] <= 1e-2 and min(worst["agree"]) >= 0.995
        torch.cuda.synchronize()
        dist.barrier()
        if hasattr(comm, "close"):
            comm.close()
    # the transport only moves bytes: both must give the same numbers
    ok = ok and results["nccl"]["parity_max_abs"] == results["peer"]["parity_max_abs"] and results["nccl"]["agree"] == results["peer"]["agree"]
    dist.barrier()
    m.unet.invalidate()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
