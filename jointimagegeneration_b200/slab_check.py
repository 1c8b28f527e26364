"""Self-check of the depth-slab decomposition (BASELINE config 5) against the unsplit computation on the same GPU.

One volume [1, C, D, H, W] is split into `world` equal depth slabs; rank r runs the SAME kernels on its slab and
exchanges halos / GroupNorm sums / attention keys+values with the other ranks (sharding.SlabComm over NCCL, or
sharding.LocalSlabGroup = all ranks emulated on one device).  The check runs

* one UNet forward: slab probabilities vs the unsplit plan's (max-abs difference, per-plane profile so that a wrong
  halo plane -- which would show at the slab boundaries -- is visible), and
* `steps` TEACHER-FORCED sampler steps: step k of the slab run starts from the unsplit run's labels of step k-1, so
  every step sees identical inputs and the drawn labels (in-kernel Philox keyed on the GLOBAL voxel index) may differ
  only where the bf16 networks' probabilities differ at a near-tie.

The reference has no slab mode; the semantics checked are those of its single-device forward
(ccdm/ddpm/models/diffusion_denoising.py:176-227, unet_openai/unet.py:758-823).  Used by tests/ (virtual ranks on
one GPU, real ranks under torchrun) and by bench.py --gpus N (N > 1) to put slab parity into the driver's record.
"""
import math
import os
from typing import List, Optional

import torch

from . import ops
from .sharding import LocalSlabGroup, slab_ranges


class SlabSession:
    """The slab plan(s) this process executes: one real rank (comm given) or all `world` virtual ranks."""

    def __init__(self, unet, spatial_full, world: int, comm=None, transport: str = "copy"):
        """comm: this process' SlabComm / PeerSlabComm (one real rank), or None = all `world` ranks emulated here with
        transport "copy" (LocalSlabGroup: collectives as device copies) or "peer" (LocalPeerGroup: the gg_peer_exchange
        kernels with every arena in this process)."""
        from .sharding import LocalPeerGroup
        from .unet_engine import UNetEngine
        self.unet, self.world, self.spatial = unet, world, tuple(spatial_full)
        self.ranges = slab_ranges(spatial_full[0], world)
        self.local_sp = (spatial_full[0] // world,) + tuple(spatial_full[1:])
        self.Vl = int(math.prod(self.local_sp))
        self.comm = comm
        if comm is not None:
            assert comm.world == world
            unet.enable_slab(comm)
            self.ranks = [comm.rank]
            self.group = None
            self.plans = [unet.plan_for(1, self.local_sp)]
        else:
            self.group = LocalPeerGroup(world) if transport == "peer" else LocalSlabGroup(world)
            self.ranks = list(range(world))
            self.engines = []
            for r in self.ranks:
                e = UNetEngine(unet, unet.dims, unet.num_heads, unet.num_head_channels)
                e.slab = self.group.comms[r]
                self.engines.append(e)
            self.plans = [e.get_plan(1, self.local_sp, unet.in_channels_padded) for e in self.engines]

    def close(self):
        if self.comm is not None:
            self.unet.enable_slab(None)
        self.plans = []
        if self.group is not None and hasattr(self.group, "close"):
            import torch
            torch.cuda.synchronize()
            self.engines = []
            self.group.close()

    def load_input(self, xin_full: torch.Tensor, t: float):
        """xin_full: the unsplit plan's input, CL bf16 [1, D, H, W, Cpad]."""
        for p, r in zip(self.plans, self.ranks):
            lo, hi = self.ranges[r]
            p.inputs["x"].copy_(xin_full[:, lo:hi])
            p.inputs["t"].fill_(float(t))

    def run(self):
        if self.group is not None:
            self.group.run(self.plans)
        else:
            self.plans[0].run()

    def probs(self, C: int) -> List[torch.Tensor]:
        return [ops.cl_to_nchw(p.outputs["head"], C, self.local_sp, softmax=True) for p in self.plans]

    def draw(self, labels_in_full: torch.Tensor, coef: torch.Tensor, C: int, seed: int, offset: int) -> List[torch.Tensor]:
        """One fused per-voxel step per local rank from the current head logits; returns uint8 labels [Vl] per rank."""
        outs = []
        for p, r in zip(self.plans, self.ranks):
            lab_in = labels_in_full.view(-1)[r * self.Vl:(r + 1) * self.Vl].contiguous()
            out = torch.empty_like(lab_in)
            ops.cat_step_cl(p.outputs["head"], lab_in, coef, out, 1, self.Vl, C, mode=ops.CAT_SAMPLE, seed=seed, offset=offset,
                            vox_base=r * self.Vl)
            outs.append(out)
        return outs


@torch.no_grad()
def unsplit_chain(model, x: torch.Tensor, cond: torch.Tensor, t_values, seed: int):
    """The unsplit resident chain on this GPU, recorded: per step the plan input (CL bf16), the probabilities of the first
    step and the uint8 labels drawn."""
    unet = model.unet
    assert unet.engine.slab is None
    B, C = x.shape[:2]
    assert B == 1
    spatial = tuple(x.shape[2:])
    V = int(math.prod(spatial))
    # the reference plan takes the kernel choices of a slab plan wherever those do not depend on the decomposition: the slab
    # plans compute GroupNorm with the three-launch form (whole-volume statistics), so the unsplit plan does here, too --
    # results then differ only through the order in which the ranks' group sums are combined
    eng = unet.engine
    prev = eng.fused_small_gn
    eng.fused_small_gn = False
    eng.plans.pop((1, tuple(spatial), unet.in_channels_padded, None, os.environ.get("GG_LANES")), None)
    plan = unet.plan_for(1, spatial)
    eng.fused_small_gn = prev
    eng.plans.pop((1, tuple(spatial), unet.in_channels_padded, None, os.environ.get("GG_LANES")), None)      # not a plan other callers should get
    xin = plan.inputs["x"]
    ops.nchw_to_cl(x.float().contiguous(), cond.float().contiguous(), c_pad=unet.in_channels_padded, out=xin)
    n_cond = cond.shape[1]
    cond_cl = ops.nchw_to_cl(cond.float().contiguous(), None, c_pad=8)[..., :n_cond].contiguous()
    lab = torch.empty((V,), dtype=torch.uint8, device=x.device)
    ops.cat_posterior_sample(x.float().contiguous(), None, None, ops.CAT_ARGMAX_GIVEN, clamp_min=0.0, labels=lab.view(1, V))
    coefs = model.diffusion.step_coef_tensor(torch.tensor(list(t_values))).to(x.device)
    rec = dict(xins=[], labels_in=[], labels_out=[], probs0=None, coefs=coefs, t_values=list(t_values), seed=seed, cond_cl=cond_cl,
               spatial=spatial, C=C)
    for i, t in enumerate(t_values):
        rec["xins"].append(xin.clone())
        rec["labels_in"].append(lab.clone())
        plan.inputs["t"].fill_(float(t))
        plan.run()
        if i == 0:
            rec["probs0"] = ops.cl_to_nchw(plan.outputs["head"], C, spatial, softmax=True).clone()
        out = torch.empty_like(lab)
        ops.cat_step_cl(plan.outputs["head"], lab, coefs[i:i + 1].contiguous(), out, 1, V, C, mode=ops.CAT_SAMPLE, cond=cond_cl,
                        n_cond=n_cond, next_x=xin, seed=seed, offset=i)
        lab = out
        rec["labels_out"].append(lab.clone())
    return rec


@torch.no_grad()
def slab_vs_unsplit(model, rec: dict, world: int, comm=None, transport: str = "copy") -> dict:
    """Runs the recorded chain through the slab plans (teacher-forced) and compares.  Returns, for the ranks this process
    holds: parity_max_abs (probabilities of step 0), bit_equal, the per-plane max-abs profile, per-step label agreement."""
    C, spatial = rec["C"], rec["spatial"]
    ses = SlabSession(model.unet, spatial, world, comm, transport)
    try:
        Dl = ses.local_sp[0]
        res = dict(world=world, ranks=list(ses.ranks), parity_max_abs=0.0, bit_equal=True, plane_max_abs=[], agree=[])
        for i, t in enumerate(rec["t_values"]):
            ses.load_input(rec["xins"][i], t)
            ses.run()
            if i == 0:
                for p, r in zip(ses.probs(C), ses.ranks):
                    lo, hi = ses.ranges[r]
                    d = (p - rec["probs0"][:, :, lo:hi]).abs()
                    res["parity_max_abs"] = max(res["parity_max_abs"], float(d.max()))
                    res["bit_equal"] = res["bit_equal"] and bool(float(d.max()) == 0.0)
                    res["plane_max_abs"].append((r, d.amax((0, 1, 3, 4)).tolist()))
            outs = ses.draw(rec["labels_in"][i], rec["coefs"][i:i + 1].contiguous(), C, rec["seed"], i)
            same, tot = 0, 0
            for o, r in zip(outs, ses.ranks):
                want = rec["labels_out"][i].view(-1)[r * ses.Vl:(r + 1) * ses.Vl]
                same += int((o == want).sum())
                tot += o.numel()
            res["agree"].append(same / tot)
        res["planes_per_rank"] = Dl
        if comm is not None:
            res["halo_exchanges_per_forward"] = comm.n_exchanges / len(rec["t_values"])
            res["gathers_per_forward"] = comm.n_gathers / len(rec["t_values"])
        return res
    finally:
        ses.close()


def reduce_over_ranks(res: dict, device) -> dict:
    """Worst case over the ranks of a real multi-GPU run (all ranks call this)."""
    import torch.distributed as dist
    mx = torch.tensor([res["parity_max_abs"], 0.0 if res["bit_equal"] else 1.0] + [-a for a in res["agree"]], device=device,
                      dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    out = dict(res)
    out["parity_max_abs"], out["bit_equal"] = float(mx[0]), bool(float(mx[1]) == 0.0)
    out["agree"] = [-float(v) for v in mx[2:]]
    return out
