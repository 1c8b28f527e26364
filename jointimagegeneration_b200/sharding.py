"""Multi-GPU partitioning of the samplers (one process per GPU, torch.distributed).

Independent chains (BASELINE configs 2-4) shard naturally: nothing in a denoising step couples
batch elements (GroupNorm is per sample), so the batch is split over the ranks with NO data-path
collective; the only communication is one gather of the final label volumes / slices
(SURVEY.md section 8e).  This module is pure host logic and runs identically over ``gloo`` (CPU
tests) and ``nccl`` (B200).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous near-equal split of n units: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: Optional[torch.Tensor], rank: int, world: int) -> Optional[torch.Tensor]:
    if t is None:
        return None
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def chain_seeds(base_seed: int, n: int) -> List[int]:
    """Per-chain Philox seeds: a chain's noise depends on its GLOBAL index only, so results do not
    change with the number of ranks."""
    return [base_seed * 1000003 + i for i in range(n)]


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather variable-length batch shards back into [total, ...] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    assert sizes[rank] == local.shape[0]
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)


def slab_ranges(depth: int, world: int, multiple: int = 16) -> List[Tuple[int, int]]:
    """Depth slabs for a single large volume (config 5): `world` EQUAL contiguous slabs, each a multiple of `multiple`
    planes (the UNet halves D four times).  Equal because the GroupNorm combine (S * R), the key/value gather
    (all_gather_into_tensor) and the global voxel index of the Philox counter (rank * V) all assume rank r holds
    planes [r D/R, (r+1) D/R); an uneven split would hang or give silently wrong statistics, so it is refused."""
    if depth % (multiple * world) != 0:
        raise ValueError(f"depth {depth} does not split into {world} equal slabs that are multiples of {multiple} planes")
    per = depth // world
    return [(r * per, (r + 1) * per) for r in range(world)]


def global_minmax(x: torch.Tensor, group=None) -> Tuple[float, float]:
    """Slice-batch min/max across ranks (sample_diffusion.py:221-222 couples the samples of one
    sample_cond call through ds.min()/ds.max()): one 2-float all-reduce per slice."""
    mn = x.min().reshape(1).float()
    mx = x.max().reshape(1).float()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return float(mn), float(mx)


class SlabComm:
    """Communication of the depth-slab decomposition of ONE volume (config 5): rank r owns a contiguous
    range of depth planes of every activation.  Three couplings per forward:
    one-plane halos before every 3-tap-in-depth conv, GroupNorm statistics over the whole volume
    (all-gather of the per-rank partial sums, combined in a fixed order on every rank -> identical
    statistics everywhere), and keys/values of the whole volume at the attention sites.
    Works over any torch.distributed backend (nccl on B200, gloo in the CPU tests); world == 1
    degenerates to zero halos / copies, which tests the padded layout on a single GPU."""

    def __init__(self, group=None):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.bytes_sent = 0
        self.n_exchanges = 0
        self.n_gathers = 0

    def exchange_halo(self, t: torch.Tensor, lead: int, depth: int, need_lo: bool = True, need_hi: bool = True):
        """t: [1, lead + depth + trail, H, W, C]; fills plane lead-1 from the previous rank's last interior
        plane and plane lead+depth from the next rank's first one (zeros at the ends of the volume)."""
        r, R = self.rank, self.world
        lo_halo, hi_halo = t[0, lead - 1], t[0, lead + depth]
        first, last = t[0, lead], t[0, lead + depth - 1]
        p2p = []
        if need_lo:
            if r > 0:
                p2p.append(dist.P2POp(dist.irecv, lo_halo, self._peer(r - 1), self.group))
            else:
                lo_halo.zero_()
            if r < R - 1:
                p2p.append(dist.P2POp(dist.isend, last, self._peer(r + 1), self.group))
        if need_hi:
            if r < R - 1:
                p2p.append(dist.P2POp(dist.irecv, hi_halo, self._peer(r + 1), self.group))
            else:
                hi_halo.zero_()
            if r > 0:
                p2p.append(dist.P2POp(dist.isend, first, self._peer(r - 1), self.group))
        self.n_exchanges += 1
        if p2p:
            self.bytes_sent += sum(op.tensor.numel() * op.tensor.element_size() for op in p2p if op.op is dist.isend)
            for w in dist.batch_isend_irecv(p2p):
                w.wait()

    def _peer(self, group_rank: int) -> int:
        return group_rank if self.group is None else dist.get_global_rank(self.group, group_rank)

    def broadcast_int(self, value: int) -> int:
        """rank 0's value on every rank (host-side scalar: the per-call Philox key of the sampler loop)."""
        if self.world == 1:
            return int(value)
        box = [int(value)]
        dist.broadcast_object_list(box, src=self._peer(0), group=self.group)
        return int(box[0])

    def all_gather(self, out: torch.Tensor, inp: torch.Tensor):
        """out (flattened) = concatenation over ranks of inp (flattened)."""
        self.n_gathers += 1
        if self.world == 1:
            out.view(-1).copy_(inp.reshape(-1))
            return
        dist.all_gather_into_tensor(out.view(-1), inp.reshape(-1), group=self.group)


class _LocalComm:
    """One virtual rank of a LocalSlabGroup: carries (rank, world) and the counters of SlabComm; its collectives are
    executed by LocalSlabGroup.run, which sees the buffers of every rank."""

    def __init__(self, group, rank: int):
        self.group_obj, self.rank, self.world = group, rank, group.world
        self.bytes_sent = self.n_exchanges = self.n_gathers = 0

    def exchange_halo(self, *a, **k):
        raise RuntimeError("virtual rank: run the plans of all ranks through LocalSlabGroup.run")

    def all_gather(self, *a, **k):
        raise RuntimeError("virtual rank: run the plans of all ranks through LocalSlabGroup.run")

    def broadcast_int(self, value: int) -> int:
        return int(value)


class LocalSlabGroup:
    """R VIRTUAL ranks of the depth-slab decomposition on ONE device: the R plans (built with comms[r] as their
    SlabComm) are executed in lock step on the current stream and every collective step is carried out directly on
    the R sets of buffers.  The plans, kernels, halo layout, GroupNorm combine, K/V gather and voxel indexing are
    exactly those of the multi-GPU run -- only the transport differs -- so slab parity at world 2/4/8 can be checked
    on a single GPU (tests/test_gpu_models.py) and a volume larger than one plan's arena can be walked slab by slab."""

    def __init__(self, world: int):
        self.world = world
        self.comms = [_LocalComm(self, r) for r in range(world)]

    def run(self, plans):
        from . import _C
        R = self.world
        assert len(plans) == R and len({len(p.steps) for p in plans}) == 1, "virtual ranks must hold identical plans"
        s = _C.stream()
        for i in range(len(plans[0].steps)):
            fn0 = plans[0].steps[i][0]
            if not hasattr(fn0, "fn"):
                for p in plans:
                    fn, args = p.steps[i]
                    st = fn(*args, s)
                    if st != 0:
                        _C.check(st, fn.__name__)
                continue
            name = fn0.fn.__name__
            argv = [p.steps[i][1] for p in plans]
            assert all(p.steps[i][0].fn.__name__ == name for p in plans)
            if name == "exchange_halo":
                self._exchange(argv)
            elif name == "all_gather":
                self._gather(argv)
            else:
                raise NotImplementedError(name)

    def _exchange(self, argv):
        R = self.world
        for r in range(R):
            t, lead, depth, need_lo, need_hi = argv[r]
            c = self.comms[r]
            c.n_exchanges += 1
            if need_lo:
                if r > 0:
                    tp, lp, dp = argv[r - 1][:3]
                    t[0, lead - 1].copy_(tp[0, lp + dp - 1])
                    c.bytes_sent += t[0, lead].numel() * t.element_size()
                else:
                    t[0, lead - 1].zero_()
            if need_hi:
                if r < R - 1:
                    tn, ln = argv[r + 1][:2]
                    t[0, lead + depth].copy_(tn[0, ln])
                    c.bytes_sent += t[0, lead].numel() * t.element_size()
                else:
                    t[0, lead + depth].zero_()

    def _gather(self, argv):
        R = self.world
        for r in range(R):
            out = argv[r][0].view(-1)
            n = argv[r][1].numel()
            self.comms[r].n_gathers += 1
            for q in range(R):
                out[q * n:(q + 1) * n].copy_(argv[q][1].reshape(-1))
