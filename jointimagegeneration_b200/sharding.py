"""Multi-GPU partitioning of the samplers (one process per GPU, torch.distributed).

Independent chains (BASELINE configs 2-4) shard naturally: nothing in a denoising step couples
batch elements (GroupNorm is per sample), so the batch is split over the ranks with NO data-path
collective; the only communication is one gather of the final label volumes / slices
(SURVEY.md section 8e).  This module is pure host logic and runs identically over ``gloo`` (CPU
tests) and ``nccl`` (B200).
"""
from typing import List, Optional, Sequence, Tuple

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
    """Depth slabs for a single large volume (config 5): contiguous planes, each slab a multiple of
    `multiple` planes (the UNet halves D four times) while planes remain."""
    units = depth // multiple
    assert units * multiple == depth, f"depth {depth} must be a multiple of {multiple}"
    out = []
    for r in range(world):
        lo, hi = shard_range(units, r, world)
        out.append((lo * multiple, hi * multiple))
    return out


def global_minmax(x: torch.Tensor, group=None) -> Tuple[float, float]:
    """Slice-batch min/max across ranks (sample_diffusion.py:221-222 couples the samples of one
    sample_cond call through ds.min()/ds.max()): one 2-float all-reduce per slice."""
    mn = x.min().reshape(1).float()
    mx = x.max().reshape(1).float()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return float(mn), float(mx)
