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


# ================================================================ depth-slab collectives over NVLink peer memory
class PeerBuf:
    """A region of this rank's peer-visible communication arena (quacks like a tensor for _C.ptr)."""

    def __init__(self, ptr: int, nbytes: int, offset: int):
        self.ptr, self.nbytes, self.offset = ptr, nbytes, offset

    def data_ptr(self) -> int:
        return self.ptr


class PeerArena:
    """This rank's peer-visible arena (cudaMalloc through the C ABI so that cudaIpc can export it) and the base addresses
    of every peer's arena as mapped into this process.  The layout is symmetric: offset X means the same slot everywhere."""

    def __init__(self, nbytes: int):
        import ctypes as C
        from . import _C
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _C.check(_C.lib().gg_peer_alloc(self.nbytes, C.byref(p)), "gg_peer_alloc")
        self.base = int(p.value)
        self.used = 0
        self.peer_base = None
        self._opened = []

    def export_handle(self) -> bytes:
        import ctypes as C
        from . import _C
        h = C.create_string_buffer(64)
        _C.check(_C.lib().gg_peer_export(C.c_void_p(self.base), h), "gg_peer_export")
        return h.raw

    def open_peers(self, handles, rank: int):
        import ctypes as C
        from . import _C
        self.peer_base = []
        for r, h in enumerate(handles):
            if r == rank:
                self.peer_base.append(self.base)
                continue
            p = C.c_void_p()
            _C.check(_C.lib().gg_peer_open(C.create_string_buffer(h, 64), C.byref(p)), "gg_peer_open")
            self.peer_base.append(int(p.value))
            self._opened.append(int(p.value))

    def alloc(self, nbytes: int, align: int = 256) -> PeerBuf:
        off = (self.used + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise MemoryError(f"peer communication arena exhausted ({self.nbytes} bytes): pass a larger arena_bytes to PeerSlabComm")
        self.used = off + nbytes
        return PeerBuf(self.base + off, nbytes, off)

    def close(self):
        import ctypes as C
        from . import _C
        for p in self._opened:
            _C.lib().gg_peer_close(C.c_void_p(p))
        self._opened = []
        if self.base:
            _C.lib().gg_peer_free(C.c_void_p(self.base))
            self.base = 0


class PeerSlabComm:
    """SlabComm whose collectives are gg_peer_exchange kernels over NVLink peer memory (csrc/peer_comm.cu) instead of
    host-enqueued NCCL calls: planned as ordinary library launches, so the slab forward is CUDA-graph capturable and a
    GroupNorm combine costs a few microseconds instead of an NCCL all-gather.  torch.distributed is used once, at
    construction, to exchange the 64-byte cudaIpc handles of the arenas.

    ``peers``: None = one process per rank (the real thing); or a list of PeerArena of the OTHER virtual ranks living in
    this process (tests on one GPU: LocalPeerGroup)."""

    transport = "nvlink-peer"
    peer = True

    def __init__(self, group=None, arena_bytes: int = 256 << 20, _virtual=None):
        import torch
        self.group = group
        self.arena = PeerArena(arena_bytes)
        if _virtual is not None:
            self.rank, self.world = _virtual
        else:
            self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, self.arena.export_handle(), group=group)
            else:
                handles = [b""]
            self.arena.open_peers(handles, self.rank)
        assert self.world <= 8
        self.bytes_sent = self.n_exchanges = self.n_gathers = 0
        self.n_forwards = 0
        self.epoch = self.arena.alloc(256)            # [0]: epoch counter, [64]: last-CTA counter
        self._sites = 0
        self._keep = []
        self._torch = torch

    # -- planning helpers ---------------------------------------------------------------------------------------
    def _peer_addr(self, r: int, buf: PeerBuf, extra: int = 0) -> int:
        return self.arena.peer_base[r] + buf.offset + extra

    def _new_args(self):
        from . import _C
        a = _C.PeerXchgArgs()
        a.epoch, a.done_counter, a.phase, a.ctas = self.epoch.ptr, self.epoch.ptr + 64, 3, 0
        self._keep.append(a)
        return a

    def alloc(self, nbytes: int) -> PeerBuf:
        """Peer-visible buffer (the destination of an all-gather must live in the arena)."""
        return self.arena.alloc(int(nbytes))

    def plan_begin(self, plan):
        """First step of a slab plan: one epoch per forward."""
        from . import _C
        plan.add(_C.lib().gg_peer_epoch_inc, self.epoch.ptr)
        self.n_forwards_planned = getattr(self, "n_forwards_planned", 0) + 1

    def plan_exchange_halo(self, plan, t, lead: int, depth: int, need_lo: bool = True, need_hi: bool = True, split: bool = False):
        """t: [1, lead + depth + trail, H, W, C] local activation.  My last interior plane -> next rank's low-halo staging slot,
        my first -> previous rank's high-halo slot; then wait for my neighbours' planes and copy them from the staging slots
        into t's halo planes (zeros at the ends of the volume)."""
        import ctypes as C
        from . import _C
        r, R = self.rank, self.world
        pb = t.shape[2] * t.shape[3] * t.shape[4] * t.element_size()
        base = t.data_ptr()
        stage = self.arena.alloc(2 * pb)              # [0]: plane arriving from rank r - 1 (low halo), [1]: from rank r + 1
        flags = self.arena.alloc(64)                  # [0]: raised by rank r - 1, [1]: by rank r + 1
        a = self._new_args()
        ns = nfo = nfi = nc = nz = 0
        if need_lo:
            if r < R - 1:                             # the next rank needs my last plane as ITS low halo
                a.src[ns], a.dst[ns], a.bytes[ns] = base + (lead + depth - 1) * pb, self._peer_addr(r + 1, stage, 0), pb
                ns += 1
                a.flag_out[nfo] = self._peer_addr(r + 1, flags, 0)
                nfo += 1
                self.bytes_sent += pb
            if r > 0:
                a.flag_in[nfi] = flags.ptr
                nfi += 1
                a.csrc[nc], a.cdst[nc], a.cbytes[nc] = stage.ptr, base + (lead - 1) * pb, pb
                nc += 1
            else:
                a.zdst[nz], a.zbytes[nz] = base + (lead - 1) * pb, pb
                nz += 1
        if need_hi:
            if r > 0:                                 # the previous rank needs my first plane as ITS high halo
                a.src[ns], a.dst[ns], a.bytes[ns] = base + lead * pb, self._peer_addr(r - 1, stage, pb), pb
                ns += 1
                a.flag_out[nfo] = self._peer_addr(r - 1, flags, 4)
                nfo += 1
                self.bytes_sent += pb
            if r < R - 1:
                a.flag_in[nfi] = flags.ptr + 4
                nfi += 1
                a.csrc[nc], a.cdst[nc], a.cbytes[nc] = stage.ptr + pb, base + (lead + depth) * pb, pb
                nc += 1
            else:
                a.zdst[nz], a.zbytes[nz] = base + (lead + depth) * pb, pb
                nz += 1
        a.nsend, a.nflag_out, a.nflag_in, a.ncopy, a.nzero = ns, nfo, nfi, nc, nz
        self.n_exchanges += 1
        if not split:
            plan.add(_C.lib().gg_peer_exchange, C.byref(a))
            plan.keep.append(a)
            return None
        # split: the push now, the wait + unpack later (returned closure) -- lets a GroupNorm combine of the same tensor share
        # the round trip (both pushes in flight before either wait)
        a1, a2 = _C.PeerXchgArgs.from_buffer_copy(a), _C.PeerXchgArgs.from_buffer_copy(a)
        a1.phase, a2.phase = 1, 2
        plan.add(_C.lib().gg_peer_exchange, C.byref(a1))
        plan.keep.extend([a1, a2])

        def finish():
            plan.add(_C.lib().gg_peer_exchange, C.byref(a2))
        return finish

    def plan_zero(self, plan, regions):
        """Local zero fills [(ptr, nbytes), ...] (halo planes beyond the ends of the volume) as one gg_peer_exchange launch
        without sends or waits; nothing is planned when there is nothing to clear."""
        import ctypes as C
        from . import _C
        regions = [(p, n) for p, n in regions if n > 0]
        if not regions:
            return
        a = self._new_args()
        a.phase = 2
        for i, (ptr, nbytes) in enumerate(regions):
            a.zdst[i], a.zbytes[i] = ptr, nbytes
        a.nzero = len(regions)
        plan.add(_C.lib().gg_peer_exchange, C.byref(a))
        plan.keep.append(a)

    def plan_all_gather(self, plan, out: PeerBuf, inp_ptr: int, nbytes: int):
        """out (in the arena, world * nbytes) = concatenation over ranks of the nbytes at inp_ptr."""
        import ctypes as C
        from . import _C
        r, R = self.rank, self.world
        assert out.nbytes >= R * nbytes and nbytes % 16 == 0
        flags = self.arena.alloc(64)                  # [q]: raised by rank q
        a = self._new_args()
        ns = nfo = nfi = 0
        for q in range(R):
            a.src[ns], a.dst[ns], a.bytes[ns] = inp_ptr, self._peer_addr(q, out, r * nbytes), nbytes
            ns += 1
            if q != r:
                a.flag_out[nfo] = self._peer_addr(q, flags, 4 * r)
                nfo += 1
                a.flag_in[nfi] = flags.ptr + 4 * q
                nfi += 1
                self.bytes_sent += nbytes
        a.nsend, a.nflag_out, a.nflag_in, a.ncopy, a.nzero = ns, nfo, nfi, 0, 0
        self.n_gathers += 1
        plan.add(_C.lib().gg_peer_exchange, C.byref(a))
        plan.keep.append(a)

    def attach_group_norm(self, fa, N: int, groups: int):
        """Turns a gg_gn_finalize call over this rank's partial rows into the whole-volume statistics: its push kernel stores
        16 bytes per (sample, group) into every rank's table, its combine kernel waits for the peers and adds the R entries
        in rank order (gg_gn_finalize_args.slab_*)."""
        r, R = self.rank, self.world
        tab = self.arena.alloc(R * N * groups * 16)
        flags = self.arena.alloc(64)
        fa.slab_world, fa.slab_rank, fa.slab_phase = R, r, 3
        for q in range(R):
            fa.slab_tables[q] = self._peer_addr(q, tab)
            if q != r:
                fa.slab_flag_out[q] = self._peer_addr(q, flags, 4 * r)
                fa.slab_flag_in[q] = flags.ptr + 4 * q
        fa.slab_epoch, fa.slab_done_counter = self.epoch.ptr, self.epoch.ptr + 64
        self.n_gathers += 1
        self.bytes_sent += (R - 1) * N * groups * 16

    def broadcast_int(self, value: int) -> int:
        if self.world == 1 or not dist.is_initialized():
            return int(value)
        box = [int(value)]
        dist.broadcast_object_list(box, src=0 if self.group is None else dist.get_global_rank(self.group, 0), group=self.group)
        return int(box[0])

    def close(self):
        self.arena.close()


class LocalPeerGroup:
    """R virtual ranks of the peer-memory transport on ONE device (tests): every rank has its own arena, all "peer" base
    addresses are plain local pointers, and the lock-step runner executes every gg_peer_exchange step as phase 1 of all
    ranks followed by phase 2 of all ranks (a single stream cannot interleave R spinning kernels)."""

    def __init__(self, world: int, arena_bytes: int = 128 << 20):
        self.world = world
        self.comms = [PeerSlabComm(arena_bytes=arena_bytes, _virtual=(r, world)) for r in range(world)]
        bases = [c.arena.base for c in self.comms]
        for c in self.comms:
            c.arena.peer_base = list(bases)

    @staticmethod
    def _collective_field(step):
        """None for a local step; else the name of the phase field of a step that talks to the other ranks."""
        fn, args = step
        name = getattr(fn, "__name__", "")
        if name == "gg_peer_exchange":
            a = args[0]._obj
            return "phase" if (a.nsend or a.nflag_out or a.nflag_in) else None
        if name == "gg_gn_finalize" and args[0]._obj.slab_world > 1:
            return "slab_phase"
        return None

    def run(self, plans):
        """Every rank runs its local steps up to its next collective step; the collective then runs as phase 1 of all ranks
        followed by phase 2 of all ranks (split steps -- a pure push or a pure wait -- run as they are).  Ranks may hold
        different numbers of LOCAL steps (edge ranks clear the halo planes beyond the volume); the collective steps match."""
        from . import _C
        s = _C.stream()
        cur = [0] * len(plans)
        while True:
            for r, p in enumerate(plans):               # local steps
                while cur[r] < len(p.steps) and self._collective_field(p.steps[cur[r]]) is None:
                    fn, args = p.steps[cur[r]]
                    _C.check(fn(*args, s), getattr(fn, "__name__", "step"))
                    cur[r] += 1
            done = [cur[r] >= len(p.steps) for r, p in enumerate(plans)]
            if all(done):
                return
            assert not any(done), "virtual ranks disagree on the number of collective steps"
            steps = [p.steps[cur[r]] for r, p in enumerate(plans)]
            fields = [self._collective_field(st) for st in steps]
            assert len(set(fields)) == 1 and len({getattr(st[0], "__name__", "") for st in steps}) == 1
            field = fields[0]
            phases = {getattr(st[1][0]._obj, field) for st in steps}
            assert len(phases) == 1, "virtual ranks disagree on the phase of a collective step"
            for phase in ((1, 2) if phases == {3} else tuple(phases)):
                for fn, args in steps:
                    a = args[0]._obj
                    old = getattr(a, field)
                    setattr(a, field, phase)
                    _C.check(fn(*args, s), getattr(fn, "__name__", "step"))
                    setattr(a, field, old)
            for r in range(len(plans)):
                cur[r] += 1

    def close(self):
        for c in self.comms:
            c.close()
