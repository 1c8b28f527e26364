"""UNetEngine: plans one denoiser forward as a static sequence of libguidegen_sm100 calls.

Replaces the torch library calls behind the reference's ``UNetModel.forward``
(ccdm/ddpm/models/unet_openai/unet.py:758-823, ldm/modules/diffusionmodules/openaimodel.py:713-745)
-- see SURVEY.md section 2.3 for the op-by-op map (K1..K15).

Data layout in HBM: every activation is channels-last bf16 ``[N, D, H, W, C]`` (2-D data: D = 1;
tokens: D = H = 1) so that a 64-channel slice of 128 neighbouring positions is one TMA box;
GroupNorm statistics, softmax, the timestep path and the head logits are fp32.  Weights are
repacked once per plan into the K-major bf16 matrices the tcgen05 kernel reads.

A plan is built per (batch, spatial shape, context shape): all buffers are allocated up front
from a small arena (liveness-based reuse), argument structs are pre-filled, and ``run()`` only
enqueues kernels -- no allocation, no host sync -- so a whole forward can be captured in a CUDA
graph (the timestep enters through a device tensor).
"""
import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _C, ops
from . import unet_modules as M


@dataclass
class Act:
    """An activation.  ``t`` is the FULL tensor [N, lead + D + trail, H, W, C]; lead/trail are halo
    planes along depth (depth-slab mode only, N == 1 there so the interior is contiguous)."""
    t: torch.Tensor          # CL bf16 (or fp32 for head logits)
    lead: int = 0
    trail: int = 0
    halo_valid: bool = False
    stats: Optional[tuple] = None      # (partial [N, n, C, 2], n): GroupNorm partial sums written by the producing conv

    @property
    def N(self): return self.t.shape[0]

    @property
    def sp(self): return (self.t.shape[1] - self.lead - self.trail, self.t.shape[2], self.t.shape[3])

    @property
    def C(self): return self.t.shape[-1]

    @property
    def S(self): return self.sp[0] * self.sp[1] * self.sp[2]

    @property
    def plane_bytes(self): return self.t.shape[2] * self.t.shape[3] * self.t.shape[4] * self.t.element_size()

    @property
    def ip(self) -> int:
        """device pointer of the interior (first non-halo plane)"""
        return self.t.data_ptr() + self.lead * self.plane_bytes

    @property
    def interior(self) -> torch.Tensor:
        return self.t[:, self.lead:self.t.shape[1] - self.trail]


class _Arena:
    """Liveness-based buffer reuse during plan construction (program order == stream order)."""

    def __init__(self, device):
        self.device = device
        self.free: List[torch.Tensor] = []
        self.owner: Dict[int, torch.Tensor] = {}
        self.stores: List[torch.Tensor] = []     # every backing allocation; the plan keeps them alive
        self.total = 0

    def alloc(self, shape, dtype) -> torch.Tensor:
        nbytes = int(math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        nbytes = max(256, (nbytes + 255) // 256 * 256)
        best = None
        for i, s in enumerate(self.free):
            if s.numel() >= nbytes and s.numel() <= 2 * nbytes + (1 << 20):
                if best is None or s.numel() < self.free[best].numel():
                    best = i
        if best is not None:
            store = self.free.pop(best)
        else:
            store = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.stores.append(store)
            self.total += nbytes
        n = int(math.prod(shape))
        es = torch.empty((), dtype=dtype).element_size()
        view = store[:n * es].view(dtype).view(shape)
        self.owner[view.data_ptr()] = store
        return view

    def release(self, t: Optional[torch.Tensor]):
        if t is None:
            return
        store = self.owner.pop(t.data_ptr(), None)
        if store is not None:
            self.free.append(store)


class _PyStep:
    def __init__(self, fn):
        self.fn = fn
        self.__name__ = getattr(fn, "__name__", "py_step")


class Plan:
    """A static launch list.  ``lanes``: independent step lists over disjoint sample ranges of the batch (one by default).
    With several lanes ``run()`` forks the current stream: lane 0 stays on it, every other lane is enqueued on its own side
    stream and joined at the end, so the many small launches of a low-resolution network (a few dozen CTAs, ~10 us each)
    overlap instead of queueing behind each other; under capture the lanes become parallel branches of ONE CUDA graph."""
    has_py = False

    def __init__(self):
        self.lanes = [[]]     # per lane: (cfunc, args tuple without the stream)
        self._cur = 0
        self.keep = []        # ctypes structs / tensors that must outlive the plan
        self.inputs: Dict[str, torch.Tensor] = {}
        self.outputs: Dict[str, torch.Tensor] = {}
        self.graph = None
        self.graph_body = None
        self.arena_bytes = 0
        self.flops = 0        # 2 * MACs actually issued by conv / attention launches
        # (gg_conv_args copy, gg_cat_epilogue) of the head conv with the sampler in its epilogue (CCDM resident loop), or None
        self.fused_head = None
        # the same per lane: [(gg_conv_args copy, gg_cat_epilogue, first sample of the lane)]
        self.fused_heads = []
        self._side = None     # side streams + fork / join events of a multi-lane plan (made on first use)

    # ---- construction
    def begin_lane(self, j: int):
        while len(self.lanes) <= j:
            self.lanes.append([])
        self._cur = j

    def add(self, fn, *args):
        self.lanes[self._cur].append((fn, args))

    def add_py(self, fn, *args):
        """A host-side step (collective / halo exchange): called as fn(*args), enqueues on the current stream."""
        self.lanes[self._cur].append((_PyStep(fn), args))
        self.has_py = True

    @property
    def steps(self):
        """All launches in lane order (per-launch timing, launch counts); THE list for a single-lane plan."""
        return self.lanes[0] if len(self.lanes) == 1 else [st for lane in self.lanes for st in lane]

    @property
    def body_steps(self):
        """Everything but each lane's last step (the head conv)."""
        return [st for lane in self.lanes for st in lane[:-1]]

    # ---- execution
    @staticmethod
    def _launch(fn, args, s):
        if isinstance(fn, _PyStep):
            fn.fn(*args)
            return
        st = fn(*args, s)
        if st != 0:
            _C.check(st, fn.__name__)

    def _enqueue(self, lanes):
        if len(lanes) == 1:
            s = _C.stream()
            for fn, args in lanes[0]:
                self._launch(fn, args, s)
            return
        assert not self.has_py, "host-side steps (NCCL slab transport) run on the current stream: single-lane plans only"
        main = torch.cuda.current_stream()
        if self._side is None:
            dev = main.device
            self._side = ([torch.cuda.Stream(device=dev) for _ in lanes[1:]], torch.cuda.Event(),
                          [torch.cuda.Event() for _ in lanes[1:]])
        side, fork, joins = self._side
        fork.record(main)
        handles = [main.cuda_stream]
        for s in side:
            s.wait_event(fork)
            handles.append(s.cuda_stream)
        for i in range(max(len(lane) for lane in lanes)):       # round robin: every stream is fed from the first launch on
            for j, lane in enumerate(lanes):
                if i < len(lane):
                    self._launch(lane[i][0], lane[i][1], handles[j])
        for s, ev in zip(side, joins):
            ev.record(s)
            main.wait_event(ev)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
            return
        self._enqueue(self.lanes)

    def run_body(self):
        """Everything but the last step of each lane (the head conv): the caller launches the head itself -- with the sampler
        epilogue whose per-step arguments (coefficients, Philox offset, label buffers) change from step to step."""
        if self.graph_body is not None:
            self.graph_body.replay()
            return
        self._enqueue([lane[:-1] for lane in self.lanes])

    def capture_body(self):
        if self.has_py or self.graph_body is not None:
            return
        self.run_body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue([lane[:-1] for lane in self.lanes])
        self.graph_body = g

    def capture(self):
        """Capture the forward into a CUDA graph (after one eager warm-up run).  Plans with host-side collectives
        (depth-slab mode over NCCL) stay eager: capturing the NCCL ops works, but measured no gain at 2 GPUs and mixing
        captured and eager collectives on one communicator dead-locked at teardown (round-1 experiment)."""
        if self.has_py:
            return
        self.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue(self.lanes)
        self.graph = g

    @property
    def num_launches(self):
        return sum(len(lane) for lane in self.lanes)


class UNetEngine:
    """kind = 'ccdm' (input_condition concat, optional softmax head, dict output) or 'ldm'."""

    def __init__(self, model: torch.nn.Module, dims: int, num_heads: int, num_head_channels: int, fused_upsample: bool = True):
        self.model = model
        self.dims = dims
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.fused_upsample = fused_upsample
        self.use_halo_conv = True
        self.use_roll_conv = True
        # one-launch GroupNorm (gg_gn_fused, shared-memory-resident form: the slice of every CTA is loaded once by bulk async
        # copies, statistics and normalisation read shared memory).  Measured inside CUDA graphs (tools/bench_gn.py,
        # profiles/r2_lanes_and_pdl.md): the three launches cost only 9-16 us per GroupNorm there, the one launch 10-13 us with
        # clusters of 2 / 4 CTAs per sample and 23-30 us with clusters of 8 (16 samples x 8 CTAs x ~200 KB do not become
        # resident together).  So: 1 = where the plan is a cluster of 2 or 4 (default), 2 = wherever a sample fits, 0 = never.
        self.fused_small_gn = int(os.environ.get("GG_FUSED_SMALL_GN", "1"))
        self.separate_skip = os.environ.get("GG_SEPARATE_SKIP", "0") != "0"     # see _resblock (measured: no gain, off)
        # GroupNorm + SiLU applied inside the depth-rolling conv (no separate pass); GG_FUSED_GN=0 is a tuning knob
        self.fused_gn_apply = os.environ.get("GG_FUSED_GN", "1") != "0"
        # ... and inside the halo-brick conv (every other stride-1 3^d conv on grids >= 16 x 8: the LDM networks, the deeper
        # CCDM levels): the same transform (eight warps) on its landed input windows, bit-identical to gg_gn_apply + conv
        # (tests/test_gpu_kernels.py::test_conv_halo_fused_groupnorm).  That kernel loads a window once per depth tap and per
        # N tile, so it re-normalises what gg_gn_apply normalises once; while every "window transformed" signal of a CTA pair
        # carried a GPU-scope fence (halo_common.cuh::mbar_arrive_remote, round 2) that cost more than the saved pass
        # (config 2 +1.6 ms).  Without the fence it wins: config 2 43.3 -> 42.6 ms, config 3 4.69 -> 4.44, config 4 10.8 -> 10.2
        # (tools/run_ab.sh), so ON by default; GG_FUSED_GN_HALO=0 restores the separate pass
        self.fused_gn_halo = os.environ.get("GG_FUSED_GN_HALO", "1") != "0"
        self.use_split_k = True
        # split-K shape: at most splitk_max K ranges per tile, at least splitk_min_kb 64-wide K blocks per range (tuning knobs)
        self.splitk_max = int(os.environ.get("GG_SPLITK_MAX", "16"))
        self.splitk_min_kb = int(os.environ.get("GG_SPLITK_MIN_KB", "6"))
        # in-kernel split-K reduction (the last split of a tile sums the partials itself, gg_conv_args.split_counters) instead of
        # the second launch: bit-identical, but only that CTA's 128 epilogue threads do the summing -- measured SLOWER
        # (config 3: 8.5 vs 4.8 ms per step), so off; kept as a tested knob
        self.splitk_fixup = os.environ.get("GG_SPLITK_FIXUP", "0") != "0"
        self.num_sms = torch.cuda.get_device_properties(next(model.parameters()).device).multi_processor_count \
            if next(model.parameters()).is_cuda else 148
        self.halo_gn_stats = os.environ.get("GG_HALO_STATS", "1") != "0"    # ... except in the halo / roll kernels (shuffle-reduced per tile)
        # per sample; below it the separate gg_gn_partial pass stays (re-measured at the end of round 2, tools/run_aq.sh: config 3
        # 4.08 ms with 8192, 4.21 with 4096, 4.44 with 1024; config 4 9.93 / 10.04 / 10.11)
        self.halo_stats_min_positions = int(os.environ.get("GG_HALO_STATS_MIN", "8192"))
        self.fused_gn_stats = False     # GroupNorm statistics from the conv epilogue: correct, but the extra epilogue
                                        # work costs more than the separate (cached) statistics pass saves on B200
        self.slab = None            # sharding.SlabComm: depth-slab decomposition of ONE volume over ranks
        self.plans: Dict[tuple, Plan] = {}
        self._wcache: Dict[tuple, torch.Tensor] = {}
        self.lib = None

    def invalidate(self):
        """Drop packed weights and plans (call after the parameters change)."""
        self.plans.clear()
        self._wcache.clear()

    # ----------------------------------------------------------------------------- helpers
    def _dev(self):
        return next(self.model.parameters()).device

    def _cached(self, key, make):
        if key not in self._wcache:
            with torch.no_grad():
                self._wcache[key] = make()
        return self._wcache[key]

    def _f32(self, p):
        return self._cached((id(p), "f32"), lambda: p.detach().float().contiguous())

    def _vec8(self, key, make, n):
        return self._cached(key, lambda: ops.pad_vec(make(), n))

    # halo planes of every activation in depth-slab mode: [spare, halo_lo | interior | halo_hi]
    # (the spare plane lets the stride-2 parity maps start one plane early, see _conv)
    @property
    def _lead(self): return 2 if self.slab is not None else 0

    @property
    def _trail(self): return 1 if self.slab is not None else 0

    def _new_act(self, ar, N, sp, Cc, dtype=torch.bfloat16) -> Act:
        lead, trail = self._lead, self._trail
        assert lead == 0 or N == 1, "depth-slab mode handles one volume (N == 1)"
        t = ar.alloc((N, sp[0] + lead + trail, sp[1], sp[2], Cc), dtype)
        return Act(t, lead, trail)

    def _exchange(self, plan, x: Act, need_lo=True, need_hi=True):
        if self.slab is None or x.lead == 0 or x.halo_valid:
            return
        if getattr(self.slab, "peer", False):       # one kernel over NVLink peer memory (csrc/peer_comm.cu), graph-capturable
            self.slab.plan_exchange_halo(plan, x.t, x.lead, x.sp[0], need_lo, need_hi)
        else:
            plan.add_py(self.slab.exchange_halo, x.t, x.lead, x.sp[0], need_lo, need_hi)
        x.halo_valid = need_lo and need_hi

    def _halo_overlap(self, plan, xs, need_lo=True, need_hi=True):
        """For _gn_scale_shift(overlap=...): plans the halo pushes of the activations `xs` now and returns the planner of the
        matching waits; marks the halos valid.  None when there is nothing to overlap (no peer-memory slab mode)."""
        if self.slab is None or not getattr(self.slab, "peer", False) or self.slab.world == 1:
            return None
        todo = [x for x in xs if x.lead and not x.halo_valid]

        def push():
            fins = [self.slab.plan_exchange_halo(plan, x.t, x.lead, x.sp[0], need_lo, need_hi, split=True) for x in todo]
            for x in todo:
                x.halo_valid = need_lo and need_hi

            def finish():
                for f in fins:
                    f()
            return finish
        return push

    # --------------------------------------------------------------------------- primitives
    def _gn_scale_shift(self, plan: Plan, ar: _Arena, x1: Act, x2: Optional[Act], norm: M.ParamNorm, overlap=None) -> torch.Tensor:
        """Per-(sample, channel) (scale, shift) of GroupNorm over cat([x1, x2], channel): fp32 [N, C1 + C2, 2].
        overlap (peer-memory slab mode): a callable that plans further pushes (the halo planes of the same tensors) and returns
        the callable planning their waits -- they are placed between this GroupNorm's push and its combine, so that both
        exchanges share one NVLink round trip."""
        lib = self.lib
        N, S = x1.N, x1.S
        C1, C2 = x1.C, (x2.C if x2 is not None else 0)
        R = self.slab.world if self.slab is not None else 1

        temps = []

        def partial(x, Cx):
            if x.stats is not None:            # written by the producing conv's epilogue
                p, n = x.stats
            else:
                n = ops.gn_num_chunks(S, Cx)
                p = ar.alloc((N, n, Cx, 2), torch.float32)
                plan.add(lib.gg_gn_partial, x.ip, N, S, Cx, _C.ptr(p))
                x.stats = (p, n)               # skip tensors are normalised twice: keep the sums with the activation
            if R > 1 and not getattr(self.slab, "peer", False):
                # every rank needs the statistics of the whole volume: gather the partial rows (NCCL transport; the peer-memory
                # transport exchanges 16 bytes per group inside gg_gn_finalize instead, see below)
                g = ar.alloc((N, R * n, Cx, 2), torch.float32)
                plan.add_py(self.slab.all_gather, g, p)
                temps.append(g)
                return g, R * n
            return p, n

        p1, n1 = partial(x1, C1)
        p2, n2 = partial(x2, C2) if x2 is not None else (None, 0)
        ss = ar.alloc((N, C1 + C2, 2), torch.float32)
        fa = _C.GnFinalizeArgs(_C.ptr(p1), C1, n1, _C.ptr(p2), C2, n2, _C.ptr(self._f32(norm.weight)),
                               _C.ptr(self._f32(norm.bias)), _C.ptr(ss), N, norm.groups, S * R, float(norm.eps))
        if R > 1 and getattr(self.slab, "peer", False):
            self.slab.attach_group_norm(fa, N, norm.groups)
            if overlap is not None:
                f1, f2 = _C.GnFinalizeArgs.from_buffer_copy(fa), _C.GnFinalizeArgs.from_buffer_copy(fa)
                f1.slab_phase, f2.slab_phase = 1, 2
                plan.keep.extend([f1, f2])
                plan.add(lib.gg_gn_finalize, C.byref(f1))      # push my group sums
                finish = overlap()                                # push my boundary planes
                plan.add(lib.gg_gn_finalize, C.byref(f2))      # wait for the peers' sums, combine
                finish()                                          # wait for the peers' planes, unpack
                for tbuf in temps:
                    ar.release(tbuf)
                return ss
        elif overlap is not None:
            overlap()()
        plan.keep.append(fa)
        plan.add(lib.gg_gn_finalize, C.byref(fa))
        for tbuf in temps:
            ar.release(tbuf)
        return ss

    def _gn(self, plan: Plan, ar: _Arena, x1: Act, x2: Optional[Act], norm: M.ParamNorm, silu: bool, with_halo: bool = False) -> Act:
        """GroupNorm (+SiLU) over cat([x1, x2], channel) as a materialised tensor.  with_halo: the consumer is a conv with three
        depth taps (depth-slab mode: the result must carry valid halo planes)."""
        C1, C2 = x1.C, (x2.C if x2 is not None else 0)
        if (self.fused_small_gn and (self.slab is None or self.slab.world == 1) and C1 + C2 <= 2048
                and x1.stats is None and (x2 is None or x2.stats is None)
                and int(self.lib.gg_gn_fused_resident(x1.S, C1 + C2)) in ((2, 4) if int(self.fused_small_gn) == 1 else (1, 2, 4, 8))):
            # the sample fits the shared memory of one cluster: statistics + apply in one launch instead of three dependent ones
            y = self._new_act(ar, x1.N, x1.sp, C1 + C2)
            plan.add(self.lib.gg_gn_fused, x1.ip, C1, x2.ip if x2 is not None else 0, C2, _C.ptr(self._f32(norm.weight)),
                     _C.ptr(self._f32(norm.bias)), y.ip, x1.N, x1.S, norm.groups, float(norm.eps), int(silu))
            return y
        xs = [x1] + ([x2] if x2 is not None else [])
        if with_halo and self.slab is not None and getattr(self.slab, "peer", False) and self.slab.world > 1 and x1.lead:
            # depth slabs over peer memory: exchange the RAW boundary planes while the group sums travel (one round trip for
            # both), then normalise the interior together with the halo planes that hold a neighbour's data; the planes beyond
            # the ends of the volume stay zero (the reference pads the NORMALISED tensor)
            ss = self._gn_scale_shift(plan, ar, x1, x2, norm, overlap=self._halo_overlap(plan, xs))
            y = self._new_act(ar, x1.N, x1.sp, C1 + C2)
            lo = 1 if self.slab.rank > 0 else 0
            hi = 1 if self.slab.rank < self.slab.world - 1 else 0
            pp = x1.sp[1] * x1.sp[2]                       # positions per plane
            plan.add(self.lib.gg_gn_apply, x1.ip - lo * x1.plane_bytes, C1, (x2.ip - lo * x2.plane_bytes) if x2 is not None else 0, C2,
                     _C.ptr(ss), y.ip - lo * y.plane_bytes, x1.N, x1.S + (lo + hi) * pp, int(silu))
            self.slab.plan_zero(plan, [(y.ip - y.plane_bytes, 0 if lo else y.plane_bytes),
                                       (y.ip + x1.sp[0] * y.plane_bytes, 0 if hi else y.plane_bytes)])
            y.halo_valid = True
            ar.release(ss)
            return y
        ss = self._gn_scale_shift(plan, ar, x1, x2, norm)
        y = self._new_act(ar, x1.N, x1.sp, C1 + C2)
        plan.add(self.lib.gg_gn_apply, x1.ip, C1, x2.ip if x2 is not None else 0, C2, _C.ptr(ss), y.ip, x1.N, x1.S, int(silu))
        ar.release(ss)
        return y

    def _gn_conv(self, plan: Plan, ar: _Arena, x1: Act, x2: Optional[Act], norm: M.ParamNorm, silu: bool, packer, cout: int, *,
                 dims: int, extra_srcs=(), **kw) -> Act:
        """conv(silu?(GroupNorm(cat([x1, x2])))) [+ un-normalised centre-only sources].  `packer(splits)` returns the
        packed-weight callable for the given channel split of the normalised input.  When the depth-rolling kernel
        takes the conv, the normalisation is applied to the halo planes inside it (no gg_gn_apply pass, no
        normalised tensor); otherwise the tensor is materialised first."""
        C1, C2 = x1.C, (x2.C if x2 is not None else 0)
        fuse = (self.fused_gn_apply and self._roll_ok(dims, 1, None, None, cout, x1.sp, None)
                and (C1 + C2) * 8 <= 4096 and C1 % 64 == 0 and C2 % 64 == 0)
        if not fuse and self.fused_gn_halo and kw.get("ksize", 3) == 3 and kw.get("stride", 1) == 1 and kw.get("taps") is None:
            # the halo-brick kernel normalises its landed windows itself (any channel count: the table is read per chunk)
            fuse = self._halo_ok((3 if dims >= 3 else 1, 3 if dims >= 2 else 1, 3), 1, x1.sp)
        taps3 = dims >= 3 and kw.get("ksize", 3) == 3 and kw.get("stride", 1) == 1 and kw.get("taps") is None
        if not fuse:
            a = self._gn(plan, ar, x1, x2, norm, silu, with_halo=taps3)
            out = self._conv(plan, ar, [(a, False)] + list(extra_srcs), packer([a.C]), cout, dims=dims, **kw)
            self._free(ar, a)
            return out
        xs = [x1] + ([x2] if x2 is not None else [])
        ss = self._gn_scale_shift(plan, ar, x1, x2, norm, overlap=self._halo_overlap(plan, xs) if taps3 else None)
        ptrs, off = [], 0
        for x in xs:
            ptrs.append(_C.ptr(ss) + 8 * off)
            off += x.C
        out = self._conv(plan, ar, [(x, False) for x in xs] + list(extra_srcs), packer([x.C for x in xs]), cout, dims=dims,
                         src_ss=ptrs + [None] * len(extra_srcs), ss_stride=2 * (C1 + C2), xf_silu=silu, **kw)
        plan.keep.append(ss)
        ar.release(ss)          # stream order: the conv that reads it is enqueued before any later writer
        return out

    def _roll_ok(self, dims, stride, taps, offsets, cout, out_spatial, y_strides) -> bool:
        """The depth-rolling kernel (conv_roll.cu): 3x3x3 stride-1 filters with a narrow output on grids that fill a
        2 x 2 group of 16 x 8 bricks."""
        return (self.use_halo_conv and self.use_roll_conv and dims == 3 and stride == 1
                and (taps is None or tuple(taps) == (3, 3, 3)) and (offsets is None or tuple(offsets) == (-1, -1, -1))
                and (cout + 15) // 16 * 16 in (16, 64) and out_spatial[1] >= 32 and out_spatial[2] >= 16 and y_strides is None)

    def _xf_z(self, D):
        """Local depth planes that hold real data for a fused-normalisation conv: the interior, plus the halo planes
        that carry a neighbouring slab's data (planes beyond the volume must stay zero: the reference pads the
        NORMALISED tensor)."""
        if self.slab is None:
            return (0, D)
        return (-1 if self.slab.rank > 0 else 0, D + 1 if self.slab.rank < self.slab.world - 1 else D)

    def _free(self, ar, a: Optional[Act]):
        if a is None:
            return
        ar.release(a.t)
        if a.stats is not None:
            ar.release(a.stats[0])

    def _halo_ok(self, taps, stride, out_spatial) -> bool:
        """The halo-brick kernel (conv_halo.cu) pays off for multi-tap stride-1 filters on grids that fill its
        16 x 8 output brick."""
        return (self.use_halo_conv and stride == 1 and taps[0] * taps[1] * taps[2] > 1 and out_spatial[1] >= 16
                and out_spatial[2] >= 8)

    def _conv(self, plan: Plan, ar: _Arena, srcs: List[Tuple[Act, bool]], w_packed, cout: int, *, dims: int,
              ksize: int = 3, stride: int = 1, bias=None, emb=None, emb_stride=0, residual: Optional[Act] = None,
              f32_out: bool = False, taps=None, offsets=None, y_ptr: Optional[int] = None, y_strides=None,
              out_spatial=None, out: Optional[Act] = None, stats: bool = False, stats_part=(0, 1),
              src_ss=None, ss_stride=0, xf_silu=True) -> Act:
        x0 = srcs[0][0]
        N, (D, H, W) = x0.N, x0.sp
        cout8 = (cout + 7) // 8 * 8
        if taps is None:
            taps = (ksize if dims >= 3 else 1, ksize if dims >= 2 else 1, ksize)
        if offsets is None:
            offsets = tuple(-(t // 2) for t in taps)
        if out_spatial is None:
            if stride == 1:
                out_spatial = (D, H, W)
            else:
                f = lambda n, on: (n - 1) // 2 + 1 if on else n
                out_spatial = (f(D, dims >= 3), f(H, dims >= 2), f(W, True))
        if out is None:
            out = self._new_act(ar, N, out_spatial, cout8, torch.float32 if f32_out else torch.bfloat16)
        halo = self._halo_ok(taps, stride, out_spatial) and callable(w_packed)
        # the halo kernels reduce GroupNorm column sums in their epilogue (bf16 outputs): one partial row per (CTA, warp),
        # i.e. up to 592 rows of C (sum, sum sq) pairs per sample -- less traffic than re-reading the tensor only when
        # a sample has well over 8 x 592 positions (the 64 x 64 latents of the LDM configs do not)
        halo_stats = halo and stats and self.halo_gn_stats and not f32_out and int(math.prod(out_spatial)) >= self.halo_stats_min_positions
        algo = 1 if halo and (halo_stats or not (stats and self.fused_gn_stats)) else 0
        if algo == 1 and self._roll_ok(dims, stride, taps, offsets, cout, out_spatial, y_strides):
            algo = 4        # narrow outputs: depth-rolling kernel (three depth taps stacked along N)
        assert src_ss is None or (algo in (1, 4) and callable(w_packed)), "fused input GroupNorm needs the halo-brick or depth-rolling kernel"
        if callable(w_packed):          # packed-weight K order depends on the kernel
            w_packed = w_packed(algo >= 1)
        else:
            algo = 0
        lead = x0.lead
        d_shift, kernel_out_sp, y_base = lead, tuple(out_spatial), (y_ptr if y_ptr is not None else out.ip)
        if lead and taps[0] > 1:
            for a, centre in srcs:
                if not centre:
                    self._exchange(plan, a, need_lo=True, need_hi=(stride == 1))
        if lead and stride == 2 and dims >= 3:
            # depth-slab stride 2: the padded buffer is [spare, halo_lo, interior...]; with the interior at
            # padded index d + 2 the reference's  2o + k - 1  becomes  2(o+1) + k - 1  in padded indices, i.e.
            # the ordinary strided conv over the WHOLE padded tensor whose output plane o' = o + 1.  Output
            # plane o' = 0 is garbage and lands in the output's halo_lo plane (overwritten by its next exchange).
            assert D % 2 == 0 and out.lead >= 1
            d_shift = 0
            kernel_out_sp = (out_spatial[0] + 1, out_spatial[1], out_spatial[2])
            y_base = out.ip - out.plane_bytes
        a = ops.make_conv_args([(s.t, c) for s, c in srcs], w_packed, cout, y_base, dims=dims, ksize=ksize, stride=stride,
                               bias=bias, emb=None, residual=residual.ip if residual is not None else None, taps=taps,
                               offsets=offsets, out_spatial=kernel_out_sp, y_strides=y_strides, d_shift=d_shift,
                               y_f32=f32_out, algo=algo, src_ss=src_ss, ss_stride=ss_stride, xf_silu=xf_silu,
                               xf_z=self._xf_z(D) if src_ss is not None else None)
        if emb is not None:
            a.emb = emb
            a.emb_stride = emb_stride
        if stats and (self.fused_gn_stats or halo_stats):
            # GroupNorm statistics of the output come out of this conv's epilogue (no separate pass over the
            # tensor); stats_part = (index, count) when several launches fill one output (folded upsample)
            per = int(self.lib.gg_conv_stats_chunks(C.byref(a)))
            if per > 0:
                idx, cnt = stats_part
                if out.stats is None:
                    out.stats = (ar.alloc((N, per * cnt, cout8, 2), torch.float32), per * cnt)
                a.gn_partial = _C.ptr(out.stats[0])
                a.gn_chunk_base, a.gn_nchunks_total = idx * per, per * cnt
                a.stats_d_min = 1 if kernel_out_sp != tuple(out_spatial) else 0
        kexp = ops.conv_packed_k(a)
        assert kexp == w_packed.shape[1], (kexp, tuple(w_packed.shape))
        if algo == 0 and self.use_split_k and a.gn_partial is None:
            # few output tiles, long reduction (low-resolution layers): spread K ranges over the idle SMs
            tiles, nkb = int(self.lib.gg_conv_num_tiles(C.byref(a))), kexp // ops.BLOCK_K
            if self.slab is not None:
                # the split factor fixes the summation order of the reduction: take the decision the UNSPLIT plan takes
                # for this layer (global depth, no halo planes), so that slab results differ from the unsplit ones only
                # through the order in which GroupNorm partial sums are combined
                g = _C.ConvArgs.from_buffer_copy(a)
                g.D, g.Do = D * self.slab.world, out_spatial[0] * self.slab.world
                tiles = int(self.lib.gg_conv_num_tiles(C.byref(g)))
            S = min(self.splitk_max, self.num_sms // max(tiles, 1), nkb // self.splitk_min_kb)
            if tiles * 2 <= self.num_sms and S >= 2:
                ws = ar.alloc((S, N * int(math.prod(kernel_out_sp)), cout8), torch.float32)
                a.split_k, a.workspace = S, _C.ptr(ws)
                if self.splitk_fixup:
                    # the last split of a tile to finish reduces it in the same launch: a few dedicated zeroed words per tile
                    # (not arena memory: nothing else may ever write them)
                    cnt = torch.zeros((int(self.lib.gg_conv_num_tiles(C.byref(a))),), dtype=torch.int32, device=ws.device)
                    a.split_counters = cnt.data_ptr()
                    plan.keep.append(cnt)
                plan.keep.append(ws)
                ar.release(ws)          # stream order: the reduce launch of this conv is done before any later kernel
        plan.keep.append(a)
        plan.add(self.lib.gg_conv_fwd, C.byref(a))
        plan.flops += 2 * N * int(math.prod(out_spatial)) * cout * w_packed.shape[1]
        return out

    def _pack(self, conv: M.ParamConv, splits, extra=(), chunk_major=False):
        key = (id(conv.weight), tuple(splits), tuple(id(e) for e in extra), chunk_major)

        def make():
            w = conv.weight.detach()
            missing = sum(splits) - w.shape[1]
            if missing > 0:      # activation channels were zero-padded to a multiple of 8 (e.g. 13 -> 16)
                assert len(splits) == 1
                w = torch.cat([w, w.new_zeros((w.shape[0], missing) + tuple(w.shape[2:]))], 1)
            return ops.pack_conv_weight(w, splits, extra=[e for e in extra], chunk_major=chunk_major)

        return self._cached(key, make)

    def _packer(self, conv, splits):
        return lambda cm: self._pack(conv, splits, chunk_major=cm)

    # ------------------------------------------------------------------------------ layers
    def _resblock(self, plan, ar, rb: M.ResBlock, x1: Act, x2: Optional[Act], emb_ptr: int, emb_stride: int) -> Act:
        dims = rb.dims
        cout = rb.out_channels
        c1 = rb.in_layers[2]
        h1 = self._gn_conv(plan, ar, x1, x2, rb.in_layers[0], True, lambda splits: self._packer(c1, splits), cout, dims=dims,
                           emb=emb_ptr, emb_stride=emb_stride, stats=True)
        c2 = rb.out_layers[3]
        if isinstance(rb.skip_connection, torch.nn.Identity):
            assert x2 is None and x1.C == cout
            b2 = self._vec8((id(c2.bias), "b"), lambda: c2.bias, cout)
            out = self._gn_conv(plan, ar, h1, None, rb.out_layers[0], True, lambda splits: self._packer(c2, splits), cout,
                                dims=dims, bias=_C.ptr(b2), residual=x1, stats=True)
        else:
            sk = rb.skip_connection
            if sk.kernel_size != 1:
                raise NotImplementedError("3x3 skip convolution (use_conv=True) is not used by any shipped config")
            xs = [x1] + ([x2] if x2 is not None else [])
            skw = sk.weight.detach().reshape(cout, -1)
            extras, c0 = [], 0
            for x in xs:
                extras.append(skw[:, c0:c0 + x.C])
                c0 += x.C
            key = (id(c2.weight), id(sk.weight), tuple(x.C for x in xs))

            def packer(splits):
                return lambda cm: self._cached(key + (tuple(splits), cm),
                                               lambda: ops.pack_conv_weight(c2.weight, list(splits), extra=extras, chunk_major=cm))

            b2 = self._vec8((id(c2.bias), id(sk.bias), "b"), lambda: c2.bias.detach() + sk.bias.detach(), cout)
            if self.separate_skip and self._roll_ok(dims, 1, None, None, cout, h1.sp, None):
                # experiment: in the depth-rolling kernel extra 1x1x1 sources take plane / weight ring slots every
                # step; a stand-alone (HBM-bound) 1x1x1 conv whose result enters conv2 as the residual measured
                # the same step time on B200 (46.4 vs 46.1 ms), so the fused form stays the default
                wsk = self._pack(sk, [x.C for x in xs])
                skip = self._conv(plan, ar, [(x, False) for x in xs], wsk, cout, dims=dims, ksize=1)
                out = self._gn_conv(plan, ar, h1, None, rb.out_layers[0], True, lambda splits: self._packer(c2, splits), cout,
                                    dims=dims, bias=_C.ptr(b2), residual=skip, stats=True)
                self._free(ar, skip)
            else:
                out = self._gn_conv(plan, ar, h1, None, rb.out_layers[0], True, packer, cout, dims=dims,
                                    extra_srcs=[(x, True) for x in xs], bias=_C.ptr(b2), stats=True)
        self._free(ar, h1)
        return out

    def _attn_workspace(self, plan, ar, aa):
        """V^T staging of the tensor-core attention kernel (gg_attn_args.workspace); released right after the launch is
        planned (stream order: nothing planned later can write it before this launch has read it)."""
        n = int(self.lib.gg_attention_workspace_bytes(C.byref(aa)))
        if n > 0:
            ws = ar.alloc((n,), torch.uint8)
            aa.workspace, aa.workspace_bytes = ws.data_ptr(), n
            plan.keep.append(ws)
            ar.release(ws)

    def _heads(self, ch):
        return self.num_heads if self.num_head_channels == -1 else ch // self.num_head_channels

    def _attention_block(self, plan, ar, ab: M.AttentionBlock, x: Act) -> Act:
        N, S, Cc = x.N, x.S, x.C
        H = ab.num_heads
        d = Cc // H
        R = self.slab.world if self.slab is not None else 1
        xn = self._gn(plan, ar, x, None, ab.norm, False)
        bq = self._vec8((id(ab.qkv.bias), "b"), lambda: ab.qkv.bias, 3 * Cc)
        qkv = self._conv(plan, ar, [(xn, False)], self._pack(ab.qkv, [Cc]), 3 * Cc, dims=3, ksize=1, bias=_C.ptr(bq))
        ar.release(xn.t)
        o = self._new_act(ar, N, x.sp, Cc)
        W3 = 3 * Cc
        qbase = kbase = qkv.ip
        Tk = S
        gathered = None
        if R > 1:
            # queries stay local, keys/values of the whole volume are gathered (slabs are contiguous in depth,
            # tokens are ordered (d, h, w), so rank order == global token order)
            if getattr(self.slab, "peer", False):
                gathered = self.slab.alloc(R * S * W3 * 2)
                self.slab.plan_all_gather(plan, gathered, qkv.ip, S * W3 * 2)
            else:
                gathered = ar.alloc((R * S, W3), torch.bfloat16)
                plan.add_py(self.slab.all_gather, gathered, qkv.interior)
            kbase, Tk = gathered.data_ptr(), R * S
        aa = _C.AttnArgs(qbase, kbase + d * 2, kbase + 2 * d * 2, o.ip, S * W3, Tk * W3, Tk * W3, S * Cc, W3, W3, W3, Cc,
                         3 * d, 3 * d, 3 * d, d, N, H, S, Tk, d, 1.0 / math.sqrt(d), None, 0)
        self._attn_workspace(plan, ar, aa)
        plan.keep.append(aa)
        plan.add(self.lib.gg_attention_fwd, C.byref(aa))
        plan.flops += 4 * N * H * S * Tk * d
        ar.release(qkv.t), ar.release(gathered)
        bp = self._vec8((id(ab.proj_out.bias), "b"), lambda: ab.proj_out.bias, Cc)
        out = self._conv(plan, ar, [(o, False)], self._pack(ab.proj_out, [Cc]), Cc, dims=3, ksize=1, bias=_C.ptr(bp), residual=x,
                         stats=True)
        ar.release(o.t)
        return out

    def _linear(self, plan, ar, x: Act, lin_w: torch.Tensor, key, cout, bias_t=None, residual=None) -> Act:
        wp = self._cached(key, lambda: ops.pack_conv_weight(lin_w.reshape(lin_w.shape[0], -1, 1), [lin_w.shape[1]]))
        b = None
        if bias_t is not None:
            b = _C.ptr(self._vec8(key + ("b",), lambda: bias_t, cout))
        return self._conv(plan, ar, [(x, False)], wp, cout, dims=3, ksize=1, bias=b, residual=residual)

    def _cross_attention(self, plan, ar, ca: M.CrossAttention, xq: Act, ctx: Optional[Act], residual: Act) -> Act:
        """to_out(attention(to_q(xq), to_k(ctx), to_v(ctx))) + residual; ctx None = self-attention."""
        H, d = ca.heads, ca.dim_head
        inner = H * d
        N, Tq = xq.N, xq.S
        if ctx is None:
            w = self._cached((id(ca.to_q.weight), "qkv"),
                             lambda: torch.cat([ca.to_q.weight, ca.to_k.weight, ca.to_v.weight], 0).detach())
            qkv = self._linear(plan, ar, xq, w, (id(ca.to_q.weight), "qkvp"), 3 * inner)
            base = qkv.ip
            Tk, rs = Tq, 3 * inner
            qp, kp, vp = base, base + inner * 2, base + 2 * inner * 2
            q_str, k_str = (Tq * rs, rs, d), (Tk * rs, rs, d)
            bufs = [qkv.t]
        else:
            q = self._linear(plan, ar, xq, ca.to_q.weight, (id(ca.to_q.weight), "qp"), inner)
            w = self._cached((id(ca.to_k.weight), "kv"), lambda: torch.cat([ca.to_k.weight, ca.to_v.weight], 0).detach())
            kv = self._linear(plan, ar, ctx, w, (id(ca.to_k.weight), "kvp"), 2 * inner)
            Tk = ctx.S
            qp, kp, vp = q.ip, kv.ip, kv.ip + inner * 2
            q_str, k_str = (Tq * inner, inner, d), (Tk * 2 * inner, 2 * inner, d)
            bufs = [q.t, kv.t]
        o = self._new_act(ar, N, xq.sp, inner)
        aa = _C.AttnArgs(qp, kp, vp, o.ip, q_str[0], k_str[0], k_str[0], Tq * inner, q_str[1], k_str[1], k_str[1], inner,
                         d, d, d, d, N, H, Tq, Tk, d, float(ca.scale), None, 0)
        self._attn_workspace(plan, ar, aa)
        plan.keep.append(aa)
        plan.add(self.lib.gg_attention_fwd, C.byref(aa))
        plan.flops += 4 * N * H * Tq * Tk * d
        for b in bufs:
            ar.release(b)
        lo = ca.to_out[0]
        out = self._linear(plan, ar, o, lo.weight, (id(lo.weight), "p"), lo.out_features, lo.bias, residual)
        ar.release(o.t)
        return out

    def _layernorm(self, plan, ar, x: Act, norm: M.ParamNorm) -> Act:
        y = self._new_act(ar, x.N, x.sp, x.C)
        plan.add(self.lib.gg_layernorm, x.ip, _C.ptr(self._f32(norm.weight)), _C.ptr(self._f32(norm.bias)), y.ip,
                 x.N * x.S, x.C, float(norm.eps))
        return y

    def _transformer_block(self, plan, ar, blk: M.BasicTransformerBlock, h: Act, ctx: Optional[Act]) -> Act:
        """BasicTransformerBlock._forward (ldm/modules/attention.py:210-214, ccdm unet_openai/attention.py:142-146):
        self-attention, cross-attention (self when ctx is None), GEGLU feed-forward, each with its residual.
        Consumes (releases) h."""
        n1 = self._layernorm(plan, ar, h, blk.norm1)
        h2 = self._cross_attention(plan, ar, blk.attn1, n1, None, h)
        ar.release(n1.t), ar.release(h.t)
        n2 = self._layernorm(plan, ar, h2, blk.norm2)
        h3 = self._cross_attention(plan, ar, blk.attn2, n2, ctx, h2)
        ar.release(n2.t), ar.release(h2.t)
        n3 = self._layernorm(plan, ar, h3, blk.norm3)
        gp = blk.ff.net[0].proj
        f1 = self._linear(plan, ar, n3, gp.weight, (id(gp.weight), "p"), gp.out_features, gp.bias)
        ar.release(n3.t)
        inner = gp.out_features // 2
        g = self._new_act(ar, h3.N, h3.sp, inner)
        plan.add(self.lib.gg_geglu, f1.ip, g.ip, h3.N * h3.S, inner)
        ar.release(f1.t)
        l2 = blk.ff.net[2]
        out = self._linear(plan, ar, g, l2.weight, (id(l2.weight), "p"), l2.out_features, l2.bias, h3)
        ar.release(g.t), ar.release(h3.t)
        return out

    def build_token_plan(self, N: int, L: int, Cc: int, blocks) -> Plan:
        """A stack of BasicTransformerBlocks over a token tensor [N, L, C] (bf16, channels-last): the text-context
        encoder of the CCDM (ccdm/ddpm/models/encoder.py:103-123).  inputs['tokens'] -> outputs['tokens']."""
        self.lib = _C.lib()
        dev = self._dev()
        plan, ar = Plan(), _Arena(dev)
        t_in = torch.zeros((N, 1, 1, L, Cc), dtype=torch.bfloat16, device=dev)
        plan.inputs["tokens"] = t_in
        plan.keep.append(t_in)
        h = Act(t_in)           # not arena memory: the first block's release of its input is a no-op
        for blk in blocks:
            h = self._transformer_block(plan, ar, blk, h, None)
        plan.outputs["tokens"] = h.interior
        plan.keep.append(ar.stores)
        plan.arena_bytes = ar.total
        return plan

    def _spatial_transformer(self, plan, ar, st: M.SpatialTransformer, x: Act, ctx: Optional[Act]) -> Act:
        if self.slab is not None:
            raise NotImplementedError("depth-slab mode covers the shipped CCDM network (AttentionBlock), not SpatialTransformer")
        xn = self._gn(plan, ar, x, None, st.norm, False)
        pin = st.proj_in
        h = self._linear(plan, ar, xn, pin.weight.reshape(pin.out_channels, -1), (id(pin.weight), "p"), pin.out_channels, pin.bias)
        ar.release(xn.t)
        for blk in st.transformer_blocks:
            h = self._transformer_block(plan, ar, blk, h, ctx)
        po = st.proj_out
        out = self._linear(plan, ar, h, po.weight.reshape(po.out_channels, -1), (id(po.weight), "p"), po.out_channels, po.bias, x)
        ar.release(h.t)
        return out

    def _downsample(self, plan, ar, ds: M.Downsample, x: Act) -> Act:
        b = self._vec8((id(ds.op.bias), "b"), lambda: ds.op.bias, ds.out_channels)
        return self._conv(plan, ar, [(x, False)], self._pack(ds.op, [x.C]), ds.out_channels, dims=ds.dims, stride=2,
                          bias=_C.ptr(b), stats=True)

    def _upsample(self, plan, ar, up: M.Upsample, x: Act) -> Act:
        dims = up.dims
        N, (D, H, W), Cc = x.N, x.sp, x.C
        fd, fh = (2 if dims >= 3 else 1), (2 if dims >= 2 else 1)
        Do, Ho, Wo = D * fd, H * fh, W * 2
        if not up.use_conv or not self.fused_upsample:
            y = self._new_act(ar, N, (Do, Ho, Wo), Cc)
            plan.add(self.lib.gg_upsample2x, x.ip, y.ip, N, D, H, W, Cc, dims)
            if not up.use_conv:
                return y
            b = self._vec8((id(up.conv.bias), "b"), lambda: up.conv.bias, up.out_channels)
            out = self._conv(plan, ar, [(y, False)], self._packer(up.conv, [Cc]), up.out_channels, dims=dims, bias=_C.ptr(b),
                             stats=True)
            ar.release(y.t)
            return out
        cout = up.out_channels
        b = self._vec8((id(up.conv.bias), "b"), lambda: up.conv.bias, cout)
        # nearest-x2 upsample folded into the conv: for output parity class pi (per dim) the 3-tap
        # filter over the upsampled grid collapses to 2 taps over the coarse grid at offsets
        # {pi - 1, pi}; taps that hit the same coarse voxel have their weights summed (exact, done in
        # fp32 before the bf16 rounding).  27 -> 8 taps = 3.4x fewer MACs; the upsampled tensor is
        # never materialised.  Each class writes a stride-2 view of the output.
        out = self._new_act(ar, N, (Do, Ho, Wo), cout)
        for pd in range(fd):
            for ph in range(fh):
                for pw in range(2):
                    wp = (lambda pd=pd, ph=ph, pw=pw: lambda cm: self._cached(
                        (id(up.conv.weight), "up", pd, ph, pw, cm),
                        lambda: ops.pack_conv_weight(_fold_upsample_weight(up.conv.weight, dims, (pd, ph, pw)), [Cc], chunk_major=cm)))()
                    taps = (2 if dims >= 3 else 1, 2 if dims >= 2 else 1, 2)
                    offs = (pd - 1 if dims >= 3 else 0, ph - 1 if dims >= 2 else 0, pw - 1)
                    yp = out.ip + ((pd * Ho + ph) * Wo + pw) * cout * 2
                    ystr = (Do * Ho * Wo * cout, fd * Ho * Wo * cout, fh * Wo * cout, 2 * cout)
                    self._conv(plan, ar, [(x, False)], wp, cout, dims=dims, bias=_C.ptr(b), taps=taps, offsets=offs, y_ptr=yp,
                               y_strides=ystr, out_spatial=(D, H, W), out=out, stats=True,
                               stats_part=((pd * fh + ph) * 2 + pw, fd * fh * 2))
        return out

    # -------------------------------------------------------------------------------- plan
    def pick_lanes(self, N: int, spatial) -> int:
        """Independent sample lanes of a plan (Plan.lanes).  GG_LANES pins the count.  Measured on B200 (round 2,
        profiles/r2_lanes_and_pdl.md): no gain where every launch already fills the GPU (config 2: 45.0 -> 45.9 ms) and none on
        the 64 x 64 latents of config 3 either (4.80 -> 5.05 ms: there a launch is a persistent 1-CTA-per-SM kernel whose
        fixed cost, not its CTA count, is the time -- lanes only multiply the launches); two samples of a 512 x 512 slice
        (config 4) gain 5 % from overlapping each other's tails.  So: two lanes for mid-sized 2-D batches, one otherwise."""
        env = os.environ.get("GG_LANES")
        if self.slab is not None:
            return 1
        if env is not None:
            k = max(1, int(env))
        else:
            positions = N * int(math.prod(spatial))
            k = 2 if (self.dims == 2 and (1 << 18) < positions <= (1 << 20)) else 1
        while k > 1 and N % k != 0:
            k -= 1
        return k

    def build_plan(self, N: int, spatial: Tuple[int, ...], in_ch_pad: int, ctx_shape=None, f32_head: bool = True,
                   lanes: Optional[int] = None) -> Plan:
        self.lib = _C.lib()
        m = self.model
        dev = self._dev()
        plan = Plan()
        lanes = self.pick_lanes(N, spatial) if lanes is None else lanes
        assert N % lanes == 0
        sp3 = (1,) * (3 - len(spatial)) + tuple(spatial)
        lead, trail = self._lead, self._trail
        x_full = torch.zeros((N, sp3[0] + lead + trail) + sp3[1:] + (in_ch_pad,), dtype=torch.bfloat16, device=dev)
        t_in = torch.zeros((N,), dtype=torch.float32, device=dev)
        plan.inputs["x"], plan.inputs["t"] = Act(x_full, lead, trail).interior, t_in     # contiguous: lead > 0 only with N == 1
        plan.keep.append(x_full)
        c_in = None
        if ctx_shape is not None:
            L, cd = ctx_shape
            c_in = torch.zeros((N, 1, 1, L, cd), dtype=torch.bfloat16, device=dev)
            plan.inputs["context"] = c_in
        oc = m.out[2]
        head_full = None
        if lanes > 1:
            # the lanes write disjoint sample ranges of ONE head tensor
            head_full = torch.zeros((N, sp3[0] + lead + trail) + sp3[1:] + ((oc.out_channels + 7) // 8 * 8,),
                                    dtype=torch.float32 if f32_head else torch.bfloat16, device=dev)
            plan.keep.append(head_full)
            plan.outputs["head"] = Act(head_full, lead, trail).interior
        per = N // lanes
        for j in range(lanes):
            plan.begin_lane(j)
            n0, n1 = j * per, (j + 1) * per
            ar = _Arena(dev)
            head = self._build_lane(plan, ar, Act(x_full[n0:n1], lead, trail), t_in[n0:n1],
                                    Act(c_in[n0:n1]) if c_in is not None else None, sp3, f32_head,
                                    Act(head_full[n0:n1], lead, trail) if head_full is not None else None)
            if lanes == 1:
                plan.outputs["head"] = head.interior
            ha = plan.lanes[j][-1][1][0]._obj
            if (kind_ccdm_head(m) and int(ha.algo) == 4 and oc.out_channels <= 16 and f32_head
                    and per * int(math.prod(sp3)) % 4 == 0):
                # the same launch with the categorical sampler in its epilogue (gg_conv_args.cat): used by the resident loop
                cat = _C.CatEpilogue()
                fa = _C.ConvArgs.from_buffer_copy(ha)
                fa.cat = C.pointer(cat)
                plan.fused_heads.append((fa, cat, n0))
                plan.keep.append((fa, cat))
            plan.keep.append(ar.stores)      # kernels hold raw pointers into these: they must outlive the plan
            plan.arena_bytes += ar.total
        if len(plan.fused_heads) == lanes:
            plan.fused_head = plan.fused_heads[0][:2]
        else:
            plan.fused_heads = []
        return plan

    def _build_lane(self, plan: Plan, ar: _Arena, x_act: Act, t_in: torch.Tensor, ctx: Optional[Act], sp3, f32_head: bool,
                    head_out: Optional[Act]) -> Act:
        """The launches of one forward over the samples of `x_act` (all of the batch, or one lane's share)."""
        m = self.model
        dev = self._dev()
        N = x_act.N
        if self.slab is not None and getattr(self.slab, "peer", False):
            self.slab.plan_begin(plan)              # one communication epoch per forward
        # ---- timestep path (unet.py:772 / :511-515) + every ResBlock's emb projection in one launch
        mc = m.model_channels
        E = 4 * mc
        temb = ar.alloc((N, mc), torch.float32)
        plan.add(self.lib.gg_timestep_embedding, _C.ptr(t_in), _C.ptr(temb), N, mc, 10000.0)
        e1 = ar.alloc((N, E), torch.float32)
        te0, te2 = m.time_embed[0], m.time_embed[2]
        plan.add(self.lib.gg_small_linear, _C.ptr(temb), _C.ptr(self._f32(te0.weight)), _C.ptr(self._f32(te0.bias)), _C.ptr(e1),
                 N, E, mc, 0, 1)
        emb = ar.alloc((N, E), torch.float32)
        # emb is only ever consumed through SiLU (every ResBlock's emb_layers = SiLU -> Linear, unet.py:205-211):
        # apply it once here instead of once per output row of the big projection below
        plan.add(self.lib.gg_small_linear, _C.ptr(e1), _C.ptr(self._f32(te2.weight)), _C.ptr(self._f32(te2.bias)), _C.ptr(emb),
                 N, E, E, 0, 1)
        rbs = [mod for mod in m.modules() if isinstance(mod, M.ResBlock)]
        offs, tot = {}, 0
        for rb in rbs:
            offs[id(rb)] = tot
            tot += (rb.out_channels + 7) // 8 * 8

        def make_emb_w():
            ws, bs = [], []
            for rb in rbs:
                c8 = (rb.out_channels + 7) // 8 * 8
                lin, c1 = rb.emb_layers[1], rb.in_layers[2]
                w = torch.zeros((c8, E), dtype=torch.float32, device=dev)
                w[:rb.out_channels] = lin.weight.detach().float()
                b = torch.zeros((c8,), dtype=torch.float32, device=dev)
                b[:rb.out_channels] = lin.bias.detach().float() + c1.bias.detach().float()   # conv1 bias folded in
                ws.append(w), bs.append(b)
            return torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()

        W_all, b_all = self._cached(("emb_all",), make_emb_w)
        emb_all = ar.alloc((N, tot), torch.float32)
        plan.add(self.lib.gg_small_linear, _C.ptr(emb), _C.ptr(W_all), _C.ptr(b_all), _C.ptr(emb_all), N, tot, E, 0, 0)

        def run_block(block, h: Act, skip: Optional[Act], protected) -> Act:
            first = True
            for layer in block:
                if isinstance(layer, M.ResBlock):
                    ep = emb_all.data_ptr() + 4 * offs[id(layer)]
                    new = self._resblock(plan, ar, layer, h, skip if first else None, ep, tot)
                elif isinstance(layer, M.AttentionBlock):
                    new = self._attention_block(plan, ar, layer, h)
                elif isinstance(layer, M.SpatialTransformer):
                    new = self._spatial_transformer(plan, ar, layer, h, ctx)
                elif isinstance(layer, M.Downsample):
                    new = self._downsample(plan, ar, layer, h)
                elif isinstance(layer, M.Upsample):
                    new = self._upsample(plan, ar, layer, h)
                elif isinstance(layer, M.ParamConv):
                    b = self._vec8((id(layer.bias), "b"), lambda: layer.bias, layer.out_channels)
                    h.halo_valid = False           # the sampler rewrites the input between forwards
                    new = self._conv(plan, ar, [(h, False)], self._packer(layer, [h.C]), layer.out_channels, dims=layer.dims,
                                     bias=_C.ptr(b), stats=True)
                else:
                    raise NotImplementedError(type(layer).__name__)
                if not any(h.t is p.t for p in protected):
                    self._free(ar, h)
                if first and skip is not None:
                    self._free(ar, skip)
                h, first = new, False
            return h

        hs: List[Act] = []
        h = x_act
        for block in m.input_blocks:
            h = run_block(block, h, None, hs + [x_act])
            hs.append(h)
        h = run_block(m.middle_block, h, None, hs)
        for block in m.output_blocks:
            skip = hs.pop()
            h = run_block(block, h, skip, hs)
        # ---- head: GN -> SiLU -> conv (-> softmax fused downstream)   unet.py:715-721
        oc = m.out[2]
        b = self._vec8((id(oc.bias), "b"), lambda: oc.bias, oc.out_channels)
        kw = dict(out=head_out) if head_out is not None else {}
        head = self._gn_conv(plan, ar, h, None, m.out[0], True, lambda splits: self._packer(oc, splits), oc.out_channels,
                             dims=oc.dims, bias=_C.ptr(b), f32_out=f32_head, **kw)
        self._free(ar, h)
        plan.keep.extend([emb_all, emb, e1, temb])
        return head

    def get_plan(self, N, spatial, in_ch_pad, ctx_shape=None) -> Plan:
        key = (N, tuple(spatial), in_ch_pad, ctx_shape, os.environ.get("GG_LANES"))
        if key not in self.plans:
            self.plans[key] = self.build_plan(N, tuple(spatial), in_ch_pad, ctx_shape)
        return self.plans[key]


def kind_ccdm_head(m) -> bool:
    """The sampler epilogue applies to the CCDM network (softmax head over classes, unet.py:715-721), dims 3."""
    return bool(getattr(m, "sofmtax_output", False)) and getattr(m, "dims", 0) == 3


def _fold_upsample_weight(w: torch.Tensor, dims: int, parity) -> torch.Tensor:
    """3-tap weights over a nearest-x2 upsampled grid -> 2-tap weights over the coarse grid for one
    output parity class.  Per dim, output o = 2j + pi reads upsampled p = o + k - 1 -> coarse
    j + floor((pi + k - 1) / 2): pi = 0 -> taps {k0} at j-1, {k1, k2} at j; pi = 1 -> {k0, k1} at j,
    {k2} at j+1."""
    w = w.detach().float()
    groups = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
    par = parity[3 - dims:]
    for ax in range(dims):
        g = groups[par[ax]]
        dim = 2 + ax
        w = torch.stack([w.index_select(dim, torch.tensor(ix, device=w.device)).sum(dim) for ix in g], dim)
    return w
