"""Micro-benchmark of the stride-1 3x3x3 conv kernels on one layer shape (tuning aid, not a bench line).
Usage: python tools/bench_conv.py [N D H W]   -- prints ms and TFLOP/s per (algo, Cin, Cout, epilogue variant)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jointimagegeneration_b200 import _C, ops
    N, D, H, W = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (8, 64, 128, 128)
    algos = [int(v) for v in os.environ.get("ALGOS", "2,3,4").split(",")]
    rs = np.random.RandomState(0)
    dev = "cuda"
    cases = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("CASES", "64:64,128:64,64:12,128:128").split(",")]
    variants = os.environ.get("VARIANTS", "plain,res,stats,res+stats").split(",")
    for Cin, Cout in cases:
        x = torch.randn((N, D, H, W, Cin), device=dev, dtype=torch.bfloat16)
        w = torch.from_numpy((rs.standard_normal((Cout, Cin, 3, 3, 3)) / math.sqrt(Cin * 27)).astype(np.float32)).to(dev)
        b = torch.zeros(Cout, device=dev)
        Cout8 = (Cout + 7) // 8 * 8
        res = torch.randn((N, D, H, W, Cout8), device=dev, dtype=torch.bfloat16)
        for algo in algos:
            if algo == 4 and Cout > 80:
                continue
            wp = ops.pack_conv_weight(w, [Cin], chunk_major=True)
            for variant in variants:
                if "stats" in variant and Cout8 != 64:
                    continue
                y = torch.empty((N, D, H, W, Cout8), device=dev, dtype=torch.bfloat16)
                bp = ops.pad_vec(b, Cout)
                ss = None
                if "xf" in variant:       # fused input GroupNorm + SiLU (algo 4 only)
                    if algo != 4:
                        continue
                    gam, bet = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
                    ss = ops.gn_finalize(ops.gn_partial(x), None, gam, bet, D * H * W, 1e-5)
                srcs, wpk, ssl = [(x, False)], wp, ([ss] if ss is not None else None)
                if "skip" in variant:       # fused 1x1x1 skip over two extra raw sources (128 + 64 channels)
                    xa = torch.randn((N, D, H, W, 128), device=dev, dtype=torch.bfloat16)
                    xb = torch.randn((N, D, H, W, 64), device=dev, dtype=torch.bfloat16)
                    ex = [torch.randn((Cout, 128), device=dev) * 0.05, torch.randn((Cout, 64), device=dev) * 0.05]
                    wpk = ops.pack_conv_weight(w, [Cin], extra=ex, chunk_major=True)
                    srcs = [(x, False), (xa, True), (xb, True)]
                    ssl = ([ss, None, None] if ss is not None else None)
                a = ops.make_conv_args(srcs, wpk, Cout, y, dims=3, ksize=3, stride=1, bias=bp,
                                       residual=res if "res" in variant else None, algo=algo,
                                       src_ss=ssl, ss_stride=2 * Cin)
                if "cat" in variant:        # sampler epilogue of the head conv
                    if algo != 4 or Cout > 16:
                        continue
                    V = D * H * W
                    lab_in = torch.randint(0, Cout, (N * V,), device=dev, dtype=torch.uint8)
                    lab_out = torch.empty_like(lab_in)
                    nx = torch.empty((N * V, 16), device=dev, dtype=torch.bfloat16)
                    coef = torch.tensor([[0.9, 0.5]] * N, device=dev)
                    cat = _C.CatEpilogue(lab_in.data_ptr(), lab_out.data_ptr(), nx.data_ptr(), None, coef.data_ptr(), Cout, 1, 16, 1, 1e-12, 7, 3, 0)
                    y = torch.empty((N, D, H, W, 16), device=dev, dtype=torch.float32)
                    a = ops.make_conv_args(srcs, wpk, Cout, y, dims=3, ksize=3, stride=1, bias=bp, algo=algo, src_ss=ssl, ss_stride=2 * Cin)
                    a.cat = C.pointer(cat)
                part = None
                if "stats" in variant:
                    per = int(_C.lib().gg_conv_stats_chunks(C.byref(a)))
                    part = torch.empty((N, per, Cout8, 2), device=dev)
                    a.gn_partial, a.gn_chunk_base, a.gn_nchunks_total = _C.ptr(part), 0, per
                for _ in range(2):
                    ops.conv_fwd(a)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    ops.conv_fwd(a)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                fl = 2.0 * N * D * H * W * Cout * Cin * 27
                print(f"Cin {Cin:4d} Cout {Cout:4d} algo {algo} {variant:10s}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        del x, res


if __name__ == "__main__":
    main()
