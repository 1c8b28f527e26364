"""Micro-benchmark of ONE parity phase of the folded nearest-x2 upsample conv (2x2x2 taps over the coarse grid, stride-2 output
view; unet_engine._upsample) -- tuning aid, not a bench line.
Usage: python tools/bench_upconv.py [N D H W C]   (coarse grid; default 8 32 64 64 128)"""
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
    from jointimagegeneration_b200.unet_engine import _fold_upsample_weight
    N, D, H, W, Cc = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else (8, 32, 64, 64, 128)
    dev = "cuda"
    rs = np.random.RandomState(0)
    x = torch.randn((N, D, H, W, Cc), device=dev, dtype=torch.bfloat16)
    w = torch.from_numpy((rs.standard_normal((Cc, Cc, 3, 3, 3)) / math.sqrt(Cc * 27)).astype(np.float32)).to(dev)
    b = ops.pad_vec(torch.zeros(Cc, device=dev), Cc)
    Do, Ho, Wo = 2 * D, 2 * H, 2 * W
    out = torch.empty((N, Do, Ho, Wo, Cc), device=dev, dtype=torch.bfloat16)
    pd, ph, pw = 1, 0, 1
    wp = ops.pack_conv_weight(_fold_upsample_weight(w, 3, (pd, ph, pw)), [Cc], chunk_major=True)
    yp = out.data_ptr() + ((pd * Ho + ph) * Wo + pw) * Cc * 2
    ystr = (Do * Ho * Wo * Cc, 2 * Ho * Wo * Cc, 2 * Wo * Cc, 2 * Cc)
    for variant in os.environ.get("VARIANTS", "plain,stats").split(","):
        a = ops.make_conv_args([(x, False)], wp, Cc, yp, dims=3, bias=b, taps=(2, 2, 2), offsets=(pd - 1, ph - 1, pw - 1),
                               out_spatial=(D, H, W), y_strides=ystr, algo=int(os.environ.get("ALGO", "1")))
        part = None
        if "stats" in variant:
            per = int(_C.lib().gg_conv_stats_chunks(C.byref(a)))
            part = torch.empty((N, per, Cc, 2), device=dev)
            a.gn_partial, a.gn_chunk_base, a.gn_nchunks_total = _C.ptr(part), 0, per
        for _ in range(2):
            ops.conv_fwd(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 8
        e0.record()
        for _ in range(reps):
            ops.conv_fwd(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * N * D * H * W * Cc * Cc * 8
        print(f"upsample phase N{N} {D}x{H}x{W} C{Cc} {variant:6s}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s issued", flush=True)


if __name__ == "__main__":
    main()
