"""GroupNorm (+SiLU) forms on the shapes of the LDM `_ae` network and the deep CCDM levels: three launches (partial, finalize,
apply) vs gg_gn_fused, each captured in a CUDA graph of 20 back-to-back calls (tuning aid, not a bench line)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=20, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters) * 1e3


def main():
    from jointimagegeneration_b200 import _C, ops
    shapes = [(16, 4096, 160, 0), (16, 4096, 160, 160), (16, 1024, 320, 0), (16, 1024, 320, 320), (16, 1024, 640, 320), (16, 256, 640, 0),
              (16, 256, 640, 640), (16, 64, 640, 0), (16, 64, 800, 640), (16, 16, 800, 0), (16, 16, 800, 800),
              (8, 2048, 256, 0), (8, 256, 320, 0), (8, 256, 320, 320), (2, 65536, 128, 0)]
    for N, S, C1, C2 in shapes:
        x1 = torch.randn((N, 1, 1, S, C1), device="cuda", dtype=torch.bfloat16)
        x2 = torch.randn((N, 1, 1, S, C2), device="cuda", dtype=torch.bfloat16) if C2 else None
        C = C1 + C2
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.empty((N, 1, 1, S, C), device="cuda", dtype=torch.bfloat16)
        t3 = timed(lambda: ops.group_norm_cl(x1, x2, gamma, beta, 1e-5, True, out=out))
        cl = int(_C.lib().gg_gn_fused_resident(S, C))
        t1 = timed(lambda: ops.gn_fused(x1, x2, gamma, beta, 1e-5, True, out=out))
        mb = N * S * C * 2 / 1e6
        print(f"N {N:2d} S {S:5d} C {C1:4d}+{C2:4d} ({mb:6.1f} MB): three launches {t3:6.1f} us   one launch {t1:6.1f} us (cluster {cl})   "
              f"{2 * mb / t1 / 1e3:5.2f} TB/s", flush=True)


if __name__ == "__main__":
    main()
