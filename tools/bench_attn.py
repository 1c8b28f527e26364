"""Micro-benchmark of gg_attention_fwd on the legacy-layout self-attention sites (tuning aid).  GG_ATTN_TC=0 selects the
mma.sync kernel.  Usage: python tools/bench_attn.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jointimagegeneration_b200 import ops
    cases = [(8, 8, 2048, 32, "cfg2 ds8"), (8, 10, 256, 32, "cfg2 ds16"), (1, 8, 16384, 32, "cfg5 ds8"), (1, 10, 2048, 32, "cfg5 ds16"),
             (16, 10, 1024, 32, "cfg3 ds2"), (16, 20, 256, 32, "cfg3 ds4"), (2, 16, 4096, 32, "cfg4 ds8"), (2, 20, 1024, 32, "cfg4 ds16")]
    for B, H, T, d, name in cases:
        qkv = torch.randn((B, T, 3 * H * d), device="cuda", dtype=torch.bfloat16)
        out = torch.empty((B, T, H * d), device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.attention_legacy(qkv, H, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.attention_legacy(qkv, H, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 4.0 * B * H * T * T * d
        print(f"{name:10s} B{B} H{H} T{T} d{d}: {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  ({B * H * T * T / ms / 1e6:8.1f} G exp/s)", flush=True)


if __name__ == "__main__":
    main()
