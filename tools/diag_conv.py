"""Run every conv test case without aborting and print error patterns (debug aid for the
tcgen05 kernel; writes gpurun_out/diag_conv.txt).  Usage: python tools/diag_conv.py [case ...]"""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from jointimagegeneration_b200 import ops
    import test_gpu_kernels as T
    names = sys.argv[1:] or list(T.CONV_CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "diag_conv.txt"), "w")

    def emit(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()

    for name in names:
        try:
            err, got, want = T._conv_case(ops, **T.CONV_CASES[name])
            d = (got - want).abs()
            scale = float(want.abs().max())
            bad = d > 0.02 * scale
            emit(f"{name}: rel_err={err:.3e} bad_frac={float(bad.float().mean()):.4f} scale={scale:.3f}")
            if bad.any():
                # which channels / positions are wrong
                per_c = bad.float().mean(dim=(0, 2, 3, 4))
                per_w = bad.float().mean(dim=(0, 1, 2, 3))
                per_h = bad.float().mean(dim=(0, 1, 2, 4))
                per_d = bad.float().mean(dim=(0, 1, 3, 4))
                per_n = bad.float().mean(dim=(1, 2, 3, 4))
                emit("  bad per channel:", [round(float(v), 2) for v in per_c[:32]], "...")
                emit("  bad per w:", [round(float(v), 2) for v in per_w[:32]])
                emit("  bad per h:", [round(float(v), 2) for v in per_h[:32]])
                emit("  bad per d:", [round(float(v), 2) for v in per_d[:32]])
                emit("  bad per n:", [round(float(v), 2) for v in per_n])
                emit("  got[0,:4,0,0,:4]:", got[0, :4, 0, 0, :4].tolist())
                emit("  want[0,:4,0,0,:4]:", want[0, :4, 0, 0, :4].tolist())
                ratio = (got.flatten()[:8] / want.flatten()[:8]).tolist()
                emit("  ratio first 8:", ratio)
        except Exception as e:  # noqa: BLE001
            emit(f"{name}: EXCEPTION {type(e).__name__}: {e}")
            traceback.print_exc()
            if "CUDA" in str(e) or "cuda" in str(e):
                emit("CUDA context likely poisoned; stopping")
                break
    out.close()


if __name__ == "__main__":
    main()
