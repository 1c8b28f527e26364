"""Times AutoencoderKL.encode / .decode at the shipped size (ch 128, mult 1-2-4-4, 512 x 512 slices, n = 2) -- tuning aid."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jointimagegeneration_b200.ldm.autoencoder import AutoencoderKL

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dd = dict(ch=128, out_ch=1, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0, in_channels=1,
              resolution=512, z_channels=4, double_z=True, dims=2)
    torch.manual_seed(0)
    ae = AutoencoderKL(dd, 4).cuda().eval()
    x = torch.randn(n, 1, 512, 512, device="cuda")
    z = torch.randn(n, 4, 64, 64, device="cuda")
    for name, fn in (("encode", lambda: ae.encode(x).mode()), ("decode", lambda: ae.decode(z))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        plan = next(iter((ae.encoder if name == "encode" else ae.decoder)._plans.values()))
        ms = e0.elapsed_time(e1) / 10
        print(f"{name}: {ms:.3f} ms for n = {n} slices of 512 x 512 ({plan.num_launches} launches, {plan.flops / 1e12:.3f} TFLOP issued, "
              f"{plan.flops / ms / 1e9:.0f} TFLOP/s, arena {plan.arena_bytes / 2**20:.0f} MiB)")

main()
