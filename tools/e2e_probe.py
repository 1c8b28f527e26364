"""Times the public DenoisingModel.forward call repeatedly (tuning aid for bench.py's e2e number)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from jointimagegeneration_b200.ccdm.builder import build_model

def main():
    dev = torch.device("cuda:0")
    wl = bench.WORKLOADS["ccdm_cfg2"]
    B, Cc, sp, T = wl["batch"], wl["C"], wl["spatial"], wl["T"]
    model = build_model(T, "cosine", {"s": 0.008}, [(1,) + sp, (Cc,) + sp], None, "unet_openai", dict(bench.CCDM_NET), "synthetic",
                        "majority", dims=3)
    bench.randomize_zero_modules(model.unet, 7)
    model = model.to(dev).eval()
    model.loop, model.use_cuda_graph, model.philox_seed = "resident", True, 99
    lab0 = torch.randint(0, Cc, (B,) + sp, device=dev)
    x_T = torch.zeros((B, Cc) + sp, dtype=torch.float32, device=dev).scatter_(1, lab0[:, None], 1.0)
    cond = torch.zeros((B, 1) + sp, dtype=torch.float32, device=dev)
    x_host, c_host = x_T.cpu().pin_memory(), cond.cpu().pin_memory()
    out_host = torch.empty((B, Cc) + sp, dtype=torch.int64).pin_memory()
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = model(x_host, c_host, t=torch.tensor(10000 + 10))["diffusion_out"]
        t1 = time.perf_counter()
        out_host.copy_(res, non_blocking=False)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"rep {rep}: forward {1e3 * (t1 - t0):.1f} ms, copy out {1e3 * (t2 - t1):.1f} ms", flush=True)

main()
