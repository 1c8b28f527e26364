"""Runs the per-voxel posterior / draw kernel at the reference interface on config-2 sized tensors (for an ncu capture)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jointimagegeneration_b200 import ops  # noqa: E402

B, C, sp = 8, 12, (64, 128, 128)
V = sp[0] * sp[1] * sp[2]
x0 = torch.softmax(torch.randn((B, C) + sp, device="cuda"), 1)
lab = torch.randint(0, C, (B,) + sp, device="cuda")
xt = torch.zeros((B, C) + sp, device="cuda").scatter_(1, lab[:, None], 1.0)
q = torch.empty((B * V, C), device="cuda").exponential_(1)
coef = torch.tensor([[0.93, 0.41]] * B, device="cuda")
out = torch.empty_like(x0)
for _ in range(3):
    ops.cat_posterior_sample(x0, xt, coef, ops.CAT_SAMPLE, q=q, out=out)
torch.cuda.synchronize()
print("ok")
