#!/usr/bin/env python
"""ncu target: a few launches of the on-chip normal operator (b4 t15 c10 200x200, then b16)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for b in (4, 16):
    t, c, h, w = 15, 10, 200, 200
    s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
    x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
    m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
    v = torch.tensor([1.0], device=dev)
    for _ in range(3): ops.raw_normal_op(x, s, m, v)
    torch.cuda.synchronize()
print("done")
