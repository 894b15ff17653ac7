#!/usr/bin/env python
"""Launch plain sens_expand and sens_expand + soft DC a few times (for ncu captures)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops
b, t, c, h, w = 4, 15, 10, 200, 200
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
refs = [torch.randn(b, t, c, h, w, 2, device=dev, generator=g) for _ in range(2)]
s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
v = torch.tensor([1.0], device=dev)
for i in range(3):
    ops.raw_sens_expand(x, s)
    ops.raw_sens_expand(x, s, 2, refs[i % 2], m, v)
    ops.raw_sens_reduce(refs[i % 2], s)
torch.cuda.synchronize()
print("done")
