#!/usr/bin/env python
"""fft2c / sens_expand+DC / sens_reduce timings of the current library for env sweeps (one process per setting)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops
from quick_bench import timeit
b, t, c, h, w = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (4, 15, 10, 200, 200))]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
ks = [torch.randn(b, t, c, h, w, 2, device=dev, generator=g) for _ in range(3)]
s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
v = torch.tensor([1.0], device=dev)
i = [0]
def nxt(): i[0] = (i[0] + 1) % 3; return i[0]
import os
warm = bool(os.environ.get("B2S_SWEEP_WARM"))       # the cascade's reference k-space is the same tensor every time
r = [timeit(lambda: ops.raw_fft2c(ks[nxt()], False, 1)), timeit(lambda: ops.raw_sens_reduce(ks[nxt()], s)),
     timeit(lambda: ops.raw_sens_expand(x, s, 2, ks[0] if warm else ks[nxt()], m, v))]
print("fft2c %.1f us  sens_reduce %.1f us  sens_expand_dc %.1f us" % tuple(1e6 * q for q in r))
