#!/usr/bin/env python
"""dev: expand+DC time vs fraction of sampled rows (separates the cost of the DC machinery from the cost of the reference traffic)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
b, t, c = 1, 4, 148
ks = [torch.randn(b, t, c, 200, 200, 2, device=dev, generator=g) for _ in range(2)]
s = torch.randn(b, c, 200, 200, 2, device=dev, generator=g)
x = torch.randn(b, t, 200, 200, 2, device=dev, generator=g)
v = torch.ones(1, device=dev)
i = [0]
def nxt(): i[0] ^= 1; return i[0]
print("plain expand", timeit(lambda: ops.raw_sens_expand(x, s)))
for frac in (0.0, 0.02, 0.1, 0.25, 0.5, 1.0):
    m = (torch.rand(b, t, 200, device=dev, generator=g) < frac).to(torch.uint8)
    print(f"sampled fraction {frac:4.2f}: expand_dc {timeit(lambda: ops.raw_sens_expand(x, s, 2, ks[nxt()], m, v)):7.1f} us   (mask mode {timeit(lambda: ops.raw_sens_expand(x, s, 1, None, m, None)):7.1f} us)")
