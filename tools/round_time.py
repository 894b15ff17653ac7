#!/usr/bin/env python
"""dev: kernel time vs number of images (whole rounds of the 148-CTA persistent grid) -> per-round time and fixed overhead."""
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
for (b, t, c) in [(1, 1, 148), (1, 2, 148), (1, 4, 148), (1, 8, 148), (4, 15, 10), (1, 4, 150), (1, 4, 160), (1, 4, 185)]:
    n = b * t * c
    ks = [torch.randn(b, t, c, 200, 200, 2, device=dev, generator=g) for _ in range(2)]
    s = torch.randn(b, c, 200, 200, 2, device=dev, generator=g)
    x = torch.randn(b, t, 200, 200, 2, device=dev, generator=g)
    m = (torch.rand(b, t, 200, device=dev, generator=g) < 0.25).to(torch.uint8)
    v = torch.ones(1, device=dev)
    i = [0]
    def nxt(): i[0] ^= 1; return i[0]
    r = {
        "fft2c": timeit(lambda: ops.raw_fft2c(ks[nxt()], False, 1)),
        "reduce": timeit(lambda: ops.raw_sens_reduce(ks[nxt()], s)),
        "expand": timeit(lambda: ops.raw_sens_expand(x, s)),
        "expand_dc": timeit(lambda: ops.raw_sens_expand(x, s, 2, ks[nxt()], m, v)),
    }
    print(f"images {n:5d} (b{b} t{t} c{c}): " + "  ".join(f"{k} {val:7.1f} us" for k, val in r.items()))
    del ks, s, x
