#!/usr/bin/env python
"""Timings of the on-chip normal operator (b2s_normal_op / b2s_normal_dc), CUDA events, median of 30."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops
from tools.quick_bench import timeit

def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    for (b, t, c, h, w) in [(4, 15, 10, 200, 200), (16, 15, 10, 200, 200), (1, 25, 20, 200, 200), (1, 15, 10, 200, 200),
                            (2, 15, 10, 256, 256), (1, 30, 32, 256, 256)]:
        if not ops.normal_op_supported(h, w): continue
        s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
        x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
        bref = torch.randn(b, t, h, w, 2, device=dev, generator=g)
        ssq = s.pow(2).sum(dim=(1, 4)).contiguous()
        m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
        v = torch.tensor([1.0], device=dev)
        I, S = x.numel() * 4, s.numel() * 4
        t0 = timeit(lambda: ops.raw_normal_op(x, s, m, v), n=30)
        t1 = timeit(lambda: ops.raw_normal_dc(x, s, m, v, ssq, bref), n=30)
        flop = b * t * c * w * (2 * 5 * h * 7.64 + 14 * h)      # two length-h transforms (5 N log2 N) + products
        print(f"b{b} t{t} c{c} {h}x{w}: normal_op {t0*1e6:7.1f} us ({(2*I+S)/t0/1e9:6.0f} GB/s of 2I+S, {flop/t0/1e12:5.1f} TFLOP/s)   "
              f"normal_dc {t1*1e6:7.1f} us")

if __name__ == "__main__":
    main()
