#!/usr/bin/env python
"""xf/yf plane glue around the regularisers: the reference's eager chain (varnet.py:215-232 + NormUnet glue,
norm_unet.py:101-113, U-Net replaced by identity) vs b2s_planes_* on the same GPU.  CUDA events, median of 30."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops
from oracle import load_reference as L
from tools.quick_bench import timeit

def main():
    L.load()
    from reconstruction.models.denoisers.norm_unet import NormUnet
    nu = NormUnet(4, 2).cuda()
    for (b, t, h, w) in [(1, 15, 200, 200), (4, 15, 200, 200), (1, 25, 200, 200), (1, 30, 256, 256)]:
        x = torch.randn(b, t, h, w, 2, device="cuda")
        def glue(z):
            z = nu.complex_to_chan_dim(z)
            z, mean, std = nu.norm(z)
            z, pads = nu.pad(z)
            z = nu.unpad(z, *pads)
            z = nu.unnorm(z, mean, std)
            return nu.chan_complex_to_last_dim(z)
        def eager():
            xf = x.clone().permute(0, 2, 3, 1, 4).reshape(b * h, 1, w, t, 2)
            yf = x.clone().permute(0, 3, 2, 1, 4).reshape(b * w, 1, h, t, 2)
            xf, yf = glue(xf), glue(yf)
            xf_r = xf.view(b, h, 1, w, t, 2).permute(0, 4, 2, 1, 3, 5)
            yf_r = yf.view(b, w, 1, h, t, 2).permute(0, 4, 2, 3, 1, 5)
            return 0.5 * (xf_r + yf_r)
        def ours():
            xf, yf, ctx = ops.raw_planes_pack(x, True, True)
            return ops.raw_planes_unpack(xf, yf, ctx)
        with torch.no_grad():
            te, to = timeit(eager, n=30), timeit(ours, n=30)
        print(f"b{b} t{t} {h}x{w}: reference eager glue {te*1e6:8.1f} us   b2s_planes_* {to*1e6:7.1f} us   ({te/to:4.1f}x)")

if __name__ == "__main__":
    main()
