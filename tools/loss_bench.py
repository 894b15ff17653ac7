#!/usr/bin/env python
"""SSIM loss forward + backward: fused kernels vs the reference's formulation in eager torch (conv2d per frame,
one host round trip per frame - utils/losses.py:25-58) on the same GPU.  CUDA events, median of 20."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
from deep_cine_cardiac_mri_b200 import metrics


def eager_loss(Xt, Yt, w, cov_norm, k1=0.01, k2=0.03):
    ssims = 0.
    nt = Xt.shape[2]
    for t in range(nt):
        X, Y = Xt[:, :, t, :], Yt[:, :, t, :]
        dr = torch.Tensor([Y.max()]).to("cuda")[:, None, None, None]
        C1, C2 = (k1 * dr) ** 2, (k2 * dr) ** 2
        ux, uy = F.conv2d(X, w), F.conv2d(Y, w)
        uxx, uyy, uxy = F.conv2d(X * X, w), F.conv2d(Y * Y, w), F.conv2d(X * Y, w)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
        ssims += 1 - ((A1 * A2) / (B1 * B2)).mean()
    return ssims / nt


def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    w = torch.ones(1, 1, 7, 7, device="cuda") / 49
    for b in (1, 4):
        g = torch.Generator(device="cuda").manual_seed(b)
        y = torch.rand(b, 1, 15, 200, 200, device="cuda", generator=g) * 3
        x = (y + 0.1 * torch.randn(b, 1, 15, 200, 200, device="cuda", generator=g)).requires_grad_(True)

        def ours():
            x.grad = None
            metrics.ssim_loss(x, y).backward()

        def eager():
            x.grad = None
            eager_loss(x, y, w, 49 / 48).backward()
        a, e = float(metrics.ssim_loss(x, y)), float(eager_loss(x, y, w, 49 / 48))
        print(f"b={b} t=15 200x200  loss fused {a:.7f} eager {e:.7f} | fwd+bwd fused {timeit(ours):8.1f} us   eager torch {timeit(eager):8.1f} us")


if __name__ == "__main__":
    main()
