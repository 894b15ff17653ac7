"""Synthetic cine k-space with the reference's shape / dtype conventions
(data/subsample.py:117-151, data/transforms.py:66-92,343).

numpy-only generators (seeded PCG64), so CPU tests, the oracle and the GPU
benches all see bit-identical inputs; `to_torch` moves a case to a device.
"""
from __future__ import annotations

import numpy as np


def random_mask(seed: int, b: int, t: int, h: int, n_center: int = 10, acc: int = 4) -> np.ndarray:
    """uint8 (b,t,1,h,1,1): per frame int(h/acc)-n_center rows drawn without replacement from the
    tail-adjusted Gaussian pdf + n_center centre rows (RandomMaskFunc semantics)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((b, t, h), dtype=np.uint8)
    pdf = np.exp(-(0.5 / (h / 10.0) ** 2) * (np.arange(h) - h / 2) ** 2) + (h / (2.0 * acc)) / h
    lo, hi = h // 2 - n_center // 2, h // 2 + n_center // 2
    pdf[lo:hi] = 0
    pdf /= pdf.sum()
    n_lines = max(int(h / acc) - n_center, 0)
    for i in range(b):
        for j in range(t):
            m[i, j, rng.choice(h, n_lines, replace=False, p=pdf)] = 1
    m[:, :, lo:hi] = 1
    return m.reshape(b, t, 1, h, 1, 1)


def sens_maps(seed: int, b: int, c: int, h: int, w: int, smooth: bool = True) -> np.ndarray:
    """(b,1,c,h,w,2) float32 coil maps with RSS == 1 at every pixel (no zero pixel:
    divide_root_sum_of_squares has no epsilon, varnet.py:58-59)."""
    rng = np.random.default_rng(seed)
    if smooth:
        yy, xx = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
        s = np.empty((b, 1, c, h, w, 2), dtype=np.float32)
        for i in range(b):
            for j in range(c):
                cy, cx = rng.uniform(-1, 1, 2)
                amp = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 1.5) + 0.05
                ph = rng.uniform(-np.pi, np.pi) + 1.5 * (yy * rng.uniform(-1, 1) + xx * rng.uniform(-1, 1))
                s[i, 0, j, ..., 0] = amp * np.cos(ph)
                s[i, 0, j, ..., 1] = amp * np.sin(ph)
    else:
        s = rng.standard_normal((b, 1, c, h, w, 2), dtype=np.float32)
    nrm = np.sqrt((s.astype(np.float64) ** 2).sum(axis=(2, 5), keepdims=True))
    return (s / nrm).astype(np.float32)


def phantom(seed: int, b: int, t: int, h: int, w: int) -> np.ndarray:
    """(b,t,1,h,w,2) smooth moving-ellipse object with a little texture."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    x = np.zeros((b, t, 1, h, w, 2), dtype=np.float32)
    for i in range(b):
        base = 0.05 * rng.standard_normal((h, w)).astype(np.float32)
        for j in range(t):
            r = 0.45 + 0.1 * np.sin(2 * np.pi * j / t)
            body = (xx ** 2 / 0.8 ** 2 + yy ** 2 / 0.9 ** 2 < 1).astype(np.float32)
            heart = ((xx - 0.1) ** 2 + (yy + 0.05) ** 2 < r ** 2).astype(np.float32)
            mag = 0.4 * body + 0.6 * heart + base * body
            ph = 0.3 * xx + 0.2 * yy
            x[i, j, 0, ..., 0] = mag * np.cos(ph)
            x[i, j, 0, ..., 1] = mag * np.sin(ph)
    return x


def cine_case(seed: int, b: int, t: int, c: int, h: int, w: int, noise: float = 0.01, randn: bool = False):
    """dict(image, sens, kspace, masked_kspace, mask) with the reference's shapes. Full k-space is
    computed with numpy (fft2c semantics) — used for inputs only, never on the product path."""
    rng = np.random.default_rng(seed + 17)
    img = rng.standard_normal((b, t, 1, h, w, 2), dtype=np.float32) if randn else phantom(seed, b, t, h, w)
    s = sens_maps(seed + 1, b, c, h, w, smooth=not randn)
    zi = img[..., 0] + 1j * img[..., 1]
    zs = s[..., 0] + 1j * s[..., 1]
    ax = (-2, -1)
    k = np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(zi * zs, axes=ax), axes=ax, norm="ortho"), axes=ax)
    k = np.stack([k.real, k.imag], axis=-1).astype(np.float32)
    k += noise * rng.standard_normal(k.shape, dtype=np.float32)
    m = random_mask(seed + 2, b, t, h)
    return dict(image=img, sens=s, kspace=k, masked_kspace=(k * m + 0.0).astype(np.float32), mask=m)


def to_torch(case: dict, device="cuda"):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in case.items()}
