"""torch-CPU port of the reference's SENSE/DC op chain — the timed CPU baseline.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product package).
The reference's hot path is eager torch on the host cores; it cannot travel to
the GPU box, so `bench.py`'s `cpu_baseline` leg and `--impl reference` arm time
this port instead (`kind: "port"`).  It keeps the reference's op structure —
shift / FFT / shift, multiply via four real products + stack, the 8-op DC blend —
so its cost is representative: utils/fftc.py:5-110, utils/math.py:5-79,
utils/coil_combine.py:21-34, models/varnet.py:58-86,143-151,181-282.
Checked against the numpy oracle in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch


def _shift(x, dims, inverse):
    # utils/fftc.py:166-213 (fftshift: n//2, ifftshift: (n+1)//2), one roll per dim like the reference
    for d in dims:
        n = x.shape[d]
        x = torch.roll(x, (n + 1) // 2 if inverse else n // 2, d)
    return x


def _centred(x, real_dims, fn, norm):
    if x.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    x = _shift(x, real_dims, True)
    cdims = tuple(d + 1 for d in real_dims)
    x = torch.view_as_real(fn(torch.view_as_complex(x.contiguous()), dim=cdims, norm=norm))
    return _shift(x, real_dims, False)


def fft2c(x, norm="ortho"):
    return _centred(x, (-3, -2), torch.fft.fftn, norm)


def ifft2c(x, norm="ortho"):
    return _centred(x, (-3, -2), torch.fft.ifftn, norm)


def fft1c(x, norm="ortho"):
    return _centred(x, (-2,), torch.fft.fftn, norm)


def ifft1c(x, norm="ortho"):
    return _centred(x, (-2,), torch.fft.ifftn, norm)


def complex_mul(x, y):
    re = x[..., 0] * y[..., 0] - x[..., 1] * y[..., 1]
    im = x[..., 0] * y[..., 1] + x[..., 1] * y[..., 0]
    return torch.stack((re, im), dim=-1)


def complex_conj(x):
    return torch.stack((x[..., 0], -x[..., 1]), dim=-1)


def complex_abs(x):
    return (x ** 2).sum(dim=-1).sqrt()


def rss_complex(x, dim=0):
    return torch.sqrt((x ** 2).sum(dim=-1).sum(dim))


def sens_expand(x, sens):
    return fft2c(complex_mul(x, sens))


def sens_reduce(k, sens):
    return complex_mul(ifft2c(k), complex_conj(sens)).sum(dim=2, keepdim=True)


def dc_blend(k, ref, mask, v):
    return (1 - mask) * k + mask * (k + v * ref) / (1 + v)


def sens_model(masked_kspace, mask, unet=None):
    """models/varnet.py:62-86 with the U-Net as an optional callable (identity when None)."""
    cent = mask.shape[-3] // 2
    line = mask[:, 0, :].squeeze()
    left = torch.nonzero(line[:cent] == 0)[-1]
    right = torch.nonzero(line[cent:] == 0)[0] + cent
    nlf = int(right - left)
    pad = (mask.shape[-3] - nlf + 1) // 2
    mean = torch.mean(masked_kspace, 1)
    x = torch.zeros_like(mean)
    x[:, :, pad:pad + nlf, :] = mean[:, :, pad:pad + nlf, :]
    x = ifft2c(x)
    if unet is not None:
        x = unet(x)
    x = x / rss_complex(x, dim=1).unsqueeze(-1).unsqueeze(1)
    return x.unsqueeze(1)


def temporal_pre(img, xf=True):
    """models/varnet.py:202-213 on (b,t,h,w,2)."""
    t = img.shape[1]
    mean = torch.stack(t * [torch.mean(img.clone(), dim=1)], dim=1)
    x = img - mean
    if xf:
        x = fft1c(x.permute(0, 2, 3, 1, 4)).permute(0, 3, 1, 2, 4)
    return x, mean


def temporal_post(out, mean, xf=True):
    """models/varnet.py:234-241: out (b,t,1,h,w,2)."""
    if xf:
        out = ifft1c(out.permute(0, 2, 3, 4, 1, 5)).permute(0, 4, 1, 2, 3, 5)
    return out + mean.unsqueeze(2)


def varnet_hot_path(masked_kspace, mask, n_cascades=12, v=1.0, xf=True, regulariser=None, sens_unet=None):
    """SENSE/DC hot path of one XF-VarNet forward (models/varnet.py:143-151, 244-282); the cuDNN
    regularisers are callables (identity when None: they are outside the hot path)."""
    sens = sens_model(masked_kspace, mask, sens_unet)
    k = masked_kspace.clone()
    v = torch.as_tensor(v, dtype=masked_kspace.dtype)
    for _ in range(n_cascades):
        img = sens_reduce(k, sens)
        x, mean = temporal_pre(img.squeeze(2), xf)
        x = x.unsqueeze(2)
        if regulariser is not None:
            x = regulariser(x)
        model_out = temporal_post(x, mean, xf)
        k = dc_blend(sens_expand(model_out, sens), masked_kspace, mask, v)
    return complex_abs(complex_mul(ifft2c(k), complex_conj(sens)).sum(dim=2, keepdim=False))
