"""Training loss and test metrics of the reference on the GPU.

* `SSIMLoss` - drop-in for `reconstruction.utils.losses.SSIMLoss` (utils/losses.py:6-58): same constructor,
  same registered buffer `w` (checkpoints keep loading), same call `loss(Xt, Yt, data_range)` on
  `(b,1,t,h,w)` tensors, same quirk: the `data_range` argument is ignored and each frame uses the maximum of
  the target frame over the whole batch (:35).  One fused kernel per direction instead of 15 x (5 convolutions
  + 20 pointwise kernels + one host round trip per frame); no host synchronisation at all.
* `mse / nmse / psnr / ssim` - `reconstruction.utils.evaluate` (utils/evaluate.py:6-49) on CUDA tensors,
  returning 0-dim device tensors (the reference round-trips through numpy + skimage).

All results come from ordered reductions: bit-identical from run to run.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .ops import _need_cuda, _p, _stream

_ERR_SCRATCH = 3 * 1024


def _f32c(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        raise TypeError(f"b200sense metrics are float32 (got {x.dtype})")
    return x.contiguous()


def raw_frame_max(y4: torch.Tensor) -> torch.Tensor:
    b, t, h, w = y4.shape
    out = torch.empty(t, dtype=torch.float32, device=y4.device)
    _lib.check(_lib.lib().b2s_frame_max(_p(y4), _p(out), b, t, h * w, _stream()), "frame_max")
    return out


def raw_ssim_fwd(x4, y4, data_range, dr_stride, win, k1, k2):
    """x4, y4 (b,t,h,w) -> out (t+1,): per-frame mean SSIM, then the loss."""
    b, t, h, w = x4.shape
    out = torch.empty(t + 1, dtype=torch.float32, device=x4.device)
    scratch = torch.empty(max(1, _lib.lib().b2s_ssim_scratch_floats(b, t, h, w)), dtype=torch.float32, device=x4.device)
    _lib.check(_lib.lib().b2s_ssim_fwd(_p(x4), _p(y4), _p(data_range), dr_stride, b, t, h, w, win, float(k1), float(k2),
                                       _p(out), _p(scratch), _stream()), "ssim_fwd")
    return out


def raw_ssim_bwd(x4, y4, data_range, dr_stride, gout, win, k1, k2):
    b, t, h, w = x4.shape
    gx = torch.empty_like(x4)
    _lib.check(_lib.lib().b2s_ssim_bwd(_p(x4), _p(y4), _p(data_range), dr_stride, _p(gout), b, t, h, w, win, float(k1),
                                       float(k2), _p(gx), _stream()), "ssim_bwd")
    return gx


class _SSIMLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x4, y4, win, k1, k2):
        x4, y4 = _f32c(x4), _f32c(y4)
        dr = raw_frame_max(y4)                                   # losses.py:35, without the host round trip
        out = raw_ssim_fwd(x4, y4, dr, 1, win, k1, k2)
        ctx.save_for_backward(x4, y4, dr)
        ctx.cfg = (win, k1, k2)
        return out[-1]

    @staticmethod
    def backward(ctx, g):
        x4, y4, dr = ctx.saved_tensors
        win, k1, k2 = ctx.cfg
        gx = raw_ssim_bwd(x4, y4, dr, 1, g.reshape(1).to(torch.float32).contiguous(), win, k1, k2)
        return gx, None, None, None, None                        # the target is a constant (it never requires grad)


class SSIMLoss(nn.Module):
    """Time-averaged SSIM loss module (utils/losses.py:6-58)."""

    def __init__(self, win_size: int = 7, k1: float = 0.01, k2: float = 0.03):
        super().__init__()
        self.win_size = win_size
        self.k1, self.k2 = k1, k2
        self.register_buffer("w", torch.ones(1, 1, win_size, win_size) / win_size ** 2)
        NP = win_size ** 2
        self.cov_norm = NP / (NP - 1)

    def forward(self, Xt: torch.Tensor, Yt: torch.Tensor, data_range: torch.Tensor = None):
        return ssim_loss(Xt, Yt, self.win_size, self.k1, self.k2)


def ssim_loss(Xt: torch.Tensor, Yt: torch.Tensor, win_size: int = 7, k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """Xt, Yt (b,1,t,h,w) (the reference passes `output.unsqueeze(1)`, varnet_module.py:110-112)."""
    _need_cuda(Xt, Yt)
    if Xt.dim() != 5 or Xt.shape[1] != 1 or Xt.shape != Yt.shape:
        raise ValueError(f"expected two (b,1,t,h,w) tensors, got {tuple(Xt.shape)} and {tuple(Yt.shape)}")
    return _SSIMLossFn.apply(Xt[:, 0], Yt[:, 0], win_size, k1, k2)


# ------------------------------------------------------------------ utils/evaluate.py
def _err_stats(gt: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    _need_cuda(gt, pred)
    if gt.shape != pred.shape:
        raise ValueError("Ground truth dimensions does not match pred.")
    gt, pred = _f32c(gt), _f32c(pred)
    out = torch.empty(4, dtype=torch.float32, device=gt.device)
    scratch = torch.empty(_ERR_SCRATCH, dtype=torch.float32, device=gt.device)
    _lib.check(_lib.lib().b2s_err_stats(_p(gt), _p(pred), gt.numel(), _p(out), _p(scratch), _stream()), "err_stats")
    return out


def mse(gt, pred):
    """utils/evaluate.py:6-8."""
    s = _err_stats(gt, pred)
    return s[0] / s[3]


def nmse(gt, pred):
    """utils/evaluate.py:11-13."""
    s = _err_stats(gt, pred)
    return s[0] / s[1]


def psnr(gt, pred, maxval=None):
    """utils/evaluate.py:16-22 (skimage peak_signal_noise_ratio: 10 log10(range^2 / mse))."""
    s = _err_stats(gt, pred)
    mx = s[2] if maxval is None else torch.as_tensor(maxval, dtype=torch.float32, device=s.device)
    return 10.0 * torch.log10(mx * mx * s[3] / s[0])


def ssim(gt, pred, maxval=None):
    """utils/evaluate.py:25-42: gt, pred (t,h,w); mean over frames of skimage's structural_similarity with
    its defaults (7x7 uniform window, sample covariance, K1 0.01, K2 0.03) and one data_range per volume."""
    if not gt.dim() == 3:
        raise ValueError("Unexpected number of dimensions in ground truth.")
    if not gt.dim() == pred.dim():
        raise ValueError("Ground truth dimensions does not match pred.")
    _need_cuda(gt, pred)
    gt, pred = _f32c(gt), _f32c(pred)
    t, h, w = gt.shape
    dr = (raw_frame_max(gt.reshape(1, 1, t * h, w)) if maxval is None
          else torch.as_tensor(maxval, dtype=torch.float32, device=gt.device).reshape(1))
    out = raw_ssim_fwd(gt.unsqueeze(0), pred.unsqueeze(0), dr.contiguous(), 0, 7, 0.01, 0.03)
    return out[:-1].mean()


METRIC_FUNCS = dict(MSE=mse, NMSE=nmse, PSNR=psnr, SSIM=ssim)
