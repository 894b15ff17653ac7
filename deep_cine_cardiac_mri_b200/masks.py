"""Undersampling masks and `apply_mask` of the reference's data front-end.

* `RandomMaskFunc` / `EquispacedMaskFunc` / `create_mask_for_mask_type` restate data/subsample.py:75-235 with the
  reference's exact random-number calls, so a given seed yields the SAME mask as the reference (tests pin this against
  the reference's own classes).  Mask generation stays on the host: it is ~60 draws per slice from numpy's legacy
  Mersenne-Twister stream (`np.random.choice` with a pdf), and reproducing that stream is what parity means here - a
  device RNG would produce different (if equally valid) masks.  Note the reference's quirk: `RandomMaskFunc` seeds
  only the choice of (center lines, acceleration) through its own `RandomState`; the rows themselves are drawn from the
  GLOBAL numpy stream (`np.random.choice`, subsample.py:139).  Here the stream is explicit (`rng=`), with the global
  module as the default so that behaviour is identical.
* `apply_mask` (data/transforms.py:66-92) multiplies on the GPU (`b2s_apply_mask`: k * m + 0.0) and returns the mask in
  the reference's float layout; `apply_mask_u8` also returns the uint8 `(b,t,1,h,1,1)` mask the models consume
  (`mask.byte()`, transforms.py:343).
"""
from __future__ import annotations

import contextlib
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, ops


@contextlib.contextmanager
def temp_seed(rng, seed):
    """data/subsample.py:16-30."""
    if seed is None:
        yield
    else:
        state = rng.get_state()
        rng.seed(seed)
        try:
            yield
        finally:
            rng.set_state(state)


class MaskFunc:
    """data/subsample.py:33-72."""

    def __init__(self, center_fractions: Sequence[float], accelerations: Sequence[int]):
        if not len(center_fractions) == len(accelerations):
            raise ValueError("Number of center fractions should match number of accelerations")
        self.center_fractions = center_fractions
        self.accelerations = accelerations
        self.rng = np.random.RandomState()

    def __call__(self, shape, seed=None):
        raise NotImplementedError

    def choose_acceleration(self):
        choice = self.rng.randint(0, len(self.accelerations))
        return self.center_fractions[choice], self.accelerations[choice]


class RandomMaskFunc(MaskFunc):
    """data/subsample.py:75-151: per frame int(Nx/acc) - n_center rows drawn without replacement from a tail-adjusted
    Gaussian pdf + the n_center centre rows.  `rng`: the stream the rows are drawn from (default: numpy's global one, as
    in the reference)."""

    def __init__(self, center_fractions, accelerations, rng=None):
        super().__init__(center_fractions, accelerations)
        self.row_rng = rng if rng is not None else np.random

    def __call__(self, shape, seed=None) -> torch.Tensor:
        if len(shape) < 3:
            raise ValueError("Shape should have 3 or more dimensions")
        with temp_seed(self.rng, seed):
            sample_n, acc = self.choose_acceleration()
        N, Nc, Nx, Ny, Nch = shape
        pdf_x = np.exp(-(0.5 / (Nx / 10.) ** 2) * (np.arange(Nx) - Nx / 2) ** 2)
        lmda = Nx / (2. * acc)
        n_lines = int(Nx / acc)
        pdf_x += lmda * 1. / Nx
        if sample_n:
            pdf_x[Nx // 2 - sample_n // 2: Nx // 2 + sample_n // 2] = 0
            pdf_x /= np.sum(pdf_x)
            n_lines -= sample_n
        mask = np.zeros((N, Nx))
        for i in range(N):
            idx = self.row_rng.choice(Nx, n_lines, False, pdf_x)
            mask[i, idx] = 1
        if sample_n:
            mask[:, Nx // 2 - sample_n // 2: Nx // 2 + sample_n // 2] = 1
        mask_shape = [1 for _ in shape]
        mask_shape[-3] = Nx
        mask_shape[0] = N
        return torch.from_numpy(mask.reshape(*mask_shape).astype(np.float32))


class EquispacedMaskFunc(MaskFunc):
    """data/subsample.py:154-215."""

    def __call__(self, shape, seed=None) -> torch.Tensor:
        if len(shape) < 3:
            raise ValueError("Shape should have 3 or more dimensions")
        with temp_seed(self.rng, seed):
            center_fraction, acceleration = self.choose_acceleration()
            num_rows = shape[-3]
            num_low_freqs = int(round(num_rows * center_fraction))
            mask = np.zeros(num_rows, dtype=np.float32)
            pad = (num_rows - num_low_freqs + 1) // 2
            mask[pad: pad + num_low_freqs] = True
            adjusted_accel = (acceleration * (num_low_freqs - num_rows)) / (num_low_freqs * acceleration - num_rows)
            offset = self.rng.randint(0, round(adjusted_accel))
            accel_samples = np.arange(offset, num_rows - 1, adjusted_accel)
            accel_samples = np.around(accel_samples).astype(np.uint)
            mask[accel_samples] = True
            mask_shape = [1 for _ in shape]
            mask_shape[-3] = num_rows
            mask = torch.from_numpy(mask.reshape(*mask_shape).astype(np.float32))
        return mask


def create_mask_for_mask_type(mask_type_str: str, center_fractions, accelerations) -> MaskFunc:
    """data/subsample.py:218-235."""
    if mask_type_str == "random":
        return RandomMaskFunc(center_fractions, accelerations)
    if mask_type_str == "equispaced":
        return EquispacedMaskFunc(center_fractions, accelerations)
    raise Exception(f"{mask_type_str} not supported")


def raw_apply_mask(kspace: torch.Tensor, mask_u8: torch.Tensor) -> torch.Tensor:
    """kspace (n, c, h, w, 2) float32 CUDA, mask_u8 (n, h) uint8 -> kspace * mask + 0.0"""
    ops._need_cuda(kspace, mask_u8)
    kspace = ops._f32c(kspace)
    n, c, h, w, _ = kspace.shape
    out = torch.empty_like(kspace)
    _lib.check(_lib.lib().b2s_apply_mask(ops._p(kspace), ops._p(mask_u8), ops._p(out), n, c, h, w, ops._stream()), "apply_mask")
    return out


def apply_mask(data: torch.Tensor, mask_func: MaskFunc, seed: Optional[Union[int, Tuple[int, ...]]] = None):
    """data/transforms.py:66-92 for a CUDA k-space `data` (t, c, h, w, 2) (the dataset layout, mri_data.py:283-303) or
    (b, t, c, h, w, 2): returns (masked_data, mask) with the reference's float mask of shape (t,1,h,1,1)."""
    shape = np.array(data.shape)
    shape[1 if data.dim() == 5 else 2] = 1
    gen_shape = shape if data.dim() == 5 else shape[1:]
    if data.dim() == 6 and data.shape[0] != 1:
        raise ValueError("apply_mask: one slice at a time, as the reference's dataset transform")
    mask = mask_func(gen_shape, seed)                                   # (t,1,h,1,1) float32, host
    t, h = int(gen_shape[0]), int(gen_shape[-3])
    m8 = (mask.reshape(t, h) != 0).to(torch.uint8).to(data.device, non_blocking=True)
    d5 = data if data.dim() == 5 else data[0]
    out = raw_apply_mask(d5, m8)
    return (out if data.dim() == 5 else out.unsqueeze(0)), mask


def apply_mask_u8(data: torch.Tensor, mask_func: MaskFunc, seed=None):
    """`apply_mask` plus the (1,t,1,h,1,1) uint8 device mask the models take (transforms.py:343 `mask.byte()`)."""
    masked, mask = apply_mask(data, mask_func, seed)
    t, h = mask.shape[0], mask.shape[-3]
    m8 = (mask.reshape(1, t, 1, h, 1, 1) != 0).to(torch.uint8).to(data.device)
    return masked, m8
