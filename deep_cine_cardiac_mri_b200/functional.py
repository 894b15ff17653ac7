"""fastMRI-style functional API of the reference, backed by the sm_100a kernels.

Signature-, shape- and error-compatible with `reconstruction/utils/__init__.py:1-25`
(fftc.py, math.py, coil_combine.py): real tensors whose last dimension is the
(re, im) pair, arbitrary leading batch dimensions, out-of-place, inputs never
mutated.  CUDA float32 only — CPU tensors raise instead of silently falling back.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops

_ERR_ONE = "Tensor does not have separate complex dim."
_ERR_TWO = "Tensors do not have separate complex dim."


# ------------------------------- fftc.py ----------------------------------- #
def fft1c(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/fftc.py:5-29 — centred 1-D FFT over dim -2."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.fft1c(data, norm, inverse=False)


def ifft1c(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/fftc.py:32-56."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.fft1c(data, norm, inverse=True)


def fft2c(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/fftc.py:59-83 — centred 2-D FFT over dims -3, -2."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    if data.dim() < 3:
        raise IndexError("Dimension out of range (fft2c needs at least 3 dimensions)")
    return ops.fft2c(data, norm, inverse=False)


def ifft2c(data: torch.Tensor, norm: str = "ortho") -> torch.Tensor:
    """utils/fftc.py:86-110."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    if data.dim() < 3:
        raise IndexError("Dimension out of range (ifft2c needs at least 3 dimensions)")
    return ops.fft2c(data, norm, inverse=True)


def roll_one_dim(x: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
    """utils/fftc.py:119-138 (index plumbing; the fused kernels never call it)."""
    shift = shift % x.size(dim)
    if shift == 0:
        return x
    return torch.roll(x, shift, dim)


def roll(x: torch.Tensor, shift: List[int], dim: List[int]) -> torch.Tensor:
    """utils/fftc.py:141-163."""
    if len(shift) != len(dim):
        raise ValueError("len(shift) must match len(dim)")
    for (s, d) in zip(shift, dim):
        x = roll_one_dim(x, s, d)
    return x


def fftshift(x: torch.Tensor, dim: Optional[List[int]] = None) -> torch.Tensor:
    """utils/fftc.py:166-188."""
    if dim is None:
        dim = list(range(x.dim()))
    return roll(x, [x.shape[d] // 2 for d in dim], dim)


def ifftshift(x: torch.Tensor, dim: Optional[List[int]] = None) -> torch.Tensor:
    """utils/fftc.py:191-213."""
    if dim is None:
        dim = list(range(x.dim()))
    return roll(x, [(x.shape[d] + 1) // 2 for d in dim], dim)


# ------------------------------- math.py ----------------------------------- #
def complex_mul(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """utils/math.py:5-25 (broadcasting)."""
    if not x.shape[-1] == y.shape[-1] == 2:
        raise ValueError(_ERR_TWO)
    return ops.ComplexMulFn.apply(x, y)


def complex_conj(x: torch.Tensor) -> torch.Tensor:
    """utils/math.py:28-45."""
    if not x.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.ComplexConjFn.apply(x)


def complex_abs(data: torch.Tensor) -> torch.Tensor:
    """utils/math.py:48-62."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.ComplexAbsFn.apply(data, False)


def complex_abs_sq(data: torch.Tensor) -> torch.Tensor:
    """utils/math.py:65-79."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.ComplexAbsFn.apply(data, True)


def tensor_to_complex_np(data: torch.Tensor) -> np.ndarray:
    """utils/math.py:82-94."""
    data = data.numpy()
    return data[..., 0] + 1j * data[..., 1]


def real_to_complex_multi_ch(x: torch.Tensor, n: int) -> torch.Tensor:
    """utils/math.py:97-118 — XPDNet buffer packing (boundary glue)."""
    if not x.shape[-1] == 2 * n:
        raise ValueError("Real and imaginary parts do not have the same size")
    return torch.complex(x[..., :n], x[..., n:])


def complex_to_real_multi_ch(x: torch.Tensor) -> torch.Tensor:
    """utils/math.py:121-135."""
    return torch.cat([x.real, x.imag], dim=-1)


# ---------------------------- coil_combine.py ------------------------------ #
def rss(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils/coil_combine.py:5-18."""
    return ops.RssFn.apply(data, dim, False)


def rss_complex(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils/coil_combine.py:21-34."""
    if not data.shape[-1] == 2:
        raise ValueError(_ERR_ONE)
    return ops.RssFn.apply(data, dim, True)


__all__ = [
    "fft1c", "ifft1c", "fft2c", "ifft2c", "fftshift", "ifftshift", "roll", "roll_one_dim",
    "complex_mul", "complex_conj", "complex_abs", "complex_abs_sq", "tensor_to_complex_np",
    "real_to_complex_multi_ch", "complex_to_real_multi_ch", "rss", "rss_complex",
]
