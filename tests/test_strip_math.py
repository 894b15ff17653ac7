"""CPU restatement of the index / sign / twiddle algebra of the strip-streamed kernels (csrc/strip_core.cuh) against the
oracle's centred transforms: two-step N = N1*N2 decomposition with the centring folded into a rotation of the N1-point
outputs and a signed twiddle table, the in-thread N2 = A*B split, and the strip-major scratch layout.  The CUDA kernels
themselves are covered by the GPU parity tests (`fused_path` fixture); this keeps their arithmetic checkable without a GPU."""
import numpy as np
import pytest

from oracle import sense_oracle as O

PLANS = {200: (8, 25, 5, 5), 256: (16, 16, 4, 4)}       # StripDim<N>: N1, N2, A, B


def dft(x):
    n = x.shape[-1]
    k = np.arange(n)
    return x @ np.exp(-2j * np.pi * np.outer(k, k) / n)


def strip_fftc_1d(x, scale=1.0):
    """Centred forward transform of the last axis the way one pass of the strip kernel computes it."""
    n = x.shape[-1]
    n1, n2, a, b = PLANS[n]
    sh = n1 // 2 if n2 % 2 else 0
    kk1, nn2 = np.meshgrid(np.arange(n1), np.arange(n2), indexing="ij")
    table = scale * (-1.0) ** (nn2 + kk1) * np.exp(-2j * np.pi * nn2 * kk1 / n)          # T[k1][n2], strip_build_table
    # step 1: N1-point codelets over stride-N2 elements, rotation by SH, twiddle -> X[k1][n2]
    xs = x.reshape(x.shape[:-1] + (n1, n2))                                               # [n1][n2], n = N2*n1 + n2
    y = dft(np.moveaxis(xs, -2, -1))                                                      # [n2][k1']
    y = np.roll(y, -sh, axis=-1)                                                          # Y[k1] = A[(k1 + SH) % N1]
    X = np.moveaxis(y, -1, -2) * table                                                    # [k1][n2]
    # step 2a: A-point codelets over n2 = B*j + b, constant twiddles W_N2^(b ka), in place at [ka*B + b]
    Xa = X.reshape(X.shape[:-1] + (a, b))                                                 # [j][b]
    Ya = dft(np.moveaxis(Xa, -2, -1))                                                     # [b][ka]
    bb, ka = np.meshgrid(np.arange(b), np.arange(a), indexing="ij")
    Ya = Ya * np.exp(-2j * np.pi * bb * ka / n2)
    Z = np.moveaxis(Ya, -1, -2)                                                           # [ka][b]
    # step 2b: B-point codelets over b -> output k2 = ka + A*kb, overall k = k1 + N1*k2
    out = dft(Z)                                                                          # [k1][ka][kb]
    res = np.zeros(x.shape, dtype=complex)
    for k1 in range(n1):
        for ia in range(a):
            for ib in range(b):
                res[..., k1 + n1 * (ia + a * ib)] = out[..., k1, ia, ib]
    return res * (-1.0) ** (n // 2)


@pytest.mark.parametrize("n", sorted(PLANS))
def test_one_pass_equals_the_centred_transform(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
    want = np.fft.fftshift(np.fft.fft(np.fft.ifftshift(x, axes=-1), axis=-1, norm="ortho"), axes=-1)
    got = strip_fftc_1d(x, scale=1.0 / np.sqrt(n))
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("n", sorted(PLANS))
def test_two_passes_through_the_strip_major_scratch_equal_fft2c(n):
    """pass R over rows -> scratch[(col // 4)][row][col % 4] -> pass C over 4-column strips == oracle fft2c; the inverse
    runs the same machinery on re/im-swapped data."""
    rng = np.random.default_rng(7 + n)
    img = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    rows = strip_fftc_1d(img)                                                             # pass R, table scale 1
    scratch = rows.reshape(n, n // 4, 4).transpose(1, 0, 2).copy()                        # [strip][row][4]
    out = np.empty_like(img)
    for s in range(n // 4):
        out[:, 4 * s:4 * s + 4] = strip_fftc_1d(scratch[s].T, scale=1.0 / n).T            # pass C carries the ortho scale
    ri = np.stack([img.real, img.imag], -1)
    want = O.fft2c(ri)
    assert np.abs(out - (want[..., 0] + 1j * want[..., 1])).max() <= 1e-12 * np.abs(want).max()
    swap = lambda z: z.imag + 1j * z.real
    inv_rows = strip_fftc_1d(swap(img))
    inv = np.empty_like(img)
    for s in range(n // 4):
        inv[:, 4 * s:4 * s + 4] = strip_fftc_1d(inv_rows[:, 4 * s:4 * s + 4].T.copy(), scale=1.0 / n).T
    want_i = O.ifft2c(ri)
    assert np.abs(swap(inv) - (want_i[..., 0] + 1j * want_i[..., 1])).max() <= 1e-12 * np.abs(want_i).max()
