"""Pin the numpy oracle against fixtures produced by the reference's own torch
code (tests/golden/make_golden.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest

import make_golden as G
from oracle import sense_oracle as O

GOLD = np.load(G.HERE / "golden_v1.npz")
TOL = 2e-6      # fp32 reference vs fp64 oracle, relative to max|ref| (~2e-7..5e-7)


def check(name, arr, tol=TOL):
    arr = np.asarray(arr)
    assert tuple(GOLD[f"{name}/shape"]) == arr.shape, name
    flat = arr.astype(np.float64).ravel()
    samp = GOLD[f"{name}/sample"].astype(np.float64)
    got = flat[G.sample_index(flat.size)]
    scale = max(np.abs(samp).max(), 1e-30)
    assert np.abs(got - samp).max() / scale <= tol, name
    # checksums over the whole tensor
    assert abs((flat ** 2).sum() - GOLD[f"{name}/sumsq"]) <= 1e-5 * GOLD[f"{name}/sumsq"] + 1e-12, name
    assert abs(flat.sum() - GOLD[f"{name}/total"]) <= 1e-4 * np.sqrt(GOLD[f"{name}/sumsq"] * flat.size) + 1e-6, name


@pytest.mark.parametrize("i", range(len(G.SMALL_FFT_SHAPES)))
@pytest.mark.parametrize("norm", ["ortho", None, "forward"])
def test_fft2c_small(i, norm):
    x = G.rng_normal(100 + i, G.SMALL_FFT_SHAPES[i]).astype(np.float64)
    check(f"fft2c/{i}/{norm}", O.fft2c(x, norm=norm))
    check(f"ifft2c/{i}/{norm}", O.ifft2c(x, norm=norm))


@pytest.mark.parametrize("i", range(len(G.FFT1_SHAPES)))
def test_fft1c(i):
    x = G.rng_normal(200 + i, G.FFT1_SHAPES[i]).astype(np.float64)
    check(f"fft1c/{i}", O.fft1c(x))
    check(f"ifft1c/{i}", O.ifft1c(x))


def test_fp32_oracle_matches_too():
    x = G.rng_normal(100, G.SMALL_FFT_SHAPES[0])
    out = O.fft2c(x)
    assert out.dtype == np.float32
    check("fft2c/0/ortho", out, tol=5e-6)


def test_pointwise():
    x = G.rng_normal(300, (2, 3, 5, 6, 2)).astype(np.float64)
    y = G.rng_normal(301, (2, 1, 5, 6, 2)).astype(np.float64)
    check("complex_mul", O.complex_mul(x, y))
    check("complex_conj", O.complex_conj(x))
    check("complex_abs", O.complex_abs(x))
    check("complex_abs_sq", O.complex_abs_sq(x))
    check("rss", O.rss(x, dim=1))
    check("rss_complex", O.rss_complex(x, dim=1))
    check("fftshift", O.fftshift(x, dim=[-3, -2]))
    check("ifftshift", O.ifftshift(G.rng_normal(302, (3, 7, 5, 2)), dim=[-3, -2]))


def test_errors_match_reference_messages():
    bad = np.zeros((4, 4, 3))
    for fn in (O.fft2c, O.ifft2c, O.fft1c, O.ifft1c, O.complex_conj, O.complex_abs, O.complex_abs_sq):
        with pytest.raises(ValueError, match="Tensor does not have separate complex dim."):
            fn(bad)
    with pytest.raises(ValueError, match="Tensors do not have separate complex dim."):
        O.complex_mul(bad, bad)
    with pytest.raises(ValueError, match="len\\(shift\\) must match len\\(dim\\)"):
        O.roll(bad, [1, 2], [0])


@pytest.mark.parametrize("tag", list(G.BIG))
def test_block_tier(tag):
    b, t, c, h, w = G.BIG[tag]
    cs = {k: (v.astype(np.float64) if v.dtype == np.float32 and v.ndim else v)
          for k, v in G.sense_case(1000 + len(tag) + h, b, t, c, h, w).items()}
    img, k, ref, sens, mask = cs["img"], cs["k"], cs["ref"], cs["sens"], cs["mask"]
    v = O.softplus(cs["lam"])
    check(f"{tag}/softplus", np.array([v]))
    check(f"{tag}/fft2c", O.fft2c(k))
    check(f"{tag}/ifft2c", O.ifft2c(k))
    check(f"{tag}/ifft2c_backward", O.ifft2c(k, norm=None))
    check(f"{tag}/sens_expand", O.sens_expand(img, sens))
    check(f"{tag}/sens_reduce", O.sens_reduce(k, sens))
    check(f"{tag}/dc_blend", O.dc_blend(k, ref, mask, v))
    check(f"{tag}/normal_op", O.normal_op(img, mask, sens, v))
    rhs = O.sens_reduce(O.apply_mask(ref, mask), sens) + v * img
    check(f"{tag}/conj_grad", O.conj_grad(img, rhs, mask, sens, v, 4), tol=2e-5)
    ibuf = np.repeat(img, 5, axis=-1)
    check(f"{tag}/xpd_forward", O.forward_operator(ibuf, mask, sens, 5, True))
    check(f"{tag}/xpd_backward", O.backward_operator(k, mask, sens, 1, True))
    if b == 1:
        check(f"{tag}/varnet_block", O.varnet_block(k, ref, mask, sens, v))
        mk = O.apply_mask(k, mask)
        pre = O.sens_model_pre(mk, mask)
        check(f"{tag}/sens_model_pre", pre[:, None])
        check(f"{tag}/sens_model", O.divide_root_sum_of_squares(pre)[:, None], tol=2e-5)
    x, mean = O.temporal_pre(img[:, :, 0])
    check(f"{tag}/temporal_pre", x)
    check(f"{tag}/temporal_post", O.temporal_post(img, mean))
    pk = np.concatenate([ibuf, ibuf[..., :1], ibuf[..., 5:6]], axis=-1)[:, :, 0]
    check(f"{tag}/xpd_tfft", O.xpd_temporal_fft(pk, 6))
    check(f"{tag}/xpd_tifft", O.xpd_temporal_ifft(ibuf[:, :, 0], 5))


def test_acs_window_default_mask():
    # 10 centre lines at h=200 -> left 94, right 105, nlf 11, rows 95..105
    m = G.make_mask(7, 1, 3, 200)
    m[0, 0, 0, 90:95] = 0
    m[0, 0, 0, 105:110] = 0
    assert O.acs_window(m) == (95, 11)


@pytest.mark.parametrize("name", sorted(G.LOSS_CASES))
def test_ssim_loss_oracle_matches_reference_golden(name):
    """oracle.ssim_loss (fp64 restatement of utils/losses.py:25-58) vs the reference's own SSIMLoss run in
    fp64 and fp32 (tests/golden/make_golden.py loss)."""
    z = np.load(Path(__file__).parent / "golden" / "golden_v2_loss.npz")
    pred, tgt = G.loss_case(name)
    loss, means = O.ssim_loss(pred[:, None], tgt[:, None])
    assert abs(loss - float(z[f"{name}/f64/loss"])) <= 1e-9
    assert abs(loss - float(z[f"{name}/f32/loss"])) <= 5e-6          # the reference's own fp32 noise
    assert means.shape == (pred.shape[1],)
    # evaluate.py's ssim is the same formula with one data range per volume
    if pred.shape[2] >= 7:
        v = O.ssim(tgt[0], pred[0])
        assert 0.0 < v <= 1.0
