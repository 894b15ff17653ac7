"""numpy restatement of the reference's SENSE / data-consistency path.

TEST INFRASTRUCTURE (checker), not product code.  Every function cites the
reference file:line it restates (paths relative to the reference checkout).
Arrays follow the reference's fastMRI convention: real dtype, trailing dimension
of size 2 = (re, im).  Pass float64 arrays to get the fp64 arbiter; float32
arrays are computed in single precision like the reference.

Parity status: unpinned by reference tests (it has none); pinned against the
reference's own torch outputs in tests/golden/ (see tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

_ERR_ONE = "Tensor does not have separate complex dim."
_ERR_TWO = "Tensors do not have separate complex dim."


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def _c(x: np.ndarray) -> np.ndarray:
    """(...,2) real -> complex view/copy of matching precision."""
    x = np.asarray(x)
    return x[..., 0] + 1j * x[..., 1] if x.dtype != np.float32 else (
        x[..., 0] + np.complex64(1j) * x[..., 1]).astype(np.complex64)


def _r(z: np.ndarray, dtype=None) -> np.ndarray:
    """complex -> (...,2) real."""
    out = np.stack((z.real, z.imag), axis=-1)
    return out.astype(dtype) if dtype is not None else out


# --------------------------------------------------------------------------- #
# utils/fftc.py
# --------------------------------------------------------------------------- #
def roll_one_dim(x, shift, dim):
    """utils/fftc.py:119-138 — circular shift of one axis (narrow + cat)."""
    n = x.shape[dim]
    shift = shift % n
    if shift == 0:
        return x
    idx = (np.arange(n) - shift) % n
    return np.take(x, idx, axis=dim)


def roll(x, shift, dim):
    """utils/fftc.py:141-163."""
    if len(shift) != len(dim):
        raise ValueError("len(shift) must match len(dim)")
    for s, d in zip(shift, dim):
        x = roll_one_dim(x, s, d)
    return x


def fftshift(x, dim=None):
    """utils/fftc.py:166-188 — shift by n//2."""
    if dim is None:
        dim = list(range(x.ndim))
    return roll(x, [x.shape[d] // 2 for d in dim], dim)


def ifftshift(x, dim=None):
    """utils/fftc.py:191-213 — shift by (n+1)//2."""
    if dim is None:
        dim = list(range(x.ndim))
    return roll(x, [(x.shape[d] + 1) // 2 for d in dim], dim)


def _np_norm(norm):
    # torch.fft: None == "backward"
    return "backward" if norm is None else norm


def _centered(data, axes_real, fn, norm):
    data = np.asarray(data)
    if data.shape[-1] != 2:
        raise ValueError(_ERR_ONE)
    x = ifftshift(data, dim=list(axes_real))
    axes_c = tuple(a + 1 for a in axes_real)        # real axis -k == complex axis -(k-1)
    z = fn(_c(x), axes=axes_c, norm=_np_norm(norm))
    return fftshift(_r(z, data.dtype), dim=list(axes_real))


def fft1c(data, norm="ortho"):
    """utils/fftc.py:5-29 — centred 1-D FFT over real-dim -2."""
    return _centered(data, (-2,), np.fft.fftn, norm)


def ifft1c(data, norm="ortho"):
    """utils/fftc.py:32-56."""
    return _centered(data, (-2,), np.fft.ifftn, norm)


def fft2c(data, norm="ortho"):
    """utils/fftc.py:59-83 — centred 2-D FFT over real-dims -3,-2."""
    return _centered(data, (-3, -2), np.fft.fftn, norm)


def ifft2c(data, norm="ortho"):
    """utils/fftc.py:86-110."""
    return _centered(data, (-3, -2), np.fft.ifftn, norm)


# --------------------------------------------------------------------------- #
# utils/math.py, utils/coil_combine.py
# --------------------------------------------------------------------------- #
def complex_mul(x, y):
    """utils/math.py:5-25."""
    x, y = np.asarray(x), np.asarray(y)
    if not (x.shape[-1] == y.shape[-1] == 2):
        raise ValueError(_ERR_TWO)
    re = x[..., 0] * y[..., 0] - x[..., 1] * y[..., 1]
    im = x[..., 0] * y[..., 1] + x[..., 1] * y[..., 0]
    return np.stack((re, im), axis=-1)


def complex_conj(x):
    """utils/math.py:28-45."""
    x = np.asarray(x)
    if x.shape[-1] != 2:
        raise ValueError(_ERR_ONE)
    return np.stack((x[..., 0], -x[..., 1]), axis=-1)


def complex_abs_sq(x):
    """utils/math.py:65-79."""
    x = np.asarray(x)
    if x.shape[-1] != 2:
        raise ValueError(_ERR_ONE)
    return (x ** 2).sum(axis=-1)


def complex_abs(x):
    """utils/math.py:48-62."""
    return np.sqrt(complex_abs_sq(x))


def rss(x, dim=0):
    """utils/coil_combine.py:5-18."""
    return np.sqrt((np.asarray(x) ** 2).sum(axis=dim))


def rss_complex(x, dim=0):
    """utils/coil_combine.py:21-34."""
    return np.sqrt(complex_abs_sq(x).sum(axis=dim))


def real_to_complex_multi_ch(x, n):
    """utils/math.py:97-118 — [re_0..re_{n-1}, im_0..im_{n-1}] -> complex (..., n)."""
    x = np.asarray(x)
    if x.shape[-1] != 2 * n:
        raise ValueError("Real and imaginary parts do not have the same size")
    return x[..., :n] + 1j * x[..., n:]


def complex_to_real_multi_ch(z):
    """utils/math.py:121-135."""
    return np.concatenate([z.real, z.imag], axis=-1)


# --------------------------------------------------------------------------- #
# SENSE operators (models/varnet.py, cinenet.py, xpdnet.py, recurrent_*.py)
# --------------------------------------------------------------------------- #
def sens_expand(x, sens_maps):
    """A: models/varnet.py:181-185 (== cinenet.py:106-110) — fft2c(S * x).

    x (b,t,1,h,w,2), sens (b,1,c,h,w,2) -> (b,t,c,h,w,2)."""
    return fft2c(complex_mul(x, sens_maps))


def sens_reduce(k, sens_maps, keepdim=True):
    """A^H: models/varnet.py:187-194 — sum_c conj(S) * ifft2c(k)."""
    x = ifft2c(k)
    return complex_mul(x, complex_conj(sens_maps)).sum(axis=2, keepdims=keepdim)


def softplus(lam):
    """nn.Softplus(beta=1) (models/varnet.py:176)."""
    lam = np.asarray(lam, dtype=np.float64)
    return np.log1p(np.exp(-np.abs(lam))) + np.maximum(lam, 0.0)


def dc_blend(model_term, ref_kspace, mask, v):
    """models/varnet.py:281-282 — (1-m)*k + m*(k + v*ref)/(1+v).

    mask is uint8 (b,t,1,h,1,1); (1 - mask) is uint8 arithmetic then promoted."""
    mask = np.asarray(mask)
    one_minus = (1 - mask).astype(model_term.dtype)
    m = mask.astype(model_term.dtype)
    v = np.asarray(v, dtype=model_term.dtype)
    return one_minus * model_term + m * (model_term + v * ref_kspace) / (1 + v)


def varnet_block(current_kspace, ref_kspace, mask, sens_maps, v, regulariser=None):
    """models/varnet.py:244-282 with the regulariser as a callable on the
    coil-combined image (b,t,1,h,w,2) -> same shape (identity if None)."""
    img = sens_reduce(current_kspace, sens_maps)
    if regulariser is not None:
        img = regulariser(img)
    return dc_blend(sens_expand(img, sens_maps), ref_kspace, mask, v)


def varnet_rnn_data_consistency(x, ref_kspace, mask, sens_maps, v):
    """models/recurrent_varnet.py:65-90 — x is (b,2,h,w,t); returns (b,2,h,w,t)."""
    img = np.transpose(x, (0, 4, 2, 3, 1))[:, :, None]           # b,t,1,h,w,2
    dc = dc_blend(sens_expand(img, sens_maps), ref_kspace, mask, v)
    out = sens_reduce(dc, sens_maps, keepdim=False)                # b,t,h,w,2
    return np.transpose(out, (0, 4, 2, 3, 1))


def apply_mask(k, mask):
    """`k * mask + 0.0` (cinenet.py:129, xpdnet.py:131,163, transforms.py:90)."""
    return k * np.asarray(mask).astype(k.dtype) + 0.0


def normal_op(x, mask, sens_maps, v):
    """models/cinenet.py:121-133 — A^H M A x + softplus(lambda) x."""
    k = apply_mask(sens_expand(x, sens_maps), mask)
    return sens_reduce(k, sens_maps) + np.asarray(v, dtype=x.dtype) * x


def conj_grad(x, b, mask, sens_maps, v, cg_iters):
    """models/cinenet.py:136-171 — CG on Hx=b; alpha/beta are detached floats,
    dots span the whole flattened batch."""
    r = b - normal_op(x, mask, sens_maps, v)
    p = r.copy()
    rs_old = float(np.dot(r.ravel(), r.ravel()))
    for _ in range(cg_iters):
        d = normal_op(p, mask, sens_maps, v)
        alpha = rs_old / float(np.dot(p.ravel(), d.ravel()))
        x = x + alpha * p
        r = r - alpha * d
        rs_new = float(np.dot(r.ravel(), r.ravel()))
        beta = rs_new / rs_old
        rs_old = rs_new
        p = r + beta * p
    return x


def forward_operator(image, mask, sens_maps, buffer_size, masked):
    """models/xpdnet.py:119-133 — acts on channels 0 and buffer_size."""
    img = np.stack([image[..., 0], image[..., buffer_size]], axis=-1)
    k = sens_expand(img, sens_maps)
    return apply_mask(k, mask) if masked else k


def backward_operator(kspace, mask, sens_maps, buffer_size, masked):
    """models/xpdnet.py:152-167."""
    k = np.stack([kspace[..., 0], kspace[..., buffer_size]], axis=-1)
    if masked:
        k = apply_mask(k, mask)
    return sens_reduce(k, sens_maps)


def measurements_residual(concat_kspace):
    """models/xpdnet.py:295-298 — packed [re_cur, re_ref, im_cur, im_ref]."""
    cur = np.stack([concat_kspace[..., 0], concat_kspace[..., 2]], axis=-1)
    ref = np.stack([concat_kspace[..., 1], concat_kspace[..., 3]], axis=-1)
    return cur - ref


def xpdnet_k_residual(image_buffer, mask, sens_maps, ref_kspace, i_buffer_size):
    """models/xpdnet.py:372-403 in primal-only mode: M A x0 - y."""
    return forward_operator(image_buffer, mask, sens_maps, i_buffer_size, True) - ref_kspace


# --------------------------------------------------------------------------- #
# SensitivityModel pre / post (models/varnet.py:58-86, xpdnet.py:69-100)
# --------------------------------------------------------------------------- #
def acs_window(mask):
    """models/varnet.py:64-68 — (pad, num_low_freqs) from frame 0 of batch 0."""
    mask = np.asarray(mask)
    h = mask.shape[-3]
    line = np.squeeze(mask[:, 0, :])                    # -> (h,) for b == 1
    cent = h // 2
    left = np.nonzero(line[:cent] == 0)[0][-1]
    right = np.nonzero(line[cent:] == 0)[0][0] + cent
    nlf = int(right - left)
    pad = (h - nlf + 1) // 2
    return pad, nlf


def mask_center(x, mask_from, mask_to):
    """data/transforms.py:95-108 — keep rows [from,to) of dim 2 of (b,c,h,w,2)."""
    out = np.zeros_like(x)
    out[:, :, mask_from:mask_to, :] = x[:, :, mask_from:mask_to, :]
    return out


def sens_model_pre(masked_kspace, mask):
    """models/varnet.py:64-74 — ACS low-pass of the time-mean, then ifft2c.
    Returns (b,c,h,w,2) coil images fed to the sensitivity U-Net."""
    pad, nlf = acs_window(mask)
    x = mask_center(masked_kspace.mean(axis=1), pad, pad + nlf)
    return ifft2c(x)


def divide_root_sum_of_squares(x):
    """models/varnet.py:58-59 — x / rss_complex(x, dim=1) (no epsilon)."""
    return x / rss_complex(x, dim=1)[..., None][:, None]


# --------------------------------------------------------------------------- #
# temporal transforms (xfyf_transform head/tail)
# --------------------------------------------------------------------------- #
def temporal_pre(image_combined, xf=True):
    """models/varnet.py:202-213 — subtract temporal mean, centred FFT over t.
    image (b,t,h,w,2) -> (x (b,t,h,w,2), mean (b,h,w,2))."""
    mean = image_combined.mean(axis=1)
    x = image_combined - mean[:, None]
    if xf:
        x = np.transpose(fft1c(np.transpose(x, (0, 2, 3, 1, 4))), (0, 3, 1, 2, 4))
    return x, mean


def temporal_post(out, mean, xf=True):
    """models/varnet.py:234-241 — out (b,t,1,h,w,2), inverse temporal FFT + mean."""
    if xf:
        o = np.transpose(out, (0, 2, 3, 4, 1, 5))
        out = np.transpose(ifft1c(o), (0, 4, 1, 2, 3, 5))
    return out + mean[:, None, None]


def xpd_temporal_fft(x, n_ch):
    """models/xpdnet.py:465-467 — ifftshift(fft(fftshift(x,1), t, 1, 'ortho'), 1)
    on packed real channels (b,t,h,w,2n)."""
    z = real_to_complex_multi_ch(x, n_ch)
    z = np.fft.ifftshift(np.fft.fft(np.fft.fftshift(z, axes=1), axis=1, norm="ortho"), axes=1)
    return complex_to_real_multi_ch(z).astype(x.dtype)


def xpd_temporal_ifft(x, n_ch):
    """models/xpdnet.py:499-501 — fftshift(ifft(ifftshift(x,1), t, 1, 'ortho'), 1)."""
    z = real_to_complex_multi_ch(x, n_ch)
    z = np.fft.fftshift(np.fft.ifft(np.fft.ifftshift(z, axes=1), axis=1, norm="ortho"), axes=1)
    return complex_to_real_multi_ch(z).astype(x.dtype)


# --------------------------------------------------------------------------- #
# metrics (utils/evaluate.py:6-42, utils/losses.py:25-58) — skimage is absent here
# --------------------------------------------------------------------------- #
def nmse(gt, pred):
    """utils/evaluate.py:11-13."""
    return np.linalg.norm(gt - pred) ** 2 / np.linalg.norm(gt) ** 2


def psnr(gt, pred, maxval=None):
    """utils/evaluate.py:16-22 (skimage peak_signal_noise_ratio formula)."""
    maxval = gt.max() if maxval is None else maxval
    err = np.mean((gt.astype(np.float64) - pred.astype(np.float64)) ** 2)
    return 10 * np.log10(maxval ** 2 / err)


def _box7(a, win):
    c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), axis=0), axis=1)
    return (c[win:, win:] - c[:-win, win:] - c[win:, :-win] + c[:-win, :-win]) / (win * win)


def ssim(gt, pred, maxval=None, win=7, k1=0.01, k2=0.03):
    """Time-averaged SSIM, 7x7 uniform window, sample covariance
    (utils/losses.py:25-58 formula == skimage structural_similarity defaults,
    utils/evaluate.py:25-42). gt, pred: (t,h,w)."""
    if gt.ndim != 3:
        raise ValueError("Unexpected number of dimensions in ground truth.")
    maxval = gt.max() if maxval is None else maxval
    cov_norm = win * win / (win * win - 1)
    c1, c2 = (k1 * maxval) ** 2, (k2 * maxval) ** 2
    tot = 0.0
    for x, y in zip(gt.astype(np.float64), pred.astype(np.float64)):
        ux, uy = _box7(x, win), _box7(y, win)
        vx = cov_norm * (_box7(x * x, win) - ux * ux)
        vy = cov_norm * (_box7(y * y, win) - uy * uy)
        vxy = cov_norm * (_box7(x * y, win) - ux * uy)
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        tot += s.mean()
    return tot / gt.shape[0]


def ssim_loss(Xt, Yt, win=7, k1=0.01, k2=0.03):
    """SSIMLoss.forward, utils/losses.py:25-58: Xt, Yt (b,1,t,h,w); per frame t the data range is the maximum
    of the target frame over the WHOLE batch (`torch.Tensor([Y.max()])`, :35 - the `data_range` argument is
    overwritten), S from five win x win uniform 'valid' convolutions with the sample-covariance factor
    NP/(NP-1), loss = mean_t (1 - mean S).  Returns (loss, per-frame mean S), fp64."""
    Xt = np.asarray(Xt, dtype=np.float64)
    Yt = np.asarray(Yt, dtype=np.float64)
    cov_norm = win * win / (win * win - 1)
    nt = Xt.shape[2]
    means = np.zeros(nt)
    for t in range(nt):
        data_range = Yt[:, :, t].max()
        c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
        acc, cnt = 0.0, 0
        for b in range(Xt.shape[0]):
            x, y = Xt[b, 0, t], Yt[b, 0, t]
            ux, uy = _box7(x, win), _box7(y, win)
            vx = cov_norm * (_box7(x * x, win) - ux * ux)
            vy = cov_norm * (_box7(y * y, win) - uy * uy)
            vxy = cov_norm * (_box7(x * y, win) - ux * uy)
            s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
            acc += s.sum()
            cnt += s.size
        means[t] = acc / cnt
    return float(np.mean(1.0 - means)), means
