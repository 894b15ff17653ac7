#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE's own torch code.

Runs only in the authoring container (needs /root/reference, which does not
exist on the GPU box).  Inputs are never stored: every case regenerates them
from `case_inputs(name)` (numpy PCG64, seeded), so tests rebuild identical
inputs anywhere.  Outputs are stored in full for small shapes and as a fixed
pseudo-random sample of 4096 entries + two checksums for 200x200 / 256x256.

    python tests/golden/make_golden.py      # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
N_SAMPLE = 4096


# --------------------------------------------------------------------------- #
# shared with the tests (tests import this module for `case_inputs`/`digest`)
# --------------------------------------------------------------------------- #
def rng_normal(seed, shape):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def make_mask(seed, b, t, h, n_center=10, acc=4):
    """Row mask with RandomMaskFunc's conventions (data/subsample.py:117-151):
    per frame int(h/acc)-n_center random rows + n_center centre rows, uint8
    (b,t,1,h,1,1)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((b, t, h), dtype=np.uint8)
    pdf = np.exp(-(0.5 / (h / 10.0) ** 2) * (np.arange(h) - h / 2) ** 2) + (h / (2.0 * acc)) / h
    n_center = min(n_center, max(h // 4, 1))
    pdf[h // 2 - n_center // 2: h // 2 + n_center // 2] = 0
    pdf /= pdf.sum()
    n_lines = max(int(h / acc) - n_center, 0)
    for i in range(b):
        for j in range(t):
            m[i, j, rng.choice(h, n_lines, replace=False, p=pdf)] = 1
    m[:, :, h // 2 - n_center // 2: h // 2 + n_center // 2] = 1
    return m.reshape(b, t, 1, h, 1, 1)


def make_sens(seed, b, c, h, w):
    """randn coil maps normalised to RSS == 1 (never zero at a pixel)."""
    s = rng_normal(seed, (b, 1, c, h, w, 2))
    nrm = np.sqrt((s ** 2).sum(axis=(2, 5), keepdims=True))
    return (s / nrm).astype(np.float32)


def sense_case(seed, b, t, c, h, w):
    return dict(
        img=rng_normal(seed, (b, t, 1, h, w, 2)),
        k=rng_normal(seed + 1, (b, t, c, h, w, 2)),
        ref=rng_normal(seed + 2, (b, t, c, h, w, 2)),
        sens=make_sens(seed + 3, b, c, h, w),
        mask=make_mask(seed + 4, b, t, h),
        lam=np.float32(0.3),
    )


def sample_index(n, seed=12345):
    if n <= N_SAMPLE:
        return np.arange(n)
    return np.sort(np.random.default_rng(seed).choice(n, N_SAMPLE, replace=False))


def digest(arr):
    """What is stored for one output tensor."""
    a = np.asarray(arr, dtype=np.float32).ravel()
    idx = sample_index(a.size)
    return dict(shape=np.array(arr.shape), sample=a[idx],
                sumsq=np.float64((a.astype(np.float64) ** 2).sum()),
                total=np.float64(a.astype(np.float64).sum()))


SMALL_FFT_SHAPES = [(2, 3, 12, 10, 2), (3, 5, 7, 2), (1, 2, 16, 16, 2), (2, 9, 8, 2), (4, 20, 30, 2)]
FFT1_SHAPES = [(3, 4, 15, 2), (2, 5, 16, 2), (2, 25, 2), (6, 17, 2), (2, 3, 30, 2)]
BIG = dict(b200=(1, 3, 4, 200, 200), b256=(1, 2, 3, 256, 256), rag=(2, 2, 3, 200, 200))


# --------------------------------------------------------------------------- #
def _load_reference():
    sys.path.insert(0, "/root/reference")
    for name in ("bart", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torch  # noqa
    import reconstruction.utils  # noqa  (rec.utils must be imported explicitly)
    import reconstruction.models as models
    import reconstruction as rec
    return rec, models


def main():
    import torch
    rec, models = _load_reference()
    U = rec.utils
    T = torch.from_numpy
    out = {}

    def put(name, tensor):
        for k, v in digest(tensor.detach().numpy()).items():
            out[f"{name}/{k}"] = v

    # --- functional tier, small shapes, all norms (utils/fftc.py) ------------
    for i, shp in enumerate(SMALL_FFT_SHAPES):
        x = T(rng_normal(100 + i, shp))
        for norm in ("ortho", None, "forward"):
            put(f"fft2c/{i}/{norm}", U.fft2c(x, norm=norm))
            put(f"ifft2c/{i}/{norm}", U.ifft2c(x, norm=norm))
    for i, shp in enumerate(FFT1_SHAPES):
        x = T(rng_normal(200 + i, shp))
        put(f"fft1c/{i}", U.fft1c(x))
        put(f"ifft1c/{i}", U.ifft1c(x))
    x = T(rng_normal(300, (2, 3, 5, 6, 2)))
    y = T(rng_normal(301, (2, 1, 5, 6, 2)))
    put("complex_mul", U.complex_mul(x, y))
    put("complex_conj", U.complex_conj(x))
    put("complex_abs", U.complex_abs(x))
    put("complex_abs_sq", U.complex_abs_sq(x))
    put("rss", U.rss(x, dim=1))
    put("rss_complex", U.rss_complex(x, dim=1))
    put("fftshift", U.fftshift(x, dim=[-3, -2]))
    put("ifftshift", U.ifftshift(T(rng_normal(302, (3, 7, 5, 2))), dim=[-3, -2]))

    # --- block tier at real sizes (models/*.py) ------------------------------
    ident = torch.nn.Identity()
    for tag, (b, t, c, h, w) in BIG.items():
        cs = sense_case(1000 + len(tag) + h, b, t, c, h, w)
        img, k, ref, sens = T(cs["img"]), T(cs["k"]), T(cs["ref"]), T(cs["sens"])
        mask = T(cs["mask"])
        vb = models.VarNetBlock(ident, "2D", False)
        with torch.no_grad():
            vb.lambda_reg.fill_(float(cs["lam"]))
            put(f"{tag}/fft2c", U.fft2c(k))
            put(f"{tag}/ifft2c", U.ifft2c(k))
            put(f"{tag}/ifft2c_backward", U.ifft2c(k, norm=None))
            put(f"{tag}/sens_expand", vb.sens_expand(img, sens))
            put(f"{tag}/sens_reduce", vb.sens_reduce(k, sens))
            v = vb.Softplus(vb.lambda_reg)
            put(f"{tag}/softplus", v)
            put(f"{tag}/dc_blend", (1 - mask) * k + mask * (k + v * ref) / (1 + v))
            if b == 1:   # '2D' VarNetBlock.forward squeezes batch (varnet.py:259-268)
                put(f"{tag}/varnet_block", vb(k, ref, mask, sens))
            cb = models.CineNetBlock(ident, 4, "2D", False)
            cb.lambda_reg.fill_(float(cs["lam"]))
            put(f"{tag}/normal_op", cb.HOperator(img, mask, sens))
            rhs = vb.sens_reduce(ref * mask + 0.0, sens) + cb.Softplus(cb.lambda_reg) * img
            put(f"{tag}/conj_grad", cb.ConjGrad(img, rhs, mask, sens, 4))
            # XPDNet operators on a 5-deep packed buffer (xpdnet.py:104-167)
            ibuf = torch.repeat_interleave(img, 5, dim=-1)
            fo, bo = models.xpdnet.ForwardOperator(masked=True), models.xpdnet.BackwardOperator(masked=True)
            put(f"{tag}/xpd_forward", fo(ibuf, mask, sens, 5))
            put(f"{tag}/xpd_backward", bo(k, mask, sens, 1))
            if b == 1:
                sm = models.SensitivityModel(8, 4)
                sm.norm_unet = ident                         # regulariser out of scope
                mk = k * mask + 0.0
                put(f"{tag}/sens_model", sm(mk, mask))
                # pre-part only (mean_t -> mask_center -> ifft2c), varnet.py:64-74
                sm2 = models.SensitivityModel(8, 4)
                sm2.norm_unet = ident
                sm2.divide_root_sum_of_squares = lambda z: z
                put(f"{tag}/sens_model_pre", sm2(mk, mask))
            # temporal transforms: varnet.py:202-213 / 234-241
            ic = img.squeeze(2)
            mean = torch.stack(t * [torch.mean(ic.clone(), dim=1)], dim=1)
            xf = U.fft1c((ic - mean).permute(0, 2, 3, 1, 4)).permute(0, 3, 1, 2, 4)
            put(f"{tag}/temporal_pre", xf)
            o = U.ifft1c(img.permute(0, 2, 3, 4, 1, 5)).permute(0, 4, 1, 2, 3, 5) + mean.unsqueeze(2)
            put(f"{tag}/temporal_post", o)
            # XPDNet XF variant call sites xpdnet.py:465-467 / 499-501 (raw torch.fft, odd t quirk)
            pk = torch.cat([ibuf, ibuf[..., :1], ibuf[..., 5:6]], dim=-1).squeeze(2)   # 12 packed ch
            z = U.real_to_complex_multi_ch(pk, 6)
            z = torch.fft.ifftshift(torch.fft.fft(torch.fft.fftshift(z, 1), t, 1, "ortho"), 1)
            put(f"{tag}/xpd_tfft", U.complex_to_real_multi_ch(z))
            z2 = U.real_to_complex_multi_ch(ibuf.squeeze(2), 5)
            z2 = torch.fft.fftshift(torch.fft.ifft(torch.fft.ifftshift(z2, 1), t, 1, "ortho"), 1)
            put(f"{tag}/xpd_tifft", U.complex_to_real_multi_ch(z2))

    np.savez_compressed(HERE / "golden_v1.npz", **out)
    print(f"wrote {len(out)} arrays -> {HERE/'golden_v1.npz'} "
          f"({(HERE/'golden_v1.npz').stat().st_size/1e6:.2f} MB)")


# --------------------------------------------------------------------------- #
# SSIM loss (utils/losses.py) - second fixture file, golden_v2_loss.npz
# --------------------------------------------------------------------------- #
LOSS_CASES = {"a": (2, 3, 40, 36), "one": (1, 1, 7, 9), "full": (1, 15, 200, 200)}   # (b, t, h, w)


def loss_case(name):
    """Seeded (prediction, target) pair with the statistics of magnitude images: target = |smooth + noise|,
    prediction = target + small error.  Shapes (b, t, h, w) float32."""
    b, t, h, w = LOSS_CASES[name]
    seed = 900 + sorted(LOSS_CASES).index(name)
    tgt = np.abs(rng_normal(seed, (b, t, h, w)) + 2.0 * np.sin(np.arange(w, dtype=np.float32) / 7.0)).astype(np.float32)
    pred = (tgt + 0.1 * rng_normal(seed + 50, (b, t, h, w))).astype(np.float32)
    return pred, tgt


def main_loss():
    """Runs the reference's own SSIMLoss (forward and autograd) on the CPU; its hard-coded `.to('cuda')`
    (losses.py:35) is neutralised for the duration of the call."""
    import torch
    rec, _ = _load_reference()
    from reconstruction.utils.losses import SSIMLoss
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        return orig_to(self, *a, **k) if (a or k) else self

    out = {}
    torch.Tensor.to = to_cpu
    try:
        for name in LOSS_CASES:
            pred, tgt = loss_case(name)
            for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
                x = torch.from_numpy(pred).to(dt).unsqueeze(1).requires_grad_(True)      # (b,1,t,h,w) as varnet_module.py:110-112
                y = torch.from_numpy(tgt).to(dt).unsqueeze(1)
                mod = SSIMLoss().to(dt)
                # losses.py:35 builds data_range with torch.Tensor([...]) (float32); promote it in the f64 run
                loss = mod(x, y, data_range=torch.tensor([float(tgt.max())]))
                loss.backward()
                out[f"{name}/{tag}/loss"] = np.asarray(loss.detach().numpy())
                g = x.grad.squeeze(1).numpy()
                if g.size > 50000:
                    idx = sample_index(g.size)
                    out[f"{name}/{tag}/grad_sample"] = g.reshape(-1)[idx]
                else:
                    out[f"{name}/{tag}/grad"] = g
    finally:
        torch.Tensor.to = orig_to
    np.savez_compressed(HERE / "golden_v2_loss.npz", **out)
    print(f"wrote {len(out)} arrays -> {HERE/'golden_v2_loss.npz'}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "loss":
        main_loss()
    else:
        main()
