"""The REAL reference models (baseline/_ref, unmodified) with and without `patch_reference()` on the GPU.

`__graft_entry__.build()` ships `/root/reference/reconstruction` to `baseline/_ref` (git-ignored, travels to the GPU box).
Every model class of `reconstruction.models` is instantiated with seeded weights and its real regularisers (U-Net /
NormUnet / MWCNN / CRNN, run by cuDNN exactly as the reference runs them); the same module object is then run

    ref32   unpatched, fp32   (the reference's own eager torch + cuFFT path on this GPU)
    ours    patched,   fp32   (SENSE / DC path on the b200sense kernels, regularisers untouched)
    ref64   unpatched, fp64   (arbiter, where the model can run in double)

and `ours` must be as close to the arbiter as the reference's own fp32 run is (the bound is stated in `close()`):
through several cascades of randomly initialised CNNs the fp32 reference itself is only ~1e-5..1e-4 from its fp64
run, so a bare `|ours - ref32| <= 1e-5 max` would test the conditioning of the regularisers, not the operators.
Operator-level parity at 1e-5 is tests/test_gpu_parity.py; here the drop-in plumbing (shapes, views, buffer packing,
dispatch on the real classes, autograd through the real modules) is what is under test.
"""
from __future__ import annotations

import copy
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import load_reference as L          # test infrastructure                       # noqa: E402
from deep_cine_cardiac_mri_b200 import blocks, metrics, patch, synth                       # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not L.available(), reason="baseline/_ref missing: run __graft_entry__.build()")]

B, T, C, H, W = 1, 6, 4, 200, 200


@pytest.fixture(scope="module")
def rec():
    torch.backends.cudnn.allow_tf32 = False          # the regularisers must compute the same in both arms
    torch.backends.cuda.matmul.allow_tf32 = False
    r = L.load()
    yield r
    patch.unpatch_reference()


def inputs(seed, t=T, c=C, h=H, w=W, dtype=torch.float32):
    case = synth.cine_case(seed, B, t, c, h, w)
    d = synth.to_torch(case, "cuda")
    return d["masked_kspace"].to(dtype), d["mask"], d["sens"].to(dtype)


def ground_truth(seed, t=T, c=C, h=H, w=W):
    """|x| of the synthetic object behind `inputs(seed)`: (t,h,w), the target of the image metrics."""
    img = synth.cine_case(seed, B, t, c, h, w)["image"]
    return torch.from_numpy(np.sqrt((img.astype(np.float64) ** 2).sum(-1))[0, :, 0]).float().cuda()


def run(model, args, patched, grad=False):
    if patched:
        patch.patch_reference()
    else:
        patch.unpatch_reference()
    try:
        if grad:
            model.zero_grad(set_to_none=True)
            out = model(*args)
            wgt = torch.linspace(0.5, 1.5, out.numel(), device=out.device, dtype=out.dtype).view_as(out)
            (out * wgt).sum().backward()
            grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
            return out.detach(), grads
        with torch.no_grad():
            return model(*args)
    finally:
        patch.unpatch_reference()


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def close(ours, ref32, ref64, what, floor=1e-5, slack=4.0):
    """ours within max(floor, slack * |ref32 - ref64|) of the fp64 arbiter, relative to max|ref64|."""
    e_ref = relmax(ref32, ref64)
    e_ours = relmax(ours, ref64)
    print(f"{what}: |ours-ref64| {e_ours:.2e}  |ref32-ref64| {e_ref:.2e}  |ours-ref32| {relmax(ours, ref32):.2e}")
    assert e_ours <= max(floor, slack * e_ref), (what, e_ours, e_ref)


def three_way(rec, make, args_of, what, image_domain=None, seed=None):
    torch.manual_seed(1234)
    model = make().cuda().eval()
    a32 = args_of(torch.float32)
    ref32 = run(model, a32, False)
    if image_domain is not None:
        blocks.set_image_domain_inference(image_domain)
    try:
        ours = run(model, a32, True)
    finally:
        blocks.set_image_domain_inference(True)
    m64 = copy.deepcopy(model).double()
    ref64 = run(m64, args_of(torch.float64), False)
    assert ours.shape == ref32.shape and ours.dtype == ref32.dtype
    close(ours, ref32, ref64, what)
    # end-to-end image metrics against the synthetic ground truth: SSIM / NMSE / PSNR of the patched run equal the
    # reference's (the networks are untrained, so the values themselves are poor - their AGREEMENT is the test)
    if seed is not None:
        gt = ground_truth(seed)
        for fn, tol in ((metrics.ssim, 1e-4), (metrics.psnr, 1e-3), (metrics.nmse, 1e-4)):
            a, b_ = float(fn(gt, ours[0])), float(fn(gt, ref32[0]))
            assert abs(a - b_) <= tol * max(1.0, abs(b_)), (fn.__name__, a, b_)
    return model


@pytest.mark.parametrize("dyn", ["XF", "XT", "2D", "3D"])
@pytest.mark.parametrize("image_domain", [True, False])
def test_varnet_real_model(rec, dyn, image_domain):
    mk = lambda: rec.models.VarNet(num_cascades=3, sens_chans=4, sens_pools=2, chans=4, pools=2, dynamic_type=dyn)   # noqa: E731
    three_way(rec, mk, lambda dt: inputs(11, dtype=dt)[:2], f"VarNet {dyn} image_domain={image_domain}", image_domain, seed=11)


@pytest.mark.parametrize("dyn", ["XF", "XT", "2D", "3D"])
def test_cinenet_real_model(rec, dyn):
    mk = lambda: rec.models.CineNet(num_cascades=2, CG_iters=4, chans=4, pools=2, dynamic_type=dyn)   # noqa: E731
    three_way(rec, mk, lambda dt: inputs(12, dtype=dt), f"CineNet {dyn}", seed=12)


@pytest.mark.parametrize("dyn,primal_only", [("XT", True), ("XF", True), ("2D", True), ("XT", False)])
def test_xpdnet_real_model(rec, dyn, primal_only):
    mk = lambda: rec.models.XPDNet(num_cascades=2, sens_chans=4, sens_pools=2, n_scales=2, n_filters_per_scale=[12, 16],   # noqa: E731
                                   n_convs_per_scale=[1, 1], n_first_convs=1, first_conv_n_filters=12, dynamic_type=dyn,
                                   primal_only=primal_only, n_dual=2)
    torch.manual_seed(1234)
    model = mk().cuda().eval()
    a32 = inputs(13)[:2]
    ref32 = run(model, a32, False)
    ours = run(model, a32, True)
    # MWCNN's IWT allocates a float32 buffer (denoisers/mwcnn.py:257): the model cannot run in double, so the bound is
    # against the fp32 reference with the conditioning of two MWCNN cascades as slack
    e = relmax(ours, ref32)
    print(f"XPDNet {dyn} primal_only={primal_only}: |ours-ref32| {e:.2e}")
    assert ours.shape == ref32.shape
    assert e <= 2e-4


def test_rnn_real_models(rec):
    for name, mk, n_args in (
            ("VarNet_RNN", lambda: rec.models.VarNet_RNN(num_cascades=2, sens_chans=4, sens_pools=2, chans=8), 2),
            ("CineNet_RNN", lambda: rec.models.CineNet_RNN(num_cascades=2, CG_iters=3, chans=8), 3),
            ("XPDNet_RNN", lambda: rec.models.XPDNet_RNN(num_cascades=2, sens_chans=4, sens_pools=2, chans=8), 2)):
        torch.manual_seed(77)
        model = mk().cuda().eval()
        a32 = inputs(14)[:n_args]
        ref32 = run(model, a32, False)
        ours = run(model, a32, True)
        e = relmax(ours, ref32)                      # (hidden states are hard-coded float32 .cuda(): no fp64 arbiter)
        print(f"{name}: |ours-ref32| {e:.2e}")
        assert ours.shape == ref32.shape
        assert e <= 2e-4, name


def test_gradients_through_real_models(rec):
    """lambda_reg and the sensitivity-net weights receive the reference's gradients (autograd through the custom ops:
    the adjoint kernels are the backward)."""
    cases = (
        ("VarNet", lambda: rec.models.VarNet(num_cascades=2, sens_chans=4, sens_pools=2, chans=4, pools=2, dynamic_type="XF"), 2),
        ("XPDNet", lambda: rec.models.XPDNet(num_cascades=2, sens_chans=4, sens_pools=2, n_scales=2, n_filters_per_scale=[12, 16],
                                             n_convs_per_scale=[1, 1], n_first_convs=1, first_conv_n_filters=12, dynamic_type="XT"), 2),
        ("CineNet", lambda: rec.models.CineNet(num_cascades=2, CG_iters=3, chans=4, pools=2, dynamic_type="XT"), 3),
    )
    for name, mk, n_args in cases:
        torch.manual_seed(5)
        model = mk().cuda().train()
        a32 = inputs(15)[:n_args]
        out_r, g_ref = run(model, a32, False, grad=True)
        out_o, g_our = run(model, a32, True, grad=True)
        assert relmax(out_o, out_r) <= 2e-4, name
        assert set(g_ref) == set(g_our), name
        checked = 0
        for n in g_ref:
            if "lambda_reg" in n or "sens_net" in n:
                scale = float(g_ref[n].abs().max())
                if scale == 0.0:
                    assert float(g_our[n].abs().max()) == 0.0, (name, n)
                    continue
                # both arms are fp32 through the same randomly initialised CNNs (no fp64 arbiter for MWCNN / CRNN models):
                # the bound is on the error of the whole gradient tensor, with the element-wise maximum as a sanity check
                err_l2 = float((g_our[n] - g_ref[n]).double().norm() / g_ref[n].double().norm())
                err_max = float((g_our[n] - g_ref[n]).abs().max()) / scale
                print(f"{name} grad {n}: rel l2 {err_l2:.2e} max {err_max:.2e}")
                # (VarNet / CineNet agree to ~2e-4; XPDNet's un-normalised sens U-Net amplifies the fp32 differences of
                # the two arms to ~1e-2 in its first layers - the operator gradients themselves are pinned at 5e-5 in
                # tests/test_gpu_parity.py::test_autograd_xpdnet_chain)
                tol = 3e-2 if name == "XPDNet" else 2e-3
                assert err_l2 <= tol and err_max <= 3 * tol, (name, n, err_l2, err_max)
                checked += 1
        assert checked >= 2, name


def test_xpdnet_block_bodies_match_reference(rec):
    """a12: k_domain_correction / i_domain_correction head of the real XPDNetBlock, patched vs unpatched, on the buffers
    the real XPDNet.forward builds (xpdnet.py:301-326)."""
    torch.manual_seed(3)
    model = rec.models.XPDNet(num_cascades=1, sens_chans=4, sens_pools=2, n_scales=2, n_filters_per_scale=[12, 16],
                              n_convs_per_scale=[1, 1], n_first_convs=1, first_conv_n_filters=12, dynamic_type="XT").cuda().eval()
    mk, mask, sens = inputs(16)
    blk = model.cascades[0]
    with torch.no_grad():
        image = model.backward_op(mk, mask, sens, 1)
        ibuf = torch.repeat_interleave(image, model.i_buffer_size, dim=-1) + 0.01 * torch.randn(1, T, 1, H, W, 10, device="cuda")
        kbuf = torch.repeat_interleave(mk, model.k_buffer_size, dim=-1)
        outs = {}
        for patched in (False, True):
            (patch.patch_reference if patched else patch.unpatch_reference)()
            k = blk.k_domain_correction(0, ibuf, kbuf, mask, sens, mk)
            captured = {}
            orig = type(blk).xfyf_transform
            type(blk).xfyf_transform = lambda self, buf, i, _c=captured: _c.setdefault("head", buf.clone())
            try:
                blk.i_domain_correction(1, ibuf, k, mask, sens)
            finally:
                type(blk).xfyf_transform = orig
            outs[patched] = (k, captured["head"])
        patch.unpatch_reference()
    # two fp32 evaluations of M A x - y (each ~5e-6 of the maximum from the exact value, test_gpu_parity.py pins ours
    # against the fp64 oracle at 1e-5): their difference may reach 2e-5
    assert relmax(outs[True][0], outs[False][0]) <= 2e-5
    assert relmax(outs[True][1], outs[False][1]) <= 2e-5


# ------------------------------------------------------------------------------------------------------------------ #
# regulariser-side layout glue: plane packing around the REAL NormUnet / Unet
# ------------------------------------------------------------------------------------------------------------------ #
@pytest.mark.parametrize("shape", [(2, 15, 200, 200), (1, 7, 24, 40), (1, 30, 256, 256), (1, 16, 32, 48), (1, 3, 18, 36), (2, 5, 10, 12), (1, 17, 20, 6)])
def test_plane_pack_unpack_against_reference_normunet_glue(rec, shape):
    """b2s_planes_stats / pack / unpack against the reference's own glue: the permute/view pairs of
    VarNetBlock.xfyf_transform (varnet.py:215-216, 228-232) and NormUnet.complex_to_chan_dim / norm / pad / unpad /
    unnorm / chan_complex_to_last_dim (norm_unet.py:48-113), run in fp64 torch as the arbiter."""
    from deep_cine_cardiac_mri_b200 import ops
    from reconstruction.models.denoisers.norm_unet import NormUnet
    b, t, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(b, t, h, w, 2, device="cuda", generator=g) * 3 + 0.7
    nu = NormUnet(4, 2).cuda().double()
    x64 = x.double()
    xf_ref = x64.clone().permute(0, 2, 3, 1, 4).reshape(b * h, 1, w, t, 2)
    yf_ref = x64.clone().permute(0, 3, 2, 1, 4).reshape(b * w, 1, h, t, 2)
    want, keep = [], []
    for z in (xf_ref, yf_ref):
        zz, mean, std = nu.norm(nu.complex_to_chan_dim(z))
        zz, pads = nu.pad(zz)
        want.append(zz); keep.append((mean, std, pads))
    xf, yf, ctx = ops.raw_planes_pack(x, normalise=True, pad=True)
    assert xf.shape == want[0].shape and yf.shape == want[1].shape
    # the statistics: fused into the pack kernels == the stand-alone entry point == NormUnet.norm's mean / std
    from deep_cine_cardiac_mri_b200 import _lib
    sx = torch.empty(b * h, 2, 2, device="cuda"); sy = torch.empty(b * w, 2, 2, device="cuda")
    _lib.check(_lib.lib().b2s_planes_stats(ops._p(x), ops._p(sx), ops._p(sy), b, t, h, w, ops._stream()), "planes_stats")
    for got_s, alone, (mean, std, _) in zip(ctx[:2], (sx, sy), keep):
        assert relmax(got_s, alone) <= 1e-6
        assert relmax(got_s[..., 0], mean.view(-1, 2)) <= 1e-6 and relmax(got_s[..., 1], std.view(-1, 2)) <= 1e-6
    assert relmax(xf, want[0]) <= 2e-6 and relmax(yf, want[1]) <= 2e-6
    # the way back: feed "U-Net outputs" u (a fixed pointwise function of the inputs) through both paths
    f = lambda z: torch.tanh(z) * 1.5 + 0.25 * z                                       # noqa: E731
    back = []
    for z, (mean, std, pads) in zip(want, keep):
        back.append(nu.chan_complex_to_last_dim(nu.unnorm(nu.unpad(f(z), *pads), mean, std)))
    xf_r = back[0].view(b, h, 1, w, t, 2).permute(0, 4, 2, 1, 3, 5)
    yf_r = back[1].view(b, w, 1, h, t, 2).permute(0, 4, 2, 3, 1, 5)
    want_out = (0.5 * (xf_r + yf_r)).squeeze(2)
    got = ops.raw_planes_unpack(f(xf), f(yf), ctx)
    assert relmax(got, want_out) <= 5e-6
    # bare permutation (CineNet's plain Unet, cinenet.py:193-212): exact
    xf2, yf2, ctx2 = ops.raw_planes_pack(x, normalise=False, pad=False)
    assert torch.equal(xf2, x.permute(0, 2, 4, 3, 1).reshape(b * h, 2, w, t))
    assert torch.equal(yf2, x.permute(0, 3, 4, 2, 1).reshape(b * w, 2, h, t))
    xr = xf2.view(b, h, 1, 2, w, t).permute(0, 5, 2, 1, 4, 3)
    yr = yf2.view(b, w, 1, 2, h, t).permute(0, 5, 2, 4, 1, 3)
    assert torch.equal(ops.raw_planes_unpack(xf2, yf2, ctx2), (0.5 * (xr + yr)).squeeze(2))


@pytest.mark.parametrize("kind", ["varnet_xf", "varnet_xt_shared", "cinenet_xf"])
def test_xfyf_transform_fast_glue_on_real_blocks(rec, kind):
    """xfyf_transform of the real VarNetBlock (NormUnet pair) / CineNetBlock (Unet pair) under no_grad: patched with the
    fused plane glue == patched with the reference's eager glue == the unpatched reference, through the real U-Nets."""
    from reconstruction.models import varnet, cinenet
    from reconstruction.models.denoisers.norm_unet import NormUnet
    from reconstruction.models.denoisers.unet import Unet
    from deep_cine_cardiac_mri_b200 import _lib
    torch.manual_seed(3)
    if kind == "varnet_xf":
        blk = varnet.VarNetBlock(torch.nn.ModuleList([NormUnet(6, 2), NormUnet(6, 2)]), "XF", False).cuda().eval()
    elif kind == "varnet_xt_shared":
        blk = varnet.VarNetBlock(NormUnet(6, 2), "XT", True).cuda().eval()
    else:
        blk = cinenet.CineNetBlock(torch.nn.ModuleList([Unet(6, 2, dims=2), Unet(6, 2, dims=2)]), 4, "XF", False).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(9)
    img = torch.randn(1, 15, 200, 200, 2, device="cuda", generator=g)
    with torch.no_grad():
        patch.unpatch_reference()
        ref = blk.xfyf_transform(img)
        patch.patch_reference()
        try:
            blocks.set_plane_glue_inference(False)
            eager = blk.xfyf_transform(img)
            blocks.set_plane_glue_inference(True)
            n0 = _lib.lib().b2s_launch_count(0)
            fast = blk.xfyf_transform(img)
            launched = _lib.lib().b2s_launch_count(0) - n0
        finally:
            blocks.set_plane_glue_inference(True)
            patch.unpatch_reference()
    assert relmax(eager, ref) <= 1e-5
    assert relmax(fast, ref) <= 1e-5
    assert launched >= 4            # temporal head + (stats) + pack + unpack + temporal tail really ran on our kernels
