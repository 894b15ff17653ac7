"""Block-tier drop-ins: the SENSE / data-consistency methods of the reference's
model classes, re-expressed over the fused kernels with unchanged signatures.

Every function takes the reference module instance as `self` (they are bound
onto the reference classes by `patch.patch_reference()`), so learnable state
(`lambda_reg`, `Softplus`, the regulariser sub-modules) stays where the
reference's checkpoints expect it.  The regularisers (U-Net / MWCNN / CRNN) are
called exactly as the reference calls them.
"""
from __future__ import annotations

import torch

from . import ops
from . import functional as F


# --------------------------------------------------------------------------- #
# shared operator helpers
# --------------------------------------------------------------------------- #
def sens_expand(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
    """VarNetBlock.sens_expand (models/varnet.py:181-185), CineNetBlock (cinenet.py:106-110),
    CineNet_RNN (recurrent_cinenet.py:59-63): fft2c(S * x).  x (b,t,1,h,w,2) -> (b,t,c,h,w,2)."""
    return ops.sens_expand(x, sens_maps)


def sens_reduce(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
    """VarNetBlock.sens_reduce (varnet.py:187-194) & copies: sum_c conj(S) ifft2c(k), keepdim."""
    return ops.sens_reduce(x, sens_maps).unsqueeze(2)


_PLANE_GLUE = True


def set_plane_glue_inference(on: bool) -> None:
    """Fused x-f / y-f plane packing around the regularisers under torch.no_grad() (default on); off = the reference's
    permute / view / NormUnet glue in eager torch (always used when autograd is recording)."""
    global _PLANE_GLUE
    _PLANE_GLUE = bool(on)


def _plane_glue_ok(x: torch.Tensor) -> bool:
    return _PLANE_GLUE and x.is_cuda and x.dtype == torch.float32 and not (torch.is_grad_enabled() and x.requires_grad) \
        and not torch.is_grad_enabled()


def _is_norm_unet(m) -> bool:
    """A reference NormUnet (denoisers/norm_unet.py:18-114): 2-channel complex planes around `.unet`."""
    return type(m).__name__ == "NormUnet" and hasattr(m, "unet") and getattr(m.unet, "in_chans", 2) == 2 \
        and getattr(m.unet, "out_chans", 2) == 2


def _xfyf(self, image_combined: torch.Tensor, run_models) -> torch.Tensor:
    """Temporal head/tail of xfyf_transform (varnet.py:196-241, cinenet.py:174-219) around the
    untouched regularisers.  image_combined (b,t,h,w,2) -> (b,t,1,h,w,2)."""
    xf = self.dynamic_type == 'XF'
    x, mean = ops.TemporalPreFn.apply(image_combined, xf)           # mean-subtract (+ fft1c over t)
    out = run_models(x)                                             # (b,t,1,h,w,2)
    return ops.TemporalPostFn.apply(out.squeeze(2), mean, xf).unsqueeze(2)


def varnet_xfyf_transform(self, image_combined: torch.Tensor) -> torch.Tensor:
    """VarNetBlock.xfyf_transform (varnet.py:196-241)."""
    b, t, h, w, ch = image_combined.shape

    def run(x):
        model_xf, model_yf = (self.model, self.model) if self.weight_sharing else self.model
        if _plane_glue_ok(x) and _is_norm_unet(model_xf) and _is_norm_unet(model_yf):
            # inference: the NCHW inputs of the two U-Nets (permute/view + NormUnet.complex_to_chan_dim / norm / pad,
            # norm_unet.py:101-105) come out of one kernel pass, the U-Nets themselves run untouched, and unpad /
            # unnorm / chan_complex_to_last_dim / the 0.5 (xf + yf) average (norm_unet.py:109-113, varnet.py:228-232)
            # are one more launch
            xf, yf, ctx = ops.raw_planes_pack(x, normalise=True, pad=True)
            return ops.raw_planes_unpack(model_xf.unet(xf), model_yf.unet(yf), ctx).unsqueeze(2)
        xf = x.permute(0, 2, 3, 1, 4).reshape(b * h, 1, w, t, 2)
        yf = x.permute(0, 3, 2, 1, 4).reshape(b * w, 1, h, t, 2)
        if self.weight_sharing:
            xf, yf = self.model(xf), self.model(yf)
        else:
            model_xf, model_yf = self.model
            xf, yf = model_xf(xf), model_yf(yf)
        xf_r = xf.view(b, h, 1, w, t, 2).permute(0, 4, 2, 1, 3, 5)
        yf_r = yf.view(b, w, 1, h, t, 2).permute(0, 4, 2, 3, 1, 5)
        return 0.5 * (xf_r + yf_r)

    return _xfyf(self, image_combined, run)


def cinenet_xfyf_transform(self, image_combined: torch.Tensor) -> torch.Tensor:
    """CineNetBlock.xfyf_transform (cinenet.py:174-219)."""
    b, t, h, w, ch = image_combined.shape

    def run(x):
        if _plane_glue_ok(x):
            # inference: the two permute/reshape copies in and the permute / average out (cinenet.py:193-212) as one
            # launch each way; the plain U-Nets take the planes as they are (no normalisation, no padding)
            model_xf, model_yf = (self.model, self.model) if self.weight_sharing else self.model
            xf, yf, ctx = ops.raw_planes_pack(x, normalise=False, pad=False)
            return ops.raw_planes_unpack(model_xf(xf), model_yf(yf), ctx).unsqueeze(2)
        xf = x.permute(0, 2, 4, 3, 1).reshape(b * h, 2, w, t)
        yf = x.permute(0, 3, 4, 2, 1).reshape(b * w, 2, h, t)
        if self.weight_sharing:
            xf, yf = self.model(xf), self.model(yf)
        else:
            model_xf, model_yf = self.model
            xf, yf = model_xf(xf), model_yf(yf)
        xf_r = xf.view(b, h, 1, 2, w, t).permute(0, 5, 2, 1, 4, 3)
        yf_r = yf.view(b, w, 1, 2, h, t).permute(0, 5, 2, 4, 1, 3)
        return 0.5 * (xf_r + yf_r)

    return _xfyf(self, image_combined, run)


# --------------------------------------------------------------------------- #
# VarNet
# --------------------------------------------------------------------------- #
def _varnet_regularise(self, image_combined):
    """The regulariser call of VarNetBlock.forward (varnet.py:255-279): (b,t,1,h,w,2) -> (b,t,1,h,w,2)."""
    if self.dynamic_type in ['XF', 'XT']:
        return self.xfyf_transform(image_combined.squeeze(2))
    if self.dynamic_type == '2D':
        return self.model(image_combined.squeeze(0)).unsqueeze(0)
    if self.dynamic_type == '3D':
        return self.model(image_combined.permute(0, 2, 1, 3, 4, 5)).permute(0, 2, 1, 3, 4, 5)
    raise ValueError(f"unknown dynamic_type {self.dynamic_type!r}")


def varnet_block_forward(self, current_kspace, ref_kspace, mask, sens_maps):
    """VarNetBlock.forward (varnet.py:244-282): A^H -> regulariser -> A fused with the soft-DC blend."""
    image_combined = ops.sens_reduce(current_kspace, sens_maps).unsqueeze(2)
    model_out = _varnet_regularise(self, image_combined)
    v = self.Softplus(self.lambda_reg)
    return ops.sens_expand(model_out, sens_maps, ops.EXPAND_DC, ref=ref_kspace, mask=mask, v=v)


_image_domain_inference = True


def set_image_domain_inference(flag: bool) -> None:
    """Inference fast path of the `VarNet.forward` drop-in (default on): between cascades the predicted k-space is only
    ever consumed by the next `sens_reduce` (varnet.py:253, 150-151), and
        A^H[ DC(A x, ref) ] = (sum_c |S_c|^2) x - eta (A^H M A x - A^H M ref),   eta = v/(1+v),
    so under `torch.no_grad()` every cascade's SENSE/DC work is ONE on-chip launch (b2s_normal_dc) and k-space is never
    materialised; the regularisers run unchanged.  Results agree with the k-space path to ~1e-6 of the maximum.  With
    autograd enabled, unsupported sizes or foreign cascade objects the k-space path runs as before."""
    global _image_domain_inference
    _image_domain_inference = bool(flag)


def _varnet_forward_image_domain(self, masked_kspace, mask, sens_maps):
    b, t, c, h, w, _ = masked_kspace.shape
    mk = ops._f32c(masked_kspace)
    s5 = ops._f32c(sens_maps.squeeze(1) if sens_maps.dim() == 6 else sens_maps)
    m8 = ops._mask_u8(mask, b, t, h)
    ssq = F.complex_abs_sq(s5).sum(dim=1).contiguous()                               # (b,h,w)  sum_c |S_c|^2
    bref = ops.raw_sens_reduce(mk, s5, ops.REDUCE_MASK, False, m8, None, 1)          # A^H M ref
    img = ops.raw_sens_reduce(mk, s5, ops.REDUCE_PLAIN, False, None, None, 1)        # cascade 0 starts from k = ref
    n = len(self.cascades)
    if n == 0:
        return F.complex_abs(img)
    for i, cascade in enumerate(self.cascades):
        model_out = _varnet_regularise(cascade, img.unsqueeze(2))
        v = cascade.Softplus(cascade.lambda_reg).detach().reshape(1).to(dtype=torch.float32)
        img = ops.raw_normal_dc(ops._f32c(model_out.squeeze(2)), s5, m8, v, ssq, bref, magnitude=(i == n - 1))   # last: |.| fused
    return img


def varnet_forward(self, masked_kspace: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """VarNet.forward (varnet.py:143-151); the clone of masked_kspace is not needed (ops never mutate)."""
    sens_maps = self.sens_net(masked_kspace, mask)
    if (_image_domain_inference and not torch.is_grad_enabled() and masked_kspace.is_cuda and masked_kspace.dim() == 6
            and ops.normal_op_supported(masked_kspace.shape[3], masked_kspace.shape[4])
            and all(hasattr(cb, "lambda_reg") and hasattr(cb, "dynamic_type") for cb in self.cascades)):
        return _varnet_forward_image_domain(self, masked_kspace, mask, sens_maps)
    kspace_pred = masked_kspace
    for cascade in self.cascades:
        kspace_pred = cascade(kspace_pred, masked_kspace, mask, sens_maps)
    return F.complex_abs(ops.sens_reduce(kspace_pred, sens_maps))


def _sens_pre(masked_kspace, mask):
    """ACS low-pass of the time mean + ifft2c (varnet.py:64-74 / xpdnet.py:75-85); the window is found
    on the device (no nonzero() host syncs)."""
    b, t, c, h, w, _ = masked_kspace.shape
    x = ops.AcsMeanFn.apply(masked_kspace, ops._mask_u8(mask, b, t, h))
    return ops.fft2c(x, "ortho", inverse=True)                       # (b,c,h,w,2)


def varnet_sens_model_forward(self, masked_kspace: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """SensitivityModel.forward (varnet.py:62-86)."""
    x = _sens_pre(masked_kspace, mask)
    x, b = self.chans_to_batch_dim(x)
    x = self.norm_unet(x)
    x = self.batch_chans_to_chan_dim(x, b)
    return ops.RssNormalizeFn.apply(x).unsqueeze(1)


def xpdnet_sens_model_forward(self, masked_kspace: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """xpdnet.SensitivityModel.forward (xpdnet.py:73-100)."""
    x = _sens_pre(masked_kspace, mask)
    b, x = self.chans_to_batch_dim(x)
    x_temp = x
    x = self.unet_model(x)
    if self.res_connection:
        x = x + x_temp
    x = self.batch_chans_to_chan_dim(x, b)
    return ops.RssNormalizeFn.apply(x).unsqueeze(1)


def divide_root_sum_of_squares(self, x: torch.Tensor) -> torch.Tensor:
    """SensitivityModel.divide_root_sum_of_squares (varnet.py:58-59)."""
    return ops.RssNormalizeFn.apply(x)


# --------------------------------------------------------------------------- #
# VarNet_RNN (recurrent_varnet.py:65-90): (b,2,h,w,t) image layout
# --------------------------------------------------------------------------- #
def varnet_rnn_sens_expand(self, x, sens_maps):
    return ops.sens_expand(x.permute(0, 4, 2, 3, 1), sens_maps)


def varnet_rnn_sens_reduce(self, x, sens_maps):
    return ops.sens_reduce(x, sens_maps).permute(0, 4, 2, 3, 1)


def varnet_rnn_data_consistency(self, x, ref_kspace, mask, sens_maps):
    v = self.Softplus(self.lambda_reg)
    dc = ops.sens_expand(x.permute(0, 4, 2, 3, 1), sens_maps, ops.EXPAND_DC, ref=ref_kspace, mask=mask, v=v)
    return ops.sens_reduce(dc, sens_maps).permute(0, 4, 2, 3, 1)


# --------------------------------------------------------------------------- #
# CineNet (cinenet.py:121-171, recurrent_cinenet.py:74-124)
# --------------------------------------------------------------------------- #
def h_operator(self, x: torch.Tensor, mask: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
    """HOperator: A^H M A x + softplus(lambda) x, k-space kept on chip.  x (b,t,1,h,w,2)."""
    v = self.Softplus(self.lambda_reg)
    return ops.normal_op(x.squeeze(2), sens_maps, mask, v).unsqueeze(2)


def conj_grad(self, x, b, mask, sens_maps, CG_iters: int):
    """ConjGrad (cinenet.py:136-171).  alpha/beta are constants for autograd exactly as in the
    reference (`.item()`), but they stay on the device: no host synchronisation."""
    v = self.Softplus(self.lambda_reg)
    squeeze = x.dim() == 6
    xs = x.squeeze(2) if squeeze else x
    bs = b.squeeze(2) if squeeze else b
    if torch.is_grad_enabled() and (xs.requires_grad or bs.requires_grad or v.requires_grad):
        out = _cg_autograd(xs, bs, mask, sens_maps, v, CG_iters)
    else:
        out = _cg_inference(xs, bs, mask, sens_maps, v, CG_iters)
    return out.unsqueeze(2) if squeeze else out


def _cg_autograd(x, b, mask, sens, v, iters):
    H = lambda z: ops.normal_op(z, sens, mask, v)                    # noqa: E731
    r = b - H(x)
    p = r.clone()
    rs_old = ops.raw_dot(r.detach().contiguous(), r.detach().contiguous())
    for _ in range(iters):
        d = H(p)
        pd = ops.raw_dot(p.detach().contiguous(), d.detach().contiguous())
        alpha = rs_old / pd
        x = x + alpha * p
        r = r - alpha * d
        rs_new = ops.raw_dot(r.detach().contiguous(), r.detach().contiguous())
        p = r + (rs_new / rs_old) * p
        rs_old = rs_new
    return x


def _cg_inference(x, b, mask, sens, v, iters):
    from . import _lib
    lib = _lib.lib()
    st = ops._stream
    P = ops._p
    x = ops._f32c(x).clone()
    b = ops._f32c(b)
    if sens.dim() == 6:
        sens = sens.squeeze(1)
    sens = ops._f32c(sens)
    bb, t, h, w, _ = x.shape
    m8 = ops._mask_u8(mask, bb, t, h)
    vd = v.detach().reshape(1).to(device=x.device, dtype=torch.float32)
    n = x.numel()
    dev = x.device
    scal = torch.empty(4, dtype=torch.float32, device=dev)           # rs_old, pd, rs_new
    scratch = torch.empty(1024, dtype=torch.float32, device=dev)
    rs_old, pd, rs_new = scal[0:1], scal[1:2], scal[2:3]

    def H(z):
        if ops.normal_op_supported(h, w):
            return ops.raw_normal_op(z, sens, m8, vd)
        k = ops.raw_sens_expand(z, sens, ops.EXPAND_MASK, None, m8, None, 1)
        return ops.raw_axpby(ops.raw_sens_reduce(k, sens, norm=1), z, vd, 1.0)

    r = ops.raw_axpby(b, H(x), None, -1.0)                           # r = b - Hx
    p = r.clone()
    _lib.check(lib.b2s_dot(P(r), P(r), P(rs_old), n, P(scratch), st()), "dot")
    if ops.normal_op_supported(h, w):
        # fused iteration: <p, Hp> comes out of the normal-operator launch as per-item partials, the two vector kernels
        # do the rest (3 launches per iteration instead of 8; same arithmetic, fixed summation order)
        n_pd = bb * t * (w // 4)
        pd_part = torch.empty(n_pd, dtype=torch.float32, device=dev)
        rr_part = torch.empty(int(lib.b2s_cg_blocks(n)), dtype=torch.float32, device=dev)
        d = torch.empty_like(p)
        c = sens.shape[1]
        for _ in range(iters):
            _lib.check(lib.b2s_normal_op_dot(P(p), P(sens), P(m8), P(vd), P(d), P(pd_part), bb, t, c, h, w, st()), "normal_op_dot")
            _lib.check(lib.b2s_cg_update(P(p), P(d), P(x), P(r), P(pd_part), n_pd, P(rs_old), P(rr_part), n, st()), "cg_update")
            _lib.check(lib.b2s_cg_direction(P(p), P(r), P(rr_part), P(rs_old), P(rs_new), n, st()), "cg_direction")
            rs_old, rs_new = rs_new, rs_old
        return x
    for _ in range(iters):
        d = H(p)
        _lib.check(lib.b2s_dot(P(p), P(d), P(pd), n, P(scratch), st()), "dot")
        _lib.check(lib.b2s_axpy_ratio(P(x), P(p), P(rs_old), P(pd), 1.0, n, st()), "axpy")     # x += alpha p
        _lib.check(lib.b2s_axpy_ratio(P(r), P(d), P(rs_old), P(pd), -1.0, n, st()), "axpy")    # r -= alpha d
        _lib.check(lib.b2s_dot(P(r), P(r), P(rs_new), n, P(scratch), st()), "dot")
        _lib.check(lib.b2s_xpay_ratio(P(p), P(r), P(rs_new), P(rs_old), n, st()), "xpay")      # p = r + beta p
        rs_old, rs_new = rs_new, rs_old
    return x


def cinenet_forward(self, masked_kspace, mask, sens_maps):
    """CineNet.forward (cinenet.py:61-73)."""
    image_pred = ops.sens_reduce(masked_kspace, sens_maps).unsqueeze(2)
    image_ref = image_pred
    for cascade in self.cascades:
        image_pred = cascade(image_pred, image_ref, mask, sens_maps)
    return F.complex_abs(image_pred.squeeze(2))


# --------------------------------------------------------------------------- #
# XPDNet operators (xpdnet.py:104-167)
# --------------------------------------------------------------------------- #
def forward_operator_forward(self, image, mask, sens_maps, buffer_size: int):
    """ForwardOperator.forward: acts on buffer channels 0 and buffer_size; optional `* mask + 0.0`."""
    img = torch.stack([image[..., 0], image[..., buffer_size]], dim=-1)
    if self.masked:
        return ops.sens_expand(img, sens_maps, ops.EXPAND_MASK, mask=mask)
    return ops.sens_expand(img, sens_maps)


def backward_operator_forward(self, kspace, mask, sens_maps, buffer_size: int):
    """BackwardOperator.forward: optional mask, ifft2c, conj(S) multiply, coil sum (keepdim)."""
    if kspace.shape[-1] != 2:
        kspace = torch.stack([kspace[..., 0], kspace[..., buffer_size]], dim=-1)
    return ops.sens_reduce(kspace, sens_maps, mask=mask if self.masked else None).unsqueeze(2)


def xpd_temporal_fft(x: torch.Tensor, n_ch: int) -> torch.Tensor:
    """xpdnet.py:465-467: packed (b,t,h,w,2n) -> ifftshift(fft(fftshift(.,1),t,1,'ortho'),1), packed."""
    b, t, h, w, _ = x.shape
    z = torch.stack([x[..., :n_ch], x[..., n_ch:]], dim=-1)          # (b,t,h,w,n,2)
    z = z.permute(0, 2, 3, 4, 1, 5)                                   # (b,h,w,n,t,2): FFT dim at -2
    z = ops.fft1c(z.contiguous(), "ortho", False, shift_in=t // 2, shift_out=(t + 1) // 2)
    z = z.permute(0, 4, 1, 2, 3, 5)
    return torch.cat([z[..., 0], z[..., 1]], dim=-1)


def xpd_temporal_ifft(x: torch.Tensor, n_ch: int) -> torch.Tensor:
    """xpdnet.py:499-501: fftshift(ifft(ifftshift(.,1),t,1,'ortho'),1) == ifft1c over t."""
    b, t, h, w, _ = x.shape
    z = torch.stack([x[..., :n_ch], x[..., n_ch:]], dim=-1).permute(0, 2, 3, 4, 1, 5)
    z = ops.fft1c(z.contiguous(), "ortho", True)
    z = z.permute(0, 4, 1, 2, 3, 5)
    return torch.cat([z[..., 0], z[..., 1]], dim=-1)


# --------------------------------------------------------------------------- #
# XPDNet cascade bodies (xpdnet.py:295-298, 372-446; recurrent_xpdnet.py:87-150)
#
# Buffers are packed [re_0 .. re_{n-1}, im_0 .. im_{n-1}] along the last dim (utils/math.py:97-135).  The reference
# moves between that packing and torch.complex through real_to_complex_multi_ch / cat / complex_to_real_multi_ch
# (five to seven copies per block); here every re-packing is ONE gather (`torch.stack` / `torch.cat` of slices) and the
# operators in between are the fused kernels.
# --------------------------------------------------------------------------- #
def _pick(buffer: torch.Tensor, n: int) -> torch.Tensor:
    """First complex entry of a packed buffer (..., 2n) as a (..., 2) tensor (the operators act on it only:
    xpdnet.py:127-128, 160-161)."""
    if buffer.shape[-1] == 2:
        return buffer
    return torch.stack([buffer[..., 0], buffer[..., n]], dim=-1)


def _append(buffer: torch.Tensor, new: torch.Tensor, n: int) -> torch.Tensor:
    """Packed buffer (..., 2n) + one complex entry (..., 2) -> packed (..., 2n + 2): what
    complex_to_real_multi_ch(cat([real_to_complex_multi_ch(buffer, n), real_to_complex_multi_ch(new, 1)])) builds."""
    return torch.cat([buffer[..., :n], new[..., 0:1], buffer[..., n:], new[..., 1:2]], dim=-1)


def _is_measurements_residual(net) -> bool:
    return getattr(net, "__name__", "") == "measurements_residual"


def xpdnet_measurements_residual(self, concat_kspace: torch.Tensor) -> torch.Tensor:
    """XPDNet.measurements_residual (xpdnet.py:295-298): packed [re_cur, re_ref, im_cur, im_ref] -> cur - ref."""
    return torch.stack([concat_kspace[..., 0] - concat_kspace[..., 1], concat_kspace[..., 2] - concat_kspace[..., 3]], dim=-1)


def _xpd_k_domain(self, index, image_buffer, kspace_buffer, mask, sens_maps, ref_kspace):
    net = self.kspace_net[index]
    img = _pick(image_buffer, self.i_buffer_size)
    if not self.k_buffer_mode and _is_measurements_residual(net):
        # primal-only: the "k-space net" is the measurement residual, so the whole K block is M A x - y in ONE launch
        # (forward operator, mask and subtraction fused: B2S_EXPAND_RESIDUAL)
        return ops.sens_expand(img, sens_maps, ops.EXPAND_RESIDUAL, ref=ref_kspace, mask=mask)
    fwd = ops.sens_expand(img, sens_maps, ops.EXPAND_MASK, mask=mask) if self.forward_op.masked else ops.sens_expand(img, sens_maps)
    if self.k_buffer_mode:
        nd = self.k_buffer_size
        packed = torch.cat([kspace_buffer[..., :nd], fwd[..., 0:1], ref_kspace[..., 0:1],
                            kspace_buffer[..., nd:], fwd[..., 1:2], ref_kspace[..., 1:2]], dim=-1)
    else:
        packed = torch.cat([fwd[..., 0:1], ref_kspace[..., 0:1], fwd[..., 1:2], ref_kspace[..., 1:2]], dim=-1)
    return net(packed)


def xpdnet_k_domain_correction(self, i_domain, image_buffer, kspace_buffer, mask, sens_maps, ref_kspace):
    """XPDNetBlock.k_domain_correction (xpdnet.py:372-403)."""
    return _xpd_k_domain(self, i_domain // 2, image_buffer, kspace_buffer, mask, sens_maps, ref_kspace)


def xpdnet_rnn_k_domain_correction(self, i_cascade, image_buffer, kspace_buffer, mask, sens_maps, ref_kspace):
    """XPDNet_RNN.k_domain_correction (recurrent_xpdnet.py:93-125)."""
    return _xpd_k_domain(self, i_cascade, image_buffer, kspace_buffer, mask, sens_maps, ref_kspace)


def xpdnet_update_image_buffer(self, image_buffer, kspace_buffer, mask, sens_maps):
    """Head of XPDNetBlock.i_domain_correction (xpdnet.py:419-429) == XPDNet_RNN.update_image_buffer
    (recurrent_xpdnet.py:128-150): A^H (M r) appended to the packed image buffer."""
    new = self.backward_op(kspace_buffer, mask, sens_maps, self.k_buffer_size)          # (b,t,1,h,w,2), fused kernel
    if not self.i_buffer_mode:
        return new
    return _append(image_buffer, new, self.i_buffer_size)


def xpdnet_i_domain_correction(self, i_domain, image_buffer, kspace_buffer, mask, sens_maps):
    """XPDNetBlock.i_domain_correction (xpdnet.py:406-446); the image nets are called exactly as the reference does."""
    image_buffer = xpdnet_update_image_buffer(self, image_buffer, kspace_buffer, mask, sens_maps)
    b, t, c, h, w, ch = image_buffer.shape
    ch_out = 2 * self.i_buffer_size
    if self.dynamic_type in ['XF', 'XT']:
        return self.xfyf_transform(image_buffer.squeeze(2), i_domain)
    if self.dynamic_type == '2D':
        image_in = image_buffer.permute(0, 1, 2, 5, 3, 4).reshape(b * t, c * ch, h, w)
        return self.image_net[i_domain // 2](image_in).reshape(b, t, c, ch_out, h, w).permute(0, 1, 2, 4, 5, 3)
    raise ValueError(f"unknown dynamic_type {self.dynamic_type!r}")
