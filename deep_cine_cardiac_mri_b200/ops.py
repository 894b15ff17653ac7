"""PyTorch custom operators over the C ABI (include/b200sense.h).

Layering:  `raw_*`  = one ABI call on the current CUDA stream, no autograd;
`*Fn` autograd Functions use the adjoint kernels as the backward
(SURVEY.md section 10); the module-level functions at the bottom are what
`functional.py` / `blocks.py` call.  Tensors must be float32 CUDA tensors —
anything else raises (there is no CPU or cuFFT fallback).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

NORM = {None: 0, "backward": 0, "ortho": 1, "forward": 2}
_ADJ_NORM = {0: 2, 1: 1, 2: 0}          # adjoint of a transform with norm n is the inverse with this norm
EXPAND_PLAIN, EXPAND_MASK, EXPAND_DC, EXPAND_RESIDUAL = 0, 1, 2, 3
REDUCE_PLAIN, REDUCE_MASK, REDUCE_DCGRAD = 0, 1, 2
REDUCE_DETERMINISTIC = 0x10       # OR-ed into the weight mode: ordered coil sum, no float atomics
DC_BWD_GV_FLOATS = 1032           # include/b200sense.h: B2S_DC_BWD_GV_FLOATS


# --------------------------------------------------------------------------- #
# plumbing
# --------------------------------------------------------------------------- #
def _norm(norm) -> int:
    try:
        return NORM[norm]
    except KeyError:
        raise RuntimeError(f"Invalid normalization mode: {norm!r}") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200sense operators run on CUDA tensors only "
                               "(there is no CPU fallback); got a tensor on " + str(t.device))
        if t.dtype != torch.float32 and t.dtype != torch.uint8:
            raise TypeError(f"b200sense operators are float32 (got {t.dtype})")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous float32 (no copy when already so)"""
    if t.dtype != torch.float32:
        raise TypeError(f"b200sense operators are float32 (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


def _mask_u8(mask: torch.Tensor, b: int, t: int, h: int) -> torch.Tensor:
    """reference mask (b,t,1,h,1,1) (uint8 / float / bool) -> contiguous uint8 (b,t,h)."""
    m = mask
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    if m.numel() != b * t * h:
        m = m.expand(b, t, 1, h, 1, 1) if m.dim() == 6 else m.expand(b, t, h)
    return m.reshape(b, t, h).contiguous()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _vdev(v, device) -> torch.Tensor:
    """softplus(lambda) as a 1-element float32 device tensor (no host sync)."""
    if isinstance(v, torch.Tensor):
        return v.detach().reshape(-1)[:1].to(device=device, dtype=torch.float32).contiguous()
    return torch.full((1,), float(v), dtype=torch.float32, device=device)


_force_deterministic = False


def set_deterministic(flag: bool) -> None:
    """Force the ordered (atomics-free) coil sum in `sens_reduce` regardless of torch's global switch."""
    global _force_deterministic
    _force_deterministic = bool(flag)


def deterministic() -> bool:
    """True when results must be run-to-run bit-identical: `torch.use_deterministic_algorithms(True)` (what
    the reference's `Trainer(deterministic=True)` sets, train_test_varnet.py:292) or `set_deterministic(True)`."""
    return _force_deterministic or torch.are_deterministic_algorithms_enabled()


def set_fused_path(path) -> None:
    """Kernel family behind the fused plan sizes (a test / measurement knob; process-wide):
    None / "auto"  the library's measured cost model chooses per launch (default),
    "half"         half-split (200x200) / quarter-split (256x256) kernels only,
    "packed"       the packed whole-image kernel (second half parked in tensor memory) wherever it exists,
    "strip"        strip-streamed kernels - only in libraries built with `make EXPERIMENTS=1` (ValueError otherwise)."""
    code = {None: 0, "auto": 0, "strip": 1, "half": 2, "onchip": 2, "packed": 3}[path]
    _lib.check(_lib.lib().b2s_set_fused_path(code), "set_fused_path")


def set_sm_reserve(n_sms: int) -> None:
    """Keep `n_sms` SMs out of the persistent fused kernels' grids; `upload_masked_kspace` then runs on exactly those
    (see include/b200sense.h: b2s_set_sm_reserve).  Affects launches (and graph captures) made afterwards."""
    _lib.check(_lib.lib().b2s_set_sm_reserve(int(n_sms)), "set_sm_reserve")


def strip_status() -> int:
    """0 when no inter-CTA dependency wait of the strip kernels ever timed out on the current device (synchronises)."""
    return int(_lib.lib().b2s_debug_strip_status())


def upload_masked_kspace(kspace_host: torch.Tensor, mask_dev: torch.Tensor, out: torch.Tensor = None, verify: bool = False) -> torch.Tensor:
    """Sparse host->device upload of a masked k-space (what data/transforms.py:66-92 apply_mask leaves: unsampled
    rows are zero).  `kspace_host` (b,t,c,h,w,2) float32 in PINNED host memory, `mask_dev` the (b,t,1,h,1,1) / (b,t,h)
    uint8 mask already on the device.  Only the sampled rows cross PCIe (the GPU reads them in place over UVA), the
    others are written as zeros; runs on the current stream.

    PRECONDITION: the unsampled rows of `kspace_host` are zero (true for `apply_mask` output, which is what the
    reference's models are fed, mri_module / transforms.py:66-92).  A k-space that is NOT masked gives a different
    result from a dense copy without any error; `verify=True` checks the precondition on the host first (reads the whole
    buffer and synchronises - for tests and debugging) and raises ValueError when it does not hold."""
    if kspace_host.is_cuda or not kspace_host.is_pinned():
        raise ValueError("upload_masked_kspace: kspace_host must be a pinned host tensor")
    if kspace_host.dtype != torch.float32 or kspace_host.dim() != 6 or kspace_host.shape[-1] != 2 or not kspace_host.is_contiguous():
        raise ValueError("upload_masked_kspace: kspace_host must be contiguous float32 (b,t,c,h,w,2)")
    _need_cuda(mask_dev)
    b, t, c, h, w, _ = kspace_host.shape
    m = _mask_u8(mask_dev, b, t, h)
    if verify:
        off = (m == 0).cpu().view(b, t, 1, h, 1, 1)
        if bool((kspace_host * off).ne(0).any()):
            raise ValueError("upload_masked_kspace: kspace_host has non-zero samples on rows the mask does not select "
                             "(the sparse upload needs apply_mask output; use a dense .to(device) copy otherwise)")
    if out is None:
        out = torch.empty(kspace_host.shape, dtype=torch.float32, device=mask_dev.device)
    _lib.check(_lib.lib().b2s_upload_rows(C.c_void_p(kspace_host.data_ptr()), _p(m), _p(out), b * t, c, h, w, _stream()), "upload_rows")
    return out


def _scratch_for(b, t, c, h, w, device, full=False):
    """Scratch for one call (deterministic coil sum, shapes without a fused plan).  Allocated per call on the current
    stream: torch's caching allocator then owns the stream / CUDA-graph-pool bookkeeping, so concurrent streams
    (pipeline.varnet_hot_path_streams) and captured graphs never share or outlive a buffer (a per-device cache did)."""
    n = b * t * c * h * w * 8 if full else _lib.lib().b2s_scratch_bytes(b, t, c, h, w)
    if n == 0:
        return None, 0
    return torch.empty(n, dtype=torch.uint8, device=device), n


# --------------------------------------------------------------------------- #
# raw calls
# --------------------------------------------------------------------------- #
def raw_fft2c(x: torch.Tensor, inverse: bool, norm: int) -> torch.Tensor:
    _need_cuda(x)
    x = _f32c(x)
    h, w = x.shape[-3], x.shape[-2]
    out = torch.empty_like(x)
    n = x.numel() // (2 * h * w) if h * w else 0
    _lib.check(_lib.lib().b2s_fft2c(_p(x), _p(out), n, h, w, int(inverse), norm, _stream()), "fft2c")
    return out


def raw_fft1c_layout(x: torch.Tensor, outer: int, n: int, inner: int, inverse: bool, norm: int,
                     shift_in: int, shift_out: int, out: torch.Tensor) -> None:
    _lib.check(_lib.lib().b2s_fft1c(_p(x), _p(out), outer, n, inner, int(inverse), norm, shift_in, shift_out,
                                    _stream()), "fft1c")


def raw_sens_expand(image, sens, mode=EXPAND_PLAIN, ref=None, mask_u8=None, v=None, norm=1):
    """image (b,t,h,w,2), sens (b,c,h,w,2) -> (b,t,c,h,w,2)."""
    _need_cuda(image, sens, ref, mask_u8, v)
    b, t, h, w, _ = image.shape
    c = sens.shape[1]
    out = torch.empty((b, t, c, h, w, 2), dtype=torch.float32, device=image.device)
    sc, nsc = _scratch_for(b, t, c, h, w, image.device)
    _lib.check(_lib.lib().b2s_sens_expand(_p(image), _p(sens), _p(out), _p(ref), _p(mask_u8), _p(v), mode,
                                          b, t, c, h, w, norm, _p(sc), nsc, _stream()), "sens_expand")
    return out


def raw_sens_reduce(kspace, mult, wmode=REDUCE_PLAIN, over_frames=False, mask_u8=None, v=None, norm=1):
    """kspace (b,t,c,h,w,2); mult = sens (b,c,h,w,2) -> (b,t,h,w,2), or (over_frames) mult = image
    (b,t,h,w,2) -> (b,c,h,w,2)."""
    _need_cuda(kspace, mult, mask_u8, v)
    b, t, c, h, w, _ = kspace.shape
    shape = (b, c, h, w, 2) if over_frames else (b, t, h, w, 2)
    out = torch.empty(shape, dtype=torch.float32, device=kspace.device)
    det = deterministic()
    if det:
        wmode |= REDUCE_DETERMINISTIC
    sc, nsc = _scratch_for(b, t, c, h, w, kspace.device, full=det)
    _lib.check(_lib.lib().b2s_sens_reduce(_p(kspace), _p(mult), _p(out), _p(mask_u8), _p(v), wmode,
                                          int(over_frames), b, t, c, h, w, norm, _p(sc), nsc, _stream()),
               "sens_reduce")
    return out


def raw_dc_blend(k, ref, mask_u8, v):
    _need_cuda(k, ref, mask_u8, v)
    b, t, c, h, w, _ = k.shape
    out = torch.empty_like(k)
    _lib.check(_lib.lib().b2s_dc_blend(_p(k), _p(ref), _p(mask_u8), _p(v), _p(out), b * t, c, h, w, _stream()),
               "dc_blend")
    return out


def raw_dc_blend_bwd(g, out, ref, mask_u8, v, want_gk, want_gref, want_gv):
    b, t, c, h, w, _ = g.shape
    gk = torch.empty_like(g) if want_gk else None
    gref = torch.empty_like(g) if want_gref else None
    gv = torch.empty(DC_BWD_GV_FLOATS, dtype=torch.float32, device=g.device) if want_gv else None
    _lib.check(_lib.lib().b2s_dc_blend_bwd(_p(g), _p(out), _p(ref), _p(mask_u8), _p(v), _p(gk), _p(gref), _p(gv),
                                           b * t, c, h, w, _stream()), "dc_blend_bwd")
    return gk, gref, (gv[:1] if want_gv else None)


def raw_normal_op(x, sens, mask_u8, v):
    """x (b,t,h,w,2) -> A^H M A x + v x (on-chip kernel, csrc/normal_warp.cuh; see normal_op_supported)."""
    _need_cuda(x, sens, mask_u8, v)
    b, t, h, w, _ = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.lib().b2s_normal_op(_p(x), _p(sens), _p(mask_u8), _p(v), _p(out), b, t, sens.shape[1], h, w,
                                        _stream()), "normal_op")
    return out


def raw_normal_dc(x, sens, mask_u8, v, ssq, bref, magnitude: bool = False):
    """One image-domain VarNet DC cascade: ssq*x - v/(1+v) (A^H M A x - bref)  (inference path, see normal_op_supported).
    `magnitude=True`: the last cascade - returns |.| as (b,t,h,w), the final complex_abs(sens_reduce(.)) of
    VarNet.forward (varnet.py:150-151) fused into the same launch."""
    _need_cuda(x, sens, mask_u8, v, ssq, bref)
    b, t, h, w, _ = x.shape
    if magnitude:
        out = torch.empty(b, t, h, w, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().b2s_normal_dc_abs(_p(x), _p(sens), _p(mask_u8), _p(v), _p(ssq), _p(bref), _p(out), b, t,
                                                sens.shape[1], h, w, _stream()), "normal_dc_abs")
        return out
    out = torch.empty_like(x)
    _lib.check(_lib.lib().b2s_normal_dc(_p(x), _p(sens), _p(mask_u8), _p(v), _p(ssq), _p(bref), _p(out), b, t,
                                        sens.shape[1], h, w, _stream()), "normal_dc")
    return out


def normal_op_supported(h: int, w: int) -> bool:
    """Shapes of the on-chip normal operator (csrc/normal_warp.cuh): 200 or 256 rows, 4-column groups."""
    return h in (200, 256) and w > 0 and w % 4 == 0


def _pad16(n: int):
    """NormUnet.pad (norm_unet.py:75-86): size rounded up to a multiple of 16, (leading, trailing) zero padding."""
    mult = ((n - 1) | 15) + 1
    return mult, (mult - n) // 2


def raw_planes_pack(x, normalise: bool = True, pad: bool = True):
    """x (b,t,h,w,2) -> (xf (b*h,2,wp,tp), yf (b*w,2,hp,tp), ctx): the NCHW inputs of the x-f / y-f U-Nets
    (varnet.py:215-216 + NormUnet.complex_to_chan_dim / norm / pad, norm_unet.py:48-86) in three launches (x is read twice).
    `ctx` carries the statistics and pad sizes for raw_planes_unpack."""
    _need_cuda(x)
    x = _f32c(x)
    b, t, h, w, _ = x.shape
    (hp, ph0), (wp, pw0), (tp, pt0) = (_pad16(h), _pad16(w), _pad16(t)) if pad else ((h, 0), (w, 0), (t, 0))
    sxf = syf = scratch = None
    nbytes = 0
    if normalise:
        sxf = torch.empty(b * h, 2, 2, dtype=torch.float32, device=x.device)
        syf = torch.empty(b * w, 2, 2, dtype=torch.float32, device=x.device)
        nbytes = int(_lib.lib().b2s_planes_scratch_bytes(b, t, h, w))
        scratch = torch.empty(max(nbytes, 16) // 8, dtype=torch.float64, device=x.device)
    xf = torch.empty(b * h, 2, wp, tp, dtype=torch.float32, device=x.device)
    yf = torch.empty(b * w, 2, hp, tp, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().b2s_planes_pack(_p(x), _p(sxf), _p(syf), _p(xf), _p(yf), b, t, h, w, hp, wp, tp, ph0, pw0, pt0,
                                          _p(scratch), nbytes, _stream()), "planes_pack")
    return xf, yf, (sxf, syf, (b, t, h, w, hp, wp, tp, ph0, pw0, pt0))


def raw_planes_unpack(uxf, uyf, ctx):
    """U-Net outputs (layouts of raw_planes_pack) -> 0.5 * (xf_r + yf_r) as (b,t,h,w,2): NormUnet.unpad / unnorm /
    chan_complex_to_last_dim (norm_unet.py:88-113) and varnet.py:228-232 in one launch."""
    sxf, syf, dims = ctx
    b, t, h, w, hp, wp, tp = dims[:7]
    if tuple(uxf.shape) != (b * h, 2, wp, tp) or tuple(uyf.shape) != (b * w, 2, hp, tp):
        raise ValueError(f"raw_planes_unpack: unexpected plane shapes {tuple(uxf.shape)}, {tuple(uyf.shape)}")
    uxf, uyf = _f32c(uxf), _f32c(uyf)
    out = torch.empty(b, t, h, w, 2, dtype=torch.float32, device=uxf.device)
    _lib.check(_lib.lib().b2s_planes_unpack(_p(uxf), _p(uyf), _p(sxf), _p(syf), _p(out), *dims, _stream()), "planes_unpack")
    return out


def raw_temporal_pre(image, xf: bool):
    """image (b,t,h,w,2) -> (x, mean (b,h,w,2))"""
    _need_cuda(image)
    b, t = image.shape[:2]
    hw = image[0, 0].numel() // 2
    x = torch.empty_like(image)
    mean = torch.empty(image.shape[:1] + image.shape[2:], dtype=torch.float32, device=image.device)
    _lib.check(_lib.lib().b2s_temporal_pre(_p(image), _p(x), _p(mean), b, t, hw, int(xf), _stream()), "temporal_pre")
    return x, mean


def raw_temporal_post(x, mean, xf: bool):
    _need_cuda(x, mean)
    b, t = x.shape[:2]
    hw = mean[0].numel() // 2
    out = torch.empty_like(x)
    _lib.check(_lib.lib().b2s_temporal_post(_p(x), _p(mean), _p(out), b, t, hw, int(xf), _stream()), "temporal_post")
    return out


def raw_axpby(a, b, v=None, scale=1.0):
    out = torch.empty_like(a)
    _lib.check(_lib.lib().b2s_axpby(_p(a), _p(b), _p(v), float(scale), _p(out), a.numel(), _stream()), "axpby")
    return out


def raw_dot(a, b, out=None, scratch=None):
    out = torch.empty(1, dtype=torch.float32, device=a.device) if out is None else out
    scratch = torch.empty(1024, dtype=torch.float32, device=a.device) if scratch is None else scratch
    _lib.check(_lib.lib().b2s_dot(_p(a), _p(b), _p(out), a.numel(), _p(scratch), _stream()), "dot")
    return out


# --------------------------------------------------------------------------- #
# autograd Functions
# --------------------------------------------------------------------------- #
class FFT2cFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse, norm):
        ctx.inverse, ctx.norm = inverse, norm
        return raw_fft2c(x, inverse, norm)

    @staticmethod
    def backward(ctx, g):
        return FFT2cFn.apply(g, not ctx.inverse, _ADJ_NORM[ctx.norm]), None, None


def _dense_layout(x: torch.Tensor):
    """(outer, n, inner) of the memory layout of x (..., n, 2) w.r.t. dim -2, or None if x is not a
    dense permutation with the complex pair innermost in memory."""
    if x.stride(-1) != 1 or x.numel() == 0:
        return None
    dims = sorted(range(x.dim() - 1), key=lambda d: (-x.stride(d), d))
    expect = 2
    for d in reversed(dims):
        if x.shape[d] != 1 and x.stride(d) != expect:
            return None
        expect *= x.shape[d]
    pos = dims.index(x.dim() - 2)
    outer = 1
    for d in dims[:pos]:
        outer *= x.shape[d]
    inner = 1
    for d in dims[pos + 1:]:
        inner *= x.shape[d]
    return outer, x.shape[-2], inner


class FFT1cFn(torch.autograd.Function):
    """Centred 1-D transform over dim -2 of (..., n, 2); handles the reference's permuted views
    (varnet.py:211-213) in place, without materialising the permutation."""

    @staticmethod
    def forward(ctx, x, inverse, norm, shift_in, shift_out):
        _need_cuda(x)
        if x.dtype != torch.float32:
            raise TypeError(f"b200sense operators are float32 (got {x.dtype})")
        lay = _dense_layout(x)
        if lay is None:
            x = x.contiguous()
            lay = (x.numel() // (2 * x.shape[-2]), x.shape[-2], 1)
        out = torch.empty_like(x)               # preserve_format keeps the dense permuted strides
        if out.stride() != x.stride():
            x = x.contiguous()
            out = torch.empty_like(x)
            lay = (x.numel() // (2 * x.shape[-2]), x.shape[-2], 1)
        ctx.args = (inverse, norm, shift_in, shift_out)
        if x.numel():
            raw_fft1c_layout(x, lay[0], lay[1], lay[2], inverse, norm, shift_in, shift_out, out)
        return out

    @staticmethod
    def backward(ctx, g):
        inverse, norm, s_in, s_out = ctx.args
        return FFT1cFn.apply(g, not inverse, _ADJ_NORM[norm], -s_out, -s_in), None, None, None, None


class SensExpandFn(torch.autograd.Function):
    """k = Epi(F(S x)); backward: gx = A^H(w g), gS = sum_t conj(x) F^H(w g), gref, gv."""

    @staticmethod
    def forward(ctx, image, sens, ref, mask_u8, v, mode, norm):
        image, sens = _f32c(image), _f32c(sens)
        ref = _f32c(ref) if ref is not None else None
        vd = v.detach() if v is not None else None
        out = raw_sens_expand(image, sens, mode, ref, mask_u8, vd, norm)
        ctx.mode, ctx.norm = mode, norm
        ctx.save_for_backward(image, sens, ref, mask_u8, vd, out if mode == EXPAND_DC else None)
        return out

    @staticmethod
    def backward(ctx, g):
        image, sens, ref, mask_u8, v, out = ctx.saved_tensors
        g = _f32c(g)
        wmode = {EXPAND_PLAIN: REDUCE_PLAIN, EXPAND_MASK: REDUCE_MASK, EXPAND_DC: REDUCE_DCGRAD,
                 EXPAND_RESIDUAL: REDUCE_MASK}[ctx.mode]
        adj = _ADJ_NORM[ctx.norm]
        need = ctx.needs_input_grad
        gx = raw_sens_reduce(g, sens, wmode, False, mask_u8, v, adj) if need[0] else None
        gs = raw_sens_reduce(g, image, wmode, True, mask_u8, v, adj) if need[1] else None
        gref = gv = None
        if ctx.mode == EXPAND_DC and (need[2] or need[4]):
            _, gref, gv = raw_dc_blend_bwd(g, out, ref, mask_u8, v, False, need[2], need[4])
        elif ctx.mode == EXPAND_RESIDUAL and need[2]:
            gref = -g
        return gx, gs, gref, None, gv, None, None


class SensReduceFn(torch.autograd.Function):
    """x = sum_c conj(S) F^H(w k); backward: gk = w F(S g), gS = sum_t conj(g) F^H(w k)."""

    @staticmethod
    def forward(ctx, kspace, sens, mask_u8, wmode, norm):
        kspace, sens = _f32c(kspace), _f32c(sens)
        ctx.wmode, ctx.norm = wmode, norm
        ctx.save_for_backward(kspace, sens, mask_u8)
        return raw_sens_reduce(kspace, sens, wmode, False, mask_u8, None, norm)

    @staticmethod
    def backward(ctx, g):
        kspace, sens, mask_u8 = ctx.saved_tensors
        g = _f32c(g)
        need = ctx.needs_input_grad
        adj = _ADJ_NORM[ctx.norm]
        gk = gs = None
        if need[0]:
            gk = SensExpandFn.apply(g, sens, None, mask_u8, None,
                                    EXPAND_MASK if ctx.wmode == REDUCE_MASK else EXPAND_PLAIN, adj)
        if need[1]:
            gs = raw_sens_reduce(kspace, g, ctx.wmode, True, mask_u8, None, ctx.norm)
        return gk, gs, None, None, None


class DCBlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, k, ref, mask_u8, v):
        k, ref = _f32c(k), _f32c(ref)
        vd = v.detach()
        out = raw_dc_blend(k, ref, mask_u8, vd)
        ctx.save_for_backward(out, ref, mask_u8, vd)
        return out

    @staticmethod
    def backward(ctx, g):
        out, ref, mask_u8, v = ctx.saved_tensors
        need = ctx.needs_input_grad
        gk, gref, gv = raw_dc_blend_bwd(_f32c(g), out, ref, mask_u8, v, need[0], need[1], need[3])
        return gk, gref, None, gv


class NormalOpFn(torch.autograd.Function):
    """H = A^H M A + v is self-adjoint: gx = H g, gv = <g, x>."""

    @staticmethod
    def forward(ctx, x, sens, mask_u8, v):
        x, sens = _f32c(x), _f32c(sens)
        vd = v.detach()
        ctx.save_for_backward(x, sens, mask_u8, vd)
        return raw_normal_op(x, sens, mask_u8, vd)

    @staticmethod
    def backward(ctx, g):
        x, sens, mask_u8, v = ctx.saved_tensors
        g = _f32c(g)
        need = ctx.needs_input_grad
        if need[1]:
            raise RuntimeError("NormalOpFn: no gradient w.r.t. the sensitivity maps here; ops.normal_op composes "
                               "sens_expand / sens_reduce when sens requires grad")
        gx = NormalOpFn.apply(g, sens, mask_u8, v) if need[0] else None
        gv = raw_dot(g, x) if need[3] else None
        return gx, None, None, gv


class TemporalPreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, xf):
        ctx.xf = xf
        x, mean = raw_temporal_pre(_f32c(image), xf)
        return x, mean

    @staticmethod
    def backward(ctx, gx, gmean):
        # x = F_t (I - 11^T/T) image ; mean = 11^T/T image  (both linear)
        t = gx.shape[1]
        gx = _f32c(gx)
        zero_mean = torch.zeros(gx.shape[:1] + gx.shape[2:], dtype=torch.float32, device=gx.device)
        tmp = raw_temporal_post(gx, zero_mean, ctx.xf) if ctx.xf else gx        # adjoint of fft1c = ifft1c
        gi, _ = raw_temporal_pre(tmp, False)                                    # subtract temporal mean
        if gmean is not None:
            gi = gi + (gmean / t).unsqueeze(1)
        return gi, None


class TemporalPostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mean, xf):
        ctx.xf = xf
        return raw_temporal_post(_f32c(x), _f32c(mean), xf)

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        t = g.shape[1]
        need = ctx.needs_input_grad
        gx = gm = None
        if need[0] or need[1]:
            centred, mean_g = raw_temporal_pre(g, False)
            if need[1]:
                gm = mean_g * t
            if need[0]:
                if ctx.xf:
                    # adjoint of ifft1c = fft1c, applied to g itself (not mean-subtracted)
                    b = g.shape[0]
                    hw = mean_g[0].numel() // 2
                    gx = torch.empty_like(g)
                    raw_fft1c_layout(g, b, t, hw, False, 1, (t + 1) // 2, t // 2, gx)
                else:
                    gx = g
        return gx, gm, None


class RssNormalizeFn(torch.autograd.Function):
    """x / rss_complex(x, coil dim 1) on (b,c,h,w,2) — varnet.py:58-59."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = _f32c(x)
        b, c = x.shape[:2]
        out = torch.empty_like(x)
        _lib.check(_lib.lib().b2s_rss_normalize(_p(x), _p(out), b, c, x[0, 0].numel() // 2, _stream()), "rss_normalize")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _f32c(g)
        b, c = x.shape[:2]
        gin = torch.empty_like(x)
        _lib.check(_lib.lib().b2s_rss_normalize_bwd(_p(g), _p(x), _p(gin), b, c, x[0, 0].numel() // 2, _stream()),
                   "rss_normalize_bwd")
        return gin


class AcsMeanFn(torch.autograd.Function):
    """mask_center(mean_t(k), ACS window) — varnet.py:64-71 (device-side window, no host sync)."""

    @staticmethod
    def forward(ctx, kspace, mask_u8):
        _need_cuda(kspace, mask_u8)
        kspace = _f32c(kspace)
        b, t, c, h, w, _ = kspace.shape
        out = torch.empty((b, c, h, w, 2), dtype=torch.float32, device=kspace.device)
        win = torch.empty((b, 2), dtype=torch.int32, device=kspace.device)
        _lib.check(_lib.lib().b2s_acs_mean(_p(kspace), _p(mask_u8), _p(out), _p(win), b, t, c, h, w, _stream()), "acs_mean")
        ctx.save_for_backward(win)
        ctx.t = t
        return out

    @staticmethod
    def backward(ctx, g):
        (win,) = ctx.saved_tensors
        b, c, h, w, _ = g.shape
        rows = torch.arange(h, device=g.device).view(1, h)
        keep = ((rows >= win[:, :1]) & (rows < win[:, :1] + win[:, 1:2])).to(g.dtype)      # (b,h)
        gk = (g * keep.view(b, 1, h, 1, 1) / ctx.t).unsqueeze(1).expand(b, ctx.t, c, h, w, 2)
        return gk.contiguous(), None


# --------------------------------------------------------------------------- #
# element-wise ops of utils/math.py / coil_combine.py (forward kernels; torch autograd via formulas)
# --------------------------------------------------------------------------- #
def _bcast_strides(t: torch.Tensor, shape):
    te = t.expand(*shape, 2)
    if te.stride(-1) != 1 or any(s % 2 for s in te.stride()[:-1]):
        te = t.contiguous().expand(*shape, 2)
    return te, [s // 2 for s in te.stride()[:-1]]


def raw_complex_mul(x, y, conj_y=False):
    _need_cuda(x, y)
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise TypeError("b200sense operators are float32")
    shape = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
    out = torch.empty(*shape, 2, dtype=torch.float32, device=x.device)
    # collapse to <= 6 dims by merging is not attempted; the reference never exceeds 5 leading dims
    if len(shape) > 6:
        raise ValueError("complex_mul: more than 6 leading dimensions")
    xe, sx = _bcast_strides(x, shape)
    ye, sy = _bcast_strides(y, shape)
    n = len(shape)
    arr = C.c_int64 * max(n, 1)
    _lib.check(_lib.lib().b2s_complex_mul(_p(xe), _p(ye), _p(out), n, arr(*shape), arr(*sx), arr(*sy), int(conj_y),
                                          _stream()), "complex_mul")
    return out


def _sum_to(g, shape):
    if tuple(g.shape) == tuple(shape):
        return g
    lead = g.dim() - len(shape)
    if lead:
        g = g.sum(dim=tuple(range(lead)))
    dims = tuple(i for i, (a, b) in enumerate(zip(g.shape, shape)) if a != b)
    return g.sum(dim=dims, keepdim=True) if dims else g


class ComplexMulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        return raw_complex_mul(x, y)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        g = _f32c(g)
        gx = _sum_to(raw_complex_mul(g, y, conj_y=True), x.shape) if ctx.needs_input_grad[0] else None
        gy = _sum_to(raw_complex_mul(g, x, conj_y=True), y.shape) if ctx.needs_input_grad[1] else None
        return gx, gy


class ComplexConjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        x = _f32c(x)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().b2s_complex_conj(_p(x), _p(out), x.numel() // 2, _stream()), "complex_conj")
        return out

    @staticmethod
    def backward(ctx, g):
        return ComplexConjFn.apply(g)


class ComplexAbsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, squared):
        _need_cuda(x)
        x = _f32c(x)
        out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().b2s_complex_abs(_p(x), _p(out), x.numel() // 2, int(squared), _stream()), "complex_abs")
        ctx.squared = squared
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        if ctx.squared:
            return 2 * x * g.unsqueeze(-1), None
        return x * (g / out).unsqueeze(-1), None


class RssFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, is_complex):
        _need_cuda(x)
        x = _f32c(x)
        nd = x.dim() - (1 if is_complex else 0)
        d = dim % x.dim()
        if is_complex and d == x.dim() - 1:
            raise ValueError("rss_complex: dim must not be the complex dimension")
        outer = 1
        for s in x.shape[:d]:
            outer *= s
        inner = 1
        for s in x.shape[d + 1:nd]:
            inner *= s
        oshape = x.shape[:d] + x.shape[d + 1:nd]
        out = torch.empty(oshape, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().b2s_rss(_p(x), _p(out), outer, x.shape[d], inner, int(is_complex), _stream()), "rss")
        ctx.d, ctx.is_complex = d, is_complex
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out = ctx.saved_tensors
        r = (g / out).unsqueeze(ctx.d)
        if ctx.is_complex:
            r = r.unsqueeze(-1)
        return x * r, None, None


# --------------------------------------------------------------------------- #
# public entry points used by functional.py / blocks.py
# --------------------------------------------------------------------------- #
def fft2c(x, norm="ortho", inverse=False):
    return FFT2cFn.apply(x, inverse, _norm(norm))


def fft1c(x, norm="ortho", inverse=False, shift_in=None, shift_out=None):
    n = x.shape[-2]
    s_in = (n + 1) // 2 if shift_in is None else shift_in
    s_out = n // 2 if shift_out is None else shift_out
    return FFT1cFn.apply(x, inverse, _norm(norm), s_in, s_out)


def sens_expand(image, sens, mode=EXPAND_PLAIN, ref=None, mask=None, v=None, norm="ortho"):
    """image (b,t,h,w,2) or (b,t,1,h,w,2); sens (b,c,h,w,2) or (b,1,c,h,w,2) -> (b,t,c,h,w,2)."""
    if image.dim() == 6:
        image = image.squeeze(2)
    if sens.dim() == 6:
        sens = sens.squeeze(1)
    b, t, h, w, _ = image.shape
    m8 = _mask_u8(mask, b, t, h) if mask is not None else None
    vd = _vdev(v, image.device) if (v is not None and not isinstance(v, torch.Tensor)) else v
    if isinstance(vd, torch.Tensor) and vd.numel() != 1:
        raise ValueError("v must be a scalar")
    if isinstance(vd, torch.Tensor):
        vd = vd.reshape(1).to(device=image.device, dtype=torch.float32)
    return SensExpandFn.apply(image, sens, ref, m8, vd, mode, _norm(norm))


def sens_reduce(kspace, sens, mask=None, norm="ortho"):
    """kspace (b,t,c,h,w,2), sens (b,c,h,w,2)|(b,1,c,h,w,2) -> (b,t,h,w,2)."""
    if sens.dim() == 6:
        sens = sens.squeeze(1)
    b, t, c, h, w, _ = kspace.shape
    m8 = _mask_u8(mask, b, t, h) if mask is not None else None
    return SensReduceFn.apply(kspace, sens, m8, REDUCE_MASK if mask is not None else REDUCE_PLAIN, _norm(norm))


def dc_blend(k, ref, mask, v):
    b, t, c, h, w, _ = k.shape
    vd = v.reshape(1).to(device=k.device, dtype=torch.float32) if isinstance(v, torch.Tensor) else _vdev(v, k.device)
    return DCBlendFn.apply(k, ref, _mask_u8(mask, b, t, h), vd)


def normal_op(x, sens, mask, v):
    """x (b,t,h,w,2) -> A^H M A x + v x."""
    if sens.dim() == 6:
        sens = sens.squeeze(1)
    b, t, h, w, _ = x.shape
    vd = v.reshape(1).to(device=x.device, dtype=torch.float32) if isinstance(v, torch.Tensor) else _vdev(v, x.device)
    if normal_op_supported(h, w) and not sens.requires_grad:
        return NormalOpFn.apply(x, sens, _mask_u8(mask, b, t, h), vd)
    k = sens_expand(x, sens, EXPAND_MASK, mask=mask)
    return sens_reduce(k, sens) + vd * x
