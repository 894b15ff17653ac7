"""Whole-slice SENSE/DC hot path of an unrolled XF/XT-VarNet forward on the fused kernels.

This is the path `models/varnet.py:143-151` (VarNet.forward) walks between the
regularisers: SensitivityModel pre/post, then per cascade A^H -> temporal head ->
[regulariser] -> temporal tail -> A fused with the soft-DC blend, then |A^H k|.
The regularisers are passed in as callables (the reference's cuDNN U-Nets, untouched);
`None` means identity, which is what the hot-path benchmark measures.  No host
synchronisation happens anywhere in here, so the whole call is CUDA-graph capturable.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Union

import torch

from . import ops
from . import functional as F
from . import blocks


def sensitivity_maps(masked_kspace: torch.Tensor, mask: torch.Tensor,
                     sens_unet: Optional[Callable] = None) -> torch.Tensor:
    """(b,t,c,h,w,2), mask (b,t,1,h,1,1) -> (b,1,c,h,w,2)  (models/varnet.py:62-86)."""
    x = blocks._sens_pre(masked_kspace, mask)
    if sens_unet is not None:
        x = sens_unet(x)
    return ops.RssNormalizeFn.apply(x).unsqueeze(1)


def varnet_hot_path(masked_kspace: torch.Tensor, mask: torch.Tensor,
                    v: Union[float, torch.Tensor, Sequence] = 1.0, n_cascades: int = 12, xf: bool = True,
                    regulariser: Optional[Callable] = None, sens_unet: Optional[Callable] = None,
                    sens_maps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns the reconstructed magnitude cine (b,t,h,w)."""
    b, t, c, h, w, _ = masked_kspace.shape
    sens = sensitivity_maps(masked_kspace, mask, sens_unet) if sens_maps is None else sens_maps
    m8 = ops._mask_u8(mask, b, t, h)
    vs = list(v) if isinstance(v, (list, tuple)) else [v] * n_cascades
    vs = [x if isinstance(x, torch.Tensor) else ops._vdev(x, masked_kspace.device) for x in vs]
    k = masked_kspace
    for i in range(n_cascades):
        img = ops.sens_reduce(k, sens)                                   # A^H k            (b,t,h,w,2)
        x, mean = ops.TemporalPreFn.apply(img, xf)                       # - mean_t, fft1c_t
        if regulariser is not None:
            x = regulariser(x.unsqueeze(2)).squeeze(2)
        model_out = ops.TemporalPostFn.apply(x, mean, xf)                # ifft1c_t, + mean_t
        k = ops.SensExpandFn.apply(model_out, sens.squeeze(1), masked_kspace, m8, vs[i], ops.EXPAND_DC, 1)
    return F.complex_abs(ops.sens_reduce(k, sens))


_side_streams = {}


def varnet_hot_path_streams(masked_kspace: torch.Tensor, mask: torch.Tensor, v=1.0, n_cascades: int = 12, xf: bool = True,
                            n_streams: int = 2, **kw) -> torch.Tensor:
    """`varnet_hot_path` with the slice batch split over `n_streams` CUDA streams.

    Slices are independent, and every fused kernel is a persistent grid of one CTA per SM whose last
    round is ragged (600 images x 2 halves on 148 SMs = 8.1 rounds -> 9): with two streams the tail of one stream's
    kernel is filled by the head of the other's.  Fork/join is by events only, so the call is CUDA-graph capturable."""
    b = masked_kspace.shape[0]
    n = max(1, min(n_streams, b))
    if n == 1:
        return varnet_hot_path(masked_kspace, mask, v, n_cascades, xf, **kw)
    cur = torch.cuda.current_stream()
    dev = masked_kspace.device
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    bounds = [(i * b) // n for i in range(n + 1)]
    outs = []
    for i in range(n):
        st = pool[i]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            o = varnet_hot_path(masked_kspace[bounds[i]:bounds[i + 1]], mask[bounds[i]:bounds[i + 1]], v, n_cascades, xf, **kw)
        o.record_stream(cur)
        outs.append(o)
    for i in range(n):
        cur.wait_stream(pool[i])
    return torch.cat(outs, 0)


def varnet_hot_path_image_domain(masked_kspace: torch.Tensor, mask: torch.Tensor,
                                 v: Union[float, torch.Tensor, Sequence] = 1.0, n_cascades: int = 12, xf: bool = True,
                                 regulariser: Optional[Callable] = None, sens_unet: Optional[Callable] = None,
                                 sens_maps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same function as `varnet_hot_path` (inference), without ever materialising k-space.

    Between cascades the predicted k-space is only consumed by the next `sens_reduce`
    (varnet.py:253) or by the final one (varnet.py:150-151), and
        A^H[ DC(A x, ref) ] = (sum_c |S_c|^2) x - eta (A^H M A x - A^H ref),   eta = v/(1+v),
    because F^H F = I and the mask commutes with the transform along w.  Each cascade is one
    launch of the on-chip normal-operator kernel (b2s_normal_dc): 2I + S bytes instead of 3K + 2I + 2S.
    """
    b, t, c, h, w, _ = masked_kspace.shape
    if not ops.normal_op_supported(h, w) or (torch.is_grad_enabled() and masked_kspace.requires_grad):
        return varnet_hot_path(masked_kspace, mask, v, n_cascades, xf, regulariser, sens_unet, sens_maps)
    sens = sensitivity_maps(masked_kspace, mask, sens_unet) if sens_maps is None else sens_maps
    s5 = ops._f32c(sens.squeeze(1))
    m8 = ops._mask_u8(mask, b, t, h)
    vs = list(v) if isinstance(v, (list, tuple)) else [v] * n_cascades
    vs = [x.detach().reshape(1) if isinstance(x, torch.Tensor) else ops._vdev(x, masked_kspace.device) for x in vs]
    ssq = F.complex_abs_sq(s5).sum(dim=1).contiguous()                     # (b,h,w)  sum_c |S_c|^2
    bref = ops.raw_sens_reduce(ops._f32c(masked_kspace), s5, ops.REDUCE_MASK, False, m8, None, 1)   # A^H M ref
    # cascade 0 starts from k = ref, i.e. from A^H k with NO mask (varnet.py:253), exactly like varnet_hot_path and the
    # reference; equal to bref only when the unsampled rows of the input really are zero
    img = ops.raw_sens_reduce(ops._f32c(masked_kspace), s5, ops.REDUCE_PLAIN, False, None, None, 1)
    if n_cascades == 0:
        return F.complex_abs(img)
    for i in range(n_cascades):
        x, mean = ops.raw_temporal_pre(img, xf)
        if regulariser is not None:
            x = regulariser(x.unsqueeze(2)).squeeze(2).contiguous()
        model_out = ops.raw_temporal_post(x, mean, xf)
        img = ops.raw_normal_dc(model_out, s5, m8, vs[i], ssq, bref, magnitude=(i == n_cascades - 1))   # last: |A^H k| fused
    return img


def cinenet_hot_path(masked_kspace: torch.Tensor, mask: torch.Tensor, sens_maps: torch.Tensor,
                     v: Union[float, torch.Tensor] = 1.0, n_cascades: int = 10, cg_iters: int = 4,
                     regulariser: Optional[Callable] = None) -> torch.Tensor:
    """SENSE/CG hot path of a CineNet / CineNet_RNN forward (models/cinenet.py:61-73, 222-257;
    recurrent_cinenet.py:127-187): x_ref = A^H y, then per cascade [regulariser] and `cg_iters` steps of CG on
    (A^H M A + v) x = x_ref + v x_reg.  Every H application is one on-chip normal-operator launch and the CG
    scalars stay on the device (the reference does 3 `.item()` host syncs per iteration).  b == 1 per call keeps
    the reference's semantics (its dot products span the whole batch, cinenet.py:148).  Returns (b,t,h,w)."""
    from . import blocks
    b, t, c, h, w, _ = masked_kspace.shape
    vd = v.detach().reshape(1) if isinstance(v, torch.Tensor) else ops._vdev(v, masked_kspace.device)
    x_ref = ops.sens_reduce(masked_kspace, sens_maps)                        # (b,t,h,w,2)
    x = x_ref
    for _ in range(n_cascades):
        model_out = x if regulariser is None else regulariser(x.unsqueeze(2)).squeeze(2).contiguous()
        rhs = ops.raw_axpby(ops._f32c(x_ref), ops._f32c(model_out), vd, 1.0)  # x_ref + v * model_out
        x = blocks._cg_inference(model_out, rhs, mask, sens_maps, vd, cg_iters)
    return F.complex_abs(x)


def run_on_streams(fn: Callable, arg_sets: Sequence[tuple], **kw) -> list:
    """`[fn(*args, **kw) for args in arg_sets]` with every call on its own CUDA stream (fork/join by events, so the whole
    thing is graph-capturable).  For independent cine slices whose kernels do not fill the GPU on their own - e.g. the
    CineNet CG chain, b = 1 per call: 250 normal-operator CTAs on 148 SMs, then dots and axpys - the slices overlap."""
    if not arg_sets:
        return []
    cur = torch.cuda.current_stream()
    dev = next(a for a in arg_sets[0] if isinstance(a, torch.Tensor)).device
    pool = _side_streams.setdefault(dev.index, [])
    while len(pool) < len(arg_sets):
        pool.append(torch.cuda.Stream(dev))
    outs = []
    for i, args in enumerate(arg_sets):
        pool[i].wait_stream(cur)
        with torch.cuda.stream(pool[i]):
            o = fn(*args, **kw)
        o.record_stream(cur)
        outs.append(o)
    for i in range(len(arg_sets)):
        cur.wait_stream(pool[i])
    return outs


class Graphed:
    """CUDA-graph capture of a hot-path call with static input buffers.

    Nothing on the path synchronises with the host (device-side ACS window, device-side CG scalars, `v` read from
    device memory), so a whole forward is capturable: `g = Graphed(varnet_hot_path, mk, mask, v, 12)`, then
    `g.inputs[0].copy_(new_kspace); out = g()` replays every kernel with one launch."""

    def __init__(self, fn: Callable, *args, warmup: int = 2, **kwargs):
        self.inputs = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(warmup):                       # plans / function attributes / scratch are set up here
                fn(*self.inputs, **kwargs)
        torch.cuda.current_stream().wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.output = fn(*self.inputs, **kwargs)

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.output


def hot_path_algorithmic_bytes(b: int, t: int, c: int, h: int, w: int, n_cascades: int) -> dict:
    """Algorithmic HBM bytes of the calls above, fp32."""
    K, I, S = b * t * c * h * w * 8, b * t * h * w * 8, b * c * h * w * 8
    return {
        "sens_reduce": K + S + I, "sens_expand_dc": I + S + 2 * K, "dc_step": 3 * K + 2 * I + 2 * S,
        "temporal": 2 * (2 * I + I // t), "total": n_cascades * (3 * K + 2 * I + 2 * S + 2 * (2 * I + I // t))
        + (K + S + I) + 2 * S + b * c * h * w * 8,
        "K": K, "I": I, "S": S,
    }
