"""Dispatcher-level PyTorch custom operators (`torch.ops.b200sense.*`).

`ops.py` exposes the kernels as `torch.autograd.Function`s (lowest call overhead; what `functional` /
`blocks` / `pipeline` use).  This module registers the same kernels with `torch.library.custom_op`, with
fake (meta) implementations and autograd formulas, so that they are first-class operators for
`torch.compile`, `torch.export`, `torch.library.opcheck` and anything else that walks the dispatcher:

    torch.ops.b200sense.fft2c(x, inverse, norm)                       # utils/fftc.py:59-110
    torch.ops.b200sense.sens_expand(image, sens, ref, mask, v, mode, norm)   # varnet.py:181-185 (+281-282)
    torch.ops.b200sense.sens_reduce(kspace, sens, mask, wmode, norm)  # varnet.py:187-194
    torch.ops.b200sense.normal_op(x, sens, mask, v)                   # cinenet.py:121-133

Tensor layouts are the C ABI's (include/b200sense.h): image (b,t,h,w,2), sens (b,c,h,w,2), k-space
(b,t,c,h,w,2), mask uint8 (b,t,h), v float32 (1,).  `norm`: 0 backward, 1 ortho, 2 forward.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

_ADJ = {0: 2, 1: 1, 2: 0}


# ------------------------------------------------------------------ fft2c
@torch.library.custom_op("b200sense::fft2c", mutates_args=())
def fft2c(x: torch.Tensor, inverse: bool, norm: int) -> torch.Tensor:
    return ops.raw_fft2c(x, inverse, norm)


@fft2c.register_fake
def _(x, inverse, norm):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _fft2c_setup(ctx, inputs, output):
    _, ctx.inverse, ctx.norm = inputs


def _fft2c_bwd(ctx, g):
    return torch.ops.b200sense.fft2c(g.contiguous(), not ctx.inverse, _ADJ[ctx.norm]), None, None


fft2c.register_autograd(_fft2c_bwd, setup_context=_fft2c_setup)


# ------------------------------------------------------------------ sens_expand
@torch.library.custom_op("b200sense::sens_expand", mutates_args=())
def sens_expand(image: torch.Tensor, sens: torch.Tensor, ref: Optional[torch.Tensor], mask: Optional[torch.Tensor],
                v: Optional[torch.Tensor], mode: int, norm: int) -> torch.Tensor:
    return ops.raw_sens_expand(image.contiguous(), sens.contiguous(), mode, None if ref is None else ref.contiguous(),
                               mask, v, norm)


@sens_expand.register_fake
def _(image, sens, ref, mask, v, mode, norm):
    b, t, h, w, _ = image.shape
    return image.new_empty((b, t, sens.shape[1], h, w, 2))


def _expand_setup(ctx, inputs, output):
    image, sens, ref, mask, v, ctx.mode, ctx.norm = inputs
    ctx.save_for_backward(image, sens, ref, mask, v, output if ctx.mode == ops.EXPAND_DC else None)


def _expand_bwd(ctx, g):
    image, sens, ref, mask, v, out = ctx.saved_tensors
    g = g.contiguous()
    wmode = {ops.EXPAND_PLAIN: ops.REDUCE_PLAIN, ops.EXPAND_MASK: ops.REDUCE_MASK, ops.EXPAND_DC: ops.REDUCE_DCGRAD,
             ops.EXPAND_RESIDUAL: ops.REDUCE_MASK}[ctx.mode]
    need = ctx.needs_input_grad
    adj = _ADJ[ctx.norm]
    gx = ops.raw_sens_reduce(g, sens.contiguous(), wmode, False, mask, v, adj) if need[0] else None
    gs = ops.raw_sens_reduce(g, image.contiguous(), wmode, True, mask, v, adj) if need[1] else None
    gref = gv = None
    if ctx.mode == ops.EXPAND_DC and (need[2] or need[4]):
        _, gref, gv = ops.raw_dc_blend_bwd(g, out, ref.contiguous(), mask, v, False, need[2], need[4])
    elif ctx.mode == ops.EXPAND_RESIDUAL and need[2]:
        gref = -g
    return gx, gs, gref, None, gv, None, None


sens_expand.register_autograd(_expand_bwd, setup_context=_expand_setup)


# ------------------------------------------------------------------ sens_reduce
@torch.library.custom_op("b200sense::sens_reduce", mutates_args=())
def sens_reduce(kspace: torch.Tensor, sens: torch.Tensor, mask: Optional[torch.Tensor], wmode: int, norm: int) -> torch.Tensor:
    return ops.raw_sens_reduce(kspace.contiguous(), sens.contiguous(), wmode, False, mask, None, norm)


@sens_reduce.register_fake
def _(kspace, sens, mask, wmode, norm):
    b, t, c, h, w, _ = kspace.shape
    return kspace.new_empty((b, t, h, w, 2))


def _reduce_setup(ctx, inputs, output):
    kspace, sens, mask, ctx.wmode, ctx.norm = inputs
    ctx.save_for_backward(kspace, sens, mask)


def _reduce_bwd(ctx, g):
    kspace, sens, mask = ctx.saved_tensors
    g = g.contiguous()
    need = ctx.needs_input_grad
    gk = gs = None
    if need[0]:
        gk = torch.ops.b200sense.sens_expand(g, sens, None, mask, None,
                                             ops.EXPAND_MASK if ctx.wmode == ops.REDUCE_MASK else ops.EXPAND_PLAIN, _ADJ[ctx.norm])
    if need[1]:
        gs = ops.raw_sens_reduce(kspace.contiguous(), g, ctx.wmode, True, mask, None, ctx.norm)
    return gk, gs, None, None, None


sens_reduce.register_autograd(_reduce_bwd, setup_context=_reduce_setup)


# ------------------------------------------------------------------ normal operator
@torch.library.custom_op("b200sense::normal_op", mutates_args=())
def normal_op(x: torch.Tensor, sens: torch.Tensor, mask: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    return ops.raw_normal_op(x.contiguous(), sens.contiguous(), mask, v)


@normal_op.register_fake
def _(x, sens, mask, v):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


def _normal_setup(ctx, inputs, output):
    x, sens, mask, v = inputs
    ctx.save_for_backward(x, sens, mask, v)


def _normal_bwd(ctx, g):
    x, sens, mask, v = ctx.saved_tensors
    need = ctx.needs_input_grad
    if need[1]:
        raise RuntimeError("b200sense.normal_op: gradient w.r.t. the sensitivity maps is not provided by the on-chip normal "
                           "operator; use ops.normal_op (it composes sens_expand / sens_reduce when sens requires grad)")
    gx = torch.ops.b200sense.normal_op(g.contiguous(), sens, mask, v) if need[0] else None     # H is self-adjoint
    gv = ops.raw_dot(g.contiguous(), x.contiguous()) if need[3] else None
    return gx, None, None, gv


normal_op.register_autograd(_normal_bwd, setup_context=_normal_setup)
