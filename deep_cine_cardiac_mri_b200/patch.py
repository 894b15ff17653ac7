"""Re-bind the reference's SENSE / data-consistency path onto the fused kernels.

The reference's models resolve `rec.utils.<fn>` at call time
(`import reconstruction as rec`, e.g. models/varnet.py:8,185), so replacing the
attributes of `reconstruction.utils` re-routes every functional call site; the
block methods are replaced on the classes.  Nothing else of the reference is
touched: regularisers, parameters, checkpoints and `forward` signatures stay.

    import reconstruction.utils, reconstruction.models      # the reference, on sys.path
    from deep_cine_cardiac_mri_b200 import patch
    patch.patch_reference()          # ... run the reference's models as usual ...
    patch.unpatch_reference()
"""
from __future__ import annotations

import importlib
import sys
import types

import torch

from . import blocks
from . import functional as F

_saved = []          # (obj, attr, old_value | _MISSING)
_MISSING = object()

FUNCTIONAL_NAMES = ["fft1c", "ifft1c", "fft2c", "ifft2c", "fftshift", "ifftshift", "roll",
                    "complex_mul", "complex_conj", "complex_abs", "complex_abs_sq", "rss", "rss_complex"]


def _set(obj, attr, value):
    _saved.append((obj, attr, getattr(obj, attr, _MISSING)))
    setattr(obj, attr, value)


def stub_optional_imports():
    """`reconstruction.models` imports `bart` and `h5py` through `reconstruction.data`
    (data/mri_data.py:19,35); neither is needed by the models themselves."""
    for name in ("bart", "h5py"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)


def _xpdnet_xfyf_factory(rec):
    def xfyf_transform(self, image_buffer, i_domain):
        """XPDNetBlock.xfyf_transform (xpdnet.py:449-509) with the temporal transforms on the kernels."""
        b, t, h, w, ch = image_buffer.shape
        ch_out = 2 * self.i_buffer_size
        image_mean = image_buffer.mean(dim=1, keepdim=True)
        x = image_buffer - image_mean
        if self.dynamic_type == 'XF':
            x = blocks.xpd_temporal_fft(x, self.i_buffer_size + 1)
        xf = x.permute(0, 2, 4, 3, 1).reshape(b * h, ch, w, t)
        yf = x.permute(0, 3, 4, 2, 1).reshape(b * w, ch, h, t)
        xf, pad_xf = rec.utils.pad_for_mwcnn(xf, self.n_scales)
        yf, pad_yf = rec.utils.pad_for_mwcnn(yf, self.n_scales)
        if self.weight_sharing:
            model = self.image_net[i_domain // 2]
            xf, yf = model(xf), model(yf)
        else:
            model_xf, model_yf = self.image_net[i_domain // 2]
            xf, yf = model_xf(xf), model_yf(yf)
        xf = rec.utils.unpad_from_mwcnn(xf, pad_xf)
        yf = rec.utils.unpad_from_mwcnn(yf, pad_yf)
        xf_r = xf.view(b, h, 1, ch_out, w, t).permute(0, 5, 2, 1, 4, 3)
        yf_r = yf.view(b, w, 1, ch_out, h, t).permute(0, 5, 2, 4, 1, 3)
        out = 0.5 * (xf_r + yf_r)
        if self.dynamic_type == 'XF':
            out = blocks.xpd_temporal_ifft(out.squeeze(2), self.i_buffer_size).unsqueeze(2)
        m = image_mean.unsqueeze(2)
        in_res = torch.cat([m[..., :self.i_buffer_size], m[..., self.i_buffer_size + 1:-1]], dim=-1)
        return out + in_res
    return xfyf_transform


def patch_reference(functional: bool = True, block_methods: bool = True):
    """Bind the b200sense operators into the already-importable `reconstruction` package."""
    if _saved:
        return
    stub_optional_imports()
    rec = importlib.import_module("reconstruction")
    utils = importlib.import_module("reconstruction.utils")
    if functional:
        for name in FUNCTIONAL_NAMES:
            _set(utils, name, getattr(F, name))
    if not block_methods:
        return
    models = importlib.import_module("reconstruction.models")
    varnet = importlib.import_module("reconstruction.models.varnet")
    cinenet = importlib.import_module("reconstruction.models.cinenet")
    xpdnet = importlib.import_module("reconstruction.models.xpdnet")
    rvar = importlib.import_module("reconstruction.models.recurrent_varnet")
    rcin = importlib.import_module("reconstruction.models.recurrent_cinenet")

    _set(varnet.VarNetBlock, "sens_expand", blocks.sens_expand)
    _set(varnet.VarNetBlock, "sens_reduce", blocks.sens_reduce)
    _set(varnet.VarNetBlock, "xfyf_transform", blocks.varnet_xfyf_transform)
    _set(varnet.VarNetBlock, "forward", blocks.varnet_block_forward)
    _set(varnet.VarNet, "forward", blocks.varnet_forward)
    _set(varnet.SensitivityModel, "forward", blocks.varnet_sens_model_forward)
    _set(varnet.SensitivityModel, "divide_root_sum_of_squares", blocks.divide_root_sum_of_squares)

    _set(cinenet.CineNetBlock, "sens_expand", blocks.sens_expand)
    _set(cinenet.CineNetBlock, "sens_reduce", blocks.sens_reduce)
    _set(cinenet.CineNetBlock, "HOperator", blocks.h_operator)
    _set(cinenet.CineNetBlock, "ConjGrad", blocks.conj_grad)
    _set(cinenet.CineNetBlock, "xfyf_transform", blocks.cinenet_xfyf_transform)
    _set(cinenet.CineNet, "forward", blocks.cinenet_forward)

    _set(xpdnet.ForwardOperator, "forward", blocks.forward_operator_forward)
    _set(xpdnet.BackwardOperator, "forward", blocks.backward_operator_forward)
    _set(xpdnet.SensitivityModel, "forward", blocks.xpdnet_sens_model_forward)
    _set(xpdnet.SensitivityModel, "divide_root_sum_of_squares", blocks.divide_root_sum_of_squares)
    _set(xpdnet.XPDNetBlock, "xfyf_transform", _xpdnet_xfyf_factory(rec))
    _set(xpdnet.XPDNetBlock, "k_domain_correction", blocks.xpdnet_k_domain_correction)
    _set(xpdnet.XPDNetBlock, "i_domain_correction", blocks.xpdnet_i_domain_correction)
    _set(xpdnet.XPDNet, "measurements_residual", blocks.xpdnet_measurements_residual)
    rxpd = importlib.import_module("reconstruction.models.recurrent_xpdnet")
    _set(rxpd.XPDNet_RNN, "k_domain_correction", blocks.xpdnet_rnn_k_domain_correction)
    _set(rxpd.XPDNet_RNN, "update_image_buffer", blocks.xpdnet_update_image_buffer)
    _set(rxpd.XPDNet_RNN, "measurements_residual", blocks.xpdnet_measurements_residual)

    _set(rvar.VarNet_RNN, "sens_expand", blocks.varnet_rnn_sens_expand)
    _set(rvar.VarNet_RNN, "sens_reduce", blocks.varnet_rnn_sens_reduce)
    _set(rvar.VarNet_RNN, "data_consistency", blocks.varnet_rnn_data_consistency)
    _set(rcin.CineNet_RNN, "sens_expand", blocks.sens_expand)
    _set(rcin.CineNet_RNN, "sens_reduce", blocks.sens_reduce)
    _set(rcin.CineNet_RNN, "HOperator", blocks.h_operator)
    _set(rcin.CineNet_RNN, "ConjGrad", blocks.conj_grad)

    # training loss (utils/losses.py:25-58): same module, same buffer; forward on the fused SSIM kernels
    try:
        losses = importlib.import_module("reconstruction.utils.losses")
    except ImportError:
        losses = None
    if losses is not None and hasattr(losses, "SSIMLoss"):
        _set(losses.SSIMLoss, "forward", _ssim_loss_forward)
    return models


def _ssim_loss_forward(self, Xt, Yt, data_range=None):
    from . import metrics
    return metrics.ssim_loss(Xt, Yt, self.win_size, self.k1, self.k2)


def unpatch_reference():
    while _saved:
        obj, attr, old = _saved.pop()
        if old is _MISSING:
            delattr(obj, attr)
        else:
            setattr(obj, attr, old)


def is_patched() -> bool:
    return bool(_saved)
