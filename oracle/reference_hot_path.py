"""The hot-path workload of bench.py executed by the UNMODIFIED reference code (baseline/_ref, see load_reference.py).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product package).

bench.py's step is "the SENSE / data-consistency path of a 12-cascade XF-VarNet forward, regularisers = identity".  The
same workload on the reference is its own `VarNet` (models/varnet.py:88-151) with the two U-Nets swapped for identity
modules: every line that then executes - `SensitivityModel.forward` (:62-86), `VarNetBlock.forward` (:244-282) with
`sens_reduce`, `xfyf_transform`'s temporal fft1c / ifft1c and permutes, `sens_expand`, the soft-DC blend, and the final
`complex_abs(sens_reduce)` (:150-151) - is the reference's code calling `reconstruction.utils`, on whatever device the
tensors live (CPU for the `cpu_baseline` / `--impl reference` arm, the GPU for the `gpu_reference` extra).
"""
from __future__ import annotations

import torch

from . import load_reference


class _Identity(torch.nn.Module):
    def forward(self, x):
        return x


def reference_varnet_identity(num_cascades: int = 12, dynamic_type: str = "XF"):
    """reconstruction.models.VarNet whose regularisers (cascade NormUnet, sensitivity NormUnet) are identities."""
    rec = load_reference.load()
    model = rec.models.VarNet(num_cascades=num_cascades, sens_chans=2, sens_pools=1, chans=2, pools=1,
                              dynamic_type=dynamic_type, weight_sharing=True)
    ident = _Identity()
    model.model = ident
    for blk in model.cascades:                       # (the reference shares one regulariser object across cascades)
        blk.model = ident
    model.sens_net.norm_unet = ident
    return model.eval()
