"""deep_cine_cardiac_mri_b200 — B200-native SENSE / data-consistency operators
for the unrolled cine-MRI cascades of f78bono/deep-cine-cardiac-mri.

Only the hot path lives here: `csrc/` (sm_100a kernels + C ABI), `ops` (custom
autograd operators over the ABI), `functional` (the reference's fastMRI-style
API), `blocks` + `patch` (drop-ins for the reference's block methods), `synth`
(synthetic cine k-space) and `dist` (slice sharding over the GPUs of one box).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def build(verbose: bool = False):
    return _lib.build(verbose)
