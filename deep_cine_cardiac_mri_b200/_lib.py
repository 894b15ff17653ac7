"""ctypes loader of libb2sense.so (C ABI: include/b200sense.h).

The library is built in-tree (``python __graft_entry__.py`` / ``make -C csrc``)
and must exist: there is NO CPU or library fallback behind these operators —
``lib()`` raises ``RuntimeError`` if the shared object is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["B2S_LIB"]) if os.environ.get("B2S_LIB") else PKG / "lib" / "libb2sense.so"
_lib = None

_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> argtypes  (restype is int unless listed in _RESTYPE)
SIGNATURES = {
    "b2s_version": [],
    "b2s_last_error": [],
    "b2s_launch_count": [_i],
    "b2s_set_sm_reserve": [_i],
    "b2s_debug_strip_status": [],
    "b2s_set_fused_path": [_i],
    "b2s_has_fused_plan": [_i, _i],
    "b2s_scratch_bytes": [_i, _i, _i, _i, _i],
    "b2s_fft2c": [_p, _p, _i64, _i, _i, _i, _i, _p],
    "b2s_fft1c": [_p, _p, _i64, _i, _i64, _i, _i, _i, _i, _p],
    "b2s_sens_expand": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p],
    "b2s_sens_reduce": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p],
    "b2s_apply_mask": [_p, _p, _p, _i64, _i, _i, _i, _p],
    "b2s_dc_blend": [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _p],
    "b2s_dc_blend_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _p],
    "b2s_complex_mul": [_p, _p, _p, _i, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), _i, _p],
    "b2s_complex_conj": [_p, _p, _i64, _p],
    "b2s_complex_abs": [_p, _p, _i64, _i, _p],
    "b2s_rss": [_p, _p, _i64, _i64, _i64, _i, _p],
    "b2s_acs_mean": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2s_rss_normalize": [_p, _p, _i, _i, _i64, _p],
    "b2s_rss_normalize_bwd": [_p, _p, _p, _i, _i, _i64, _p],
    "b2s_temporal_pre": [_p, _p, _p, _i, _i, _i64, _i, _p],
    "b2s_temporal_post": [_p, _p, _p, _i, _i, _i64, _i, _p],
    "b2s_planes_stats": [_p, _p, _p, _i, _i, _i, _i, _p],
    "b2s_planes_scratch_bytes": [_i, _i, _i, _i],
    "b2s_planes_pack": [_p, _p, _p, _p, _p] + [_i] * 10 + [_p, _sz, _p],
    "b2s_planes_unpack": [_p, _p, _p, _p, _p] + [_i] * 10 + [_p],
    "b2s_normal_op": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2s_normal_dc": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2s_normal_dc_abs": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2s_normal_op_dot": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2s_cg_blocks": [_i64],
    "b2s_cg_update": [_p, _p, _p, _p, _p, _i, _p, _p, _i64, _p],
    "b2s_cg_direction": [_p, _p, _p, _p, _p, _i64, _p],
    "b2s_dot": [_p, _p, _p, _i64, _p, _p],
    "b2s_axpy_ratio": [_p, _p, _p, _p, _f, _i64, _p],
    "b2s_xpay_ratio": [_p, _p, _p, _p, _i64, _p],
    "b2s_axpby": [_p, _p, _p, _f, _p, _i64, _p],
    "b2s_dc_step_ws_bytes": [_i, _i, _i, _i, _i],
    "b2s_dc_step_host": [_p, _p, _p, _p, _f, _p, _i, _i, _i, _i, _i, _p, _sz, _p],
    "b2s_upload_rows": [_p, _p, _p, _i64, _i, _i, _i, _p],
    "b2s_ssim_scratch_floats": [_i, _i, _i, _i],
    "b2s_ssim_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p],
    "b2s_ssim_bwd": [_p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _f, _f, _p, _p],
    "b2s_frame_max": [_p, _p, _i, _i, _i64, _p],
    "b2s_err_stats": [_p, _p, _i64, _p, _p, _p],
}
_RESTYPE = {"b2s_launch_count": C.c_ulonglong, "b2s_last_error": C.c_char_p, "b2s_scratch_bytes": _sz, "b2s_dc_step_ws_bytes": _sz,
            "b2s_ssim_scratch_floats": _sz, "b2s_planes_scratch_bytes": _sz}


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into lib/libb2sense.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(PKG / "csrc"), "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libb2sense.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: the b200sense CUDA library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU/cuFFT fallback for these operators.")
        handle = C.CDLL(str(LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI and the header diverge
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, _i)
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    """Map ABI status codes to the exceptions the reference's callers expect."""
    if rc == 0:
        return
    msg = lib().b2s_last_error().decode() or what
    if rc in (1, 2):
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
