"""Locate and import the UNMODIFIED reference package (`reconstruction`) for tests and the reference arms of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product package).

The reference is pure Python without packaging metadata (no setup.py / pyproject: `pip install /root/reference`
has nothing to install), so `__graft_entry__.build()` copies its `reconstruction/` package verbatim to
`baseline/_ref/reconstruction` (git-ignored, shipped to the GPU box by gpurun).  `reconstruction.models` imports `bart`
and `h5py` through `reconstruction.data` (data/mri_data.py:19,35, data/transforms.py:29); neither is used by the model
classes, so empty stand-in modules are registered when the real ones are missing.
"""
from __future__ import annotations

import importlib
import shutil
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SHIPPED = ROOT / "baseline" / "_ref"
SOURCE = Path("/root/reference")


def ship_reference(verbose: bool = False) -> bool:
    """Copy /root/reference/reconstruction (and the CLI scripts) to baseline/_ref when the source is present."""
    if not (SOURCE / "reconstruction").is_dir():
        return (SHIPPED / "reconstruction").is_dir()
    SHIPPED.mkdir(parents=True, exist_ok=True)
    for sub in ("reconstruction", "traintest_scripts"):
        src, dst = SOURCE / sub, SHIPPED / sub
        if src.is_dir():
            if dst.exists():
                shutil.rmtree(dst)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    if verbose:
        print(f"shipped the reference to {SHIPPED}")
    return True


def reference_root() -> Path | None:
    for cand in (SHIPPED, SOURCE):
        if (cand / "reconstruction" / "utils" / "fftc.py").is_file():
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def load(models: bool = True):
    """Returns the imported `reconstruction` package (with .utils and, if asked, .models loaded)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not available: run __graft_entry__.build() where /root/reference exists")
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    loaded = sys.modules.get("reconstruction")
    if loaded is not None and not str(getattr(loaded, "__file__", "") or "").startswith(str(root)):
        for name in [n for n in sys.modules if n == "reconstruction" or n.startswith("reconstruction.")]:
            del sys.modules[name]                                    # a different `reconstruction` (e.g. a test double) was imported
        sys.path.remove(str(root)); sys.path.insert(0, str(root))
    for name in ("bart", "h5py"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    rec = importlib.import_module("reconstruction")
    importlib.import_module("reconstruction.utils")
    if models:
        importlib.import_module("reconstruction.models")
    return rec
