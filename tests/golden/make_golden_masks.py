#!/usr/bin/env python
"""Golden undersampling masks from the REFERENCE's own mask functions (data/subsample.py:75-215).

Runs only where /root/reference (or baseline/_ref) exists; writes tests/golden/golden_v3_masks.npz.  Each entry:
the mask the reference returns for `shape` with numpy's global stream and the mask function both seeded with `seed`.

    python tests/golden/make_golden_masks.py
"""
import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
CASES = [("random", [10], [4], (15, 1, 200, 200, 2)), ("random", [10], [4], (25, 1, 200, 200, 2)), ("random", [8], [8], (30, 1, 256, 256, 2)),
         ("random", [4, 6], [4, 8], (6, 1, 64, 48, 2)), ("equispaced", [0.08], [4], (15, 1, 200, 200, 2)), ("equispaced", [0.04, 0.08], [8, 4], (3, 1, 256, 256, 2))]
SEEDS = (0, 1, 12345)


def main():
    from oracle import load_reference
    load_reference.load(models=False)
    R = importlib.import_module("reconstruction.data.subsample")
    out = {}
    for ci, (kind, cf, acc, shape) in enumerate(CASES):
        for seed in SEEDS:
            np.random.seed(seed)
            m = R.create_mask_for_mask_type(kind, cf, acc)(shape, seed)
            out[f"c{ci}_s{seed}"] = m.numpy()
    np.savez_compressed(HERE / "golden_v3_masks.npz", **out)
    print("wrote", len(out), "masks")


if __name__ == "__main__":
    main()
