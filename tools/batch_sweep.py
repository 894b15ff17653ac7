#!/usr/bin/env python
"""Headline step (12-cascade k-space hot path in one CUDA graph) against slices per step and streams: the persistent grids
quantise (150 images per slice on 148 SMs) while small batches keep the k-space between expand+DC and the next reduce in L2."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from deep_cine_cardiac_mri_b200 import pipeline, synth
dev = torch.device("cuda", 0)
v = torch.ones(1, device=dev)
cases = [synth.cine_case(1000 + i, 1, 15, 10, 200, 200) for i in range(4)]
def inputs(n):
    mk = np.concatenate([cases[i % 4]["masked_kspace"] for i in range(n)], 0)
    mask = np.concatenate([cases[i % 4]["mask"] for i in range(n)], 0)
    return torch.from_numpy(mk).to(dev), torch.from_numpy(mask).to(dev)
for n in (4, 8, 12, 16, 24, 32):
    mk, mask = inputs(n)
    for ns in (1, 2, 3, 4, 6, 8):
        if ns > n: continue
        step = lambda: pipeline.varnet_hot_path_streams(mk, mask, v, 12, xf=True, n_streams=ns)
        g = pipeline.Graphed(step, warmup=2)
        for _ in range(3): g()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(4, 64 // n)
        a.record()
        for _ in range(reps): g()
        b.record(); torch.cuda.synchronize()
        sec = a.elapsed_time(b) * 1e-3
        print(f"slices {n:2d} streams {ns}: {n * reps / sec:7.1f} slices/s  ({sec / reps * 1e3:.2f} ms per step)", flush=True)
        del g
    del mk, mask
    torch.cuda.empty_cache()
