#!/usr/bin/env python
"""CineNet SENSE/CG hot path (b = 1 per call): slices/s against the number of independent slices in flight (streams in one CUDA graph)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import pipeline, synth
dev = torch.device("cuda", 0)
v = torch.ones(1, device=dev)
sets = []
for i in range(8):
    cc = synth.cine_case(5000 + i, 1, 25, 20, 200, 200)
    sets.append(tuple(torch.from_numpy(cc[k]).to(dev) for k in ("masked_kspace", "mask", "sens")))
for n in (1, 2, 4, 6, 8):
    cs = sets[:n]
    step = lambda: torch.cat(pipeline.run_on_streams(lambda a, m_, s_: pipeline.cinenet_hot_path(a, m_, s_, v, 10, 4), cs), 0)
    g = pipeline.Graphed(step, warmup=1)
    for _ in range(2): g()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4): g()
    b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    print(f"{n} slices in flight: {4 * n / sec:7.1f} slices/s  ({sec / (4 * n) * 1e3:.3f} ms per slice)")
