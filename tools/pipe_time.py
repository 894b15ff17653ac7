#!/usr/bin/env python
"""Whole 12-cascade hot path (eager, one stream) on synthetic slices of a given shape: ms per step and the mean
duration of the fused expand+DC launches inside it (the context bench.py measures; operator microbenchmarks on
random data can rank kernel variants differently)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from deep_cine_cardiac_mri_b200 import ops, pipeline, synth
b, t, c, h, w = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (4, 15, 10, 200, 200))]
cases = [synth.cine_case(100 + i, 1, t, c, h, w) for i in range(b)]
mk = torch.from_numpy(np.concatenate([q["masked_kspace"] for q in cases], 0)).cuda()
mask = torch.from_numpy(np.concatenate([q["mask"] for q in cases], 0)).cuda()
v = torch.ones(1, device="cuda")
ev = []
orig = ops.raw_sens_expand
def timed(*a, **kw):
    if timed.on and len(a) > 2 and a[2] == ops.EXPAND_DC:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = orig(*a, **kw); e1.record(); ev.append((e0, e1)); return o
    return orig(*a, **kw)
timed.on = False
ops.raw_sens_expand = timed
with torch.no_grad():
    for _ in range(3): pipeline.varnet_hot_path(mk, mask, v, 12, xf=True)
    torch.cuda.synchronize()
    timed.on = True
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(10): pipeline.varnet_hot_path(mk, mask, v, 12, xf=True)
    s1.record(); torch.cuda.synchronize()
print(f"b{b} t{t} c{c} {h}x{w}: {s0.elapsed_time(s1) / 10:.3f} ms per step, expand+DC {1e3 * sum(a.elapsed_time(z) for a, z in ev) / len(ev):.1f} us per launch")
