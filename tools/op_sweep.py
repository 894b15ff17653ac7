#!/usr/bin/env python
"""BASELINE.json configs[4]: SENSE forward/adjoint operator sweep (coils x frames x size) on one GPU.
Writes a CSV (CUDA-event medians, rotating buffers) with algorithmic GB/s and the roofline fraction."""
import csv, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from deep_cine_cardiac_mri_b200 import ops

def timeit(fn, n=12, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e-3

def main(out_csv):
    peak = 6553.0
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists(): peak = float(json.loads(p.read_text())["hbm_gbs"])
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for hw in (200, 256):
        for c in (8, 16, 32):
            for t in (15, 30):
                b = max(1, int(round(400e6 / (t * c * hw * hw * 8))))          # ~400 MB of k-space per tensor (> L2 with the rotation)
                ks = [torch.randn(b, t, c, hw, hw, 2, device=dev, generator=g) for _ in range(2)]
                s = torch.randn(b, c, hw, hw, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
                x = torch.randn(b, t, hw, hw, 2, device=dev, generator=g)
                m = (torch.rand(b, t, hw, device=dev, generator=g) < 0.25).to(torch.uint8)
                v = torch.ones(1, device=dev)
                K, I, S = ks[0].numel() * 4, x.numel() * 4, s.numel() * 4
                i = [0]
                def nxt(): i[0] ^= 1; return i[0]
                t_red = timeit(lambda: ops.raw_sens_reduce(ks[nxt()], s))
                t_exp = timeit(lambda: ops.raw_sens_expand(x, s))
                t_dc = timeit(lambda: ops.raw_sens_expand(x, s, 2, ks[nxt()], m, v))
                for name, sec, byt in (("sens_reduce", t_red, K + S + I), ("sens_expand", t_exp, K + S + I), ("sens_expand_dc", t_dc, 2 * K + S + I),
                                       ("dc_step", t_red + t_dc, 3 * K + 2 * S + 2 * I)):
                    rows.append(dict(h=hw, w=hw, coils=c, frames=t, batch=b, op=name, us=round(sec * 1e6, 1), algorithmic_MB=round(byt / 1e6, 1),
                                     GBps=round(byt / sec / 1e9, 1), frac_of_measured_hbm=round(byt / sec / 1e9 / peak, 3)))
                del ks, s, x
                torch.cuda.empty_cache()
    with open(out_csv, "w", newline="") as f:
        wr = csv.DictWriter(f, fieldnames=list(rows[0].keys())); wr.writeheader(); wr.writerows(rows)
    for r in rows:
        if r["op"] == "dc_step": print(r)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r1_op_sweep.csv")
