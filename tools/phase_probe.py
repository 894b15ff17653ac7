import ctypes, sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops, _lib
lib = _lib.lib()
b, t, c, h, w = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (4, 15, 10, 200, 200))]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
k = torch.randn(b, t, c, h, w, 2, device=dev, generator=g)
ref = torch.randn(b, t, c, h, w, 2, device=dev, generator=g)
s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
v = torch.tensor([1.0], device=dev)
buf = (ctypes.c_ulonglong * 8)()
lib.b2s_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
def probe(name, fn, n=5):
    fn(); lib.b2s_debug_phase_cycles(buf, 1)
    for _ in range(n): fn()
    lib.b2s_debug_phase_cycles(buf, 1)
    items = n * b * t * c * 2
    vals = [buf[i] / items for i in range(6)]
    print(f"{name:16s} cycles/item: A={vals[0]:8.0f} Bread+dft={vals[1]:8.0f} Bwrite={vals[2]:8.0f} C={vals[3]:8.0f} fix={vals[4]:8.0f} unpark={vals[5]:8.0f} total={sum(vals):8.0f}")
probe("fft2c", lambda: ops.raw_fft2c(k, False, 1))
probe("sens_reduce", lambda: ops.raw_sens_reduce(k, s))
probe("sens_expand", lambda: ops.raw_sens_expand(x, s))
probe("sens_expand_dc", lambda: ops.raw_sens_expand(x, s, 2, ref, m, v))
