#!/usr/bin/env python
"""Quick kernel timings (CUDA events, rotating buffers > L2) for development."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from deep_cine_cardiac_mri_b200 import ops

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2] * 1e-3

def main():
    b, t, c, h, w = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (4, 15, 10, 200, 200))]
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    nrot = 3
    ks = [torch.randn(b, t, c, h, w, 2, device=dev, generator=g) for _ in range(nrot)]
    refs = [torch.randn(b, t, c, h, w, 2, device=dev, generator=g) for _ in range(nrot)]
    s = torch.randn(b, c, h, w, 2, device=dev, generator=g); s = s / s.pow(2).sum(dim=(1, 4), keepdim=True).sqrt()
    x = torch.randn(b, t, h, w, 2, device=dev, generator=g)
    m = (torch.rand(b, t, h, device=dev, generator=g) < 0.25).to(torch.uint8)
    v = torch.tensor([1.0], device=dev)
    K, I, S = ks[0].numel() * 4, x.numel() * 4, s.numel() * 4
    i = [0]
    def nxt(): i[0] = (i[0] + 1) % nrot; return i[0]
    res = {}
    res["fft2c"] = (timeit(lambda: ops.raw_fft2c(ks[nxt()], False, 1)), 2 * K)
    res["ifft2c"] = (timeit(lambda: ops.raw_fft2c(ks[nxt()], True, 1)), 2 * K)
    res["sens_reduce"] = (timeit(lambda: ops.raw_sens_reduce(ks[nxt()], s)), K + S + I)
    res["sens_expand"] = (timeit(lambda: ops.raw_sens_expand(x, s)), K + S + I)
    res["sens_expand_dc"] = (timeit(lambda: ops.raw_sens_expand(x, s, 2, refs[nxt()], m, v)), 2 * K + S + I)
    def dc_step():
        j = nxt()
        img = ops.raw_sens_reduce(ks[j], s)
        return ops.raw_sens_expand(img, s, 2, refs[j], m, v)
    res["dc_step"] = (timeit(dc_step), 3 * K + 2 * S + 2 * I)
    if ops.normal_op_supported(h, w):
        res["normal_op"] = (timeit(lambda: ops.raw_normal_op(x, s, m, v)), 2 * I + S)
    res["dc_blend"] = (timeit(lambda: ops.raw_dc_blend(ks[nxt()], refs[i[0]], m, v)), 3 * K)
    big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    big2 = torch.empty_like(big)
    res["copy_1GiB"] = (timeit(lambda: big2.copy_(big)), 2 * big.numel() * 4)
    l2buf = torch.empty(8 << 20, dtype=torch.float32, device=dev)   # 32 MB, L2 resident
    l2out = torch.empty_like(l2buf)
    res["copy_32MB_L2"] = (timeit(lambda: l2out.copy_(l2buf), n=50), 2 * l2buf.numel() * 4)
    # reference eager (cuFFT) dc step on the same GPU for context
    def ref_fft2c(z, inv=False):
        zc = torch.view_as_complex(z)
        f = torch.fft.ifftn if inv else torch.fft.fftn
        return torch.view_as_real(torch.fft.fftshift(f(torch.fft.ifftshift(zc, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1)))
    def cmul(a, b_): return torch.stack((a[..., 0]*b_[..., 0]-a[..., 1]*b_[..., 1], a[..., 0]*b_[..., 1]+a[..., 1]*b_[..., 0]), -1)
    s6 = s.unsqueeze(1); m6 = m.view(b, t, 1, h, 1, 1)
    def eager_dc():
        j = nxt()
        img = cmul(ref_fft2c(ks[j], True), torch.stack((s6[..., 0], -s6[..., 1]), -1)).sum(2, keepdim=True)
        kx = ref_fft2c(cmul(img, s6))
        return (1 - m6) * kx + m6 * (kx + v * refs[j]) / (1 + v)
    res["eager_cufft_dc_step"] = (timeit(eager_dc, n=10), 3 * K + 2 * S + 2 * I)
    print(f"config b{b} t{t} c{c} {h}x{w}  K={K/1e6:.1f}MB")
    for k_, (sec, byt) in res.items():
        print(f"{k_:22s} {sec*1e6:10.1f} us  {byt/sec/1e9:9.1f} GB/s (algorithmic)")

if __name__ == "__main__":
    main()
