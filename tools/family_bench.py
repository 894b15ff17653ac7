import sys; sys.path.insert(0, "/root/repo")
import torch
from deep_cine_cardiac_mri_b200 import ops
from tools.quick_bench import timeit
dev="cuda"; g=torch.Generator(device=dev).manual_seed(0)
for b in (4, 16):
    t,c,h,w=15,10,200,200
    ks=[torch.randn(b,t,c,h,w,2,device=dev,generator=g) for _ in range(2)]
    s=torch.randn(b,c,h,w,2,device=dev,generator=g); x=torch.randn(b,t,h,w,2,device=dev,generator=g)
    m=(torch.rand(b,t,h,device=dev,generator=g)<0.25).to(torch.uint8); v=torch.tensor([1.0],device=dev)
    i=[0]
    def nxt(): i[0]^=1; return i[0]
    for fam in ("auto","half","packed"):
        ops.set_fused_path(fam)
        r=timeit(lambda: ops.raw_sens_reduce(ks[nxt()], s))
        f=timeit(lambda: ops.raw_fft2c(ks[nxt()], False, 1))
        e=timeit(lambda: ops.raw_sens_expand(x, s))
        d=timeit(lambda: ops.raw_sens_expand(x, s, 2, ks[nxt()], m, v))
        print(f"b{b} {fam:6s}: reduce {r*1e6:7.1f}  fft2c {f*1e6:7.1f}  expand {e*1e6:7.1f}  expand_dc {d*1e6:7.1f} us")
    ops.set_fused_path(None)
