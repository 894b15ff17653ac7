#!/usr/bin/env python
"""Training-step smoke test under DDP (BASELINE configs[3] shape of the problem, tiny regulariser):
unrolled cascades = [A^H -> conv regulariser -> A fused with soft-DC (learnable lambda)], SSIM-free L1 loss,
Adam, gradient all-reduce over NCCL.  Run with torchrun on >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_step.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.nn as nn
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from deep_cine_cardiac_mri_b200 import ops, synth, dist as bdist, functional as F


class TinyCascadeNet(nn.Module):
    def __init__(self, n_cascades=3, chans=8):
        super().__init__()
        self.regs = nn.ModuleList([nn.Sequential(nn.Conv2d(2, chans, 3, padding=1), nn.ReLU(), nn.Conv2d(chans, 2, 3, padding=1))
                                   for _ in range(n_cascades)])
        self.lambdas = nn.ParameterList([nn.Parameter(torch.full((1,), 0.5413)) for _ in range(n_cascades)])
        self.softplus = nn.Softplus(1.0)

    def forward(self, masked_kspace, mask, sens):
        k = masked_kspace
        b, t, c, h, w, _ = k.shape
        for reg, lam in zip(self.regs, self.lambdas):
            img = ops.sens_reduce(k, sens)                                           # (b,t,h,w,2)
            x = img.permute(0, 1, 4, 2, 3).reshape(b * t, 2, h, w)
            x = (x + reg(x)).reshape(b, t, 2, h, w).permute(0, 1, 3, 4, 2)
            k = ops.sens_expand(x, sens, ops.EXPAND_DC, ref=masked_kspace, mask=mask, v=self.softplus(lam))
        return F.complex_abs(ops.sens_reduce(k, sens))


def main():
    rank, world, local = bdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = TinyCascadeNet().to(dev)
    ddp = DDP(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-3)
    case = synth.to_torch(synth.cine_case(100 + rank, 1, 15, 10, 200, 200), dev)      # per-rank slice
    target = F.complex_abs(case["image"].squeeze(2))
    times = []
    for step in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = ddp(case["masked_kspace"], case["mask"], case["sens"])
        loss = (out - target).abs().mean()
        loss.backward()
        opt.step()
        torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        if rank == 0:
            print(f"step {step}: loss {loss.item():.6f}  {times[-1]*1e3:.1f} ms", flush=True)
    # gradients (after all-reduce) must be identical on every rank, lambda must receive gradient
    g = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.isfinite(g).all() and float(model.lambdas[0].grad.abs()) > 0
    if world > 1:
        gs = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        assert all(torch.equal(gs[0], x) for x in gs), "gradients differ across ranks"
    if rank == 0:
        print(f"OK world={world} params={sum(p.numel() for p in model.parameters())} median step {sorted(times)[len(times)//2]*1e3:.1f} ms")
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
