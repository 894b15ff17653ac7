#!/usr/bin/env python
"""BASELINE configs[3]: XT-XPDNet (MWCNN wavelet regulariser) training step under DDP, 1/2/4/8 B200.

The model is the UNMODIFIED reference `reconstruction.models.XPDNet` (baseline/_ref) with the training script's
configuration (traintest_scripts/xpdnet/train_test_xpdnet.py:258-277: 9 cascades, XT, MWCNN n_scales 3, primal-only,
5.57 M parameters = 22.3 MB of fp32 gradient), its SENSE / DC path re-bound onto the b200sense kernels by
`patch_reference()`, the fused SSIM loss (metrics.SSIMLoss), Adam, and stock DistributedDataParallel
(`gradient_as_bucket_view=True`): the gradient all-reduce is NCCL over NVLink and overlaps the adjoint kernels of the
backward pass bucket by bucket.  `--unpatched` runs the same step on the reference's own eager ops for comparison.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_step.py [--unpatched]

Rank 0 prints one JSON line: step ms (device-timed, max over ranks), forward / backward+all-reduce / optimiser split,
the time of an isolated all-reduce of the same gradient size, and a gradient-consistency check across ranks.
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

from deep_cine_cardiac_mri_b200 import dist as bdist, functional as F, metrics, patch, synth
from oracle import load_reference


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--unpatched", action="store_true")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--cascades", type=int, default=9)
    ap.add_argument("--coils", type=int, default=10)
    ap.add_argument("--frames", type=int, default=15)
    args = ap.parse_args()
    rank, world, local = bdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    rec = load_reference.load()
    if not args.unpatched:
        patch.patch_reference()
    torch.manual_seed(0)
    model = rec.models.XPDNet(num_cascades=args.cascades, sens_chans=8, sens_pools=4, n_scales=3, dynamic_type="XT",
                              weight_sharing=False, primal_only=True, n_primal=5).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    ddp = DDP(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-4)
    loss_fn = metrics.SSIMLoss().to(dev)
    case = synth.to_torch(synth.cine_case(100 + rank, 1, args.frames, args.coils, 200, 200), dev)      # one slice per rank
    target = F.complex_abs(case["image"].squeeze(2))                                  # (1,t,h,w)

    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e

    rows = []
    for step in range(args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0 = ev()
        opt.zero_grad(set_to_none=True)
        out = ddp(case["masked_kspace"], case["mask"])
        loss = loss_fn(out.unsqueeze(1), target.unsqueeze(1))                        # (b,1,t,h,w), varnet_module.py:110-112
        e1 = ev()
        loss.backward()
        e2 = ev()
        opt.step()
        e3 = ev()
        torch.cuda.synchronize()
        rows.append((e0.elapsed_time(e3), e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), float(loss)))
    rows = rows[2:]                                                                   # warm-up (cudnn.benchmark, allocator)
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(4)]
    step_ms = bdist.max_over_ranks(med[0], dev) if world > 1 else med[0]

    # an isolated all-reduce of the same payload (what the backward pass has to hide)
    ar_ms = None
    if world > 1:
        buf = torch.zeros(n_params, device=dev)
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        a0 = ev()
        for _ in range(10):
            dist.all_reduce(buf)
        a1 = ev()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 10

    g = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
    ok = bool(torch.isfinite(g).all())
    lam = [float(p.grad.abs().sum()) for n, p in model.named_parameters() if p.grad is not None and "sens_net" in n][:1]
    same = True
    if world > 1:
        gs = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        same = all(torch.equal(gs[0], x) for x in gs)
    if rank == 0:
        print(json.dumps({
            "config": f"XT-XPDNet {args.cascades} cascades, MWCNN n_scales 3, primal-only, {args.coils}-coil {args.frames}-frame 200x200, one slice per rank",
            "ops": "reference eager" if args.unpatched else "b200sense (patch_reference)", "world": world, "params": n_params,
            "grad_bytes": 4 * n_params, "step_ms": round(step_ms, 2), "forward_ms": round(med[1], 2),
            "backward_incl_allreduce_ms": round(med[2], 2), "optimizer_ms": round(med[3], 2),
            "isolated_allreduce_ms": None if ar_ms is None else round(ar_ms, 3),
            "slices_per_sec": round(world / (step_ms * 1e-3), 2), "loss": rows[-1][4], "grads_finite": ok, "sens_net_grad_nonzero": bool(lam and lam[0] > 0),
            "grads_identical_across_ranks": same}), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
